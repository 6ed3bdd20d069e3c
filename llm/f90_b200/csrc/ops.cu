// ops.cu -- the inner subroutines of llama2.f90's transformer() as separate sm_100a kernels.
//
// These back the granular C-ABI operators (llmf90_b200_matvec/_rmsnorm/_softmax/_rope), the
// LLMF90_FLAG_GRANULAR forward (one kernel per step, the shape a tensor-parallel run needs
// around its NCCL all-reduces) and the one-time upload re-layout.  The single-GPU hot path is
// the fused streaming kernel in stream.cu.
#include "kernels.cuh"

namespace llmf90 {

// ------------------------------------------------------------------ rmsnorm (llama2.f90:450-457)
__global__ void __launch_bounds__(1024) rmsnorm_kernel(const float *__restrict__ x,
                                                       const float *__restrict__ w,
                                                       float *__restrict__ out, int n)
{
    __shared__ float red[32];
    __shared__ float s_xn;
    float ss = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float v = x[i];
        ss = fmaf(v, v, ss);
    }
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
        t = warp_sum(t);
        if (threadIdx.x == 0) s_xn = sqrtf(t / (float)n + 1e-5f);
    }
    __syncthreads();
    const float xn = s_xn;
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = x[i] * w[i] / xn;
}

cudaError_t launch_rmsnorm(const float *x, const float *w, float *out, int n, cudaStream_t st)
{
    rmsnorm_kernel<<<1, 1024, 0, st>>>(x, w, out, n);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ softmax (llama2.f90:468-478)
__global__ void __launch_bounds__(1024) softmax_kernel(const float *__restrict__ x,
                                                       float *__restrict__ p, int n, int s)
{
    __shared__ float red[32];
    __shared__ float bcast;
    const int nw = blockDim.x >> 5;
    float mx = -INFINITY;
    for (int i = threadIdx.x; i < s; i += blockDim.x) mx = fmaxf(mx, x[i]);
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = threadIdx.x < nw ? red[threadIdx.x] : -INFINITY;
        t = warp_max(t);
        if (threadIdx.x == 0) bcast = t;
    }
    __syncthreads();
    mx = bcast;
    float sum = 0.f;
    for (int i = threadIdx.x; i < s; i += blockDim.x) {
        const float e = expf(x[i] - mx);
        p[i] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = threadIdx.x < nw ? red[threadIdx.x] : 0.f;
        t = warp_sum(t);
        if (threadIdx.x == 0) bcast = t;
    }
    __syncthreads();
    sum = bcast;
    for (int i = threadIdx.x; i < s; i += blockDim.x) p[i] = p[i] / sum;
    for (int i = s + threadIdx.x; i < n; i += blockDim.x) p[i] = 0.f;
}

cudaError_t launch_softmax(const float *x, float *p, int n, int s, cudaStream_t st)
{
    softmax_kernel<<<1, 1024, 0, st>>>(x, p, n, s);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ mat-vec
// One warp per NR consecutive rows, activation vector staged once per CTA in shared memory,
// 128-bit coalesced weight loads with the dequantisation fused into the load, warp-shuffle
// row reduction (llama2.f90:529-531 and the four other inline dot_product loops).
constexpr int MV_WARPS = 8;

template <int WT, int NR>
__global__ void __launch_bounds__(MV_WARPS * 32) matvec_kernel(const uint8_t *__restrict__ W,
                                                               size_t rs, int rows, int cols,
                                                               const float *__restrict__ x,
                                                               const float *residual, float *y)
{
    extern __shared__ __align__(16) float xs[];
    for (int e = threadIdx.x; e < cols; e += blockDim.x) xs[xs_index<WT>(e)] = x[e];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r0 = (blockIdx.x * MV_WARPS + warp) * NR;
    if (r0 >= rows) return;
    float acc[NR];
#pragma unroll
    for (int i = 0; i < NR; i++) acc[i] = 0.f;
    if (r0 + NR <= rows) {
        dot_rows<WT, NR>(W + (size_t)r0 * rs, rs, xs, cols, lane, acc);
    } else {
        for (int i = 0; i < rows - r0; i++) {
            float a1[1] = {0.f};
            dot_rows<WT, 1>(W + (size_t)(r0 + i) * rs, rs, xs, cols, lane, a1);
#pragma unroll
            for (int k = 0; k < NR; k++)
                if (k == i) acc[k] = a1[0];
        }
    }
#pragma unroll
    for (int i = 0; i < NR; i++) {
        const float v = warp_sum(acc[i]);
        if (lane == 0 && r0 + i < rows) y[r0 + i] = (residual ? residual[r0 + i] : 0.f) + v;
    }
}

template <int WT, int NR>
static cudaError_t launch_matvec_t(const uint8_t *W, int rows, int cols, const float *x,
                                   const float *residual, float *y, cudaStream_t st)
{
    const size_t smem = (size_t)cols * sizeof(float);
    // (function attributes are per device: one flag per device, not one per process)
    static bool attr_done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(matvec_kernel<WT, NR>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
    const int grid = (rows + MV_WARPS * NR - 1) / (MV_WARPS * NR);
    matvec_kernel<WT, NR><<<grid, MV_WARPS * 32, smem, st>>>(W, row_stride_bytes(WT, cols), rows,
                                                             cols, x, residual, y);
    return cudaGetLastError();
}

cudaError_t launch_matvec(const uint8_t *W, int wtype, int rows, int cols, const float *x,
                          const float *residual, float *y, cudaStream_t st)
{
    if (wtype == WT_F32) return launch_matvec_t<WT_F32, 2>(W, rows, cols, x, residual, y, st);
    if (wtype == WT_F16) return launch_matvec_t<WT_F16, 2>(W, rows, cols, x, residual, y, st);
    if (wtype == WT_Q4_0) return launch_matvec_t<WT_Q4_0, 4>(W, rows, cols, x, residual, y, st);
    return cudaErrorInvalidValue;
}

// ------------------------------------------------------------------ Q6_K mat-vec (classifier of stock llama.cpp q4_0 files)
// W is `rows` rows of cols / 256 ggml block_q6_K super-blocks, 210 bytes each, exactly as in the file:
// ql[128] (low 4 bits) | qh[64] (high 2 bits) | int8 scales[16] | f16 d;  weight = d * scale * (q - 32).
// One warp per row; for each super-block lane l takes the elements l, l + 32, l + 64, l + 96 of both halves (the
// loop of ggml's dequantize_row_q6_K with l = lane): byte loads that are contiguous across the warp.  The product
// d * scale * q is exact in f32 (11 + 7 + 6 bits), so this is the f32 dot product of the dequantised row.
__global__ void __launch_bounds__(MV_WARPS * 32) matvec_q6k_kernel(const uint8_t *__restrict__ W, int rows, int cols,
                                                                   const float *__restrict__ x, float *__restrict__ y)
{
    extern __shared__ __align__(16) float xs[];
    for (int e = threadIdx.x; e < cols; e += blockDim.x) xs[e] = x[e];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * MV_WARPS + warp;
    if (r >= rows) return;
    const int nb = cols >> 8, is = lane >> 4;
    const uint8_t *blk = W + (size_t)r * nb * 210;
    float acc = 0.f;
    for (int b = 0; b < nb; b++, blk += 210) {
        const float d = __half2float(__ushort_as_half((unsigned short)(blk[208] | (blk[209] << 8))));
#pragma unroll
        for (int half = 0; half < 2; half++) {
            const uint8_t *ql = blk + 64 * half, *qh = blk + 128 + 32 * half;
            const int8_t *sc = reinterpret_cast<const int8_t *>(blk + 192 + 8 * half);
            const float *xv = xs + b * 256 + 128 * half + lane;
            const int l0 = ql[lane], l1 = ql[lane + 32], h = qh[lane];
            const int q1 = ((l0 & 0xF) | (((h >> 0) & 3) << 4)) - 32;
            const int q2 = ((l1 & 0xF) | (((h >> 2) & 3) << 4)) - 32;
            const int q3 = ((l0 >> 4) | (((h >> 4) & 3) << 4)) - 32;
            const int q4 = ((l1 >> 4) | (((h >> 6) & 3) << 4)) - 32;
            acc = fmaf(d * (float)sc[is] * (float)q1, xv[0], acc);
            acc = fmaf(d * (float)sc[is + 2] * (float)q2, xv[32], acc);
            acc = fmaf(d * (float)sc[is + 4] * (float)q3, xv[64], acc);
            acc = fmaf(d * (float)sc[is + 6] * (float)q4, xv[96], acc);
        }
    }
    acc = warp_sum(acc);
    if (lane == 0) y[r] = acc;
}

cudaError_t launch_matvec_q6k(const uint8_t *W, int rows, int cols, const float *x, float *y, cudaStream_t st)
{
    if (cols % 256 || (size_t)cols * 4 > 48 * 1024) return cudaErrorInvalidValue;  // 12288 columns of activations in default shared memory
    matvec_q6k_kernel<<<(rows + MV_WARPS - 1) / MV_WARPS, MV_WARPS * 32, (size_t)cols * 4, st>>>(W, rows, cols, x, y);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ RoPE (llama2.f90:543-559)
// Q1: exponent (2j+1)/hs (the reference's 1-based odd loop index through mod(i,head_size));
// Q2: angle = pos * freq with the 1-based pos.  Accurate powf/cosf/sinf (no fast-math).
__device__ __forceinline__ float2 rope_cs(int j, int hs, int pos1)
{
    const float freq = 1.0f / powf(10000.0f, (float)(2 * j + 1) / (float)hs);
    const float rval = (float)pos1 * freq;
    return make_float2(cosf(rval), sinf(rval));
}

__global__ void rope_table_kernel(float2 *tab, int seq, int hs)
{
    const int half = hs >> 1;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= seq * half) return;
    tab[i] = rope_cs(i % half, hs, i / half + 1);
}

cudaError_t launch_rope_table(float2 *tab, int seq, int hs, cudaStream_t st)
{
    const int n = seq * (hs / 2);
    rope_table_kernel<<<(n + 255) / 256, 256, 0, st>>>(tab, seq, hs);
    return cudaGetLastError();
}

__global__ void rope_kernel(float *q, float *k, int emb, int kv, int hs, int pos)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;  // pair index
    if (2 * p >= emb) return;
    const float2 cs = rope_cs(p % (hs >> 1), hs, pos);
    const float q0 = q[2 * p], q1 = q[2 * p + 1];
    q[2 * p] = q0 * cs.x - q1 * cs.y;
    q[2 * p + 1] = q0 * cs.y + q1 * cs.x;
    if (2 * p + 1 < kv) {  // i < kv_head_size with i = 2p+1 (llama2.f90:553)
        const float k0 = k[2 * p], k1 = k[2 * p + 1];
        k[2 * p] = k0 * cs.x - k1 * cs.y;
        k[2 * p + 1] = k0 * cs.y + k1 * cs.x;
    }
}

cudaError_t launch_rope(float *q, float *k, int emb, int kv, int hs, int pos, cudaStream_t st)
{
    const int n = emb / 2;
    rope_kernel<<<(n + 127) / 128, 128, 0, st>>>(q, k, emb, kv, hs, pos);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ embedding gather (llama2.f90:520)
__global__ void embed_kernel(const uint8_t *__restrict__ table, int wtype, int cols, size_t rs,
                             const int *__restrict__ tokpos, float *__restrict__ x)
{
    const uint8_t *row = table + (size_t)(tokpos[0] - 1) * rs;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < cols; e += gridDim.x * blockDim.x)
        x[e] = row_elem(row, wtype, cols, e);
}

cudaError_t launch_embed(const uint8_t *table, int wtype, int cols, const int *tokpos, float *x,
                         cudaStream_t st)
{
    embed_kernel<<<(cols + 255) / 256, 256, 0, st>>>(table, wtype, cols,
                                                     row_stride_bytes(wtype, cols), tokpos, x);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ RoPE + KV append (llama2.f90:543-565)
__global__ void rope_kv_kernel(float *__restrict__ qkv, int emb, int kv, int hs,
                               const float2 *__restrict__ tab, const int *__restrict__ tokpos,
                               float *__restrict__ kc_layer, float *__restrict__ vc_layer)
{
    const int pos0 = tokpos[1] - 1;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;  // pair index over q | k | v
    const int half = hs >> 1;
    const int nq = emb >> 1, nk = kv >> 1;
    if (p < nq) {
        const float2 cs = tab[pos0 * half + p % half];
        const float a = qkv[2 * p], b = qkv[2 * p + 1];
        qkv[2 * p] = a * cs.x - b * cs.y;
        qkv[2 * p + 1] = a * cs.y + b * cs.x;
    } else if (p < nq + nk) {
        const int pk = p - nq;
        const float2 cs = tab[pos0 * half + pk % half];
        const float a = qkv[emb + 2 * pk], b = qkv[emb + 2 * pk + 1];
        float *dst = kc_layer + (size_t)pos0 * kv;
        dst[2 * pk] = a * cs.x - b * cs.y;
        dst[2 * pk + 1] = a * cs.y + b * cs.x;
    } else if (p < nq + 2 * nk) {
        const int pv = p - nq - nk;
        float *dst = vc_layer + (size_t)pos0 * kv;
        dst[2 * pv] = qkv[emb + kv + 2 * pv];
        dst[2 * pv + 1] = qkv[emb + kv + 2 * pv + 1];
    }
}

cudaError_t launch_rope_kv(float *qkv, int emb, int kv, int hs, const float2 *tab,
                           const int *tokpos, float *kc_layer, float *vc_layer, cudaStream_t st)
{
    const int n = emb / 2 + kv;
    rope_kv_kernel<<<(n + 127) / 128, 128, 0, st>>>(qkv, emb, kv, hs, tab, tokpos, kc_layer, vc_layer);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ attention (llama2.f90:574-598)
// One CTA per query head: scores over t = 1..pos (kv head = h / kv_mul, quirk Q3), softmax,
// weighted sum of the value rows.  Same three steps as the reference.
constexpr int ATT_THREADS = 128;

__global__ void __launch_bounds__(ATT_THREADS) attention_kernel(
    const float *__restrict__ q, const float *__restrict__ kc, const float *__restrict__ vc,
    const int *__restrict__ tokpos, float *__restrict__ out, int kv_mul, int hs, int kv)
{
    extern __shared__ float att[];  // [pos]
    __shared__ float red[ATT_THREADS / 32];
    __shared__ float bcast;
    const int pos = tokpos[1];
    const int h = blockIdx.x, g = h / kv_mul;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = ATT_THREADS / 32;
    const float *qh = q + (size_t)h * hs;
    const float scale = sqrtf((float)hs);

    for (int t = warp; t < pos; t += nw) {
        const float *kt = kc + (size_t)t * kv + (size_t)g * hs;
        float s = 0.f;
        for (int d = lane; d < hs; d += 32) s = fmaf(qh[d], kt[d], s);
        s = warp_sum(s);
        if (lane == 0) att[t] = s / scale;
    }
    __syncthreads();
    float mx = -INFINITY;
    for (int t = threadIdx.x; t < pos; t += ATT_THREADS) mx = fmaxf(mx, att[t]);
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = red[0];
        for (int i = 1; i < nw; i++) m = fmaxf(m, red[i]);
        bcast = m;
    }
    __syncthreads();
    mx = bcast;
    float sum = 0.f;
    for (int t = threadIdx.x; t < pos; t += ATT_THREADS) {
        const float e = expf(att[t] - mx);
        att[t] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    __syncthreads();
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < nw; i++) s += red[i];
        bcast = s;
    }
    __syncthreads();
    sum = bcast;
    for (int t = threadIdx.x; t < pos; t += ATT_THREADS) att[t] = att[t] / sum;  // softmax, :476
    __syncthreads();
    for (int d = threadIdx.x; d < hs; d += ATT_THREADS) {
        float a = 0.f;
        for (int t = 0; t < pos; t++) a = fmaf(att[t], vc[(size_t)t * kv + (size_t)g * hs + d], a);
        out[(size_t)h * hs + d] = a;
    }
}

cudaError_t launch_attention(const float *q, const float *kc_layer, const float *vc_layer,
                             const int *tokpos, float *out, int n_heads, int kv_mul, int hs, int kv,
                             int seq, cudaStream_t st)
{
    const size_t smem = (size_t)seq * sizeof(float);  // one score per position
    if (smem > 48 * 1024) {
        // beyond the default 48 KB a kernel has to opt in (seq_len > 12288); the attribute is per device
        if (smem > 200 * 1024) return cudaErrorInvalidValue;
        static bool attr_done[64] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64 || !attr_done[dev]) {
            cudaError_t e = cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            if (e != cudaSuccess) return e;
            if (dev >= 0 && dev < 64) attr_done[dev] = true;
        }
    }
    attention_kernel<<<n_heads, ATT_THREADS, smem, st>>>(q, kc_layer, vc_layer, tokpos, out, kv_mul, hs, kv);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ SwiGLU (llama2.f90:613-616)
__global__ void swiglu_kernel(const float *__restrict__ h13, float *__restrict__ hb, int hid)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hid) return;
    const float g = h13[2 * i], u = h13[2 * i + 1];
    hb[i] = (g * (1.0f / (1.0f + expf(-g)))) * u;
}

cudaError_t launch_swiglu(const float *h13, float *hb, int hid, cudaStream_t st)
{
    swiglu_kernel<<<(hid + 255) / 256, 256, 0, st>>>(h13, hb, hid);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ maxloc (llama2.f90:388)
__global__ void __launch_bounds__(1024) argmax_kernel(const float *__restrict__ v, int n,
                                                      int *__restrict__ out)
{
    __shared__ float bv[32];
    __shared__ int bi[32];
    float best = -INFINITY;
    int idx = 0x7fffffff;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float a = v[i];
        if (a > best) { best = a; idx = i; }  // ascending i per thread keeps the first maximum
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ob > best || (ob == best && oi < idx)) { best = ob; idx = oi; }
    }
    if ((threadIdx.x & 31) == 0) { bv[threadIdx.x >> 5] = best; bi[threadIdx.x >> 5] = idx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); w++)
            if (bv[w] > best || (bv[w] == best && bi[w] < idx)) { best = bv[w]; idx = bi[w]; }
        out[0] = idx + 1;
    }
}

cudaError_t launch_argmax(const float *v, int n, int *out_token, cudaStream_t st)
{
    argmax_kernel<<<1, 1024, 0, st>>>(v, n, out_token);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ softmax(logits / T) + CDF walk (llama2.f90:390-391, :428-447)
// The reference divides the logits by the temperature, takes the softmax and returns the first index whose
// running sum of probabilities exceeds a uniform r (the last index if none does).  Here one CTA does it next to
// the logits, so a sampled token costs a 4-byte copy instead of the vocabulary's logits: max and sum block-wide,
// then every thread sums the probabilities of its CONTIGUOUS chunk, a block-wide scan of the 1024 chunk sums finds
// the chunk in which the running sum crosses r, and that thread walks its chunk element by element.  The sums
// are f32 like the reference's but associate differently (chunk sums instead of one long chain): the pick can
// differ from the sequential walk only when r lies within rounding (~1e-6) of a boundary of the CDF.
__global__ void __launch_bounds__(1024) sample_kernel(const float *__restrict__ logits, int n, float temperature, float r,
                                                      int *__restrict__ out)
{
    __shared__ float red[32];
    __shared__ float bcast;
    __shared__ float chunk_incl[1024];
    __shared__ int pick;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float mx = -INFINITY;
    for (int i = tid; i < n; i += 1024) mx = fmaxf(mx, logits[i] / temperature);
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    if (warp == 0) {
        float t = warp_max(red[lane]);
        if (lane == 0) bcast = t;
    }
    __syncthreads();
    mx = bcast;
    float sum = 0.f;
    for (int i = tid; i < n; i += 1024) sum += expf(logits[i] / temperature - mx);
    sum = warp_sum(sum);
    __syncthreads();
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    if (warp == 0) {
        float t = warp_sum(red[lane]);
        if (lane == 0) bcast = t;
    }
    if (tid == 0) pick = n;  // the reference's fallback: the last index (llama2.f90:444)
    __syncthreads();
    sum = bcast;
    // chunk sums and their inclusive scan (warp scan, then the 32 warp totals)
    const int c = (n + 1023) / 1024, i0 = tid * c, i1 = min(n, i0 + c);
    float cs = 0.f;
    for (int i = i0; i < i1; i++) cs += expf(logits[i] / temperature - mx) / sum;
    float incl = cs;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    __syncthreads();
    if (lane == 31) red[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        float w = red[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float v = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += v;
        }
        red[lane] = w;  // inclusive totals of warps 0 .. lane
    }
    __syncthreads();
    incl += warp > 0 ? red[warp - 1] : 0.f;
    chunk_incl[tid] = incl;
    __syncthreads();
    // the first chunk whose inclusive sum exceeds r holds the answer (the running sum is non-decreasing)
    const float before = tid > 0 ? chunk_incl[tid - 1] : 0.f;
    if (r < incl && !(r < before) && i0 < n) {
        float cdf = before;
        int ans = i1;  // not reached: r < incl means the walk crosses inside the chunk (up to rounding)
        for (int i = i0; i < i1; i++) {
            cdf += expf(logits[i] / temperature - mx) / sum;
            if (r < cdf) { ans = i + 1; break; }
        }
        pick = min(ans, n);
    }
    __syncthreads();
    if (tid == 0) out[0] = pick;
}

cudaError_t launch_sample(const float *logits, int n, float temperature, float r, int *out_token, cudaStream_t st)
{
    sample_kernel<<<1, 1024, 0, st>>>(logits, n, temperature, r, out_token);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ upload re-layout
__device__ __forceinline__ int map_row(int r, int map_kind, int row0, int half)
{
    if (map_kind == 1) return (r & 1) ? half + row0 + (r >> 1) : row0 + (r >> 1);
    return row0 + r;
}

// f32 / f16: one thread per element
template <typename T>
__global__ void repack_plain_kernel(const T *__restrict__ src, int src_cols, T *__restrict__ dst,
                                    int dst_rows, int col0, int ncols, int map_kind, int row0,
                                    int half)
{
    const size_t n = (size_t)dst_rows * ncols;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / ncols), c = (int)(i % ncols);
        dst[i] = src[(size_t)map_row(r, map_kind, row0, half) * src_cols + col0 + c];
    }
}

// q4_0: one thread per block; ggml 18-byte block -> nibble plane + scale plane
__global__ void repack_q4_kernel(const uint8_t *__restrict__ src, int src_cols,
                                 uint8_t *__restrict__ dst, int dst_rows, int col0, int ncols,
                                 int map_kind, int row0, int half)
{
    const int nb = ncols >> 5, src_nb = src_cols >> 5, b0 = col0 >> 5;
    const size_t rs = row_stride_bytes(WT_Q4_0, ncols);
    const size_t n = (size_t)dst_rows * nb;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / nb), j = (int)(i % nb);
        const uint8_t *s = src + ((size_t)map_row(r, map_kind, row0, half) * src_nb + b0 + j) * 18;
        uint8_t *d = dst + (size_t)r * rs;
        for (int k = 0; k < 16; k++) d[(size_t)j * 16 + k] = s[2 + k];
        d[(ncols >> 1) + 2 * j] = s[0];
        d[(ncols >> 1) + 2 * j + 1] = s[1];
    }
}

// q4_0 -> tiled mma format (common.cuh): one thread per (row, block)
__global__ void repack_q4_tiled_kernel(const uint8_t *__restrict__ src, int src_cols,
                                       uint8_t *__restrict__ dst, int dst_rows, int col0, int ncols,
                                       int map_kind, int row0, int half)
{
    const int nb = ncols >> 5, src_nb = src_cols >> 5, b0 = col0 >> 5, ngrp = q4t_groups(ncols);
    const size_t n = (size_t)dst_rows * nb;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / nb), j = (int)(i % nb);
        const uint8_t *s = src + ((size_t)map_row(r, map_kind, row0, half) * src_nb + b0 + j) * 18;
        uint8_t *grp = dst + ((size_t)(r >> 4) * ngrp + (j >> 3)) * Q4T_GROUP_BYTES;
        const int g = r & 7, hi_row = (r >> 3) & 1, jb = j & 7;
        const int chunk = 2 * hi_row + (jb >> 2);
        for (int t = 0; t < 4; t++) {
            uint8_t *d = grp + chunk * 512 + (g * 4 + t) * 16 + (jb & 3) * 4;
            for (int k = 0; k < 4; k++) d[k] = s[2 + 4 * t + k];
        }
        uint8_t *sc = grp + 2048 + (g * 4 + (jb >> 1)) * 8 + (2 * hi_row + (jb & 1)) * 2;
        sc[0] = s[0];
        sc[1] = s[1];
    }
}

cudaError_t launch_repack_q4_tiled(const uint8_t *src, int src_cols, uint8_t *dst, int dst_rows, int col0,
                                   int ncols, int map_kind, int row0, int half, cudaStream_t st)
{
    if (dst_rows == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(dst, 0, q4t_matrix_bytes(dst_rows, ncols), st);
    if (e != cudaSuccess) return e;
    repack_q4_tiled_kernel<<<148 * 8, 256, 0, st>>>(src, src_cols, dst, dst_rows, col0, ncols, map_kind, row0, half);
    return cudaGetLastError();
}

cudaError_t launch_repack(const uint8_t *src, int wtype, int src_cols, uint8_t *dst, int dst_rows,
                          int col0, int ncols, int map_kind, int row0, int half, cudaStream_t st)
{
    if (dst_rows == 0) return cudaSuccess;
    const int grid = 148 * 8;
    if (wtype == WT_F32)
        repack_plain_kernel<float><<<grid, 256, 0, st>>>((const float *)src, src_cols, (float *)dst,
                                                         dst_rows, col0, ncols, map_kind, row0, half);
    else if (wtype == WT_F16)
        repack_plain_kernel<__half><<<grid, 256, 0, st>>>((const __half *)src, src_cols,
                                                          (__half *)dst, dst_rows, col0, ncols,
                                                          map_kind, row0, half);
    else {
        // zero the scale-plane padding so whole rows are deterministic
        cudaError_t e = cudaMemsetAsync(dst, 0, (size_t)dst_rows * row_stride_bytes(WT_Q4_0, ncols), st);
        if (e != cudaSuccess) return e;
        repack_q4_kernel<<<grid, 256, 0, st>>>(src, src_cols, dst, dst_rows, col0, ncols, map_kind,
                                               row0, half);
    }
    return cudaGetLastError();
}

}  // namespace llmf90

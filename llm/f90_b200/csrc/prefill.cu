// prefill.cu -- batched prompt pass: the P forced prompt positions of
// llama2.f90:379-385 in ONE pass over the weights instead of P per-token forwards.
//
// What the reference does with a prompt: positions 1..n_prompt run through transformer() one at a time and
// their logits are thrown away (llama2.f90:383-385 overwrites the pick with the prompt token); the only thing
// those calls leave behind is the KV cache.  So the batched pass computes exactly that -- the key / value rows
// of P positions in every layer -- and the first position whose logits matter goes through the decode kernel.
//
// Shape of the work: per layer four GEMMs  Y[P x N] = X[P x K] . W^T  (QKV, Wo, W1|W3, W2) with P <= 128.  Each
// weight byte is still read once, so the pass costs about one token time whatever P is.  The GEMMs run on the
// 5th-generation tensor cores (tcgen05.mma, accumulators in tensor memory):
//   * "swap AB": the WEIGHT tile is the A operand (M = 128 weight rows = the 128 TMEM lanes), the activations are
//     the B operand (N = P padded to a multiple of 16), D[128 x P] f32 in TMEM; one CTA per 128-row tile and
//     K split (grid.y) so that small matrices still cover the SMs; the K splits' partial products are summed in
//     a fixed order by the next (element-wise) kernel -- no atomics, bit-reproducible;
//   * precision: f16 tensor-core products of f16 hi + lo planes accumulate in f32.  Activations are always
//     split x = hi + lo; f32 and q4_0 weights are stored as exact hi + lo planes too (a q4_0 weight d.(q-8) has
//     at most 15 significant bits), f16 weights are one plane: D = Wh.Xh + Wh.Xl (+ Wl.Xh), which keeps the K / V
//     rows within the 1e-4 the per-token path is held to;
//   * operands live in HBM already in the order the tensor core wants them in shared memory (K-major, no
//     swizzle: 8-row x 16-byte core matrices, [k-chunk of 64][plane][k-core 8][row][16 B]), so a pipeline stage
//     is two 1-D bulk copies (weights 16 / 32 KB, activations P x 256 B) onto an mbarrier -- the same TMA path
//     the decode kernel streams with, no tensor maps;
//   * roles per CTA (192 threads): warp 0 = producer (one lane issues the bulk copies), warp 1 = MMA (allocates
//     TMEM, one lane issues tcgen05.mma and commits stages back to the producer), warps 2-5 = epilogue
//     (tcgen05.ld of their 32 lanes, coalesced stores of Y^T).
// RoPE (quirks Q1 / Q2 through the same table as decode), the KV append, causal attention over the cache
// (Q3: kv head = h / kv_mul), rmsnorm, SwiGLU and the residual adds are small batched kernels around the GEMMs.
//
// This costs a second copy of the four layer matrices (f16 planes in operand order): a B200's 180 GB hold
// Llama-2-7B twice over; it is opt-in (LLMF90_FLAG_PREFILL) and single-GPU.
#include <algorithm>
#include <cstdlib>

#include "kernels.cuh"

namespace llmf90 {

namespace {

constexpr int PF_MAXP = 128;      // positions per pass (= the UMMA N limit we use; longer prompts go in chunks)
constexpr int PF_BK = 64;         // contraction elements per pipeline stage
constexpr int PF_TILE_M = 128;    // weight rows per CTA = TMEM lanes
constexpr int PF_A_PLANE = PF_TILE_M * PF_BK * 2;  // 16384 bytes of one weight plane per stage
constexpr int PF_THREADS = 192;
constexpr int PF_MAX_STAGES = 8;
constexpr int PF_MAX_SPLIT = 16;

// ------------------------------------------------------------------ tcgen05 wrappers
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t *slot_smem, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// shared-memory matrix descriptor, K-major, no swizzle: core matrices (8 rows x 16 bytes, 128 contiguous bytes)
// `lbo` bytes apart along K (the "leading" offset) and `sbo` bytes apart along M / N (the "stride" offset); version 1
// (sm_100).  The first hardware run tried the other reading of the two offsets as well: it faults.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// D[tmem] (+)= A[smem desc] . B[smem desc], kind::f16 (f16 inputs, f32 accumulate), issued by ONE thread
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier when every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s_plain(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ------------------------------------------------------------------ the GEMM
struct GemmArgs {
    const uint8_t *A;      // packed weights of this layer: [m-tile][k-chunk][plane][k-core 8][128 rows][16 B]
    const uint8_t *B;      // packed activations: [k-chunk][plane hi, lo][k-core 8][ppad rows][16 B]
    float *Y;              // partial products [split][ppad][N]
    int N, P, ppad;        // weight rows, positions, positions padded to a multiple of 16
    int kc, cps;           // k-chunks in total, k-chunks per split (grid.y = ceil(kc / cps))
    int nst;               // pipeline stages
    int tmem_cols;         // power of two >= max(32, ppad)
};

template <int NPW>  // weight planes: 1 (f16 weights) or 2 (hi + lo of f32 / q4_0 weights)
__global__ void __launch_bounds__(PF_THREADS) umma_gemm_kernel(const GemmArgs a)
{
    extern __shared__ __align__(128) uint8_t pf_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mtile = blockIdx.x, split = blockIdx.y;
    const int c0 = split * a.cps, c1 = min(a.kc, c0 + a.cps);
    const uint32_t a_bytes = NPW * PF_A_PLANE, b_plane = (uint32_t)a.ppad * 128u, b_bytes = 2u * b_plane;
    const uint32_t st_bytes = a_bytes + b_bytes;
    uint64_t *full = reinterpret_cast<uint64_t *>(pf_smem + (size_t)a.nst * st_bytes);
    uint64_t *empty = full + PF_MAX_STAGES, *done = empty + PF_MAX_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(done + 1);

    if (threadIdx.x == 0) {
        for (int s = 0; s < a.nst; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(done, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)a.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            const uint64_t pol = l2_policy_evict_first();  // weights are read once; the activations stay in L2
            const uint8_t *asrc = a.A + ((size_t)mtile * a.kc + c0) * a_bytes;
            const uint8_t *bsrc = a.B + (size_t)c0 * b_bytes;
            for (int c = c0, i = 0; c < c1; c++, i++, asrc += a_bytes, bsrc += b_bytes) {
                const int s = i % a.nst, u = i / a.nst;
                if (u > 0) mbar_wait(&empty[s], (uint32_t)(u - 1) & 1u, 41);
                uint8_t *dst = pf_smem + (size_t)s * st_bytes;
                mbar_arrive_expect_tx(&full[s], st_bytes);
                bulk_g2s(dst, asrc, a_bytes, &full[s], pol);
                bulk_g2s_plain(dst + a_bytes, bsrc, b_bytes, &full[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor: D f32 (bit 4), A / B f16 K-major (zeros), N >> 3 at bit 17, M >> 4 at bit 24
            const uint32_t idesc = (1u << 4) | ((uint32_t)(a.ppad >> 3) << 17) | ((uint32_t)(PF_TILE_M >> 4) << 24);
            // core matrices: along K (the "leading" offset) a whole column of rows apart, along M / N 128 bytes
            const uint32_t a_lbo = PF_TILE_M * 16, a_sbo = 128, b_lbo = (uint32_t)a.ppad * 16u, b_sbo = 128;
            uint32_t acc = 0;
            for (int c = c0, i = 0; c < c1; c++, i++) {
                const int s = i % a.nst, u = i / a.nst;
                mbar_wait(&full[s], (uint32_t)u & 1u, 42);
                tc_fence_after();
                const uint32_t sa = smem_u32(pf_smem + (size_t)s * st_bytes), sb = sa + a_bytes;
#pragma unroll
                for (int k = 0; k < PF_BK / 16; k++) {  // one tcgen05.mma covers 16 contraction elements = 2 core matrices
                    const uint32_t ak = sa + (uint32_t)k * 2u * (PF_TILE_M * 16), bk = sb + (uint32_t)k * 2u * ((uint32_t)a.ppad * 16u);
                    const uint64_t a_hi = umma_desc(ak, a_lbo, a_sbo);
                    const uint64_t b_hi = umma_desc(bk, b_lbo, b_sbo), b_lo = umma_desc(bk + b_plane, b_lbo, b_sbo);
                    umma_f16(tmem, a_hi, b_lo, idesc, acc);
                    acc = 1;
                    if (NPW == 2) umma_f16(tmem, umma_desc(ak + PF_A_PLANE, a_lbo, a_sbo), b_hi, idesc, 1u);
                    umma_f16(tmem, a_hi, b_hi, idesc, 1u);
                }
                umma_commit(&empty[s]);  // the stage's operands have been read when these MMAs complete
            }
            umma_commit(done);
        }
    } else {
        // epilogue: a warp may read the 32 TMEM lanes of its quarter (warp id mod 4); lane = weight row, column = position
        mbar_wait(done, 0, 43);
        tc_fence_after();
        const int q = warp & 3;
        const int row = mtile * PF_TILE_M + 32 * q + lane;
        float *y = a.Y + (size_t)split * a.ppad * a.N + row;
        for (int col = 0; col < a.ppad; col += 16) {
            uint32_t r[16];
            tmem_ld16(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)col, r);
            if (row < a.N) {
#pragma unroll
                for (int j = 0; j < 16; j++)
                    if (col + j < a.P) y[(size_t)(col + j) * a.N] = __uint_as_float(r[j]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, (uint32_t)a.tmem_cols);
}

// ------------------------------------------------------------------ packing
__device__ __forceinline__ float clamp_h(float v) { return fminf(fmaxf(v, -65504.f), 65504.f); }
// x = hi + lo in f16 (lo exact to ~2^-22 relative)
__device__ __forceinline__ void split_h(float v, __half &hi, __half &lo)
{
    v = clamp_h(v);
    hi = __float2half_rn(v);
    lo = __float2half_rn(v - __half2float(hi));
}
// one element of a HOST-format matrix on the device (rows of `cols` weights; q4_0: ggml 18-byte blocks)
__device__ __forceinline__ float host_elem(const uint8_t *m, int wtype, int cols, int r, int c)
{
    if (wtype == WT_F32) return reinterpret_cast<const float *>(m)[(size_t)r * cols + c];
    if (wtype == WT_F16) return __half2float(reinterpret_cast<const __half *>(m)[(size_t)r * cols + c]);
    const uint8_t *b = m + ((size_t)r * (cols >> 5) + (c >> 5)) * 18;
    const int i = c & 31;
    const uint8_t v = b[2 + (i & 15)];
    const int qv = i < 16 ? (v & 0x0f) : (v >> 4);
    const __half d = __ushort_as_half((unsigned short)(b[0] | (b[1] << 8)));
    return __half2float(d) * (float)(qv - 8);
}

// host-format weights (N rows x K) -> operand order, NPW planes; one thread per 16-byte unit (8 elements of one row)
__global__ void pack_weights_kernel(const uint8_t *__restrict__ src, int wtype, int N, int K, uint8_t *__restrict__ dst,
                                    int npw, int mt, int kc)
{
    const size_t units = (size_t)mt * kc * 8 * PF_TILE_M;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < units; i += (size_t)gridDim.x * blockDim.x) {
        const int rr = (int)(i % PF_TILE_M), kcore = (int)((i / PF_TILE_M) % 8);
        const size_t blk = i / (PF_TILE_M * 8);  // m-tile * kc + k-chunk
        const int c = (int)(blk % kc), m = (int)(blk / kc);
        const int row = m * PF_TILE_M + rr, k0 = c * PF_BK + kcore * 8;
        __align__(16) __half hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const float v = (row < N && k0 + j < K) ? host_elem(src, wtype, K, row, k0 + j) : 0.f;
            split_h(v, hi[j], lo[j]);
        }
        uint8_t *d = dst + blk * ((size_t)npw * PF_A_PLANE) + (size_t)kcore * (PF_TILE_M * 16) + (size_t)rr * 16;
        *reinterpret_cast<uint4 *>(d) = *reinterpret_cast<const uint4 *>(hi);
        if (npw == 2) *reinterpret_cast<uint4 *>(d + PF_A_PLANE) = *reinterpret_cast<const uint4 *>(lo);
    }
}

// 8 activations of position p, contraction index k8 * 8.., into the packed B operand
__device__ __forceinline__ void store_b8(uint8_t *B, int ppad, int p, int k8, const float (&v)[8])
{
    __align__(16) __half hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; j++) split_h(v[j], hi[j], lo[j]);
    uint8_t *d = B + (size_t)(k8 >> 3) * ((size_t)ppad * 256) + (size_t)(k8 & 7) * ((size_t)ppad * 16) + (size_t)p * 16;
    *reinterpret_cast<uint4 *>(d) = *reinterpret_cast<const uint4 *>(hi);
    *reinterpret_cast<uint4 *>(d + (size_t)ppad * 128) = *reinterpret_cast<const uint4 *>(lo);
}

// sum of the K-split partial products of element (p, n), in split order
__device__ __forceinline__ float sum_parts(const float *Y, int nsplit, int ppad, int N, int p, int n)
{
    // four loads in flight per step (a partial is an L2 round trip); the additions keep the split order
    const float *y = Y + (size_t)p * N + n;
    const size_t st = (size_t)ppad * N;
    float v = 0.f;
    int s = 0;
    for (; s + 4 <= nsplit; s += 4, y += 4 * st) {
        const float a0 = y[0], a1 = y[st], a2 = y[2 * st], a3 = y[3 * st];
        v += a0; v += a1; v += a2; v += a3;
    }
    for (; s < nsplit; s++, y += st) v += y[0];
    return v;
}

// x[p][:] = embedding row of tokens[p] (llama2.f90:520); the table is in device row format
__global__ void pf_embed_kernel(const uint8_t *__restrict__ table, int wtype, int emb, size_t rs,
                                const int *__restrict__ tokens, float *__restrict__ X)
{
    const int p = blockIdx.x;
    const uint8_t *row = table + (size_t)(tokens[p] - 1) * rs;
    for (int e = threadIdx.x; e < emb; e += blockDim.x) X[(size_t)p * emb + e] = row_elem(row, wtype, emb, e);
}

// x[p] += sum of the partial products of the previous GEMM (residual add, llama2.f90:606, :621), then
// B operand = rmsnorm(x[p]) * w (llama2.f90:450-457).  One CTA per padded position; rows >= P are zeros.
constexpr int PF_NORM_THREADS = 1024;
__global__ void __launch_bounds__(PF_NORM_THREADS) pf_rmsnorm_pack_kernel(float *__restrict__ X, const float *__restrict__ w, int emb,
                                                              const float *__restrict__ Y, int nsplit, int P, int ppad,
                                                              int kpad, uint8_t *__restrict__ B)
{
    __shared__ float red[PF_NORM_THREADS / 32];
    __shared__ float s_xn;
    const int p = blockIdx.x;
    if (p >= P) {
        const float z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int k8 = threadIdx.x; k8 < kpad / 8; k8 += blockDim.x) store_b8(B, ppad, p, k8, z);
        return;
    }
    float *x = X + (size_t)p * emb;
    float ss = 0.f;
    for (int e = threadIdx.x; e < emb; e += blockDim.x) {
        float v = x[e];
        if (nsplit > 0) { v += sum_parts(Y, nsplit, ppad, emb, p, e); x[e] = v; }
        ss = fmaf(v, v, ss);
    }
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < PF_NORM_THREADS / 32; i++) t += red[i];
        s_xn = sqrtf(t / (float)emb + 1e-5f);
    }
    __syncthreads();
    const float xn = s_xn;
    for (int k8 = threadIdx.x; k8 < kpad / 8; k8 += blockDim.x) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int e = k8 * 8 + j;
            v[j] = e < emb ? x[e] * w[e] / xn : 0.f;   // the thread re-reads elements other threads wrote: after the barriers
        }
        store_b8(B, ppad, p, k8, v);
    }
}

// RoPE on q and k with the decode path's table (quirks Q1 / Q2), KV append at cache row pos0 - 1 + p
// (llama2.f90:543-565); q goes to Q[p][emb].  grid = (positions, chunks of pairs), one thread per pair of q | k | v.
__global__ void pf_rope_kv_kernel(const float *__restrict__ Y, int nsplit, int ppad, int emb, int kv, int hs,
                                  const float2 *__restrict__ tab, int pos0, float *__restrict__ Q,
                                  float *__restrict__ kc_layer, float *__restrict__ vc_layer)
{
    const int p = blockIdx.x, N = emb + 2 * kv, half = hs >> 1;
    const int row = pos0 - 1 + p;  // 0-based cache row
    const int nq = emb >> 1, nk = kv >> 1;
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < nq + 2 * nk; i += gridDim.y * blockDim.x) {
        if (i < nq) {
            const float2 cs = tab[(size_t)row * half + i % half];
            const float a = sum_parts(Y, nsplit, ppad, N, p, 2 * i), b = sum_parts(Y, nsplit, ppad, N, p, 2 * i + 1);
            Q[(size_t)p * emb + 2 * i] = a * cs.x - b * cs.y;
            Q[(size_t)p * emb + 2 * i + 1] = a * cs.y + b * cs.x;
        } else if (i < nq + nk) {
            const int pk = i - nq;
            const float2 cs = tab[(size_t)row * half + pk % half];
            const float a = sum_parts(Y, nsplit, ppad, N, p, emb + 2 * pk), b = sum_parts(Y, nsplit, ppad, N, p, emb + 2 * pk + 1);
            float *dst = kc_layer + (size_t)row * kv;
            dst[2 * pk] = a * cs.x - b * cs.y;
            dst[2 * pk + 1] = a * cs.y + b * cs.x;
        } else {
            const int pv = i - nq - nk;
            float *dst = vc_layer + (size_t)row * kv;
            dst[2 * pv] = sum_parts(Y, nsplit, ppad, N, p, emb + kv + 2 * pv);
            dst[2 * pv + 1] = sum_parts(Y, nsplit, ppad, N, p, emb + kv + 2 * pv + 1);
        }
    }
}

// causal attention of position p (cache rows 0 .. pos0 - 1 + p) for head h, same three steps as the reference
// (llama2.f90:574-598: scores / sqrt(hs), softmax, weighted sum of the value rows; kv head = h / kv_mul), written
// straight into the packed B operand of the Wo GEMM.  grid = (heads, ppad).
constexpr int PF_ATT_THREADS = 128;
__global__ void __launch_bounds__(PF_ATT_THREADS) pf_attention_kernel(
    const float *__restrict__ Q, const float *__restrict__ kc, const float *__restrict__ vc, int pos0, int P, int ppad,
    int emb, int kv_mul, int hs, int kv, uint8_t *__restrict__ B)
{
    extern __shared__ float att[];  // one score per cached position
    __shared__ float red[PF_ATT_THREADS / 32];
    __shared__ float bcast;
    __shared__ float outv[128];
    const int h = blockIdx.x, p = blockIdx.y, g = h / kv_mul;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = PF_ATT_THREADS / 32;
    if (p >= P) {
        const float z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int j = threadIdx.x; j < hs / 8; j += PF_ATT_THREADS) store_b8(B, ppad, p, h * (hs / 8) + j, z);
        return;
    }
    const int n = pos0 + p;  // positions attended to (1-based pos of this token)
    const float *qh = Q + (size_t)p * emb + (size_t)h * hs;
    const float scale = sqrtf((float)hs);
    for (int t = warp; t < n; t += nw) {
        const float *kt = kc + (size_t)t * kv + (size_t)g * hs;
        float s = 0.f;
        for (int d = lane; d < hs; d += 32) s = fmaf(qh[d], kt[d], s);
        s = warp_sum(s);
        if (lane == 0) att[t] = s / scale;
    }
    __syncthreads();
    float mx = -INFINITY;
    for (int t = threadIdx.x; t < n; t += PF_ATT_THREADS) mx = fmaxf(mx, att[t]);
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
    for (int t = threadIdx.x; t < n; t += PF_ATT_THREADS) {
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
    for (int t = threadIdx.x; t < n; t += PF_ATT_THREADS) att[t] = att[t] / sum;  // softmax, llama2.f90:476
    __syncthreads();
    for (int d = threadIdx.x; d < hs; d += PF_ATT_THREADS) {
        float acc = 0.f;
        for (int t = 0; t < n; t++) acc = fmaf(att[t], vc[(size_t)t * kv + (size_t)g * hs + d], acc);
        outv[d] = acc;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < hs / 8; j += PF_ATT_THREADS) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = outv[j * 8 + i];
        store_b8(B, ppad, p, h * (hs / 8) + j, v);
    }
}

// hb = silu(W1 x) * (W3 x) (llama2.f90:613-616) from the partial products of the W1|W3 GEMM (rows W1 then W3,
// the host order) into the packed B operand of the W2 GEMM.  grid = (ppad, chunks of 8-element units).
__global__ void __launch_bounds__(256) pf_swiglu_pack_kernel(const float *__restrict__ Y, int nsplit, int P, int ppad, int hid,
                                                             int kpad, uint8_t *__restrict__ B)
{
    const int p = blockIdx.x;
    for (int k8 = blockIdx.y * blockDim.x + threadIdx.x; k8 < kpad / 8; k8 += gridDim.y * blockDim.x) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int i = k8 * 8 + j;
            v[j] = 0.f;
            if (p < P && i < hid) {
                const float g = sum_parts(Y, nsplit, ppad, 2 * hid, p, i), u = sum_parts(Y, nsplit, ppad, 2 * hid, p, hid + i);
                v[j] = (g * (1.0f / (1.0f + expf(-g)))) * u;
            }
        }
        store_b8(B, ppad, p, k8, v);
    }
}

// plain rows x[P][K] -> packed B operand (rows >= P and columns >= K zero); grid = ppad
__global__ void pf_pack_rows_kernel(const float *__restrict__ X, int P, int ppad, int K, int kpad, uint8_t *__restrict__ B)
{
    const int p = blockIdx.x;
    for (int k8 = threadIdx.x; k8 < kpad / 8; k8 += blockDim.x) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = (p < P && k8 * 8 + j < K) ? X[(size_t)p * K + k8 * 8 + j] : 0.f;
        store_b8(B, ppad, p, k8, v);
    }
}
// out[p][n] = sum of the K-split partial products
__global__ void pf_sum_kernel(const float *__restrict__ Y, int nsplit, int P, int ppad, int N, float *__restrict__ out)
{
    const size_t total = (size_t)P * N;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
        out[i] = sum_parts(Y, nsplit, ppad, N, (int)(i / N), (int)(i % N));
}

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

}  // namespace

// ------------------------------------------------------------------ host side
void prefill_gemm_geometry(int rows, int cols, int wtype, int n_sms, int n_pos, PrefillGemmGeom *g)
{
    g->rows = rows; g->cols = cols;
    g->planes = wtype == WT_F16 ? 1 : 2;
    g->m_tiles = cdiv(rows, PF_TILE_M); g->k_chunks = cdiv(cols, PF_BK);
    // K splits: enough CTAs to cover the SMs (rounded to the nearest whole number of splits), at most PF_MAX_SPLIT,
    // every split non-empty
    const int want = std::max(1, std::min({(n_sms + g->m_tiles / 2) / g->m_tiles, PF_MAX_SPLIT, g->k_chunks}));
    g->chunks_per_split = cdiv(g->k_chunks, want);
    g->n_splits = cdiv(g->k_chunks, g->chunks_per_split);
    g->ppad = (std::max(1, std::min(n_pos, PF_MAXP)) + 15) & ~15;
    g->tmem_cols = 32;
    while (g->tmem_cols < g->ppad) g->tmem_cols *= 2;
    g->stage_bytes = g->planes * PF_A_PLANE + g->ppad * 256;
    g->stages = std::max(2, std::min({PF_MAX_STAGES, (200 * 1024) / g->stage_bytes, g->chunks_per_split}));
    g->smem_bytes = g->stages * g->stage_bytes + (2 * PF_MAX_STAGES + 1) * 8 + 16;
    g->weight_bytes = (unsigned long long)g->m_tiles * g->k_chunks * g->planes * PF_A_PLANE;
    g->partial_bytes = (unsigned long long)g->n_splits * g->ppad * rows * 4ull;
}

struct Prefill {
    PrefillDims d{};
    int npw = 2, n_sms = 148;
    // per matrix (0 QKV, 1 Wo, 2 W1|W3, 3 W2): rows, contraction length, tiles, k-chunks, bytes per layer, K splits
    int N[4] = {}, K[4] = {}, mt[4] = {}, kc[4] = {}, cps[4] = {}, nsplit[4] = {};
    size_t layer_bytes[4] = {};
    uint8_t *W[4] = {};
    uint8_t *B = nullptr;   // packed activations (one GEMM input at a time)
    float *X = nullptr, *Q = nullptr, *Y = nullptr;
    int *tokens = nullptr;
    size_t b_bytes = 0;
};

size_t prefill_weight_bytes(const PrefillDims &d)
{
    const int npw = d.wtype == WT_F16 ? 1 : 2;
    const int N[4] = {d.nqkv, d.emb, 2 * d.hid, d.emb}, K[4] = {d.emb, d.emb, d.emb, d.hid};
    size_t t = 0;
    for (int i = 0; i < 4; i++) t += (size_t)cdiv(N[i], PF_TILE_M) * cdiv(K[i], PF_BK) * npw * PF_A_PLANE;
    return t * d.L;
}

cudaError_t prefill_create(Prefill **out, const PrefillDims &d, int n_sms)
{
    Prefill *pf = new Prefill;
    pf->d = d; pf->n_sms = n_sms;
    pf->npw = d.wtype == WT_F16 ? 1 : 2;
    const int N[4] = {d.nqkv, d.emb, 2 * d.hid, d.emb}, K[4] = {d.emb, d.emb, d.emb, d.hid};
    size_t ymax = 0;
    int kmax = 0;
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < 4 && e == cudaSuccess; i++) {
        PrefillGemmGeom g;
        prefill_gemm_geometry(N[i], K[i], d.wtype, n_sms, PF_MAXP, &g);
        pf->N[i] = N[i]; pf->K[i] = K[i];
        pf->mt[i] = g.m_tiles; pf->kc[i] = g.k_chunks;
        pf->cps[i] = g.chunks_per_split; pf->nsplit[i] = g.n_splits;
        pf->layer_bytes[i] = (size_t)g.weight_bytes;
        ymax = std::max(ymax, (size_t)pf->nsplit[i] * PF_MAXP * N[i]);
        kmax = std::max(kmax, pf->kc[i] * PF_BK);
        e = cudaMalloc((void **)&pf->W[i], pf->layer_bytes[i] * d.L);
    }
    pf->b_bytes = (size_t)(kmax / PF_BK) * PF_MAXP * 256;
    if (e == cudaSuccess) e = cudaMalloc((void **)&pf->B, pf->b_bytes);
    // zero once: the contraction padding past K is never written afterwards (0 x stale NaN would poison a sum)
    if (e == cudaSuccess) e = cudaMemset(pf->B, 0, pf->b_bytes);
    if (e == cudaSuccess) e = cudaMalloc((void **)&pf->X, (size_t)PF_MAXP * d.emb * 4);
    if (e == cudaSuccess) e = cudaMalloc((void **)&pf->Q, (size_t)PF_MAXP * d.emb * 4);
    if (e == cudaSuccess) e = cudaMalloc((void **)&pf->Y, ymax * 4);
    if (e == cudaSuccess) e = cudaMalloc((void **)&pf->tokens, PF_MAXP * 4);
    // the largest configuration: 3 stages of 32 KB weights + 32 KB activations
    if (e == cudaSuccess) e = cudaFuncSetAttribute(umma_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(umma_gemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(pf_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) { prefill_destroy(pf); return e; }
    *out = pf;
    return cudaSuccess;
}

void prefill_destroy(Prefill *pf)
{
    if (!pf) return;
    for (uint8_t *w : pf->W) if (w) cudaFree(w);
    void *ptrs[] = {pf->B, pf->X, pf->Q, pf->Y, pf->tokens};
    for (void *p : ptrs) if (p) cudaFree(p);
    delete pf;
}

int prefill_max_positions() { return PF_MAXP; }

cudaError_t prefill_pack_weights(Prefill *pf, int matrix, int layer, const uint8_t *src_host_format, cudaStream_t st)
{
    uint8_t *dst = pf->W[matrix] + (size_t)layer * pf->layer_bytes[matrix];
    pack_weights_kernel<<<148 * 8, 256, 0, st>>>(src_host_format, pf->d.wtype, pf->N[matrix], pf->K[matrix], dst, pf->npw,
                                                 pf->mt[matrix], pf->kc[matrix]);
    return cudaGetLastError();
}

static cudaError_t run_gemm(Prefill *pf, int matrix, int layer, int P, int ppad, cudaStream_t st)
{
    GemmArgs a{};
    a.A = pf->W[matrix] + (size_t)layer * pf->layer_bytes[matrix];
    a.B = pf->B; a.Y = pf->Y;
    a.N = pf->N[matrix]; a.P = P; a.ppad = ppad;
    PrefillGemmGeom g;
    prefill_gemm_geometry(pf->N[matrix], pf->K[matrix], pf->d.wtype, pf->n_sms, P, &g);
    a.kc = g.k_chunks; a.cps = g.chunks_per_split;
    a.tmem_cols = g.tmem_cols;
    a.nst = g.stages;
    const size_t smem = (size_t)g.smem_bytes;
    const dim3 grid(g.m_tiles, g.n_splits);
    if (pf->npw == 1) umma_gemm_kernel<1><<<grid, PF_THREADS, smem, st>>>(a);
    else umma_gemm_kernel<2><<<grid, PF_THREADS, smem, st>>>(a);
    return cudaGetLastError();
}

// positions pos0 .. pos0 + n - 1 (1-based) with input tokens d_tokens[0..n) (already on the device, in pf->tokens)
cudaError_t prefill_run(Prefill *pf, const PrefillRun &r, int n, int pos0, cudaStream_t st, int *launches)
{
    const PrefillDims &d = pf->d;
    const int ppad = (n + 15) & ~15, kv_mul = d.H / d.KVH;
    const int kpad_e = pf->kc[0] * PF_BK, kpad_h = pf->kc[3] * PF_BK;
    int k = 0;
    cudaError_t e = cudaSuccess;
#define PF_CK(x) do { e = (x); if (e != cudaSuccess) return e; k++; } while (0)
    pf_embed_kernel<<<n, 256, 0, st>>>(r.emb_table, d.wtype, d.emb, row_stride_bytes(d.wtype, d.emb), pf->tokens, pf->X);
    PF_CK(cudaGetLastError());
    for (int l = 0; l < d.L; l++) {
        float *kc = r.kc + (size_t)l * d.seq * d.kv, *vc = r.vc + (size_t)l * d.seq * d.kv;
        // rmsnorm (+ the previous layer's W2 residual) -> QKV
        pf_rmsnorm_pack_kernel<<<ppad, PF_NORM_THREADS, 0, st>>>(pf->X, r.rms_att + (size_t)l * d.emb, d.emb, pf->Y, l ? pf->nsplit[3] : 0, n,
                                                     ppad, kpad_e, pf->B);
        PF_CK(cudaGetLastError());
        PF_CK(run_gemm(pf, 0, l, n, ppad, st));
        pf_rope_kv_kernel<<<dim3(n, cdiv(d.emb / 2 + d.kv, 256)), 256, 0, st>>>(pf->Y, pf->nsplit[0], ppad, d.emb, d.kv, d.hs, r.rope_tab, pos0, pf->Q, kc, vc);
        PF_CK(cudaGetLastError());
        if (l == d.L - 1) break;  // the last layer's attention / FFN only feed logits nobody reads (llama2.f90:383-385)
        pf_attention_kernel<<<dim3(d.H, ppad), PF_ATT_THREADS, (size_t)(pos0 + n) * 4, st>>>(pf->Q, kc, vc, pos0, n, ppad, d.emb,
                                                                                           kv_mul, d.hs, d.kv, pf->B);
        PF_CK(cudaGetLastError());
        PF_CK(run_gemm(pf, 1, l, n, ppad, st));
        // residual + rmsnorm -> W1|W3 -> SwiGLU -> W2
        pf_rmsnorm_pack_kernel<<<ppad, PF_NORM_THREADS, 0, st>>>(pf->X, r.rms_ffn + (size_t)l * d.emb, d.emb, pf->Y, pf->nsplit[1], n, ppad,
                                                     kpad_e, pf->B);
        PF_CK(cudaGetLastError());
        PF_CK(run_gemm(pf, 2, l, n, ppad, st));
        pf_swiglu_pack_kernel<<<dim3(ppad, cdiv(kpad_h / 8, 256)), 256, 0, st>>>(pf->Y, pf->nsplit[2], n, ppad, d.hid, kpad_h, pf->B);
        PF_CK(cudaGetLastError());
        PF_CK(run_gemm(pf, 3, l, n, ppad, st));
    }
#undef PF_CK
    if (launches) *launches = k;
    return cudaSuccess;
}

int *prefill_token_buffer(Prefill *pf) { return pf->tokens; }

// The GEMM on its own (test operator): y[P][N] = x[P][K] . W^T with W in HOST format on the device.
cudaError_t prefill_gemm_op(const uint8_t *w_host_format, int wtype, int N, int K, const float *x, int P, float *y,
                            int n_sms, cudaStream_t st)
{
    if (P < 1 || P > PF_MAXP) return cudaErrorInvalidValue;
    PrefillGemmGeom g;
    prefill_gemm_geometry(N, K, wtype, n_sms, P, &g);
    const int npw = g.planes, mt = g.m_tiles, kc = g.k_chunks, ppad = g.ppad, cps = g.chunks_per_split, nsplit = g.n_splits;
    uint8_t *A = nullptr, *B = nullptr;
    float *Y = nullptr;
    cudaError_t e = cudaMalloc((void **)&A, (size_t)mt * kc * npw * PF_A_PLANE);
    if (e == cudaSuccess) e = cudaMalloc((void **)&B, (size_t)kc * ppad * 256);
    if (e == cudaSuccess) e = cudaMalloc((void **)&Y, (size_t)nsplit * ppad * N * 4);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(umma_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(umma_gemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e == cudaSuccess) {
        pack_weights_kernel<<<148 * 8, 256, 0, st>>>(w_host_format, wtype, N, K, A, npw, mt, kc);
        pf_pack_rows_kernel<<<ppad, 256, 0, st>>>(x, P, ppad, K, kc * PF_BK, B);
        GemmArgs a{};
        a.A = A; a.B = B; a.Y = Y; a.N = N; a.P = P; a.ppad = ppad; a.kc = kc; a.cps = cps;
        a.tmem_cols = g.tmem_cols;
        a.nst = g.stages;
        const size_t smem = (size_t)g.smem_bytes;
        if (npw == 1) umma_gemm_kernel<1><<<dim3(mt, nsplit), PF_THREADS, smem, st>>>(a);
        else umma_gemm_kernel<2><<<dim3(mt, nsplit), PF_THREADS, smem, st>>>(a);
        pf_sum_kernel<<<148, 256, 0, st>>>(Y, nsplit, P, ppad, N, y);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
    cudaFree(A); cudaFree(B); cudaFree(Y);
    return e;
}

}  // namespace llmf90

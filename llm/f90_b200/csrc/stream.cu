// stream.cu -- the fused weight-streaming decode kernel (one launch = one token).
//
// Replaces `function transformer(token,pos,s,w)` (llama2.f90:480-640) on one B200.
//
// Design: decode at batch 1 reads every weight byte exactly once per
// token and does ~0.5 flop per byte, so the only thing that matters is keeping HBM busy
// across the ~110 dependent mat-vec phases of a token.  One persistent cooperative CTA per SM:
//
//   * a PRODUCER warp walks the CTA's private, fully static weight schedule -- for every layer
//     its contiguous row range of Wqkv, Wo, W13 (gate/up rows interleaved at upload), W2 and
//     finally Wcls -- and streams it with 1-D TMA bulk copies (cp.async.bulk, mbarrier
//     complete_tx) into a ring of shared-memory slots.  Weights do not depend on activations,
//     so the producer never waits for a grid barrier: while the consumers finish a phase,
//     synchronise the grid and rebuild the activation vector, the ring keeps filling.
//   * CONSUMER warps each own one ring slot: wait on its `full` mbarrier, do the dot products
//     of the rows in the slot against the activation vector held in shared memory (f16 and
//     q4_0 dequantisation fused into the load, f32 accumulation, warp-shuffle reduction),
//     release the slot with an `empty` mbarrier arrive.
//   * Between phases the consumers run the tiny epilogues in place -- RoPE + KV-cache append,
//     SwiGLU, residual add -- publish their slice to global memory and meet at a grid barrier
//     (release/acquire counter in L2); the next phase's prologue re-reads the full vector
//     (rmsnorm recomputed redundantly per CTA: 8-16 KB from L2).
//   * Attention (scores, softmax, value gather) is a phase of the same kernel: (head, split)
//     items over the CTAs, online softmax, combined in the Wo prologue.
#include <cooperative_groups.h>

#include "kernels.cuh"

namespace llmf90 {

constexpr int MAX_SLOTS = 16;
constexpr int MAX_CONS_WARPS = 15;  // + 1 producer warp = 512 threads -> 128 registers/thread
constexpr int CONS_BAR = 1;  // named barrier id used by the consumer warps

struct SmemView {
    uint8_t *ring;
    float *xs, *res, *red;
    uint64_t *full, *empty;
};

__device__ __forceinline__ SmemView carve(uint8_t *smem, const StreamParams &P)
{
    SmemView v;
    size_t off = 0;
    v.ring = smem;
    off += (size_t)P.n_slots * P.slot_bytes;
    v.xs = reinterpret_cast<float *>(smem + off);
    off += (size_t)P.xs_floats * 4;
    v.res = reinterpret_cast<float *>(smem + off);
    off += (size_t)P.res_floats * 4;
    v.red = reinterpret_cast<float *>(smem + off);
    off += 64 * 4;
    v.full = reinterpret_cast<uint64_t *>(smem + off);
    v.empty = v.full + MAX_SLOTS;
    return v;
}

static size_t smem_bytes_for(int n_slots, int slot_bytes, int xs_floats, int res_floats)
{
    return (size_t)n_slots * slot_bytes + (size_t)xs_floats * 4 + (size_t)res_floats * 4 + 64 * 4 +
           2 * MAX_SLOTS * 8;
}

__host__ __device__ inline void cta_rows(const PhaseW &ph, int cta, int G, int &r0, int &r1)
{
    const long long U = ph.rows / ph.unit;
    r0 = (int)((long long)cta * U / G) * ph.unit;
    r1 = (int)((long long)(cta + 1) * U / G) * ph.unit;
}

// ------------------------------------------------------------------ producer
__device__ __forceinline__ void produce_phase(const PhaseW &ph, int layer, const StreamParams &P,
                                              const SmemView &sv, uint32_t &s, uint64_t pol)
{
    int r0, r1;
    cta_rows(ph, blockIdx.x, gridDim.x, r0, r1);
    const int nrows = r1 - r0;
    const uint8_t *src = ph.base + (size_t)layer * ph.layer_stride + (size_t)r0 * ph.rs;
    for (int r = 0; r < nrows; r += ph.rps) {
        const int n = min(ph.rps, nrows - r);
        const uint32_t bytes = (uint32_t)n * ph.rs;
        const uint32_t slot = s % (uint32_t)P.n_slots, k = s / (uint32_t)P.n_slots;
        mbar_wait(&sv.empty[slot], (k & 1u) ^ 1u);
        mbar_arrive_expect_tx(&sv.full[slot], bytes);
        bulk_g2s(sv.ring + (size_t)slot * P.slot_bytes, src + (size_t)r * ph.rs, bytes,
                 &sv.full[slot], pol);
        s++;
    }
}

// ------------------------------------------------------------------ consumer helpers
struct Cons {
    int tid, warp, lane, nt, nw;  // within the consumer group
    int slot, sub;                // ring slot this warp serves, sub-warp index within the slot
};

__device__ __forceinline__ void cons_sync(const Cons &c) { named_bar_sync(CONS_BAR, c.nt); }

__device__ __forceinline__ float cons_sum(float v, const Cons &c, float *red)
{
    v = warp_sum(v);
    if (c.lane == 0) red[c.warp] = v;
    cons_sync(c);
    float t = 0.f;
    for (int i = 0; i < c.nw; i++) t += red[i];
    cons_sync(c);
    return t;
}

template <int WT, int NR>
__device__ __forceinline__ void rows_to_res(const uint8_t *sp, size_t rs, const float *xs, int cols,
                                            int lane, float *res)
{
    float acc[NR];
#pragma unroll
    for (int i = 0; i < NR; i++) acc[i] = 0.f;
    dot_rows<WT, NR>(sp, rs, xs, cols, lane, acc);
#pragma unroll
    for (int i = 0; i < NR; i++) {
        const float v = warp_sum(acc[i]);
        if (lane == 0) res[i] = v;
    }
}

template <int WT>
__device__ __forceinline__ void consume_phase(const PhaseW &ph, const StreamParams &P,
                                              const SmemView &sv, const Cons &c, uint32_t &s)
{
    int r0, r1;
    cta_rows(ph, blockIdx.x, gridDim.x, r0, r1);
    const int nrows = r1 - r0;
    const uint32_t nst = (uint32_t)((nrows + ph.rps - 1) / ph.rps);
    const uint32_t ns = (uint32_t)P.n_slots;
    uint32_t st = s + (((uint32_t)c.slot + ns - s % ns) % ns);
    const uint8_t *slot_ptr = sv.ring + (size_t)c.slot * P.slot_bytes;
    for (; st < s + nst; st += ns) {
        mbar_wait(&sv.full[c.slot], (st / ns) & 1u);
        const int rbase = (int)(st - s) * ph.rps;
        const int n = min(ph.rps, nrows - rbase);
        // rows of this stage are split between the wps warps that share the slot
        const int per = (n + P.wps - 1) / P.wps;
        int i = c.sub * per;
        const int iend = min(n, i + per);
        for (; i + 4 <= iend; i += 4)
            rows_to_res<WT, 4>(slot_ptr + (size_t)i * ph.rs, ph.rs, sv.xs, ph.cols, c.lane,
                               sv.res + rbase + i);
        if (i + 2 <= iend) {
            rows_to_res<WT, 2>(slot_ptr + (size_t)i * ph.rs, ph.rs, sv.xs, ph.cols, c.lane,
                               sv.res + rbase + i);
            i += 2;
        }
        if (i < iend)
            rows_to_res<WT, 1>(slot_ptr + (size_t)i * ph.rs, ph.rs, sv.xs, ph.cols, c.lane,
                               sv.res + rbase + i);
        __syncwarp();
        if (c.lane == 0) mbar_arrive(&sv.empty[c.slot]);
    }
    s += nst;
}

// grid-wide barrier over the consumer groups of all CTAs (the producers never join)
__device__ __forceinline__ void grid_sync(const StreamParams &P, const Cons &c, uint32_t &nbar)
{
    cons_sync(c);
    if (c.tid == 0) {
        __threadfence();
        red_release_add_u64(P.bar_ctr, 1ull);
        const unsigned long long target = P.bar_base + (unsigned long long)(nbar + 1) * gridDim.x;
        while (ld_acquire_u64(P.bar_ctr) < target) {
        }
        __threadfence();
    }
    cons_sync(c);
    nbar++;
}

// xs = rmsnorm(src) * w   (llama2.f90:450-457); src is read through L2 (written by other CTAs)
template <int WT>
__device__ __forceinline__ void load_x_norm(const float *src, const uint8_t *emb_row,
                                            const float *__restrict__ wn, const StreamParams &P,
                                            const SmemView &sv, const Cons &c)
{
    const int n = P.emb;
    float ss = 0.f;
    for (int e = c.tid; e < n; e += c.nt) {
        const float v = emb_row ? row_elem(emb_row, P.wtype, n, e) : __ldcg(src + e);
        sv.xs[xs_index<WT>(e)] = v;
        ss = fmaf(v, v, ss);
    }
    const float tot = cons_sum(ss, c, sv.red);
    const float xn = sqrtf(tot / (float)n + 1e-5f);
    for (int e = c.tid; e < n; e += c.nt) {
        const int ix = xs_index<WT>(e);
        sv.xs[ix] = sv.xs[ix] * wn[e] / xn;
    }
    cons_sync(c);
}

// ------------------------------------------------------------------ attention phase
// (head, split) items over the CTAs.  Within an item the consumer warps take positions
// round-robin, keep an online-softmax state (m, l, acc) and merge through shared memory.
// Partial result layout in global memory: att_part[(h*S + sp)*(hs+2)] = {m, l, acc[hs]}.
__device__ __forceinline__ void attention_phase(const StreamParams &P, const SmemView &sv,
                                                const Cons &c, int layer, int pos)
{
    const int hs = P.hs, S = P.n_splits, vec = hs >> 5;  // hs in {32, 64, 128}
    const int items = P.H * S;
    const int chunk = (pos + S - 1) / S;
    const float scale = sqrtf((float)hs);
    float *sc = sv.xs;  // [nw][hs + 2] scratch (xs is dead between weight phases)
    const float *kc = P.kc + (size_t)layer * P.seq * P.kv;
    const float *vc = P.vc + (size_t)layer * P.seq * P.kv;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int h = item / S, sp = item % S, g = h / P.kv_mul;
        const int t0 = sp * chunk, t1 = min(pos, t0 + chunk);
        float qv[4], acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 4; i++)
            qv[i] = i < vec ? __ldcg(P.q + (size_t)h * hs + c.lane * vec + i) : 0.f;
        float m = -INFINITY, l = 0.f;
        for (int t = t0 + c.warp; t < t1; t += c.nw) {
            const float *kt = kc + (size_t)t * P.kv + (size_t)g * hs + c.lane * vec;
            const float *vt = vc + (size_t)t * P.kv + (size_t)g * hs + c.lane * vec;
            float kk[4], vv[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                kk[i] = i < vec ? __ldcg(kt + i) : 0.f;
                vv[i] = i < vec ? __ldcg(vt + i) : 0.f;
            }
            float sdot = 0.f;
#pragma unroll
            for (int i = 0; i < 4; i++) sdot = fmaf(qv[i], kk[i], sdot);
            sdot = warp_sum(sdot) / scale;  // dot_product(q_t,k_t)/sqrt(head_size), :582
            const float mn = fmaxf(m, sdot);
            const float corr = expf(m - mn), p = expf(sdot - mn);
            l = l * corr + p;
#pragma unroll
            for (int i = 0; i < 4; i++) acc[i] = acc[i] * corr + p * vv[i];
            m = mn;
        }
        float *mine = sc + (size_t)c.warp * (hs + 2);
        if (c.lane == 0) { mine[0] = m; mine[1] = l; }
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (i < vec) mine[2 + c.lane * vec + i] = acc[i];
        cons_sync(c);
        float *out = P.att_part + (size_t)(h * S + sp) * (hs + 2);
        for (int d = c.tid; d < hs; d += c.nt) {
            float M = -INFINITY;
            for (int w = 0; w < c.nw; w++) M = fmaxf(M, sc[(size_t)w * (hs + 2)]);
            float L = 0.f, A = 0.f;
            for (int w = 0; w < c.nw; w++) {
                const float mw = sc[(size_t)w * (hs + 2)];
                if (mw > -INFINITY) {
                    const float e = expf(mw - M);
                    L = fmaf(sc[(size_t)w * (hs + 2) + 1], e, L);
                    A = fmaf(sc[(size_t)w * (hs + 2) + 2 + d], e, A);
                }
            }
            out[2 + d] = A;
            if (d == 0) { out[0] = M; out[1] = L; }
        }
        cons_sync(c);
    }
}

// xs = attention output (all heads), merging the position splits
template <int WT>
__device__ __forceinline__ void load_x_attn(const StreamParams &P, const SmemView &sv, const Cons &c)
{
    const int hs = P.hs, S = P.n_splits;
    for (int e = c.tid; e < P.emb; e += c.nt) {
        const int h = e / hs, d = e % hs;
        const float *part = P.att_part + (size_t)h * S * (hs + 2);
        float M = -INFINITY;
        for (int s = 0; s < S; s++) M = fmaxf(M, __ldcg(part + (size_t)s * (hs + 2)));
        float num = 0.f, den = 0.f;
        for (int s = 0; s < S; s++) {
            const float ms = __ldcg(part + (size_t)s * (hs + 2));
            if (ms > -INFINITY) {
                const float w = expf(ms - M);
                den = fmaf(__ldcg(part + (size_t)s * (hs + 2) + 1), w, den);
                num = fmaf(__ldcg(part + (size_t)s * (hs + 2) + 2 + d), w, num);
            }
        }
        sv.xs[xs_index<WT>(e)] = num / den;
    }
    cons_sync(c);
}

template <int WT>
__device__ __forceinline__ void load_x_plain(const float *src, int n, const SmemView &sv, const Cons &c)
{
    for (int e = c.tid; e < n; e += c.nt) sv.xs[xs_index<WT>(e)] = __ldcg(src + e);
    cons_sync(c);
}

// ------------------------------------------------------------------ the kernel
template <int WT>
__global__ void __launch_bounds__((MAX_CONS_WARPS + 1) * 32, 1)
stream_decode_kernel(const __grid_constant__ StreamParams P)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const SmemView sv = carve(smem, P);
    const int n_cons_warps = P.n_slots * P.wps;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < P.n_slots; i++) {
            mbar_init(&sv.full[i], 1);
            mbar_init(&sv.empty[i], (uint32_t)P.wps);
        }
        fence_mbar_init();
    }
    __syncthreads();

    const int token = P.token > 0 ? P.token : P.tokpos[0];
    const int pos = P.token > 0 ? P.pos : P.tokpos[1];

    if (warp == n_cons_warps) {
        // ===================== producer warp =====================
        if (lane == 0) {
            const uint64_t pol = l2_policy_evict_first();
            uint32_t s = 0;
            for (int l = 0; l < P.L; l++)
                for (int ph = 0; ph < 4; ph++) produce_phase(P.ph[ph], l, P, sv, s, pol);
            produce_phase(P.ph[4], 0, P, sv, s, pol);
        }
        return;
    }

    // ===================== consumer warps =====================
    Cons c;
    c.tid = threadIdx.x; c.warp = warp; c.lane = lane;
    c.nw = n_cons_warps; c.nt = n_cons_warps * 32;
    c.slot = warp / P.wps; c.sub = warp % P.wps;
    uint32_t s = 0, nbar = 0;
    const bool timer = (blockIdx.x == 0 && c.tid == 0);
    unsigned long long tmark = timer ? globaltimer_ns() : 0ull;
    float tacc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    auto lap = [&](int bucket) {
        if (timer) {
            const unsigned long long now = globaltimer_ns();
            tacc[bucket] += (float)(now - tmark) * 1e-6f;
            tmark = now;
        }
    };

    const uint8_t *emb_row = P.emb_table + (size_t)(token - 1) * P.ph[0].rs;
    const float2 *rope = P.rope_tab + (size_t)(pos - 1) * (P.hs >> 1);
    int r0, r1;

    for (int l = 0; l < P.L; l++) {
        // ---- phase A: rmsnorm + fused QKV mat-vec + RoPE + KV append (llama2.f90:527-565)
        load_x_norm<WT>(P.x, l == 0 ? emb_row : nullptr, P.rms_att + (size_t)l * P.emb, P, sv, c);
        consume_phase<WT>(P.ph[0], P, sv, c, s);
        cons_sync(c);
        cta_rows(P.ph[0], blockIdx.x, gridDim.x, r0, r1);
        {
            float *kc = P.kc + ((size_t)l * P.seq + (pos - 1)) * P.kv;
            float *vc = P.vc + ((size_t)l * P.seq + (pos - 1)) * P.kv;
            const int half = P.hs >> 1;
            for (int i = 2 * c.tid; i < r1 - r0; i += 2 * c.nt) {
                const int r = r0 + i;
                const float a = sv.res[i], b = sv.res[i + 1];
                if (r < P.emb) {
                    const float2 cs = rope[(r >> 1) % half];
                    P.q[r] = a * cs.x - b * cs.y;
                    P.q[r + 1] = a * cs.y + b * cs.x;
                } else if (r < P.emb + P.kv) {
                    const int rk = r - P.emb;
                    const float2 cs = rope[(rk >> 1) % half];
                    kc[rk] = a * cs.x - b * cs.y;
                    kc[rk + 1] = a * cs.y + b * cs.x;
                } else {
                    const int rv = r - P.emb - P.kv;
                    vc[rv] = a;
                    vc[rv + 1] = b;
                }
            }
        }
        grid_sync(P, c, nbar);
        lap(0);

        // ---- phase B: attention (llama2.f90:574-598)
        attention_phase(P, sv, c, l, pos);
        grid_sync(P, c, nbar);
        lap(2);

        // ---- phase C: x += Wo * att (llama2.f90:603-605)
        load_x_attn<WT>(P, sv, c);
        consume_phase<WT>(P.ph[1], P, sv, c, s);
        cons_sync(c);
        cta_rows(P.ph[1], blockIdx.x, gridDim.x, r0, r1);
        for (int i = c.tid; i < r1 - r0; i += c.nt) {
            const int r = r0 + i;
            const float base = (l == 0) ? row_elem(emb_row, P.wtype, P.emb, r) : __ldcg(P.x + r);
            P.x[r] = base + sv.res[i];
        }
        grid_sync(P, c, nbar);

        // ---- phase D: rmsnorm + fused W1|W3 mat-vec + SwiGLU (llama2.f90:608-616)
        load_x_norm<WT>(P.x, nullptr, P.rms_ffn + (size_t)l * P.emb, P, sv, c);
        consume_phase<WT>(P.ph[2], P, sv, c, s);
        cons_sync(c);
        cta_rows(P.ph[2], blockIdx.x, gridDim.x, r0, r1);
        for (int i = 2 * c.tid; i < r1 - r0; i += 2 * c.nt) {
            const float g = sv.res[i], u = sv.res[i + 1];
            P.hb[(r0 + i) >> 1] = (g * (1.0f / (1.0f + expf(-g)))) * u;
        }
        grid_sync(P, c, nbar);

        // ---- phase E: x += W2 * hb (llama2.f90:618-620)
        load_x_plain<WT>(P.hb, P.hid, sv, c);
        consume_phase<WT>(P.ph[3], P, sv, c, s);
        cons_sync(c);
        cta_rows(P.ph[3], blockIdx.x, gridDim.x, r0, r1);
        for (int i = c.tid; i < r1 - r0; i += c.nt) {
            const int r = r0 + i;
            P.x[r] = __ldcg(P.x + r) + sv.res[i];
        }
        grid_sync(P, c, nbar);
        lap(3);
    }

    // ---- final rmsnorm + classifier (llama2.f90:627-636)
    load_x_norm<WT>(P.x, nullptr, P.rms_final, P, sv, c);
    consume_phase<WT>(P.ph[4], P, sv, c, s);
    cons_sync(c);
    cta_rows(P.ph[4], blockIdx.x, gridDim.x, r0, r1);
    float best = -INFINITY;
    int bidx = 0x7fffffff;
    for (int i = c.tid; i < r1 - r0; i += c.nt) {
        const float v = sv.res[i];
        P.logits[r0 + i] = v;
        if (v > best) { best = v; bidx = r0 + i; }
    }
    lap(4);

    if (P.do_argmax) {
        // maxloc(logits) (llama2.f90:388): first maximum wins at every reduction level
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
            if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
        }
        float *rv = sv.red;
        int *ri = reinterpret_cast<int *>(sv.red + 32);
        if (c.lane == 0) { rv[c.warp] = best; ri[c.warp] = bidx; }
        cons_sync(c);
        if (c.tid == 0) {
            for (int w = 1; w < c.nw; w++)
                if (rv[w] > best || (rv[w] == best && ri[w] < bidx)) { best = rv[w]; bidx = ri[w]; }
            P.amax_scratch[2 * blockIdx.x] = __float_as_int(best);
            P.amax_scratch[2 * blockIdx.x + 1] = bidx;
        }
        grid_sync(P, c, nbar);
        if (blockIdx.x == 0 && c.warp == 0) {
            best = -INFINITY; bidx = 0x7fffffff;
            for (int i = c.lane; i < (int)gridDim.x; i += 32) {
                const float v = __int_as_float(__ldcg(P.amax_scratch + 2 * i));
                const int ix = __ldcg(P.amax_scratch + 2 * i + 1);
                if (v > best || (v == best && ix < bidx)) { best = v; bidx = ix; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
                if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
            }
            if (c.lane == 0) {
                int next = bidx + 1;
                if (P.forced && P.forced[pos - 1] > 0) next = P.forced[pos - 1];
                if (P.out_tokens) P.out_tokens[pos - 1] = next;
                int *tp = const_cast<int *>(P.tokpos);
                tp[0] = next;
                tp[1] = pos + 1;
            }
        }
    }
    if (timer)
        for (int i = 0; i < 5; i++) P.times_dev[i] += tacc[i];
}

// ------------------------------------------------------------------ host side
int stream_barriers_per_launch(const StreamParams &p) { return 5 * p.L + (p.do_argmax ? 1 : 0); }

int plan_stream(const StreamParams &p, int n_sms, int max_smem_optin, int target_slot_bytes,
                int max_slots, StreamPlan *out)
{
    unsigned int rs_max = 0;
    int max_rows = 0;
    for (int i = 0; i < 5; i++) {
        if (p.ph[i].rs > rs_max) rs_max = p.ph[i].rs;
        const int U = p.ph[i].rows / p.ph[i].unit;
        const int per = ((U + n_sms - 1) / n_sms + 1) * p.ph[i].unit;
        if (per > max_rows) max_rows = per;
    }
    int slot = target_slot_bytes > (int)rs_max ? target_slot_bytes : (int)rs_max;
    slot = (slot + 127) & ~127;
    int xs_floats = p.emb > p.hid ? p.emb : p.hid;
    // attention scratch [consumer warps][hs+2] also lives in xs
    const int att_scratch = MAX_SLOTS * (p.hs + 2);
    if (att_scratch > xs_floats) xs_floats = att_scratch;
    xs_floats = (xs_floats + 31) & ~31;
    const int res_floats = (max_rows + 31) & ~31;
    if (max_slots > MAX_CONS_WARPS) max_slots = MAX_CONS_WARPS;
    int n_slots = max_slots;
    while (n_slots > 0 &&
           smem_bytes_for(n_slots, slot, xs_floats, res_floats) > (size_t)max_smem_optin)
        n_slots--;
    if (n_slots < 2) return 1;
    out->n_slots = n_slots;
    out->slot_bytes = slot;
    out->wps = (p.wtype == WT_Q4_0 && 2 * n_slots <= MAX_CONS_WARPS) ? 2 : 1;
    out->threads = (n_slots * out->wps + 1) * 32;
    out->smem_bytes = (int)smem_bytes_for(n_slots, slot, xs_floats, res_floats);
    out->grid = n_sms;
    out->xs_floats = xs_floats;
    out->res_floats = res_floats;
    return 0;
}

template <int WT>
static cudaError_t prepare_t(int smem_bytes)
{
    return cudaFuncSetAttribute(stream_decode_kernel<WT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                smem_bytes);
}

cudaError_t prepare_stream_kernel(int wtype, int smem_bytes)
{
    if (wtype == WT_F32) return prepare_t<WT_F32>(smem_bytes);
    if (wtype == WT_F16) return prepare_t<WT_F16>(smem_bytes);
    if (wtype == WT_Q4_0) return prepare_t<WT_Q4_0>(smem_bytes);
    return cudaErrorInvalidValue;
}

cudaError_t launch_stream(const StreamParams &p, const StreamPlan &plan, cudaStream_t st)
{
    StreamParams q = p;
    void *args[] = {(void *)&q};
    const void *fn = nullptr;
    if (p.wtype == WT_F32) fn = (const void *)stream_decode_kernel<WT_F32>;
    else if (p.wtype == WT_F16) fn = (const void *)stream_decode_kernel<WT_F16>;
    else if (p.wtype == WT_Q4_0) fn = (const void *)stream_decode_kernel<WT_Q4_0>;
    else return cudaErrorInvalidValue;
    // cooperative launch: guarantees all CTAs are co-resident (the grid barrier needs it)
    return cudaLaunchCooperativeKernel(fn, dim3(plan.grid), dim3(plan.threads), args,
                                       (size_t)plan.smem_bytes, st);
}

}  // namespace llmf90

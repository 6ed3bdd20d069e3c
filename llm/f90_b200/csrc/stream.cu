// stream.cu -- the fused weight-streaming decode kernel (one launch = one token).
//
// Replaces `function transformer(token,pos,s,w)` (llama2.f90:480-640) on one B200.
//
// Design: decode at batch 1 reads every weight byte exactly once per
// token and does ~0.5 flop per byte, so the only thing that matters is keeping HBM busy
// across the ~110 dependent mat-vec phases of a token.  One persistent cooperative CTA per SM:
//
//   * a PRODUCER warp walks the CTA's private, fully static schedule -- the token's embedding
//     row, then for every layer the rms_att vector, its row range of Wqkv and Wo, the rms_ffn
//     vector, its rows of W13 (gate/up rows interleaved at upload) and W2, finally the rms_final
//     vector and its rows of Wcls -- and streams it with 1-D TMA bulk copies (cp.async.bulk,
//     mbarrier complete_tx) into a ring of shared-memory slots.  Nothing in the schedule depends
//     on activations, so the producer never waits for a hand-over: while the consumers finish a
//     phase and rebuild the activation vector, the ring keeps filling ("banked" weights).
//   * 12 CONSUMER warps in groups of G (a divisor of 12, per phase).  The rows of a phase are cut
//     into TILES of R rows (4 f32, 8 f16, 16 q4_0) and a tile's contraction range into chunks
//     of one ring stage each.  A tile belongs to ONE group (tile t -> group t mod 12/G): its warps
//     wait on the stage's `full` mbarrier, each dots its share of the columns of the R row segments
//     with the activation vector in shared memory (one activation load serves R rows; f16 / q4_0
//     on mma.sync with the dequantisation fused, f32 accumulation), they hand the slot back, and after the tile's last
//     chunk each reduces across its lanes, the group adds its G partial results through shared
//     memory (a named barrier of the group only), and the group's first warp runs the tile's
//     epilogue ITSELF -- RoPE + KV-cache append, SwiGLU, residual partials, logits -- publishing
//     its rows at once.  No CTA-wide barrier follows a mat-vec: the hand-over latency is the
//     slowest group's, not the slowest warp's plus a barrier plus a second pass; every ring slot
//     has a group working on it, so banked stages drain at shared-memory speed (3-4x the SM's
//     share of HBM bandwidth), which is what turns the ring's head start into time saved.
//   * Between phases values travel as {value, epoch} 64-bit words ("LL" buffers, the protocol
//     NCCL uses for small messages): the next phase's prologue polls the whole vector straight out
//     of L2 until every word carries the expected epoch.  There is NO grid barrier and no fence
//     anywhere in the token: a phase hand-over costs one store->load trip through L2 (rmsnorm is
//     recomputed redundantly per CTA).
//   * Attention (scores, softmax, value gather) is a phase of the same kernel: (head, split)
//     items over the CTAs, online softmax, merged in the Wo prologue when there are splits.
//
// CODE SIZE IS A FIRST-CLASS CONSTRAINT.  Every piece of this kernel runs once per phase, i.e. its
// instructions are cold each time unless one layer's worth of code fits the instruction caches.
// Hence: one generic prologue / consume / epilogue for all phases (data-driven, not specialised),
// modest unrolling, profiling hooks out of line, cold paths in non-inlined functions.
#include <cooperative_groups.h>
#include <cstdlib>

#include "kernels.cuh"

namespace llmf90 {

constexpr int MAX_SLOTS = 16;
constexpr int NCW = 12;             // consumer warps (+ 1 producer warp = 416 threads)
constexpr int NCT = NCW * 32;       // consumer threads
constexpr int CONS_BAR = 1;         // named barrier id used by the consumer warps (ids 2.. : one per tile group)
// `full` barriers are per STAGE NUMBER, not per slot: stage s (ring order, per CTA) completes barrier
// s mod NBAR in its phase (s / NBAR) mod 2.  A warp may wait for a stage long before the stages in
// front of it have landed (its next tile is 11 tiles ahead); the one-bit mbarrier phase parity is
// only unambiguous if the previous use of the same barrier (stage s - NBAR) has completed by then.
// At most n_slots stages are in flight and a warp looks at most 11 tiles (the other groups') x MAX_NCH chunks
// ahead of the oldest one: 11 x 20 + 16 slots < NBAR.
constexpr int NBAR = 256;
constexpr int MAX_NCH = 20;         // chunks per tile (plan_stream enforces it)
static_assert((NCW - 1) * MAX_NCH + MAX_SLOTS < NBAR, "a full barrier could be waited on before its previous use completed");
constexpr int ATT_PSTRIDE_PAD = 4;  // attention partial record = {m, l, -, -, acc[hs]}
constexpr int ATT_MAX_CHUNK = 288;  // positions of one attention split: <= 256 rounded up to groups of 16, + slack

// per-launch constants of this CTA in shared memory: everything the hot loop needs comes from here
// with one LDS (kernel parameters reached from a non-inlined function are generic loads)
struct CtaPlan {
    PhaseW ph[5];
    int r0[5], nrows[5], ntiles[5], nst[5];
    unsigned long long lstride[SK_COUNT];  // bytes between layers per stage kind
    unsigned int sstride[SK_COUNT];        // bytes between the segments of a stage per stage kind
    int off_xs, off_xres, off_red, off_att, off_grp, off_full, off_empty, off_sched;
    int G[5];                              // consumer warps per tile group, per phase (1, 2, 3, 4, 6 or 12)
    int slot_bytes, n_slots;
    unsigned int slot_magic;               // ceil(2^32 / n_slots): s mod n_slots without a division
    int wtype, emb, hid, kv, att_dim, hs, tp, rank, ll_rep, v_off, seq, H, kv_mul, L;
    int rep;                               // the replica of the LL vectors this CTA polls (blockIdx % ll_rep)
    int pos, n_splits;                     // this launch's position (1-based) and attention split count
    unsigned int ep_base;                  // epoch of layer l = ep_base + l + 1
    // ring stage numbers (per CTA): stage 0 is the embedding row, layer l's section starts at 1 + l * n_layer;
    // voff / toff: offsets of a phase's vector stage / first tile stage inside its section (classifier: after
    // the last layer)
    int n_layer, voff[5], toff[5];
    unsigned long long *trace;             // profiling kernel: this CTA's 128 trace words while the traced layer runs, else null
    unsigned long long *trace_base, *phase_cycles;  // profiling kernel: the trace buffer [grid][128] (or null), the timer accumulators
    int trace_layer;
    unsigned int kvmul_inv16;              // ceil(65536 / kv_mul): h / kv_mul without a division
    float inv_emb;                         // 1 / emb (rmsnorm mean)
    float *kc, *vc;
    unsigned long long *ll_q, *ll_kv, *ll_att, *ll_part, *ll_hb;
    unsigned long long *part1[MAX_TP], *part2[MAX_TP];
    float *logits[MAX_TP];
    // token tail
    unsigned long long *amax[MAX_TP], *done[MAX_TP];
    const int *forced;
    int *out_tokens, *tokpos;
    unsigned int ep_last;                  // epoch of the token tail's records
    int do_argmax;
    volatile int prod_issued, pf_issued;   // stages copied into the ring / prefetched into L2 so far (trace)
    unsigned long long *gx_trace;          // where gather_x leaves its clock stamps (profiling kernel), or null
    volatile int abort;                    // a tensor-parallel poll timed out: the peers are gone, stop waiting for them
    int *err_flag;                         // host-visible error word (0 = ok)
    volatile int qw[NCW];                  // per warp: the phase it is in (the phase loop keeps NOTHING in registers across calls)
    float tail_best[NCW];                  // per-warp maxloc of the classifier epilogue
    int tail_idx[NCW];
    float2 rope[64];                       // this position's RoPE row
};

// profiling state of a CTA (shared memory): phase timers of the timer thread
struct Prof {
    long long tacc[PH_COUNT];
    long long tmark;
};

// One plan / one timer block per CTA at FILE scope: their shared-memory addresses are compile-time constants, so no
// device function needs a register (or, across a call, a local-memory slot) to find them.
__shared__ CtaPlan g_cp;
__shared__ Prof g_pf;
#define cp (&g_cp)
#define pf (&g_pf)

static size_t smem_bytes_for(int n_slots, int slot_bytes, int xs_floats, int emb, int hs, int sched_entries)
{
    return (size_t)n_slots * slot_bytes + (size_t)xs_floats * 4 + (size_t)emb * 4 + 64 * 4 +
           (size_t)(ATT_MAX_CHUNK + NCW * hs) * 4 + 2 * NCW * 16 * 4 + (NBAR + MAX_SLOTS) * 8 +
           (size_t)sched_entries * sizeof(SchedStage);
}

__host__ __device__ inline void cta_rows(const PhaseW &ph, int cta, int G, int &r0, int &r1)
{
    const long long U = ph.rows / ph.unit;
    r0 = (int)((long long)cta * U / G) * ph.unit;
    r1 = (int)((long long)(cta + 1) * U / G) * ph.unit;
}

__device__ __forceinline__ uint8_t *smem_base()
{
    extern __shared__ __align__(128) uint8_t smem[];
    return smem;
}
__device__ __forceinline__ uint64_t *full_bar(const CtaPlan *, uint32_t s)
{
    return reinterpret_cast<uint64_t *>(smem_base() + cp->off_full) + (s & (NBAR - 1));
}
__device__ __forceinline__ uint32_t full_par(uint32_t s) { return (s >> 8) & 1u; }
static_assert(NBAR == 256, "full_par assumes 256 barriers");
__device__ __forceinline__ uint64_t *empty_bar(const CtaPlan *, uint32_t slot)
{
    return reinterpret_cast<uint64_t *>(smem_base() + cp->off_empty) + slot;
}
__device__ __forceinline__ uint32_t slot_of(const CtaPlan *, uint32_t s)
{
    return s - __umulhi(s, cp->slot_magic) * (uint32_t)cp->n_slots;
}
__device__ __forceinline__ void cons_sync() { named_bar_sync(CONS_BAR, NCT); }

// ------------------------------------------------------------------ producer
// Walks this CTA's stage list: [embedding row][one layer's stages] x L [final norm + classifier].
// The small f32 norm vectors and the embedding row travel through the ring like weights ("vector
// stages", read by all consumer warps in the prologue) so that no prologue waits for a demand miss
// queued behind megabytes of in-flight weight requests.
//
// Two cursors walk the list.  The RING cursor copies stage s into its slot as soon as the consumers
// have handed the slot back.  The PREFETCH cursor runs up to `pf_lead` stages ahead of it and pulls
// those stages into L2 (cp.async.bulk.prefetch.L2): the ring holds ~1 us of HBM time per SM, a phase
// hand-over takes several, and while the consumers sit in one nothing else would be fetching.  L2
// (126 MB) is the elastic buffer that keeps HBM streaming through the hand-overs; the ring then
// refills from L2 at several times the SM's share of HBM bandwidth, and the consumers drain it at
// shared-memory speed.  Every weight byte still crosses HBM once (the ring copy of a prefetched
// stage hits L2, or merges with the fill in flight) and leaves L2 after its single use (evict_first
// on the ring copy).
// Pacing: the cursor that generates the HBM traffic issues at most one KB per `pace` SM cycles (0 =
// unpaced).  Every byte in flight beyond bandwidth x latency only adds queueing delay in front of
// the latency-critical LL traffic of the phase hand-overs; a paced producer keeps the queues short.
struct SchedCursor {
    int e, l;  // entry of the CTA's stage list, layer
};
__device__ __forceinline__ unsigned long long sched_stage(const CtaPlan *, const uint4 *tab, SchedCursor &c, int e_layer_end,
                                                          int token, uint32_t &bytes, uint32_t &nseg, uint32_t &sstr)
{
    const uint4 st = tab[c.e];
    const uint32_t kind = (st.w >> 8) & 0xfu;
    bytes = st.z; nseg = st.w & 0xffu; sstr = cp->sstride[kind];
    unsigned long long src = ((unsigned long long)st.y << 32 | st.x) + cp->lstride[kind] * (unsigned)c.l;
    if (kind == SK_EMB_ROW) src += (unsigned long long)(token - 1) * bytes;
    if (++c.e == e_layer_end && c.l + 1 < cp->L) { c.e = 1; c.l++; }
    return src;
}
__device__ __noinline__ void producer_loop(CtaPlan *, int token, int pace, int pf_lead)
{
    const uint64_t pol = l2_policy_evict_first();
    const uint4 *tab = reinterpret_cast<const uint4 *>(smem_base() + cp->off_sched);
    const uint32_t ns = (uint32_t)cp->n_slots;
    const int n_layer = 2 + cp->nst[0] + cp->nst[1] + cp->nst[2] + cp->nst[3];
    const int e_layer_end = 1 + n_layer, total = 1 + cp->L * n_layer + 1 + cp->nst[4];
    uint32_t slot = 0, par = 1;  // par: parity of the empty barrier to wait for
    SchedCursor rc{0, 0}, pc{0, 0};
    int s = 0, ps = 0;           // next stage of the ring cursor / of the prefetch cursor
    int unfetched = -1;          // the stage the prefetch cursor skipped (the ring cursor had caught up with it)
    long long next_ok = clock64();
#pragma unroll 1
    while (s < total) {
        bool busy = false;
        if (pf_lead > 0 && ps < total && ps <= s + pf_lead) {
            uint32_t b, n, ss;
            if (ps <= s) {
                // a stage the ring cursor is about to copy anyway is not worth a prefetch: stay ahead of it
                sched_stage(cp, tab, pc, e_layer_end, token, b, n, ss);
                unfetched = ps++;
                cp->pf_issued = ps;
                busy = true;
            } else if (clock64() >= next_ok) {
                const unsigned long long src = sched_stage(cp, tab, pc, e_layer_end, token, b, n, ss);
#pragma unroll 1
                for (uint32_t i = 0; i < n; i++) bulk_prefetch_l2(reinterpret_cast<const void *>(src + (unsigned long long)i * ss), b);
                next_ok = max(next_ok, clock64() - 2000) + (((long long)(b * n) * pace) >> 10);
                ps++;
                cp->pf_issued = ps;
                busy = true;
            }
        }
        // a copy that goes to HBM (not prefetched) shares the paced budget; one that hits L2 does not
        const bool to_hbm = pf_lead == 0 || s == unfetched;
        if (mbar_test(empty_bar(cp, slot), par) && !(to_hbm && pace > 0 && clock64() < next_ok)) {
            uint32_t bytes, nseg, sstr;
            const unsigned long long src = sched_stage(cp, tab, rc, e_layer_end, token, bytes, nseg, sstr);
            if (to_hbm) next_ok = max(next_ok, clock64() - 2000) + (((long long)(bytes * nseg) * pace) >> 10);
            uint64_t *fb = full_bar(cp, (uint32_t)s);
            mbar_arrive_expect_tx(fb, bytes * nseg);
            uint8_t *dst = smem_base() + (size_t)slot * cp->slot_bytes;
#pragma unroll 1
            for (uint32_t i = 0; i < nseg; i++)
                bulk_g2s(dst + i * bytes, reinterpret_cast<const void *>(src + (unsigned long long)i * sstr), bytes, fb, pol);
            if (++slot == ns) { slot = 0; par ^= 1u; }
            s++;
            cp->prod_issued = s;
            busy = true;
        }
        // Nothing to do right now: sleep instead of spinning.  This warp has the highest index of its scheduler
        // partition, i.e. issue priority over three consumer warps -- a tight polling loop here steals their slots.
        if (!busy) __nanosleep(64);
    }
}

// ------------------------------------------------------------------ LL buffers
// One float per 64-bit word: low half = value bits, high half = epoch.  64-bit scalar accesses
// are single-copy atomic, so value and epoch always arrive together; relaxed gpu-scope accesses
// go to L2 (never a stale L1 line).  A buffer is rewritten one layer later at the earliest, and
// a CTA can only get there after it has seen every other CTA's output of the phases in between,
// i.e. after every reader of the old contents is done (every prologue ends with a CTA-wide
// barrier, so "a CTA has published phase p" implies all its warps finished reading phase p - 1's
// inputs) -- no write-after-read hazard.
__device__ __forceinline__ void ll_store(unsigned long long *buf, int i, float v, uint32_t ep)
{
    const unsigned long long w = (unsigned long long)__float_as_uint(v) | ((unsigned long long)ep << 32);
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(buf + i), "l"(w) : "memory");
}
// system-scope variant: the target may be a peer GPU's buffer mapped over NVLink
__device__ __forceinline__ void ll_store_sys(unsigned long long *buf, int i, float v, uint32_t ep)
{
    const unsigned long long w = (unsigned long long)__float_as_uint(v) | ((unsigned long long)ep << 32);
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(buf + i), "l"(w) : "memory");
}
// two consecutive words (i even) with one 16-byte store; each 8-byte half is still self-validating
__device__ __forceinline__ void ll_store2(unsigned long long *buf, int i, float v0, float v1, uint32_t ep)
{
    const unsigned long long e = (unsigned long long)ep << 32;
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(buf + i), "l"(e | __float_as_uint(v0)),
                 "l"(e | __float_as_uint(v1))
                 : "memory");
}
__device__ __forceinline__ void ll_store2_sys(unsigned long long *buf, int i, float v0, float v1, uint32_t ep)
{
    const unsigned long long e = (unsigned long long)ep << 32;
    asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(buf + i), "l"(e | __float_as_uint(v0)),
                 "l"(e | __float_as_uint(v1))
                 : "memory");
}
__device__ __forceinline__ void ll_load2(const unsigned long long *p, unsigned long long &a, unsigned long long &b)
{
    asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ unsigned long long ll_load1(const unsigned long long *p)
{
    unsigned long long a;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(a) : "l"(p) : "memory");
    return a;
}
__device__ __forceinline__ float ll_val(unsigned long long w) { return __uint_as_float((uint32_t)w); }
__device__ __forceinline__ bool ll_ok(unsigned long long w, uint32_t ep) { return (uint32_t)(w >> 32) == ep; }
// poll four consecutive floats (i % 4 == 0)
__device__ __forceinline__ float4 ll_wait4(const unsigned long long *buf, int i, uint32_t ep)
{
    unsigned long long a, b, c, d;
    LLMF90_WD_DECL;
    do {
        ll_load2(buf + i, a, b);
        ll_load2(buf + i + 2, c, d);
        LLMF90_WD_CHECK(101, i, ep)
    } while (!(ll_ok(a, ep) && ll_ok(b, ep) && ll_ok(c, ep) && ll_ok(d, ep)));
    return make_float4(ll_val(a), ll_val(b), ll_val(c), ll_val(d));
}
// poll N consecutive float4 (i % 4 == 0); values are unpacked as they are checked, so that the raw
// {value, epoch} words do not all stay live (register pressure of the attention phase)
template <int N>
__device__ __forceinline__ void ll_wait4n(const unsigned long long *buf, int i, uint32_t ep, float4 (&o)[N])
{
    bool ok;
    LLMF90_WD_DECL;
    do {
        LLMF90_WD_CHECK(102, i, ep)
        ok = true;
#pragma unroll
        for (int k = 0; k < N; k++) {
            unsigned long long a, b, c, d;
            ll_load2(buf + i + 4 * k, a, b);
            ll_load2(buf + i + 4 * k + 2, c, d);
            ok = ok && ll_ok(a, ep) && ll_ok(b, ep) && ll_ok(c, ep) && ll_ok(d, ep);
            o[k] = make_float4(ll_val(a), ll_val(b), ll_val(c), ll_val(d));
        }
    } while (!ok);
}
// A thread's batch of PV float4 positions j = base + tid + k * NCT of an LL vector: all requests of
// a polling round are issued before the first check (one L2 round trip per round).  Positions
// past the end are clamped to the last one (a harmless duplicate request).
template <int PV>
__device__ __forceinline__ void ll_gather(const unsigned long long *buf, int n4, int base, uint32_t ep, float4 (&v)[PV])
{
    unsigned long long w[PV][4];
    int jj[PV];
#pragma unroll
    for (int k = 0; k < PV; k++) jj[k] = min(base + (int)threadIdx.x + k * NCT, n4 - 1);
    bool ok;
    LLMF90_WD_DECL;
    do {
        LLMF90_WD_CHECK(105, base, ep)
#pragma unroll
        for (int k = 0; k < PV; k++) {
            ll_load2(buf + 4 * jj[k], w[k][0], w[k][1]);
            ll_load2(buf + 4 * jj[k] + 2, w[k][2], w[k][3]);
        }
        ok = true;
#pragma unroll
        for (int k = 0; k < PV; k++)
            ok = ok && ll_ok(w[k][0], ep) && ll_ok(w[k][1], ep) && ll_ok(w[k][2], ep) && ll_ok(w[k][3], ep);
    } while (!ok);
#pragma unroll
    for (int k = 0; k < PV; k++) v[k] = make_float4(ll_val(w[k][0]), ll_val(w[k][1]), ll_val(w[k][2]), ll_val(w[k][3]));
}

// ---- vector stages: every consumer thread waits for the stage and reads what it needs; after
// the consumer-wide barrier that ends the prologue one thread hands the slot back
__device__ __forceinline__ const uint8_t *vec_stage_wait(const CtaPlan *, uint32_t s)
{
    mbar_wait(full_bar(cp, s), full_par(s), 3);
    return smem_base() + (size_t)slot_of(cp, s) * cp->slot_bytes;
}
__device__ __forceinline__ void vec_stage_release(const CtaPlan *, uint32_t s)
{
    if (threadIdx.x == 0) mbar_arrive_n(empty_bar(cp, slot_of(cp, s)), (uint32_t)NCW);
}

// ---- activation-vector prologue, one routine for all phases.  Every CTA needs the whole vector;
// it is 8-44 KB and was just written by the other CTAs, so it comes from L2 (LL words).
//   Wo / W2 (norm == false):  xs = the attention output / the SwiGLU output.
//   QKV / W13 / classifier:   x += sum over the tp ranks of their partial Wo / W2 outputs (the
//     fused all-reduce; ranks are added in rank order on every GPU, so the replicated stream stays
//     bit-identical), xs = x * w.  The common rmsnorm factor 1 / sqrt(mean(x^2) + 1e-5)
//     (llama2.f90:450-457) is returned and multiplies the phase's results in the epilogue (the
//     mat-vec is linear): no second pass over xs.  x is this CTA's copy of the residual stream in
//     shared memory; for the very first phase it is the embedding row from a ring slot (:520).
// f32 / f16: xs is the plain float vector.  q4_0 (consume_q4): xs holds the activations as f16
// pairs x = hi + lo in mma B-fragment order -- for every half group of 4 blocks, [hi | lo][block][t]
// records of 16 bytes: {x[4t], x[4t+2]}, {x[4t+1], x[4t+3]} of the low 16 elements of the block,
// then the same of the high 16 elements pre-scaled by 1/16 -- followed by the float corrections
// C[hi | lo][block] = 1032 sum(low elements) + 72 sum(high elements) (sums of the ROUNDED values, so
// that the offset 1024 + q -> q - 8 cancels exactly).  All 32 lanes call it together: an aligned
// group of 8 lanes holds the 8 float4 of one block.
// f16 activations: byte offset of the lo plane behind the hi plane of an n-vector -- 64 bytes past a multiple of 128,
// so that the four lanes reading hi and the four reading lo in one quarter-warp hit different banks (one wavefront)
__host__ __device__ __forceinline__ uint32_t f16_lo_off(int n) { return (((uint32_t)n * 2u + 127u) & ~127u) + 64u; }
template <int WT>
__device__ __forceinline__ void store_x4(float *xs, int n, int j4, const float4 v, bool valid)
{
    if constexpr (WT == WT_F32) {
        if (valid) reinterpret_cast<float4 *>(xs)[j4] = v;
    } else if constexpr (WT == WT_F16) {
        // f16 weights run on the tensor cores (tile_dot_f16): the activations are two f16 planes x = hi + lo
        // (hi = rn(x), lo = rn(x - hi): 22 bits of x, the products with f16 weights are exact in the f32
        // accumulators), hi[n] then lo[n] (f16_lo_off bytes further), each in column order
        if (valid) {
            const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
            const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
            const __half2 l01 = __floats2half2_rn(v.x - f01.x, v.y - f01.y), l23 = __floats2half2_rn(v.z - f23.x, v.w - f23.y);
            uint2 o;
            o.x = *reinterpret_cast<const uint32_t *>(&h01); o.y = *reinterpret_cast<const uint32_t *>(&h23);
            reinterpret_cast<uint2 *>(xs)[j4] = o;
            o.x = *reinterpret_cast<const uint32_t *>(&l01); o.y = *reinterpret_cast<const uint32_t *>(&l23);
            reinterpret_cast<uint2 *>(reinterpret_cast<uint8_t *>(xs) + f16_lo_off(n))[j4] = o;
        }
    } else {
        const int B = j4 >> 3, i = j4 & 7, t = i & 3, ngrp = q4t_groups(n);
        const bool hi_nib = i >= 4;
        const float sc = hi_nib ? 0.0625f : 1.f, cf = hi_nib ? 72.f * 16.f : 1032.f;
        const float x0 = v.x * sc, x1 = v.z * sc, x2 = v.y * sc, x3 = v.w * sc;
        const __half2 h01 = __floats2half2_rn(x0, x1), h23 = __floats2half2_rn(x2, x3);
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn(x0 - f01.x, x1 - f01.y), l23 = __floats2half2_rn(x2 - f23.x, x3 - f23.y);
        const float2 g01 = __half22float2(l01), g23 = __half22float2(l23);
        float ch = valid ? ((f01.x + f01.y) + (f23.x + f23.y)) * cf : 0.f;
        float cl = valid ? ((g01.x + g01.y) + (g23.x + g23.y)) * cf : 0.f;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            ch += __shfl_xor_sync(0xffffffffu, ch, o);
            cl += __shfl_xor_sync(0xffffffffu, cl, o);
        }
        if (valid) {
            uint8_t *rec = reinterpret_cast<uint8_t *>(xs) + ((size_t)((B >> 2) * 8 + (B & 3)) * 4 + t) * 16 + (hi_nib ? 8 : 0);
            uint2 o;
            o.x = *reinterpret_cast<const uint32_t *>(&h01); o.y = *reinterpret_cast<const uint32_t *>(&h23);
            *reinterpret_cast<uint2 *>(rec) = o;                  // p = 0: hi part
            o.x = *reinterpret_cast<const uint32_t *>(&l01); o.y = *reinterpret_cast<const uint32_t *>(&l23);
            *reinterpret_cast<uint2 *>(rec + 4 * 4 * 16) = o;     // p = 1: lo part, 4 blocks x 4 t further
            if (i == 0) {
                xs[(size_t)ngrp * 256 + B] = ch;
                xs[(size_t)ngrp * 256 + ngrp * 8 + B] = cl;
            }
        }
    }
}
// q4_0: blocks past the end of the vector (the last group of 8 may be partial) read as zero
__device__ __forceinline__ void q4_zero_tail(float *xs, int n)
{
    const int nblk = n >> 5, ngrp = q4t_groups(n);
    for (int b = nblk + (int)threadIdx.x; b < ngrp * 8; b += NCT) {
        uint8_t *rec = reinterpret_cast<uint8_t *>(xs) + (size_t)((b >> 2) * 8 + (b & 3)) * 64;
        uint4 *p = reinterpret_cast<uint4 *>(rec), *q = reinterpret_cast<uint4 *>(rec + 256);
        p[0] = p[1] = p[2] = p[3] = make_uint4(0u, 0u, 0u, 0u);
        q[0] = q[1] = q[2] = q[3] = make_uint4(0u, 0u, 0u, 0u);
        xs[(size_t)ngrp * 256 + b] = 0.f;
        xs[(size_t)ngrp * 256 + ngrp * 8 + b] = 0.f;
    }
}

__device__ __forceinline__ float4 emb_row4(const uint8_t *row, int wtype, int cols, int j4)
{
    return make_float4(row_elem(row, wtype, cols, 4 * j4), row_elem(row, wtype, cols, 4 * j4 + 1),
                       row_elem(row, wtype, cols, 4 * j4 + 2), row_elem(row, wtype, cols, 4 * j4 + 3));
}

// Tensor parallelism: x += the partial Wo / W2 outputs of ALL ranks (the fused all-reduce).  Every rank stored
// its partial vector into this GPU's [tp][emb] buffer (peer stores over NVLink); a thread takes its float4
// positions one at a time, requests that position from all tp vectors at once (one local L2 round trip per
// polling round, whatever tp is) and adds them in rank order -- every GPU adds the same numbers in the same
// order, the replicated residual stream stays bit-identical.  The sums go to the CTA's copy of x in shared
// memory.  Out of line: the 8 x 4 LL words in flight per lane would otherwise set the register budget of the
// single-GPU phase loop.  A poll that sees nothing for seconds means a peer is gone: flag it and stop waiting.
constexpr long long TP_POLL_TIMEOUT_CYCLES = 8000000000ll;  // ~4 s at 2 GHz
// poll position j of NS consecutive source vectors (n floats apart) and add them to v in source order
template <int NS>
__device__ __forceinline__ void tp_poll_add(const unsigned long long *src, int n, int j, uint32_t ep, float4 &v)
{
    unsigned long long w[NS][4];
    bool ok;
    const long long t0 = clock64();
    do {
#pragma unroll
        for (int r = 0; r < NS; r++) {
            ll_load2(src + (size_t)r * n + 4 * j, w[r][0], w[r][1]);
            ll_load2(src + (size_t)r * n + 4 * j + 2, w[r][2], w[r][3]);
        }
        ok = true;
#pragma unroll
        for (int r = 0; r < NS; r++) ok = ok && ll_ok(w[r][0], ep) && ll_ok(w[r][1], ep) && ll_ok(w[r][2], ep) && ll_ok(w[r][3], ep);
        if (!ok && (cp->abort || clock64() - t0 > TP_POLL_TIMEOUT_CYCLES)) {
            cp->abort = 1;
            if (cp->err_flag) *cp->err_flag = 1;
            break;
        }
    } while (!ok);
#pragma unroll
    for (int r = 0; r < NS; r++) { v.x += ll_val(w[r][0]); v.y += ll_val(w[r][1]); v.z += ll_val(w[r][2]); v.w += ll_val(w[r][3]); }
}
// (one instantiation per world size: the words in flight stay in registers; 8 sources are polled as two groups of
// four -- 64 registers of LL words do not fit beside the rest -- in rank order, so the sum is the same on every GPU)
template <int NS>
__device__ __noinline__ void gather_tp_n(const unsigned long long *src, uint32_t ep, int n4)
{
    float4 *xr4 = reinterpret_cast<float4 *>(smem_base() + cp->off_xres);
    const int n = n4 << 2;
#pragma unroll 1
    for (int j = (int)threadIdx.x; j < n4; j += NCT) {
        float4 v = xr4[j];
        if constexpr (NS == 8) {
            tp_poll_add<4>(src, n, j, ep, v);
            tp_poll_add<4>(src + (size_t)4 * n, n, j, ep, v);
        } else {
            tp_poll_add<NS>(src, n, j, ep, v);
        }
        xr4[j] = v;
    }
}
__device__ __forceinline__ void gather_tp(CtaPlan *, const unsigned long long *src, int nsrc, uint32_t ep, int n4)
{
    if (nsrc == 2) gather_tp_n<2>(src, ep, n4);
    else if (nsrc == 4) gather_tp_n<4>(src, ep, n4);
    else gather_tp_n<8>(src, ep, n4);
}

// The previous phase's mat-vec has no closing barrier (each warp publishes its tiles and moves on), so
// xs may still be read by a slower warp when a faster one gets here: poll first (that is the long
// part), then one consumer-wide barrier before the first write to xs, one after the last.
template <int WT>
__device__ __forceinline__ float gather_x(const CtaPlan *, const unsigned long long *src, int nsrc, uint32_t ep, int n,
                                       int norm, const uint8_t *emb_row, const float *wn /* shared */)
{
    float *xs = reinterpret_cast<float *>(smem_base() + cp->off_xs);
    float *red = reinterpret_cast<float *>(smem_base() + cp->off_red);
    float4 *xr4 = reinterpret_cast<float4 *>(smem_base() + cp->off_xres);
    const int tid = (int)threadIdx.x;
    constexpr int PV = 2;
    const int n4 = n >> 2;
    const float4 *wn4 = reinterpret_cast<const float4 *>(wn);
    float ss = 0.f;
    unsigned long long *gtr = cp->gx_trace;
#define GSTAMP(k_) do { if (gtr && tid == 0) gtr[k_] = (unsigned long long)clock64(); } while (0)
    GSTAMP(0);
    if (nsrc > 1 && !emb_row) gather_tp(const_cast<CtaPlan *>(cp), src, nsrc, ep, n4);
#pragma unroll 1
    for (int base = 0; base < n4; base += PV * NCT) {
        float4 v[PV];
        if (emb_row) {
#pragma unroll
            for (int k = 0; k < PV; k++) v[k] = emb_row4(emb_row, cp->wtype, n, min(base + tid + k * NCT, n4 - 1));
        } else if (nsrc > 1) {
            // tensor parallel: gather_tp has already added every rank's partials to this thread's positions of x
#pragma unroll
            for (int k = 0; k < PV; k++) v[k] = xr4[min(base + tid + k * NCT, n4 - 1)];
        } else {
            ll_gather<PV>(src, n4, base, ep, v);
            if (norm) {
#pragma unroll
                for (int k = 0; k < PV; k++) {
                    const float4 x = xr4[min(base + tid + k * NCT, n4 - 1)];
                    v[k].x += x.x; v[k].y += x.y; v[k].z += x.z; v[k].w += x.w;
                }
            }
        }
        if (base == 0) {
            GSTAMP(1);
            cons_sync();  // xs is free: every warp is past the previous phase's mat-vec
            GSTAMP(2);
        }
#pragma unroll
        for (int k = 0; k < PV; k++) {
            const int j = base + tid + k * NCT;
            const bool valid = j < n4;
            float4 t = v[k];
            if (valid && norm) {
                xr4[j] = t;
                ss = fmaf(t.x, t.x, ss); ss = fmaf(t.y, t.y, ss); ss = fmaf(t.z, t.z, ss); ss = fmaf(t.w, t.w, ss);
                const float4 w = wn4[j];
                t.x *= w.x; t.y *= w.y; t.z *= w.z; t.w *= w.w;
            }
            store_x4<WT>(xs, n, j, t, valid);
        }
    }
    if (WT == WT_Q4_0) q4_zero_tail(xs, n);
    GSTAMP(3);
    if (!norm) {
        cons_sync();
        GSTAMP(4);
        return 1.f;
    }
    ss = warp_sum(ss);
    if ((tid & 31) == 0) red[tid >> 5] = ss;
    cons_sync();
    GSTAMP(4);
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < NCW; i++) tot += red[i];
    return rsqrtf(fmaf(tot, cp->inv_emb, 1e-5f));  // 1 / sqrt(mean(x^2) + 1e-5), llama2.f90:454-456 (n == emb)
#undef GSTAMP
}

// ------------------------------------------------------------------ attention phase
// (head, split) items over the CTAs (llama2.f90:574-598); a split is a run of positions, chosen so that its
// cached rows fit one group per warp where the grid allows (engine.cu: n_splits_for; <= 256 in any case).  The
// phase is a pure latency chain between the QKV tiles and the Wo prologue, so it is organised for few dependent
// steps, not for throughput -- three stages separated by CTA barriers:
//   scores : a position is held by hs / 16 lanes (16 head dimensions each), a warp takes 512 / hs positions at a
//            time; all K / V rows of the warp are pulled into L2 first (prefetch: no destination registers while
//            the 32-register poll for q is in flight), then q is polled, the K quarters are loaded and dotted,
//            shuffles finish the dot, the score goes to shared memory; the V quarters of the warp's first group
//            are requested right after and arrive while the CTA meets at the barrier.  The last position is this
//            launch's own: the last warp polls its q, k and v from the LL buffers (attention_current).
//   softmax: every warp computes the maximum and the normaliser of ALL scores of the item itself (a few
//            shared-memory reads and two warp reductions: cheaper than exchanging them) -- no running
//            maximum, no rescaling (softmax :468-478); then acc = sum_t exp(s_t - max) v_t over the warp's
//            positions in the same lane mapping, the position groups folded with shuffles, one partial vector
//            per warp in shared memory.
//   publish: thread d adds the 12 partials of dimension pair d.  With one split the normalised head output goes
//            straight to ll_att; otherwise {max, normaliser, acc} partials go to ll_part and the Wo prologue
//            merges (load_x_attn_s).
// Positions past the end of a group read a clamped (valid) row and get weight 0.

__device__ __forceinline__ void prefetch_l2_line(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// The current position of an item: its query and key quarters and its value dims all come from the LL
// buffers (one polling loop: one L2 round trip).  Run by one warp; writes the score to *sc_out and parks
// the value dims in v_out.  (Inlined: a call inside the attention phase would spill everything that is live
// across it -- and it is a branch of its own, so its 2 x 16 LL words per lane do not add to the others.)
template <int HS>
__device__ __forceinline__ void attention_current(const unsigned long long *pq, const unsigned long long *pk,
                                               const unsigned long long *pv, uint32_t ep, float *sc_out, float *v_out)
{
    constexpr int LP = HS >> 4, vec = HS >> 5;  // lanes per position (16 dims each); value dims per lane
    constexpr float rscale = HS == 32 ? 0.17677669529663687f : (HS == 64 ? 0.125f : 0.08838834764831845f);  // 1 / sqrt(hs)
    const int lane = threadIdx.x & 31, dq = lane & (LP - 1);
    pq += dq * 16; pk += dq * 16; pv += lane * vec;
    float sdot, v[vec];
    bool ok;
    LLMF90_WD_DECL;
    do {
        LLMF90_WD_CHECK(102, 0, ep)
        ok = true;
        sdot = 0.f;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            unsigned long long a, b, c, d, e, f, g, h;
            ll_load2(pq + 4 * k, a, b);
            ll_load2(pq + 4 * k + 2, c, d);
            ll_load2(pk + 4 * k, e, f);
            ll_load2(pk + 4 * k + 2, g, h);
            ok = ok && ll_ok(a, ep) && ll_ok(b, ep) && ll_ok(c, ep) && ll_ok(d, ep) && ll_ok(e, ep) && ll_ok(f, ep) &&
                 ll_ok(g, ep) && ll_ok(h, ep);
            sdot = fmaf(ll_val(a), ll_val(e), sdot); sdot = fmaf(ll_val(b), ll_val(f), sdot);
            sdot = fmaf(ll_val(c), ll_val(g), sdot); sdot = fmaf(ll_val(d), ll_val(h), sdot);
        }
#pragma unroll
        for (int k = 0; k < vec; k++) {
            const unsigned long long a = ll_load1(pv + k);
            ok = ok && ll_ok(a, ep);
            v[k] = ll_val(a);
        }
    } while (!ok);
#pragma unroll
    for (int o = 1; o < LP; o <<= 1) sdot += __shfl_xor_sync(0xffffffffu, sdot, o);
    if (lane == 0) *sc_out = sdot * rscale;
#pragma unroll
    for (int k = 0; k < vec; k++) v_out[lane * vec + k] = v[k];
}

// Fields of the CTA plan read with a VOLATILE shared-memory load: the value cannot be hoisted above the asm
// volatile barrier in front of it nor merged with an earlier load of the same field.  The attention phase is three
// stages separated by CTA barriers; each stage re-derives what it needs from (item, layer) and fresh plan loads,
// so nothing but the loop counter is live across a barrier -- left to itself the compiler computes every pointer
// and index at the top, spills them around the 32-register poll for q and reloads them one by one after each
// barrier, and a local-memory load is an L2 round trip here (the ring leaves ~20 KB of L1): measured 3.7 K
// cycles for the 12-term sum + publish of the last stage alone.
__device__ __forceinline__ uint32_t plan_u32(const void *p)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long plan_u64(const void *p)
{
    unsigned long long v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
#define PLAN_I(f_) ((int)plan_u32(&cp->f_))
#define PLAN_P(T_, f_) (reinterpret_cast<T_>(plan_u64(&cp->f_)))

// the positions of attention item (head h, split sp) at position `pos` with S splits
struct AttItem {
    int h, sp, t0, t1, tc1, npast;
    bool cur_here;
};
__device__ __forceinline__ AttItem att_item(int item, int S, int pos)
{
    AttItem a;
    const int s_shift = 31 - __clz(S);
    const int chunk = (((pos + S - 1) >> s_shift) + 15) & ~15;
    a.h = item >> s_shift; a.sp = item & (S - 1);
    a.npast = pos - 1;  // positions 0 .. pos-2 come from the cache (earlier launches); npast is this launch's
    a.t0 = a.sp * chunk; a.t1 = min(pos, a.t0 + chunk);  // this split's positions
    a.tc1 = min(a.t1, a.npast);                           // ... of which [t0, tc1) are cached rows
    a.cur_here = a.npast >= a.t0 && a.npast < a.t1;       // the split ends with this launch's own position
    return a;
}

template <int HS, bool PROF>
__device__ __noinline__ void attention_phase_t(const CtaPlan *, int layer)
{
    // 8 trace words (profiling kernel)
#define ASTAMP(k_) do { if constexpr (PROF) { unsigned long long *tr_ = PLAN_P(unsigned long long *, trace); if (tr_ && threadIdx.x == 0) tr_[48 + (k_)] = (unsigned long long)clock64(); } } while (0)
    ASTAMP(0);
    // a position is held by LP lanes of 16 head dimensions each; a warp takes PG positions at a time
    constexpr int hs = HS, vec = HS >> 5, LP = HS >> 4, PG = 32 / LP;  // HS in {32, 64, 128}: LP 2 / 4 / 8, PG 16 / 8 / 4
    constexpr int lp_shift = HS == 32 ? 1 : (HS == 64 ? 2 : 3);
    constexpr float rscale = HS == 32 ? 0.17677669529663687f : (HS == 64 ? 0.125f : 0.08838834764831845f);  // 1 / sqrt(hs)
#pragma unroll 1
    for (int item = blockIdx.x; item < PLAN_I(H) * PLAN_I(n_splits); item += gridDim.x) {
        // value row quarter (position tb0 + pl, dims 16 dq ..) of the warp's first group: requested in stage 1, used in
        // stage 2 (same lane mapping as the key rows: 4 x float4 per lane)
        float4 vq[4];
        // ================================================================ stage 1: scores
        {
            const int warp = (int)threadIdx.x >> 5, lane = (int)threadIdx.x & 31;
            const int pl = lane >> lp_shift, dq = lane & (LP - 1);
            const int kv = PLAN_I(kv);
            const AttItem a = att_item(item, PLAN_I(n_splits), PLAN_I(pos));
            // (h / kv_mul by multiplication: llama2.f90:581 maps head h to KV head h / kv_mul)
            const int g = (int)(((uint32_t)a.h * plan_u32(&cp->kvmul_inv16)) >> 16);
            const int nwg = a.cur_here ? NCW - 1 : NCW;  // the last warp takes the current position, the others the cached rows
            const bool cur_warp = a.cur_here && warp == NCW - 1;
            const uint32_t ep = plan_u32(&cp->ep_base) + (uint32_t)layer + 1u;
            float *sc = reinterpret_cast<float *>(smem_base() + PLAN_I(off_att));  // [ATT_MAX_CHUNK] scores
            const size_t row0 = (size_t)layer * PLAN_I(seq) * kv + g * hs;
            const float *kbase = PLAN_P(const float *, kc) + row0 + dq * 16;
            const float *vc = PLAN_P(const float *, vc) + row0;
            const unsigned long long *pq = PLAN_P(const unsigned long long *, ll_q) + (size_t)PLAN_I(rep) * PLAN_I(att_dim) + a.h * hs;
            const int tb0 = a.t0 + PG * warp;  // the warp's first group
            if (cur_warp) {
                const unsigned long long *pk = PLAN_P(const unsigned long long *, ll_kv) + (size_t)PLAN_I(rep) * 2 * kv + g * hs;
                attention_current<HS>(pq, pk, pk + kv, ep, sc + (a.npast - a.t0), sc + ATT_MAX_CHUNK + warp * hs);
            } else if (tb0 < a.tc1) {
                // all K / V rows of the warp are pulled into L2 now (no destination register); only the query poll's
                // 32 registers are live while waiting
#pragma unroll 1
                for (int tb = tb0; tb < a.tc1; tb += PG * nwg) {
                    const uint32_t ro = (uint32_t)(min(tb + pl, a.tc1 - 1) * kv);
                    prefetch_l2_line(kbase + ro);
                    prefetch_l2_line(vc + dq * 16 + ro);
                }
                float4 qq[4];
                ASTAMP(1);
                ll_wait4n<4>(pq, dq * 16, ep, qq);  // (written by the QKV epilogues of this launch)
                ASTAMP(2);
#pragma unroll 1
                for (int tb = tb0; tb < a.tc1; tb += PG * nwg) {
                    const float4 *kr = reinterpret_cast<const float4 *>(kbase + (uint32_t)(min(tb + pl, a.tc1 - 1) * kv));
                    float4 kk[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) kk[i] = __ldcg(kr + i);
                    float sdot = 0.f;
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        sdot = fmaf(qq[i].x, kk[i].x, sdot); sdot = fmaf(qq[i].y, kk[i].y, sdot);
                        sdot = fmaf(qq[i].z, kk[i].z, sdot); sdot = fmaf(qq[i].w, kk[i].w, sdot);
                    }
#pragma unroll
                    for (int o = 1; o < LP; o <<= 1) sdot += __shfl_xor_sync(0xffffffffu, sdot, o);
                    if (dq == 0 && tb + pl < a.tc1) sc[tb + pl - a.t0] = sdot * rscale;  // dot_product(q_t,k_t)/sqrt(head_size), :582
                }
                const float4 *vr = reinterpret_cast<const float4 *>(vc + dq * 16 + (uint32_t)(min(tb0 + pl, a.tc1 - 1) * kv));
#pragma unroll
                for (int i = 0; i < 4; i++) vq[i] = __ldcg(vr + i);
            }
        }
        ASTAMP(3);
        cons_sync();
        // ================================================================ stage 2: softmax statistics, values
        float M = -INFINITY, L = 0.f;
        {
            const int warp = (int)threadIdx.x >> 5, lane = (int)threadIdx.x & 31;
            const AttItem a = att_item(item, PLAN_I(n_splits), PLAN_I(pos));
            const float *sc = reinterpret_cast<const float *>(smem_base() + PLAN_I(off_att));
            float *part = reinterpret_cast<float *>(smem_base() + PLAN_I(off_att)) + ATT_MAX_CHUNK;  // [NCW][hs] partial outputs
            // the statistics of the whole item, in every warp (:468-478)
            const int n = a.t1 - a.t0;
#pragma unroll 1
            for (int i = lane; i < n; i += 32) M = fmaxf(M, sc[i]);
            M = warp_max(M);
#pragma unroll 1
            for (int i = lane; i < n; i += 32) L += __expf(sc[i] - M);
            L = warp_sum(L);
            // acc = sum over the warp's positions of exp(s_t - M) v_t: lane (pl, dq) adds up its positions' quarter
            // rows, the PG position groups are folded with shuffles, the lanes pl = 0 hold the warp's partial vector
            if (a.cur_here && warp == NCW - 1) {
                const float p = __expf(sc[a.npast - a.t0] - M);
#pragma unroll
                for (int i = 0; i < vec; i++) part[warp * hs + lane * vec + i] *= p;  // (parked by attention_current)
            } else {
                const int pl = lane >> lp_shift, dq = lane & (LP - 1);
                const int nwg = a.cur_here ? NCW - 1 : NCW;
                const int tb0 = a.t0 + PG * warp;
                float4 acc[4];
#pragma unroll
                for (int i = 0; i < 4; i++) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
                for (int tb = tb0; tb < a.tc1; tb += PG * nwg) {
                    if (tb != tb0) {
                        const int kv = PLAN_I(kv);
                        const int g = (int)(((uint32_t)a.h * plan_u32(&cp->kvmul_inv16)) >> 16);
                        const float4 *vr = reinterpret_cast<const float4 *>(
                            PLAN_P(const float *, vc) + ((size_t)layer * PLAN_I(seq) * kv + g * hs + dq * 16) + (uint32_t)(min(tb + pl, a.tc1 - 1) * kv));
#pragma unroll
                        for (int i = 0; i < 4; i++) vq[i] = __ldcg(vr + i);
                    }
                    const float p = tb + pl < a.tc1 ? __expf(sc[tb + pl - a.t0] - M) : 0.f;
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        acc[i].x = fmaf(p, vq[i].x, acc[i].x); acc[i].y = fmaf(p, vq[i].y, acc[i].y);
                        acc[i].z = fmaf(p, vq[i].z, acc[i].z); acc[i].w = fmaf(p, vq[i].w, acc[i].w);
                    }
                }
#pragma unroll
                for (int o = LP; o < 32; o <<= 1)
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        acc[i].x += __shfl_xor_sync(0xffffffffu, acc[i].x, o); acc[i].y += __shfl_xor_sync(0xffffffffu, acc[i].y, o);
                        acc[i].z += __shfl_xor_sync(0xffffffffu, acc[i].z, o); acc[i].w += __shfl_xor_sync(0xffffffffu, acc[i].w, o);
                    }
                if (pl == 0) {
                    float4 *dst = reinterpret_cast<float4 *>(part + warp * hs + dq * 16);
#pragma unroll
                    for (int i = 0; i < 4; i++) dst[i] = acc[i];
                }
            }
        }
        cons_sync();
        ASTAMP(5);
        // ================================================================ stage 3: thread (dimension pair, LL replica): add the 12 partial vectors, publish
        {
            const int tid = (int)threadIdx.x;
            const int S = PLAN_I(n_splits);
            const int nout = S == 1 ? PLAN_I(ll_rep) : 1;
            if (tid < (hs >> 1) * nout) {
                const float *part = reinterpret_cast<const float *>(smem_base() + PLAN_I(off_att)) + ATT_MAX_CHUNK;
                const int d = (tid & ((hs >> 1) - 1)) << 1, rr = tid >> (HS == 32 ? 4 : (HS == 64 ? 5 : 6));
                float A0 = 0.f, A1 = 0.f;
#pragma unroll
                for (int w = 0; w < NCW; w++) {
                    const float2 p2 = *reinterpret_cast<const float2 *>(part + w * hs + d);
                    A0 += p2.x; A1 += p2.y;
                }
                const uint32_t ep = plan_u32(&cp->ep_base) + (uint32_t)layer + 1u;
                const int s_shift = 31 - __clz(S), h = item >> s_shift, sp = item & (S - 1);
                if (S == 1) {
                    const float invL = __fdividef(1.0f, L);
                    ll_store2(PLAN_P(unsigned long long *, ll_att) + (size_t)rr * PLAN_I(att_dim), h * hs + d, A0 * invL, A1 * invL, ep);
                } else {
                    unsigned long long *out = PLAN_P(unsigned long long *, ll_part) + (size_t)(h * S + sp) * (hs + ATT_PSTRIDE_PAD);
                    ll_store2(out, ATT_PSTRIDE_PAD + d, A0, A1, ep);
                    if (d == 0) ll_store2(out, 0, M, L, ep);  // (a split always has at least one position: M is finite)
                }
            }
        }
        ASTAMP(6);
        cons_sync();
    }
    ASTAMP(7);
#undef ASTAMP
}

// xs = attention output (all heads), merging the position splits (n_splits > 1).  A thread owns one
// float4 of the output; the {m, l} pair and the float4 of all S partial records are requested
// together (one L2 round trip per polling round).
template <int WT, int S>
__device__ __noinline__ void load_x_attn_s(const CtaPlan *, uint32_t ep)
{
    float *xs = reinterpret_cast<float *>(smem_base() + cp->off_xs);
    const int tid = (int)threadIdx.x;
    const int hs = cp->hs, pstride = hs + ATT_PSTRIDE_PAD, n4 = cp->att_dim >> 2;
    const int hs_shift = hs == 64 ? 6 : (hs == 128 ? 7 : 5);
    bool first = true;
    for (int jj = tid; jj < ((n4 + 31) & ~31) || first; jj += NCT) {
        const bool valid = jj < n4;
        const int j = valid ? jj : n4 - 1;
        const int h = (4 * j) >> hs_shift, d = (4 * j) & (hs - 1);
        const unsigned long long *part = cp->ll_part + (size_t)h * S * pstride;
        unsigned long long ml[S][2], av[S][4];
        bool ok;
        LLMF90_WD_DECL;
        do {
            LLMF90_WD_CHECK(107, j, ep)
#pragma unroll
            for (int s = 0; s < S; s++) {
                const unsigned long long *rec = part + (size_t)s * pstride;
                ll_load2(rec, ml[s][0], ml[s][1]);
                ll_load2(rec + ATT_PSTRIDE_PAD + d, av[s][0], av[s][1]);
                ll_load2(rec + ATT_PSTRIDE_PAD + d + 2, av[s][2], av[s][3]);
            }
            ok = true;
#pragma unroll
            for (int s = 0; s < S; s++)
                ok = ok && ll_ok(ml[s][0], ep) && ll_ok(ml[s][1], ep) && ll_ok(av[s][0], ep) && ll_ok(av[s][1], ep) &&
                     ll_ok(av[s][2], ep) && ll_ok(av[s][3], ep);
        } while (!ok);
        float M = -INFINITY;
#pragma unroll
        for (int s = 0; s < S; s++) M = fmaxf(M, ll_val(ml[s][0]));
        float den = 0.f;
        float4 num = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int s = 0; s < S; s++)
            if (ll_val(ml[s][0]) > -INFINITY) {
                const float w = __expf(ll_val(ml[s][0]) - M);
                den = fmaf(ll_val(ml[s][1]), w, den);
                num.x = fmaf(ll_val(av[s][0]), w, num.x); num.y = fmaf(ll_val(av[s][1]), w, num.y);
                num.z = fmaf(ll_val(av[s][2]), w, num.z); num.w = fmaf(ll_val(av[s][3]), w, num.w);
            }
        if (first) { cons_sync(); first = false; }  // xs is free (see gather_x); every thread gets here once
        const float rden = __fdividef(1.0f, den);
        if (jj < ((n4 + 31) & ~31))
            store_x4<WT>(xs, cp->att_dim, jj, make_float4(num.x * rden, num.y * rden, num.z * rden, num.w * rden), valid);
    }
    if (WT == WT_Q4_0) q4_zero_tail(xs, cp->att_dim);
    cons_sync();
}

// (one instantiation per split count: the S records of a thread's float4 stay in registers -- arrays sized for the
// maximum and guarded by a run-time count went to local memory)
template <int WT>
__device__ __noinline__ void load_x_attn(const CtaPlan *, int S, uint32_t ep)
{
    if (S == 2) load_x_attn_s<WT, 2>(cp, ep);
    else if (S == 4) load_x_attn_s<WT, 4>(cp, ep);
    else load_x_attn_s<WT, 8>(cp, ep);
}

// ------------------------------------------------------------------ profiling hooks (out of line)
__device__ __noinline__ void prof_lap(Prof *, int bucket)
{
    const long long now = clock64();
    pf->tacc[bucket] += now - pf->tmark;
    pf->tmark = now;
}
// per-CTA trace of one layer: globaltimer stamp at phase edge k
__device__ __noinline__ void prof_stamp(unsigned long long *trace, int k)
{
    trace[(size_t)blockIdx.x * 128 + k] = globaltimer_ns();
}

// ------------------------------------------------------------------ token tail (once per launch)
// all-gathered logits must have landed on every rank before any rank's kernel ends (tp > 1), then
// maxloc(logits) (llama2.f90:388): first maximum wins at every reduction level
__device__ __noinline__ void token_tail(const CtaPlan *, int pos)
{
    const int tid = (int)threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t epl = cp->ep_last;
    const int tp = cp->tp, rank = cp->rank;
    const int G = (int)gridDim.x;
    if (tp > 1) {
        // each CTA flags every rank once its rows are stored, CTA 0 of every rank collects the flags
        cons_sync();
        if (tid == 0) {
            asm volatile("fence.acq_rel.sys;" ::: "memory");
            for (int k = 0; k < tp; k++) ll_store_sys(cp->done[k], rank * G + (int)blockIdx.x, 0.f, epl);
        }
        if (blockIdx.x == 0) {
            for (int i = tid; i < tp * G; i += NCT) {
                const long long t0 = clock64();
                while (!ll_ok(ll_load1(cp->done[rank] + i), epl)) {
                    if (cp->abort || clock64() - t0 > TP_POLL_TIMEOUT_CYCLES) {  // a peer is gone (see gather_tp)
                        if (cp->err_flag) *cp->err_flag = 1;
                        break;
                    }
                }
            }
            asm volatile("fence.acq_rel.sys;" ::: "memory");
        }
    }
    if (!cp->do_argmax) return;
    cons_sync();  // every warp's maxloc record (run_tiles) is in shared memory
    if (tid == 0) {
        float best = cp->tail_best[0];
        int bidx = cp->tail_idx[0];
        for (int w = 1; w < NCW; w++)
            if (cp->tail_best[w] > best || (cp->tail_best[w] == best && cp->tail_idx[w] < bidx)) { best = cp->tail_best[w]; bidx = cp->tail_idx[w]; }
        for (int k = 0; k < tp; k++) {
            unsigned long long *dst = cp->amax[k] + (size_t)(rank * G + (int)blockIdx.x) * 2;
            ll_store_sys(dst, 0, best, epl);
            ll_store_sys(dst, 1, __int_as_float(bidx), epl);
        }
    }
    if (blockIdx.x == 0 && warp == 0) {
        // every rank reduces the same tp * grid records in the same order -> the same token
        float best = -INFINITY;
        int bidx = 0x7fffffff;
        for (int i = lane; i < tp * G; i += 32) {
            unsigned long long ra, rb;
            const long long t0 = clock64();
            for (;;) {
                ll_load2(cp->amax[rank] + 2 * i, ra, rb);
                if (ll_ok(ra, epl) && ll_ok(rb, epl)) break;
                if (tp > 1 && (cp->abort || clock64() - t0 > TP_POLL_TIMEOUT_CYCLES)) {
                    if (cp->err_flag) *cp->err_flag = 1;
                    break;
                }
            }
            const float v = ll_val(ra);
            const int ix = __float_as_int(ll_val(rb));
            if (v > best || (v == best && ix < bidx)) { best = v; bidx = ix; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
            if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
        }
        if (lane == 0) {
            // (NaN logits compare false everywhere: bidx stays the sentinel -- fall back to token 1 instead
            // of indexing the embedding table out of bounds in the next launch)
            int next = bidx == 0x7fffffff ? 1 : bidx + 1;
            if (cp->forced && cp->forced[pos - 1] > 0) next = cp->forced[pos - 1];
            if (cp->out_tokens) cp->out_tokens[pos - 1] = next;
            cp->tokpos[0] = next;
            cp->tokpos[1] = pos + 1;
        }
    }
}

// ------------------------------------------------------------------ the mat-vec of one tile chunk
// f32 / f16: the warp's 32 lanes split the chunk's 16-byte units; one activation load (two for f16: 8
// weights need 8 floats) serves the four rows of the tile.  Rows past the tile's valid count hold stale
// shared memory: their sums are computed and thrown away (no branch in the loop).
// The loads are volatile asm on purpose: the compiler otherwise sinks every row's load next to its
// FMAs and reuses one register quad for all rows -- five serialised shared-memory latencies per step
// (measured: 300 cycles per step, 7 bytes/cycle/warp).  In program order all loads of a step (and of
// the next step: the loop is software-pipelined by hand) are issued before the FMAs of this one.
__device__ __forceinline__ uint4 lds128(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
template <int WT>
struct TileRegs {
    uint4 w[4];
    uint4 x[WT == WT_F16 ? 2 : 1];
};
template <int WT>
__device__ __forceinline__ void tile_load(TileRegs<WT> &r, uint32_t wp, uint32_t cs, uint32_t xp)
{
    r.w[0] = lds128(wp); r.w[1] = lds128(wp + cs); r.w[2] = lds128(wp + 2 * cs); r.w[3] = lds128(wp + 3 * cs);
    r.x[0] = lds128(xp);
    if constexpr (WT == WT_F16) r.x[1] = lds128(xp + 16);
}
template <int WT>
__device__ __forceinline__ void tile_fma(const TileRegs<WT> &r, float (&acc)[4])
{
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint4 w = r.w[i];
        float a = acc[i];
        if constexpr (WT == WT_F32) {
            a = fmaf(__uint_as_float(w.x), __uint_as_float(r.x[0].x), a);
            a = fmaf(__uint_as_float(w.y), __uint_as_float(r.x[0].y), a);
            a = fmaf(__uint_as_float(w.z), __uint_as_float(r.x[0].z), a);
            a = fmaf(__uint_as_float(w.w), __uint_as_float(r.x[0].w), a);
        } else {
            const float2 f0 = __half22float2(*reinterpret_cast<const __half2 *>(&w.x));
            const float2 f1 = __half22float2(*reinterpret_cast<const __half2 *>(&w.y));
            const float2 f2 = __half22float2(*reinterpret_cast<const __half2 *>(&w.z));
            const float2 f3 = __half22float2(*reinterpret_cast<const __half2 *>(&w.w));
            a = fmaf(f0.x, __uint_as_float(r.x[0].x), a); a = fmaf(f0.y, __uint_as_float(r.x[0].y), a);
            a = fmaf(f1.x, __uint_as_float(r.x[0].z), a); a = fmaf(f1.y, __uint_as_float(r.x[0].w), a);
            a = fmaf(f2.x, __uint_as_float(r.x[1].x), a); a = fmaf(f2.y, __uint_as_float(r.x[1].y), a);
            a = fmaf(f3.x, __uint_as_float(r.x[1].z), a); a = fmaf(f3.y, __uint_as_float(r.x[1].w), a);
        }
        acc[i] = a;
    }
}
// sp: the stage in its ring slot (rows cs bytes apart); x: the chunk's first activation; nu: 16-byte
// weight units per row in this chunk.  Two register sets in rotation: a set is reloaded right after its
// FMAs and used one set of FMAs later (the rest of the shared-memory latency is covered by the other
// warps; a third set costs 20 registers, which this function does not have: it must not spill).
template <int WT>
__device__ __forceinline__ void tile_dot(const uint8_t *sp, uint32_t cs, const float *x, int nu, int lane, float (&acc)[4])
{
    constexpr uint32_t XB = WT == WT_F16 ? 32u : 16u;  // activation bytes per weight unit
    uint32_t wp = smem_u32(sp) + (uint32_t)lane * 16u, xp = smem_u32(x) + (uint32_t)lane * XB;
    const int nfull = nu >> 5;  // steps in which every lane has a unit
    TileRegs<WT> r0, r1;
    if (nfull > 0) tile_load<WT>(r0, wp, cs, xp);
    if (nfull > 1) tile_load<WT>(r1, wp + 512u, cs, xp + 32u * XB);
#pragma unroll 1
    for (int i = 0; i < nfull; i += 2) {
        tile_fma<WT>(r0, acc);
        if (i + 2 < nfull) tile_load<WT>(r0, wp + 1024u, cs, xp + 64u * XB);
        if (i + 1 < nfull) {
            tile_fma<WT>(r1, acc);
            if (i + 3 < nfull) tile_load<WT>(r1, wp + 1536u, cs, xp + 96u * XB);
        }
        wp += 1024u; xp += 64u * XB;
    }
    if (lane < (nu & 31)) {  // the last, partial step of a row whose unit count is not a multiple of 32
        const uint32_t o = (uint32_t)nfull;
        tile_load<WT>(r0, smem_u32(sp) + (o * 32u + (uint32_t)lane) * 16u, cs, smem_u32(x) + (o * 32u + (uint32_t)lane) * XB);
        tile_fma<WT>(r0, acc);
    }
}

__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1)
{
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// f16 on the tensor cores (legacy mma.sync m16n8k16, f16 x f16 -> f32).  On the FMA pipe an f16 weight costs a
// conversion and an FMA (13 instructions per 16 bytes of weights: the mat-vec is issue-bound at a third of the
// shared-memory rate); here the WEIGHTS are the B operand -- n = the 8 rows of the tile, k = 16 columns -- and
// the activations the A operand: row 0 = the hi plane, row 1 = the lo plane of x = hi + lo (store_x4), the other
// 14 rows zero.  D[0][n] + D[1][n] = dot(row n, x): two instructions per 512 bytes of weights.  The contraction
// index inside a k-step may be permuted freely as long as both operands agree, so lane (g, t) takes 16
// CONTIGUOUS bytes of row g -- columns c0 .. c0 + 7, c0 = 32 step + 8 t -- for two mmas (b0 / b1 of the first =
// halves 0,1 / 2,3, of the second = 4,5 / 6,7) and the lanes g < 2 the same 8 columns of the hi / lo plane
// (a0 / a2 likewise): one 128-bit load each.
// Layout of a stage (tile_pass_kernel wrote the matrix that way, the bulk copy moves it verbatim): k-step major,
// [step][row][64 bytes], so a warp's load is 512 contiguous bytes -- no bank conflicts; a row whose unit count
// is not a multiple of 4 ends with a short step [row][rem x 16 bytes].  A tile of v < 8 rows has v rows per step;
// the lanes g >= v then read the following steps' bytes: finite or not, they only reach the D columns n >= v,
// whose rows do not exist and are dropped in the epilogue.
// sp: the stage; xh: the hi plane at the chunk's first column, the lo plane xl_off bytes further; this warp takes
// the k-steps [s0, s1) of the chunk's (nu + 3) / 4.
__device__ __forceinline__ void tile_dot_f16(const uint8_t *sp, uint32_t v, const __half *xh, uint32_t xl_off, int nu,
                                             int s0, int s1, int lane, float (&dA)[4], float (&dB)[4], uint32_t zero16)
{
    const int g = lane >> 2, t = lane & 3, nfull = nu >> 2;
    const uint32_t sstep = v * 64u;
    uint32_t wp = smem_u32(sp) + (uint32_t)s0 * sstep + (uint32_t)lane * 16u;
    // the lanes g >= 2 hold the zero rows 2 .. 15 of the A operand: they read 16 zero bytes (zero16, stride 0)
    // instead of branching on the lane in the loop -- every value the loop needs then fits a register that is
    // only live here (the compiler kept `lane` in local memory and reloaded it every step)
    const bool ax = g < 2;
    uint32_t xp = ax ? smem_u32(xh) + (g == 1 ? xl_off : 0u) + (uint32_t)(s0 * 4 + t) * 16u : zero16;
    const uint32_t xstep = ax ? 64u : 0u;
    const int e = min(s1, nfull);
    // software-pipelined by hand (the loads are volatile asm: in program order the next step's are issued before
    // this step's mmas wait for theirs)
    uint4 w = make_uint4(0u, 0u, 0u, 0u), x = make_uint4(0u, 0u, 0u, 0u);
    if (s0 < e) { w = lds128(wp); x = lds128(xp); }
#pragma unroll 2
    for (int st = s0; st < e; st++) {
        wp += sstep; xp += xstep;
        uint4 wn = w, xn = x;
        if (st + 1 < e) { wn = lds128(wp); xn = lds128(xp); }
        mma16816(dA, x.x, 0u, x.y, 0u, w.x, w.y);
        mma16816(dB, x.z, 0u, x.w, 0u, w.z, w.w);
        w = wn; x = xn;
    }
    if (s1 > nfull) {  // the short last step of the row: rem units per row, rows rem x 16 bytes apart
        const int rem = nu & 3;
        w = make_uint4(0u, 0u, 0u, 0u); x = make_uint4(0u, 0u, 0u, 0u);
        if (t < rem) {
            w = lds128(smem_u32(sp) + (uint32_t)nfull * sstep + (uint32_t)(g * rem + t) * 16u);
            if (ax) x = lds128(smem_u32(xh) + (g == 1 ? xl_off : 0u) + (uint32_t)(nfull * 4 + t) * 16u);
        }
        mma16816(dA, x.x, 0u, x.y, 0u, w.x, w.y);
        mma16816(dB, x.z, 0u, x.w, 0u, w.z, w.w);
    }
}

// q4_0 on the tensor cores (legacy mma.sync m16n8k16, f16 x f16 -> f32): the dequantisation is the
// instruction bottleneck of a q4_0 mat-vec at B200's HBM rate, and on CUDA cores it costs >= 2
// instructions per weight.  Here a nibble pair becomes a half2 {1024 + q} with ONE lop3 (the 0x6400
// exponent trick; high nibbles give 1024 + 16 q and meet activations pre-scaled by 1/16, exact),
// the products run on the tensor pipe, and the offsets are removed per block with the pre-computed
// C[b] = 1032 sum(x over the low nibbles) + 72 sum(x over the high nibbles).  The block scale
// cannot be applied inside the mma, so the n dimension separates blocks: the B operand
// (activations) of the mma pair of block b is non-zero only in columns b and 4 + b, and D[row][b],
// D[row][4 + b] end up holding the unscaled sums of block b against the two f16 halves x = hi + lo
// of the activations (f16 alone would cost 3 digits: greedy tokens flip at near ties).
// Tiled weight format: common.cuh.  A stage holds the 8-block groups [g0, g0 + ng) of one row group of
// 16 rows; lane = 4 g + t accumulates rows g and g + 8 over its blocks.
// (w & mask) | bias in ONE instruction: two nibbles of w -> half2 {1024 + q, 1024 + q'} (or 1024 + 16 q for the high mask)
__device__ __forceinline__ uint32_t nib_half2(uint32_t w, uint32_t mask, uint32_t bias)
{
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(w), "r"(mask), "r"(bias));
    return d;
}
__device__ __forceinline__ uint32_t uint4_word(const uint4 &v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }

__device__ __forceinline__ void tile_dot_q4(const uint8_t *sp, const float *xs, int ngrp, int g0, int ng, int lane,
                                            float &acc0, float &acc1, uint32_t zero16)
{
    const int g = lane >> 2, t = lane & 3;
    // activations as f16 pairs x = hi + lo in B-fragment order, [half group of 4 blocks][hi | lo][block][t],
    // then the per-block offset corrections [hi | lo][block] (store_x4)
    const uint4 *xh4 = reinterpret_cast<const uint4 *>(xs);
    const float *C = xs + (size_t)ngrp * 256 + (size_t)(t >> 1) * ngrp * 8;
    // the two nibble masks and the exponent bias live in registers (opaque to the compiler, which would otherwise
    // fold them into immediates and need an AND and an OR per half2: LOP3 takes one immediate)
    uint32_t mlo, mhi, bias;
    asm volatile("mov.b32 %0, 0x000f000f;" : "=r"(mlo));
    asm volatile("mov.b32 %0, 0x00f000f0;" : "=r"(mhi));
    asm volatile("mov.b32 %0, 0x64006400;" : "=r"(bias));
#pragma unroll 1
    for (int gi = g0; gi < g0 + ng; gi++) {
        const uint8_t *gp = sp + (size_t)(gi - g0) * Q4T_GROUP_BYTES;
        uint4 cg[2], c8[2];
        cg[0] = *reinterpret_cast<const uint4 *>(gp + lane * 16);
        cg[1] = *reinterpret_cast<const uint4 *>(gp + 512 + lane * 16);
        c8[0] = *reinterpret_cast<const uint4 *>(gp + 1024 + lane * 16);
        c8[1] = *reinterpret_cast<const uint4 *>(gp + 1536 + lane * 16);
#pragma unroll
        for (int hb = 0; hb < 2; hb++) {
            // four blocks per accumulation: column n = 4 p + b of D holds block b times the hi (p = 0) /
            // lo (p = 1) part of x; this lane's B column is n = g, its D columns are 2t, 2t + 1
            // this lane's B column is n = g: in the mma pair of block j it is x (hi for g < 4, lo for g >= 4) if
            // j == g & 3 and zero otherwise -- the lanes of the other blocks read a 16-byte zero block instead of
            // masking four registers per block
            const uint32_t xaddr = smem_u32(xh4 + ((gi * 2 + hb) * 8 + g) * 4 + t);
            const uint2 sc = *reinterpret_cast<const uint2 *>(gp + 2048 + (g * 4 + 2 * hb + (t & 1)) * 8);
            const float2 cc = *reinterpret_cast<const float2 *>(C + gi * 8 + 4 * hb + 2 * (t & 1));
            // two accumulators (low / high nibbles): two independent mma chains of four
            float d[4] = {0.f, 0.f, 0.f, 0.f}, e[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t wg = uint4_word(cg[hb], j), w8 = uint4_word(c8[hb], j);
                const uint4 xb = lds128((g & 3) == j ? xaddr : zero16);
                const uint32_t wgs = wg >> 8, w8s = w8 >> 8;
                mma16816(d, nib_half2(wg, mlo, bias), nib_half2(w8, mlo, bias), nib_half2(wgs, mlo, bias), nib_half2(w8s, mlo, bias),
                         xb.x, xb.y);
                mma16816(e, nib_half2(wg, mhi, bias), nib_half2(w8, mhi, bias), nib_half2(wgs, mhi, bias), nib_half2(w8s, mhi, bias),
                         xb.z, xb.w);
            }
#pragma unroll
            for (int k = 0; k < 4; k++) d[k] += e[k];
            const float2 s01 = __half22float2(*reinterpret_cast<const __half2 *>(&sc.x));
            const float2 s23 = __half22float2(*reinterpret_cast<const __half2 *>(&sc.y));
            acc0 = fmaf(s01.x, d[0] - cc.x, acc0); acc0 = fmaf(s01.y, d[1] - cc.y, acc0);
            acc1 = fmaf(s23.x, d[2] - cc.x, acc1); acc1 = fmaf(s23.y, d[3] - cc.y, acc1);
        }
    }
}

// ------------------------------------------------------------------ the prologue of one phase
// The activation vector of phase q into shared memory; returns the 1 / rms factor of a norm phase (1 else).
//   QKV / W13 / classifier: rmsnorm (llama2.f90:527, :608, :627) of the residual stream plus the tp partial
//     outputs of the phase before (Wo of this layer / W2 of the previous one); layer 0 starts from the
//     embedding row (:520).  The norm weights (and the embedding row) arrive through the ring.
//   Wo / W2: the attention output (merging position splits) / the SwiGLU output.  A CTA without rows in
//     such a phase does not need its input vector: skipping the poll also keeps it from ever lagging
//     behind on a buffer nobody waits for it to have read.
template <int WT>
__device__ __forceinline__ float phase_prologue(CtaPlan *, int q)
{
    const int ph = q < 4 * cp->L ? (q & 3) : 4, l = q >> 2;
    const uint32_t ep = cp->ep_base + (uint32_t)l + 1u;
    const int rep = cp->rep, tp = cp->tp, emb = cp->emb;
    if constexpr (true) {
        if (cp->trace && threadIdx.x == 0) cp->gx_trace = (ph == 2 || ph == 3) ? cp->trace + 72 + 8 * (ph - 2) : nullptr;
    }
    if (!(ph & 1)) {
        const uint32_t sv = 1u + (uint32_t)((ph < 4 ? l : cp->L) * cp->n_layer + cp->voff[ph]);  // the norm vector's stage
        const uint8_t *emb_row = q == 0 ? vec_stage_wait(cp, 0u) : nullptr;
        const float *wn = reinterpret_cast<const float *>(vec_stage_wait(cp, sv));
        const float rs = gather_x<WT>(cp, (ph == 2 ? cp->part1[cp->rank] : cp->part2[cp->rank]) + (size_t)rep * tp * emb, tp,
                                      ph == 2 ? ep : ep - 1u, emb, 1, emb_row, wn);
        if (q == 0) vec_stage_release(cp, 0u);
        vec_stage_release(cp, sv);
        return rs;
    }
    if (cp->nrows[ph] > 0) {
        if (ph == 1 && cp->n_splits > 1) load_x_attn<WT>(cp, cp->n_splits, ep);
        else gather_x<WT>(cp, ph == 1 ? cp->ll_att + (size_t)rep * cp->att_dim : cp->ll_hb + (size_t)rep * ((cp->hid + 1) & ~1), 1, ep,
                          ph == 1 ? cp->att_dim : cp->hid, 0, nullptr, nullptr);
    }
    return 1.f;
}

// ------------------------------------------------------------------ the tiles of one phase
// This warp's tiles of phase `ph` (tile t -> warp t mod 12): the mat-vec over the tile's chunks as they
// land in the ring, the cross-lane reduction, and the tile's epilogue -- every lane ends up holding one
// row PAIR (rows 2i, 2i + 1: what RoPE and SwiGLU combine, and one 16-byte LL store), and the lanes
// that share a pair split its destinations (LL replicas x tensor-parallel ranks) between them.
// Out of line on purpose: its own register allocation, nothing live across it in the phase loop (the
// loop carries the phase number and nothing else: whatever is live across a call is spilled around it,
// and local memory is an L2 round trip in this kernel), all constants come from the plan in shared
// memory.  rscale = the 1 / rms factor of the phase's rmsnorm (the mat-vec is linear).
template <int WT, bool PROF>
__device__ __forceinline__ void run_tiles(CtaPlan *, int q, float rscale)
{
    const int ph = q < 4 * cp->L ? (q & 3) : 4, l = q >> 2, pos = cp->pos;
    const uint32_t ep = cp->ep_base + (uint32_t)l + 1u;  // epoch of everything layer l publishes
    const uint32_t s0 = 1u + (uint32_t)((ph < 4 ? l : cp->L) * cp->n_layer + cp->toff[ph]);  // the phase's first tile stage
    unsigned long long *tr = nullptr;  // PROF: this CTA's 8 trace words of the phase
    if constexpr (PROF) { if (cp->trace && ph < 4) tr = cp->trace + 16 + 8 * ph; }
#define TSTAMP(k_) do { if constexpr (PROF) { if (tr && threadIdx.x == 0 && !tdone) tr[k_] = (unsigned long long)clock64(); } } while (0)
    bool tdone = false;
    uint8_t *smem = smem_base();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const PhaseW &W = cp->ph[ph];
    const int ntiles = cp->ntiles[ph], nch = W.nch, rows_real = W.rows_real, cu = W.cu, nu_row = W.nu;
    const int r0 = cp->r0[ph], nr = cp->nrows[ph];
    const float *xs = reinterpret_cast<const float *>(smem + cp->off_xs);
    const int nrep = cp->ll_rep, tp = cp->tp, att_dim = cp->att_dim, kv = cp->kv;
    float best = -INFINITY;  // running maxloc of this lane's logits (classifier epilogue)
    int bidx = 0x7fffffff;
    if constexpr (PROF) {
        if (tr && threadIdx.x == 0) {
            tr[5] = (unsigned long long)(unsigned)cp->prod_issued | ((unsigned long long)(unsigned)cp->pf_issued << 32);
            tr[6] = s0;
        }
    }
    TSTAMP(0);
    const int G = cp->G[ph];
    const uint32_t ginv = (65535u + (uint32_t)G) / (uint32_t)G;  // ceil(65536 / G): x / G = (x * ginv) >> 16 for x < 32768
#define DIV_G(x_) ((int)(((uint32_t)(x_) * ginv) >> 16))
    const int grp = DIV_G(warp), cl = warp - grp * G, ngrp = DIV_G(NCW);
    const uint32_t gweight = (uint32_t)ngrp;  // this warp's weight in the slot's empty barrier (12 arrivals in all)
    float *gsc = reinterpret_cast<float *>(smem + cp->off_grp);  // [2][NCW][16]
    int buf = 0;
#pragma unroll 1
    for (int t = grp; t < ntiles; t += ngrp) {
        uint32_t s = s0 + (uint32_t)(t * nch);
        float va, vb;  // this lane's row pair after the reduction
        int i, sub, nsl;  // first row of the pair (within the CTA's range); destination lane index / count
        if constexpr (WT == WT_F16) {
            float dA[4] = {0.f, 0.f, 0.f, 0.f}, dB[4] = {0.f, 0.f, 0.f, 0.f};
            int u0 = 0;
            const int vrows = min(8, min(nr, rows_real - r0) - t * 8);  // rows of this tile (the schedule's `valid`)
#pragma unroll 1
            for (int c = 0; c < nch; c++, s++) {
                const int nu = min(cu, nu_row - u0);
                const int nsteps = (nu + 3) >> 2, b0 = DIV_G(cl * nsteps), b1 = DIV_G((cl + 1) * nsteps);  // this warp's k-steps
                const uint32_t slot = slot_of(cp, s);
                mbar_wait(full_bar(cp, s), full_par(s), 2);
                if (c == 0) TSTAMP(1);
                tile_dot_f16(smem + (size_t)slot * cp->slot_bytes, (uint32_t)vrows,
                             reinterpret_cast<const __half *>(xs) + 8 * u0, f16_lo_off(W.cols), nu, b0, b1, lane, dA, dB,
                             smem_u32(smem + cp->off_red) + 192u);
                __syncwarp();
                if (lane == 0) mbar_arrive_n(empty_bar(cp, slot), gweight);
                u0 += nu;
            }
            TSTAMP(2);
            // D rows 0 (lanes g = 0) and 1 (g = 1) hold the hi / lo parts of rows 2t, 2t + 1 of the tile: add them,
            // then every lane of column t takes the pair (lane g serves destination g)
            float k0 = dA[0] + dB[0], k1 = dA[1] + dB[1];
            k0 += __shfl_xor_sync(0xffffffffu, k0, 4);
            k1 += __shfl_xor_sync(0xffffffffu, k1, 4);
            va = __shfl_sync(0xffffffffu, k0, lane & 3);
            vb = __shfl_sync(0xffffffffu, k1, lane & 3);
            if (G > 1) {
                float *mine = gsc + (buf * NCW + warp) * 16;
                if (cl > 0 && lane < 4) *reinterpret_cast<float2 *>(mine + 2 * lane) = make_float2(va, vb);
                named_bar_sync(2 + grp, 32 * G);
                if (cl > 0) { buf ^= 1; continue; }
#pragma unroll 1
                for (int w = 1; w < G; w++) {
                    const float2 pp = *reinterpret_cast<const float2 *>(mine + w * 16 + 2 * (lane & 3));
                    va += pp.x; vb += pp.y;
                }
                buf ^= 1;
            }
            i = t * 8 + 2 * (lane & 3); sub = lane >> 2; nsl = 8;
        } else if constexpr (WT == WT_F32) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            int u0 = 0;
#pragma unroll 1
            for (int c = 0; c < nch; c++, s++) {
                const int nu = min(cu, nu_row - u0);
                // this warp's share of the chunk's columns: whole lane rounds [b0, b1) of 32 units
                const int rounds = (nu + 31) >> 5, b0 = DIV_G(cl * rounds), b1 = DIV_G((cl + 1) * rounds);
                const int ub = b0 << 5, un = min(nu, b1 << 5) - ub;
                const uint32_t slot = slot_of(cp, s);
                mbar_wait(full_bar(cp, s), full_par(s), 2);
                if (c == 0) TSTAMP(1);
                if (un > 0)
                    tile_dot<WT>(smem + (size_t)slot * cp->slot_bytes + (size_t)ub * 16, (uint32_t)nu * 16u,
                                 xs + (WT == WT_F16 ? 8 : 4) * (u0 + ub), un, lane, acc);
                __syncwarp();
                if (lane == 0) mbar_arrive_n(empty_bar(cp, slot), gweight);
                u0 += nu;
            }
            TSTAMP(2);
            // transposing reduction: 4 lane-partial sums -> row pair h = lane / 16 in every lane
            const bool hi16 = lane & 16, hi8 = lane & 8;
            float k0 = hi16 ? acc[2] : acc[0], k1 = hi16 ? acc[3] : acc[1];
            k0 += __shfl_xor_sync(0xffffffffu, hi16 ? acc[0] : acc[2], 16);
            k1 += __shfl_xor_sync(0xffffffffu, hi16 ? acc[1] : acc[3], 16);
            float k = hi8 ? k1 : k0;
            k += __shfl_xor_sync(0xffffffffu, hi8 ? k0 : k1, 8);
            k += __shfl_xor_sync(0xffffffffu, k, 4);
            k += __shfl_xor_sync(0xffffffffu, k, 2);
            k += __shfl_xor_sync(0xffffffffu, k, 1);
            const float o = __shfl_xor_sync(0xffffffffu, k, 8);
            va = hi8 ? o : k; vb = hi8 ? k : o;
            if (G > 1) {
                // the group's partial results meet in shared memory: warp cl > 0 leaves its two pairs, the
                // group's named barrier, warp 0 adds them (in warp order: deterministic) and goes on alone
                float *mine = gsc + (buf * NCW + warp) * 16;
                if (cl > 0 && (lane & 15) == 0) *reinterpret_cast<float2 *>(mine + (lane >> 3)) = make_float2(va, vb);
                named_bar_sync(2 + grp, 32 * G);
                if (cl > 0) { buf ^= 1; continue; }
#pragma unroll 1
                for (int w = 1; w < G; w++) {
                    const float2 pp = *reinterpret_cast<const float2 *>(mine + w * 16 + ((lane >> 4) << 1));
                    va += pp.x; vb += pp.y;
                }
                buf ^= 1;
            }
            i = t * 4 + 2 * (lane >> 4); sub = lane & 15; nsl = 16;
        } else {
            float acc0 = 0.f, acc1 = 0.f;
            int g0 = 0;
#pragma unroll 1
            for (int c = 0; c < nch; c++, s++) {
                const int ng = min(cu, nu_row - g0);
                const int b0 = DIV_G(cl * ng), b1 = DIV_G((cl + 1) * ng);  // this warp's 8-block groups of the chunk
                const uint32_t slot = slot_of(cp, s);
                mbar_wait(full_bar(cp, s), full_par(s), 2);
                if (c == 0) TSTAMP(1);
                tile_dot_q4(smem + (size_t)slot * cp->slot_bytes + (size_t)b0 * Q4T_GROUP_BYTES, xs, nu_row, g0 + b0, b1 - b0, lane, acc0, acc1,
                            smem_u32(smem + cp->off_red) + 192u);
                __syncwarp();
                if (lane == 0) mbar_arrive_n(empty_bar(cp, slot), gweight);
                g0 += ng;
            }
            TSTAMP(2);
            // sum over t (blocks 0,1 | 2,3 and the hi | lo parts): rows g, g + 8 in all four t lanes
            acc0 += __shfl_xor_sync(0xffffffffu, acc0, 1); acc0 += __shfl_xor_sync(0xffffffffu, acc0, 2);
            acc1 += __shfl_xor_sync(0xffffffffu, acc1, 1); acc1 += __shfl_xor_sync(0xffffffffu, acc1, 2);
            const int g = lane >> 2;
            if (G > 1) {
                float *mine = gsc + (buf * NCW + warp) * 16;
                if (cl > 0 && (lane & 3) == 0) { mine[g] = acc0; mine[g + 8] = acc1; }
                named_bar_sync(2 + grp, 32 * G);
                if (cl > 0) { buf ^= 1; continue; }
#pragma unroll 1
                for (int w = 1; w < G; w++) { acc0 += mine[w * 16 + g]; acc1 += mine[w * 16 + g + 8]; }
                buf ^= 1;
            }
            // the even-g lanes take the pair (g, g + 1), the odd-g lanes the pair (g + 7, g + 8)
            const float p0 = __shfl_xor_sync(0xffffffffu, acc0, 4), p1 = __shfl_xor_sync(0xffffffffu, acc1, 4);
            const bool odd = g & 1;
            va = odd ? p1 : acc0; vb = odd ? acc1 : p0;
            i = t * 16 + (odd ? g + 7 : g); sub = lane & 3; nsl = 4;
        }

        // ---- epilogue of the row pair (i, i + 1): `nsl` lanes hold it, lane `sub` serves the
        // destinations sub, sub + nsl, ... (LL replicas x ranks)
        const int r = r0 + i;
        TSTAMP(3);
        if (i >= nr || r >= rows_real) continue;  // rows past the CTA's range / padding rows of the matrix
        const bool two = r + 1 < rows_real;
        const float a = va * rscale, b = vb * rscale;
        if (ph == 0) {
            // RoPE on q and k (reference quirks Q1/Q2 are in the table), KV append (llama2.f90:543-565)
            const bool isq = r < att_dim, isv = r >= att_dim + kv;
            const int rk = isq ? r : (isv ? r - att_dim - kv : r - att_dim);
            const float2 cs2 = isv ? make_float2(1.f, 0.f) : cp->rope[(rk >> 1) & ((cp->hs >> 1) - 1)];
            const float o0 = a * cs2.x - b * cs2.y, o1 = a * cs2.y + b * cs2.x;
#pragma unroll 1
            for (int d = sub; d < nrep; d += nsl) {
                unsigned long long *dst = isq ? cp->ll_q + (size_t)d * att_dim
                                              : cp->ll_kv + (size_t)d * 2 * kv + (isv ? kv : 0);
                ll_store2(dst, rk, o0, o1, ep);
            }
            if (!isq && sub == nsl - 1) {
                // the cache row serves later launches, the LL copy this launch's attention
                float *cache = (isv ? cp->vc : cp->kc) + ((size_t)l * cp->seq + (pos - 1)) * kv;
                *reinterpret_cast<float2 *>(cache + rk) = make_float2(o0, o1);
            }
        } else if (ph == 2) {
            // SwiGLU on the interleaved gate/up rows (llama2.f90:613-616)
            const float hv = (a * __fdividef(1.0f, 1.0f + __expf(-a))) * b;
            const int hb_stride = (cp->hid + 1) & ~1;
#pragma unroll 1
            for (int d = sub; d < nrep; d += nsl) ll_store(cp->ll_hb + (size_t)d * hb_stride, r >> 1, hv, ep);
        } else if (ph == 4) {
            // logits rows of this rank go to every rank's full logits buffer (all-gather)
            const int gi = cp->v_off + r;
#pragma unroll 1
            for (int k = sub; k < tp; k += nsl) {
                cp->logits[k][gi] = a;
                if (two) cp->logits[k][gi + 1] = b;
            }
            if (sub == 0) {
                if (a > best) { best = a; bidx = gi; }
                if (two && b > best) { best = b; bidx = gi + 1; }
            }
        } else {
            // Wo / W2 (llama2.f90:603-605, :618-620): publish this rank's partial sums to every
            // rank; the residual add happens in the next norm prologue, on every CTA's copy of x
            const int rep_shift = 31 - __clz(nrep);
#pragma unroll 1
            for (int d = sub; d < nrep * tp; d += nsl) {
                const int k = d >> rep_shift, rr = d & (nrep - 1);  // destination rank, replica
                unsigned long long *dst = (ph == 1 ? cp->part1[k] : cp->part2[k]) + ((size_t)rr * tp + cp->rank) * cp->emb;
                if (two) ll_store2_sys(dst, r, a, b, ep);
                else ll_store_sys(dst, r, a, ep);
            }
        }
        TSTAMP(4);
        tdone = true;
    }
    tdone = false;
    TSTAMP(7);
    if constexpr (PROF) {
        if (tr && threadIdx.x == 0)
            tr[3] = (tr[3] & 0xffffffffull) | ((unsigned long long)(unsigned)(cp->pf_issued - (int)s0) << 48) |
                    ((unsigned long long)(unsigned)(cp->prod_issued - (int)s0) << 32);  // cursors when warp 0 is done
    }
#undef TSTAMP
#undef DIV_G
    if (ph == 4) {
        // maxloc of this warp's logits (llama2.f90:388: the first maximum wins), for the token tail
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
            if (ob > best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
        }
        if (lane == 0) { cp->tail_best[warp] = best; cp->tail_idx[warp] = bidx; }
    }
}

// ------------------------------------------------------------------ the consumer warps' phase loop
// One loop over the 4 L + 1 weight phases (q = 4 l + {0 QKV, 1 WO, 2 W13, 3 W2}; q = 4 L is the
// classifier): prologue -> tiles (mat-vec + epilogue per group) (-> attention).  A single copy of every
// piece serves all phases.  The loop is a function of its own and carries nothing but q: the pieces are
// allocated the registers their caller does not hold, and what they cannot keep in registers they spill to
// local memory -- which is an L2 round trip in this kernel (shared memory leaves the L1 next to nothing).
// PROF = true adds the phase timers of CTA 0 (SM cycles per bucket, PH_* in kernels.cuh) and the optional
// per-CTA trace of one layer.
template <int WT, bool PROF>
__device__ __noinline__ void consumer_main(CtaPlan *, Prof *)
{
    const bool timer = PROF && (blockIdx.x == 0 && threadIdx.x == 0);
    const bool tracing = PROF && cp->trace_base != nullptr;
    const unsigned long long t_ns0 = timer ? globaltimer_ns() : 0ull;
    const long long t_c0 = timer ? clock64() : 0ll;
    if (timer) {
        for (int i = 0; i < PH_COUNT; i++) pf->tacc[i] = 0;
        pf->tmark = t_c0;
    }
#define LAP(b) do { if constexpr (PROF) { if (timer) prof_lap(pf, (b)); } } while (0)
#define STAMP(l_, k_) do { if constexpr (PROF) { if (tracing && threadIdx.x == 0 && (l_) == cp->trace_layer) prof_stamp(cp->trace_base, (k_)); } } while (0)
    // The phase number lives in shared memory (one word per warp) and is re-read after every call: a value
    // kept in a register across a call has to be saved and restored around it by somebody, through local
    // memory, and local memory is an L2 round trip here.
#define Q_NOW (cp->qw[threadIdx.x >> 5])
    if ((threadIdx.x & 31) == 0) Q_NOW = 0;
    __syncwarp();
#pragma unroll 1
    for (;;) {
        {
            const int q = Q_NOW, nq = 4 * cp->L + 1;
            if (q >= nq) break;
            if constexpr (PROF) {
                const int ph = q < nq - 1 ? (q & 3) : 4, l = q >> 2;
                if (threadIdx.x == 0 && (q & 3) == 0) cp->trace = (tracing && l == cp->trace_layer && ph < 4) ? cp->trace_base + (size_t)blockIdx.x * 128 : nullptr;
                if ((q & 3) == 0) cons_sync();
                if (ph == 0) STAMP(l, 0);
            }
        }
        const float rscale = phase_prologue<WT>(cp, Q_NOW);
        if constexpr (PROF) {
            const int q = Q_NOW, ph = q < 4 * cp->L ? (q & 3) : 4;
            LAP(ph == 0 ? 0 : 2 + 3 * ph);
            STAMP(q >> 2, ph == 0 ? 1 : 3 + 3 * ph);
        }
        run_tiles<WT, PROF>(cp, Q_NOW, rscale);
        if constexpr (PROF) {
            const int q = Q_NOW, ph = q < 4 * cp->L ? (q & 3) : 4;
            LAP((ph == 0 ? 0 : 2 + 3 * ph) + 1);
            STAMP(q >> 2, ph == 0 ? 2 : 4 + 3 * ph);
        }
        if ((Q_NOW & 3) == 0 && Q_NOW < 4 * cp->L) {
            // ---- attention (llama2.f90:574-598)
            if (cp->hs == 64) attention_phase_t<64, PROF>(cp, Q_NOW >> 2);
            else if (cp->hs == 128) attention_phase_t<128, PROF>(cp, Q_NOW >> 2);
            else attention_phase_t<32, PROF>(cp, Q_NOW >> 2);
            LAP(PH_ATT);
            STAMP(Q_NOW >> 2, 4);
        }
        __syncwarp();
        if ((threadIdx.x & 31) == 0) Q_NOW = Q_NOW + 1;
        __syncwarp();
    }
#undef Q_NOW
    token_tail(cp, cp->pos);
    if (timer) {
        prof_lap(pf, PH_ARGMAX);
        for (int i = 0; i < PH_COUNT; i++) cp->phase_cycles[i] += (unsigned long long)pf->tacc[i];
        cp->phase_cycles[PH_COUNT] += (unsigned long long)(clock64() - t_c0);
        cp->phase_cycles[PH_COUNT + 1] += globaltimer_ns() - t_ns0;
    }
#undef LAP
#undef STAMP
}

// ------------------------------------------------------------------ the kernel
// PROF = false is the production kernel; PROF = true adds the phase timers of CTA 0 and the optional
// per-CTA phase-edge trace.
template <int WT, bool PROF>
__global__ void __launch_bounds__(416, 1)
stream_decode_kernel(const __grid_constant__ StreamParams P)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- shared-memory map: ring | xs | xres | red | attention scratch | full[NBAR] | empty[MAX_SLOTS] | stage list
    const int off_xs = P.n_slots * P.slot_bytes, off_xres = off_xs + P.xs_floats * 4, off_red = off_xres + P.emb * 4;
    const int off_att = off_red + 64 * 4, off_grp = off_att + (ATT_MAX_CHUNK + NCW * P.hs) * 4;
    const int off_full = off_grp + 2 * NCW * 16 * 4;  // [2][NCW][16] partial results of a tile group's warps
    const int off_empty = off_full + NBAR * 8, off_sched = off_empty + MAX_SLOTS * 8;
    if (threadIdx.x == 5) {
        g_cp.off_xs = off_xs; g_cp.off_xres = off_xres; g_cp.off_red = off_red; g_cp.off_att = off_att; g_cp.off_grp = off_grp;
        for (int i = 0; i < 5; i++) g_cp.G[i] = P.tile_warps[i];
        g_cp.off_full = off_full; g_cp.off_empty = off_empty; g_cp.off_sched = off_sched;
        g_cp.slot_bytes = P.slot_bytes; g_cp.n_slots = P.n_slots;
        g_cp.slot_magic = 0xffffffffu / (uint32_t)P.n_slots + 1u;
        g_cp.wtype = P.wtype; g_cp.emb = P.emb; g_cp.hid = P.hid; g_cp.kv = P.kv; g_cp.att_dim = P.att_dim; g_cp.hs = P.hs;
        g_cp.tp = P.tp; g_cp.rank = P.rank; g_cp.ll_rep = P.ll_rep; g_cp.v_off = P.v_off; g_cp.seq = P.seq; g_cp.H = P.H;
        g_cp.kv_mul = P.kv_mul; g_cp.L = P.L;
        g_cp.rep = (int)(blockIdx.x % (unsigned)P.ll_rep);
        g_cp.n_splits = P.n_splits; g_cp.ep_base = P.ep_base; g_cp.trace = nullptr;
        g_cp.trace_base = P.trace; g_cp.trace_layer = P.trace_layer; g_cp.phase_cycles = P.phase_cycles;
        g_cp.kvmul_inv16 = (65535u + (unsigned)P.kv_mul) / (unsigned)P.kv_mul;
        g_cp.inv_emb = 1.0f / (float)P.emb;
        g_cp.kc = P.kc; g_cp.vc = P.vc; g_cp.ll_q = P.ll_q; g_cp.ll_kv = P.ll_kv; g_cp.ll_att = P.ll_att;
        g_cp.ll_part = P.ll_part; g_cp.ll_hb = P.ll_hb;
    }
    if (threadIdx.x >= 8 && threadIdx.x < 8 + MAX_TP) {
        const int k = threadIdx.x - 8;
        g_cp.part1[k] = P.part1[k]; g_cp.part2[k] = P.part2[k]; g_cp.logits[k] = P.logits[k];
        g_cp.amax[k] = P.amax[k]; g_cp.done[k] = P.done[k];
    }
    if (threadIdx.x == 7) {
        g_cp.gx_trace = nullptr; g_cp.abort = 0; g_cp.err_flag = P.err_flag;
        g_cp.forced = P.forced; g_cp.out_tokens = P.out_tokens; g_cp.tokpos = const_cast<int *>(P.tokpos);
        g_cp.ep_last = P.ep_base + (uint32_t)P.L + 1u; g_cp.do_argmax = P.do_argmax;
    }
    if (threadIdx.x < 5) {
        const int i = threadIdx.x;
        g_cp.ph[i] = P.ph[i];
        int r0, r1;
        cta_rows(P.ph[i], blockIdx.x, gridDim.x, r0, r1);
        g_cp.r0[i] = r0; g_cp.nrows[i] = r1 - r0;
        const int nt = (r1 - r0 + P.ph[i].R - 1) / P.ph[i].R;
        g_cp.ntiles[i] = nt; g_cp.nst[i] = nt * P.ph[i].nch;
        g_cp.lstride[i] = i < 4 ? P.ph[i].layer_stride : 0ull;
        g_cp.sstride[i] = P.ph[i].rs;
    }
    if (threadIdx.x == 6) {
        g_cp.lstride[SK_RMS_ATT] = g_cp.lstride[SK_RMS_FFN] = (unsigned long long)P.emb * 4u;
        g_cp.lstride[SK_RMS_FINAL] = g_cp.lstride[SK_EMB_ROW] = 0ull;
        g_cp.sstride[SK_RMS_ATT] = g_cp.sstride[SK_RMS_FFN] = g_cp.sstride[SK_RMS_FINAL] = g_cp.sstride[SK_EMB_ROW] = 0u;
    }
    {
        const uint4 *g = reinterpret_cast<const uint4 *>(P.sched + (size_t)blockIdx.x * P.sched_stride);
        uint4 *d = reinterpret_cast<uint4 *>(smem + off_sched);
        for (int i = threadIdx.x; i < P.sched_stride; i += blockDim.x) d[i] = __ldg(g + i);
    }
    if (threadIdx.x < NBAR + P.n_slots) {
        uint64_t *bars = reinterpret_cast<uint64_t *>(smem + off_full);
        if (threadIdx.x < NBAR) mbar_init(&bars[threadIdx.x], 1);
        else mbar_init(&bars[threadIdx.x], (uint32_t)NCW);  // empty[slot]: a group of G warps arrives with weight 12 / G each
        fence_mbar_init();
    }
    const int token = P.token > 0 ? P.token : P.tokpos[0];
    const int pos = P.token > 0 ? P.pos : P.tokpos[1];
    if (threadIdx.x == 6) {
        g_cp.pos = pos;
        int nst[5];
        for (int i = 0; i < 5; i++) {
            int r0, r1;
            cta_rows(P.ph[i], blockIdx.x, gridDim.x, r0, r1);
            nst[i] = ((r1 - r0 + P.ph[i].R - 1) / P.ph[i].R) * P.ph[i].nch;
        }
        g_cp.n_layer = 2 + nst[0] + nst[1] + nst[2] + nst[3];
        g_cp.voff[0] = 0; g_cp.toff[0] = 1; g_cp.voff[1] = 0; g_cp.toff[1] = 1 + nst[0];
        g_cp.voff[2] = 1 + nst[0] + nst[1]; g_cp.toff[2] = g_cp.voff[2] + 1; g_cp.voff[3] = 0; g_cp.toff[3] = g_cp.toff[2] + nst[2];
        g_cp.voff[4] = 0; g_cp.toff[4] = 1;
    }
    if (threadIdx.x >= 32 && threadIdx.x < 36)  // 16 zero bytes (the zero rows of tile_dot_f16's A operand)
        reinterpret_cast<float *>(smem_base() + off_red)[48 + threadIdx.x - 32] = 0.f;
    // this position's RoPE row (a cold HBM read, issued first thing, used after the QKV phase)
    if (threadIdx.x >= 64 && threadIdx.x < 64 + (P.hs >> 1))
        g_cp.rope[threadIdx.x - 64] = P.rope_tab[(size_t)(pos - 1) * (P.hs >> 1) + threadIdx.x - 64];
    __syncthreads();

    if (warp == NCW) {
        // ===================== producer warp =====================
        if (lane == 0) producer_loop(cp, token, P.pace, P.pf_lead);
        return;
    }

    // ===================== consumer warps =====================
    consumer_main<WT, PROF>(cp, pf);
}

// ------------------------------------------------------------------ host side
// stages a CTA can have in the layer section / after it (upper bounds over all CTAs)
static int tiles_of(const PhaseW &w, int nrows) { return (nrows + w.R - 1) / w.R; }
static void sched_caps(const StreamParams &p, int *layer, int *post)
{
    int per_layer = 2, cls = 0;
    for (int i = 0; i < 5; i++) {
        const int st = tiles_of(p.ph[i], p.ph[i].rows_cap) * p.ph[i].nch;
        if (i < 4) per_layer += st; else cls = st;
    }
    *layer = per_layer;
    *post = 1 + cls;
}

int plan_stream(StreamParams &p, int grid, int max_smem_optin, int target_slot_bytes, int max_slots, StreamPlan *out)
{
    const bool tiled = p.wtype == WT_Q4_0;  // q4_0 in the tiled mma format
    // a slot holds one stage: a chunk of a tile, a norm vector, or an embedding row
    unsigned int slot = (unsigned)p.emb * 4u;
    if ((unsigned)row_stride_bytes(p.wtype, p.emb) > slot) slot = (unsigned)row_stride_bytes(p.wtype, p.emb);
    const unsigned min_chunk = tiled ? Q4T_GROUP_BYTES : (p.wtype == WT_F16 ? 8u * 64u : 4u * 32u * 16u);  // one group / one k-step of 8 rows / one unit per lane of 4 rows
    if (slot < min_chunk) slot = min_chunk;
    if ((unsigned)target_slot_bytes > slot) slot = (unsigned)target_slot_bytes;
    slot = (slot + 127u) & ~127u;
    for (int i = 0; i < 5; i++) {
        PhaseW &w = p.ph[i];
        const int U = w.rows / w.unit;
        w.rows_cap = ((U + grid - 1) / grid) * w.unit;
        if (tiled) {
            w.R = 16;
            w.nu = q4t_groups(w.cols);
            const int cu_max = (int)(slot / Q4T_GROUP_BYTES);
            w.nch = (w.nu + cu_max - 1) / cu_max;
            w.cu = (w.nu + w.nch - 1) / w.nch;
        } else if (p.wtype == WT_F16) {
            // f16 on the tensor cores (tile_dot_f16): 8 rows = the n dimension of one mma; chunks of whole
            // k-steps (4 units = 32 columns)
            w.R = 8;
            w.nu = (int)(row_stride_bytes(p.wtype, w.cols) / 16);
            const int cu_max = (int)(slot / (8u * 16u)) & ~3;
            w.nch = (w.nu + cu_max - 1) / cu_max;
            w.cu = (((w.nu + w.nch - 1) / w.nch) + 3) & ~3;
        } else {
            w.R = 4;
            w.nu = (int)(row_stride_bytes(p.wtype, w.cols) / 16);
            const int cu_max = (int)(slot / (4u * 16u)) & ~31;  // whole lane rounds
            w.nch = (w.nu + cu_max - 1) / cu_max;
            w.cu = (((w.nu + w.nch - 1) / w.nch) + 31) & ~31;
        }
        if (w.nch > MAX_NCH) return 1;
    }
    int xs_floats = p.emb > p.hid ? p.emb : p.hid;
    if (p.wtype == WT_F16) xs_floats += 64;  // (the lo plane starts up to 192 bytes past the hi plane: f16_lo_off)
    if (tiled)  // f16 hi + lo activations in fragment order + two corrections per block (store_x4)
        for (int i = 0; i < 5; i++) xs_floats = xs_floats > p.ph[i].nu * 272 ? xs_floats : p.ph[i].nu * 272;
    xs_floats = (xs_floats + 31) & ~31;
    if (max_slots > MAX_SLOTS) max_slots = MAX_SLOTS;
    int cap_layer, cap_post;
    sched_caps(p, &cap_layer, &cap_post);
    const int sched_entries = 1 + cap_layer + cap_post + 1;
    int n_slots = max_slots;
    while (n_slots > 0 && smem_bytes_for(n_slots, (int)slot, xs_floats, p.emb, p.hs, sched_entries) > (size_t)max_smem_optin)
        n_slots--;
    if (n_slots < 3) return 1;
    // Shared memory and L1 share 256 KB per SM, and the carve-out comes in steps (.. 164, 196, 228 KB).  The kernel
    // is capped at 128 registers (13 warps: four on one scheduler partition) and the compiler parks a few dozen
    // words per thread in local memory; with the 228 KB carve-out the ~28 KB of L1 left do not hold them and
    // every reload is an L2 round trip.  Giving up one ring slot to fall under the 196 KB step (60 KB of L1)
    // pays whenever enough ring is left -- measured, ms per token with n / n - 1 slots: TinyLlama f32 0.904 /
    // 0.889, f16 0.765 / 0.695, Llama-2-7B q4_0 1.871 / 1.768, but f16 (4 x 32 KB -> 3) 3.40 / 3.53.
    if (max_slots >= MAX_SLOTS) {  // (an explicit LLMF90_MAX_SLOTS is taken as given)
        const size_t step = 196 * 1024 - 1024 - 3072;  // the step, minus the per-CTA reservation and the static part
        const size_t min_ring = tiled ? 100 * 1024 : 120 * 1024;
        if (smem_bytes_for(n_slots, (int)slot, xs_floats, p.emb, p.hs, sched_entries) > step &&
            smem_bytes_for(n_slots - 1, (int)slot, xs_floats, p.emb, p.hs, sched_entries) <= step &&
            (size_t)(n_slots - 1) * slot >= min_ring)
            n_slots--;
    }
    out->n_slots = n_slots;
    out->slot_bytes = (int)slot;
    out->threads = (NCW + 1) * 32;
    out->smem_bytes = (int)smem_bytes_for(n_slots, (int)slot, xs_floats, p.emb, p.hs, sched_entries);
    out->grid = grid;
    out->xs_floats = xs_floats;
    // Warps per tile group, per phase.  Measured on B200 with one value for all phases: 3 (four groups) is
    // the best or within 1 % of it for every benchmark configuration (TinyLlama f32 1.006 / 0.949 / 0.930 /
    // 0.951 ms per token with 1 / 2 / 3 / 4 warps, Llama-2-7B f16 4.29 / 3.34 / 3.19 / 3.20, q4_0 3.02 / 2.30 /
    // 2.24 / 2.15): a stage is drained in a third of the time, so a slot spends its life in flight, not
    // waiting for its warp.  A short phase (a handful of tiles per CTA, all banked in the ring before it
    // starts) is a latency chain instead: its time is rounds x (fixed part + mat-vec / G), so prefer the
    // group size that takes the fewest rounds.
    for (int i = 0; i < 5; i++) {
        // (f16 / q4_0 run on the tensor pipe: few instructions per byte, so a tile's warps are latency-bound and more
        // of them per tile pay -- Llama-2-7B f16 3.40 / 3.09 / 2.99 ms per token with 3 / 4 / 6, q4_0 1.93 / 1.77 / 1.75)
        const int tiles = tiles_of(p.ph[i], p.ph[i].rows_cap);
        const int big = tiled ? (p.emb >= 4096 ? 6 : 4) : (p.wtype == WT_F16 ? (p.ph[i].nch >= 2 ? 6 : 4) : 3);
        int best = big;
        if (tiles <= NCW && tiles * p.ph[i].nch <= n_slots) {  // the whole phase of a CTA fits the ring
            long best_cost = -1;
            const int cand[4] = {4, 3, 2, 1};
            for (int k = 0; k < 4; k++) {
                const int g = cand[k], rounds = (tiles + NCW / g - 1) / (NCW / g);
                const long cost = (long)rounds * (1500 + 4500 / g);
                if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = g; }
            }
        }
        out->tile_warps[i] = best;
    }
    return 0;
}

void build_schedule(StreamParams &p, int grid, SchedStage **out)
{
    // section sizes are per CTA (they follow from its row ranges: the kernel derives them from the same
    // cta_rows()); only the padded stride is shared
    int cap_layer, cap_post;
    sched_caps(p, &cap_layer, &cap_post);
    const int cap = 1 + cap_layer + cap_post + 1;
    const bool tiled = p.wtype == WT_Q4_0;
    SchedStage *tab = (SchedStage *)calloc((size_t)grid * cap, sizeof(SchedStage));
    for (int cta = 0; cta < grid; cta++) {
        SchedStage *t = tab + (size_t)cta * cap;
        int n = 0;
        bool start = true;  // the next stage pushed is the first of a phase (its vector stage, if it has one)
        auto push = [&](const void *ptr, unsigned seg_bytes, unsigned nseg, unsigned kind) {
            t[n].src = (unsigned long long)ptr; t[n].seg_bytes = seg_bytes;
            t[n].meta = nseg | (kind << 8) | (start ? SCHED_PHASE_START : 0u);
            start = false;
            n++;
        };
        auto rows = [&](int ph) {
            int r0, r1;
            cta_rows(p.ph[ph], cta, grid, r0, r1);
            const PhaseW &w = p.ph[ph];
            if (ph == 1 || ph == 3) start = true;  // Wo / W2 have no vector stage: the phase starts with its rows
            for (int r = r0; r < r1; r += w.R)
                for (int c = 0; c < w.nch; c++) {
                    const int u0 = c * w.cu, nu = (w.nu - u0) < w.cu ? (w.nu - u0) : w.cu;
                    if (tiled) {
                        // one run of whole 8-block groups of the row group
                        push(w.base + ((size_t)(r >> 4) * w.nu + u0) * Q4T_GROUP_BYTES, (unsigned)nu * Q4T_GROUP_BYTES, 1u, (unsigned)ph);
                    } else {
                        // tile-major matrix (tile_pass_kernel): the chunk's segments of the tile's valid rows are
                        // one contiguous run -- ONE bulk copy per stage
                        int valid = (r1 < w.rows_real ? r1 : w.rows_real) - r;
                        if (valid > w.R) valid = w.R;
                        push(w.base + ((size_t)r * w.nu + (size_t)valid * u0) * 16, (unsigned)(valid * nu) * 16u, 1u, (unsigned)ph);
                    }
                }
        };
        push(p.emb_table, (unsigned)row_stride_bytes(p.wtype, p.emb), 1u, SK_EMB_ROW);  // row 0; the kernel adds (token - 1) rows
        start = true;
        push(p.rms_att, (unsigned)p.emb * 4u, 1u, SK_RMS_ATT);
        rows(0);
        rows(1);
        start = true;
        push(p.rms_ffn, (unsigned)p.emb * 4u, 1u, SK_RMS_FFN);
        rows(2);
        rows(3);
        start = true;
        push(p.rms_final, (unsigned)p.emb * 4u, 1u, SK_RMS_FINAL);
        rows(4);
    }
    p.sched_stride = cap;
    *out = tab;
}

// ------------------------------------------------------------------ tile-major weights
// A bulk copy costs its issuing thread ~100 SM cycles whatever its size (measured: tools/ubench/bulk_issue.cu), and an L2
// prefetch as much again: with one copy per ROW of a stage the producer thread -- not HBM, not the consumers --
// set the pace of the kernel (8 rows per f16 stage: ~2200 cycles per stage against ~1400 cycles of HBM time for
// its 32 KB).  So the streamed matrices are re-laid out once at init in the order the kernel consumes them: for
// every CTA's row range, tile after tile, chunk after chunk, the chunk's segment of every valid row -- f32 row
// after row, f16 k-step after k-step (tile_dot_f16) -- and a stage is ONE contiguous run.  The permutation stays
// inside a tile's rows x row_bytes block, so a CTA's data still starts at row r0 and no byte is added.
// src: one layer of the matrix, plain rows; dst: the same bytes, tile-major.  grid (CTAs of the decode kernel, y).
__global__ void tile_pass_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ dst, PhaseW w, int grid, int step_major)
{
    int r0, r1;
    cta_rows(w, blockIdx.x, grid, r0, r1);
    const int nr = min(r1, w.rows_real) - r0;
    const long long total = (long long)nr * w.nu;
    for (long long idx = (long long)blockIdx.y * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.y * blockDim.x) {
        const int jr = (int)(idx / w.nu), u = (int)(idx - (long long)jr * w.nu);
        const int t = jr / w.R, j = jr - t * w.R, v = min(w.R, nr - t * w.R);
        const int c = u / w.cu, u0 = c * w.cu, nuc = min(w.cu, w.nu - u0), uu = u - u0;
        const size_t base = (size_t)(r0 + t * w.R) * w.nu + (size_t)v * u0;
        size_t off;
        if (!step_major) {
            off = base + (size_t)j * nuc + uu;
        } else {
            const int st = uu >> 2, nfull = nuc >> 2;
            off = st < nfull ? base + (size_t)(st * v + j) * 4 + (uu & 3) : base + (size_t)nfull * v * 4 + (size_t)j * (nuc & 3) + (uu & 3);
        }
        dst[off] = src[(size_t)(r0 + jr) * w.nu + u];
    }
}

cudaError_t launch_tile_pass(const StreamParams &p, int phase, int grid, const uint8_t *src, uint8_t *dst, cudaStream_t st)
{
    if (p.wtype == WT_Q4_0) return cudaErrorInvalidValue;  // q4_0 has its own tiled format (repack_q4_tiled_kernel)
    tile_pass_kernel<<<dim3(grid, 16), 256, 0, st>>>(reinterpret_cast<const uint4 *>(src), reinterpret_cast<uint4 *>(dst), p.ph[phase], grid,
                                                    p.wtype == WT_F16 ? 1 : 0);
    return cudaGetLastError();
}

static const void *kernel_for(int wtype, bool prof)
{
    if (wtype == WT_F32) return prof ? (const void *)stream_decode_kernel<WT_F32, true> : (const void *)stream_decode_kernel<WT_F32, false>;
    if (wtype == WT_F16) return prof ? (const void *)stream_decode_kernel<WT_F16, true> : (const void *)stream_decode_kernel<WT_F16, false>;
    if (wtype == WT_Q4_0) return prof ? (const void *)stream_decode_kernel<WT_Q4_0, true> : (const void *)stream_decode_kernel<WT_Q4_0, false>;
    return nullptr;
}

cudaError_t prepare_stream_kernel(int wtype, int threads, int smem_bytes)
{
    (void)threads;
    for (int prof = 0; prof < 2; prof++) {
        const void *fn = kernel_for(wtype, prof != 0);
        if (!fn) return cudaErrorInvalidValue;
        cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t launch_stream(const StreamParams &p, const StreamPlan &plan, bool prof, cudaStream_t st)
{
    StreamParams q = p;
    void *args[] = {(void *)&q};
    const void *fn = kernel_for(p.wtype, prof);
    if (!fn) return cudaErrorInvalidValue;
    // cooperative launch: guarantees all CTAs are co-resident (the LL hand-over polls across CTAs)
    return cudaLaunchCooperativeKernel(fn, dim3(plan.grid), dim3(plan.threads), args,
                                       (size_t)plan.smem_bytes, st);
}

}  // namespace llmf90

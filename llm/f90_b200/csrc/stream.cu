// stream.cu -- the fused weight-streaming decode kernel (one launch = one token).
//
// Replaces `function transformer(token,pos,s,w)` (llama2.f90:480-640) on one B200.
//
// Design: decode at batch 1 reads every weight byte exactly once per
// token and does ~0.5 flop per byte, so the only thing that matters is keeping HBM busy
// across the ~110 dependent mat-vec phases of a token.  One persistent cooperative CTA per SM:
//
//   * a PRODUCER warp walks the CTA's private, fully static schedule -- the token's embedding
//     row, then for every layer the rms_att vector, its contiguous row range of Wqkv and Wo, the
//     rms_ffn vector, its rows of W13 (gate/up rows interleaved at upload) and W2, finally the
//     rms_final vector and its rows of Wcls -- and streams it with 1-D TMA bulk copies
//     (cp.async.bulk, mbarrier complete_tx) into a ring of shared-memory slots.  Nothing in the
//     schedule depends on activations, so the producer never waits for a hand-over: while the
//     consumers finish a phase and rebuild the activation vector, the ring keeps filling.
//   * 12 CONSUMER warps in NG groups of GW; a group owns every NG-th stage of the ring: wait on
//     its `full` mbarrier, dot the rows in the slot with the activation vector (kept in
//     registers for the phase; f16 / q4_0 dequantisation fused into the load, f32 accumulation,
//     batched warp-shuffle reduction), release the slot with an `empty` mbarrier arrive.
//   * Between phases the consumers run the tiny epilogues in place -- RoPE + KV-cache append,
//     SwiGLU, residual add -- and publish their slice as {value, epoch} 64-bit words ("LL"
//     buffers, the protocol NCCL uses for small messages): the next phase's prologue polls the
//     whole vector straight out of L2 until every word carries the expected epoch.  There is NO
//     grid barrier and no fence anywhere in the token: a phase hand-over costs one store->load
//     trip through L2 (rmsnorm is recomputed redundantly per CTA).
//   * Attention (scores, softmax, value gather) is a phase of the same kernel: (head, split)
//     items over the CTAs, online softmax, merged in the Wo prologue when there are splits.
//
// CODE SIZE IS A FIRST-CLASS CONSTRAINT.  Every piece of this kernel runs once per phase, i.e. its
// instructions are cold each time unless one layer's worth of code fits the 32 KB L1.5
// instruction cache; a cold 128-byte line (8 instructions) costs ~300 cycles (measured: a
// 228 KB build of this kernel spent most of every hand-over fetching instructions).  Hence: one
// generic prologue / consume / epilogue for all phases (data-driven, not specialised), modest
// unrolling, profiling hooks out of line, cold paths (position splits, tensor-parallel tails)
// in non-inlined functions.
#include <cooperative_groups.h>
#include <cstdlib>

#include "kernels.cuh"

namespace llmf90 {

constexpr int MAX_SLOTS = 16;
constexpr int MAX_CONS_WARPS = 12;  // + 1 producer warp = 416 threads -> 128 registers/thread
constexpr int CONS_BAR = 1;         // named barrier id used by the consumer warps
constexpr int GW = 4;               // consumer warps per group: a group of GW warps consumes one ring stage
constexpr int NG = MAX_CONS_WARPS / GW;  // groups; group g owns the stages whose schedule index is g (mod NG)
// Successive uses of a slot may belong to different groups, and a group may start waiting for use
// k + 1 of a slot before use k has landed: with ONE full barrier per slot the one-bit mbarrier phase
// parity would alias (a wait on the parity of use k + 1 returns at once while use k is in flight).
// Every slot therefore has TWO full barriers, for its even and its odd uses: a waiter for use k + 1
// can only be confused with use k - 1, which was consumed before use k could even be issued.
// Use k of a slot: barrier (k & 1), parity ((k >> 1) & 1).  (The empty barriers have one waiter,
// the producer, that sees the uses of a slot in order.)

struct SmemView {
    uint8_t *ring;
    float *xs, *res, *xres, *red;  // xres: this CTA's copy of the residual stream x (llama2.f90:520,605,620)
    uint64_t *full, *empty;
    const SchedStage *sched;  // this CTA's stage list (copied from global memory at kernel start)
};

// per-launch constants of this CTA, computed once into shared memory
struct CtaPlan {
    PhaseW ph[5];
    int r0[5], r1[5], nst[5];
    // shared-memory geometry for the out-of-line pieces (they take this plan instead of a stack copy of
    // SmemView: local memory goes through what little L1 the ring leaves)
    int off_xs, off_res, off_xres, off_red, off_full, slot_bytes, n_slots, n_cons_warps;
    // what the attention phase needs of the kernel parameters (in a device function they would be loads
    // through a generic pointer that the compiler hoists into registers -- and spills)
    struct Att {
        float *kc, *vc;
        unsigned long long *ll_q, *ll_kv, *ll_att, *ll_part;
        int H, seq, kv, kv_mul, att_dim, ll_rep;
    } att;
};

// profiling state of a CTA (shared memory): phase timers of the timer thread, trace scratch
struct Prof {
    long long tacc[PH_COUNT];
    long long tmark;
    long long twait[12 + 32];  // 12 accumulators + 8 clock stamps for each of the 4 layer phases
    volatile int prod_issued;  // stages issued by the producer so far
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
    v.xres = reinterpret_cast<float *>(smem + off);
    off += (size_t)P.emb * 4;
    v.red = reinterpret_cast<float *>(smem + off);
    off += 64 * 4;
    v.full = reinterpret_cast<uint64_t *>(smem + off);  // [2][MAX_SLOTS]: even / odd uses of a slot
    v.empty = v.full + 2 * MAX_SLOTS;
    off += 3 * MAX_SLOTS * 8;
    v.sched = reinterpret_cast<const SchedStage *>(smem + off);
    return v;
}

static size_t smem_bytes_for(int n_slots, int slot_bytes, int xs_floats, int res_floats, int emb, int sched_entries)
{
    return (size_t)n_slots * slot_bytes + (size_t)xs_floats * 4 + (size_t)res_floats * 4 + (size_t)emb * 4 + 64 * 4 +
           3 * MAX_SLOTS * 8 + (size_t)sched_entries * sizeof(SchedStage);
}

__host__ __device__ inline void cta_rows(const PhaseW &ph, int cta, int G, int &r0, int &r1)
{
    const long long U = ph.rows / ph.unit;
    r0 = (int)((long long)cta * U / G) * ph.unit;
    r1 = (int)((long long)(cta + 1) * U / G) * ph.unit;
}

// ring stages that n (a multiple of the phase's unit) rows of a phase take
__host__ __device__ inline int phase_stages(const PhaseW &ph, int n)
{
    if (ph.spg > 0) return (n >> 4) * ph.spg;  // tiled q4_0: spg stages per row group of 16
    return (n + ph.rps - 1) / ph.rps;
}

// stage index = div * n_slots + mod, advanced without division
struct RingPos {
    uint32_t mod, div;
};
__device__ __forceinline__ void ring_advance(RingPos &p, uint32_t n, uint32_t ns)
{
    p.mod += n;
    while (p.mod >= ns) { p.mod -= ns; p.div++; }
}

__device__ __forceinline__ uint64_t *full_bar(uint64_t *full, uint32_t slot, uint32_t use) { return full + (use & 1u) * MAX_SLOTS + slot; }
__device__ __forceinline__ uint32_t full_par(uint32_t use) { return (use >> 1) & 1u; }

// ------------------------------------------------------------------ producer
// Segment order of a CTA's schedule: the embedding row of the token; per layer: rms_att vector,
// its rows of QKV, of WO, rms_ffn vector, its rows of W13, of W2; then rms_final vector, CLS.
// The small f32 vectors and the embedding row travel through the ring like weights ("vector
// stages", one stage each, read by all consumer warps in the prologue) so that no prologue waits
// for a demand miss queued behind megabytes of in-flight weight requests.  The list is static
// (nothing in it depends on activations): the host builds it once (build_schedule), the
// producer warp walks it -- a handful of instructions per stage.
__device__ __noinline__ void producer_loop(const StreamParams &P, const SmemView sv, const CtaPlan *cp, int token,
                                           volatile int *issued)
{
    const uint64_t pol = l2_policy_evict_first();
    const uint4 *tab = reinterpret_cast<const uint4 *>(sv.sched);
    const uint32_t ns = (uint32_t)P.n_slots;
    // this CTA's section sizes follow from its row ranges (the host list was built from the same cta_rows)
    const int n_layer = 2 + cp->nst[0] + cp->nst[1] + cp->nst[2] + cp->nst[3];
    const int e_layer_end = 1 + n_layer, total = 1 + P.L * n_layer + 1 + cp->nst[4];
    uint32_t slot = 0, par = 1, use = 0;  // par: parity of the empty barrier to wait for; use: use count of the slots
    int e = 0, l = 0;
    // L2 prefetch cursor (pf_stages > 0): while the ring is full AND everything issued has landed --
    // the consumers sit in a hand-over and HBM would idle -- the stages beyond the ring are pulled
    // into L2, so that the ring later refills at L2 speed.
    int ps = 0, pe = 0, pl = 0;
    uint32_t last_slot = 0, last_use = 0;
    uint32_t bslot = 0xffffffffu, bpar = 0, par_of_last = 0;  // empty barrier (slot, parity) of the last stage of the previous phase
    int ahead = 0;
    // pacing: issue at most one KB per `pace` SM cycles (0 = unpaced).  Every byte in flight
    // beyond bandwidth x latency only adds queueing delay in front of the latency-critical LL
    // traffic of the phase hand-overs; a paced producer keeps the queues short.
    long long next_ok = clock64();
    auto stage_addr = [&](int ee, int ll, int ss, uint32_t &bytes) {
        const uint4 st = tab[ee];
        bytes = st.z;
        unsigned long long src = ((unsigned long long)st.y << 32 | st.x) +
                                 ((unsigned long long)(st.w & ~SCHED_PHASE_START) << 4) * (unsigned)ll;
        if (ss == 0) src += (unsigned long long)(token - 1) * bytes;  // the token's embedding row
        return src;
    };
#pragma unroll 1
    for (int s = 0; s < total; s++) {
        uint32_t bytes;
        const unsigned long long src = stage_addr(e, l, s, bytes);
        // hand-over protection: at most `lookahead` stages of the next phase are issued before the last
        // stage of the current phase has been released by the consumers
        if (P.lookahead > 0 && s > 0) {
            if (tab[e].w & SCHED_PHASE_START) { ahead = 0; bslot = last_slot; bpar = par_of_last; }
            if (++ahead > P.lookahead && bslot != 0xffffffffu) {
                while (!mbar_test(&sv.empty[bslot], bpar)) { }
                bslot = 0xffffffffu;
            }
        }
        if (ps <= s) { ps = s + 1; pe = e; pl = l; if (++pe == e_layer_end && pl + 1 < P.L) { pe = 1; pl++; } }
        if (++e == e_layer_end && l + 1 < P.L) { e = 1; l++; }
        if (P.pf_stages > 0) {
            while (!mbar_test(&sv.empty[slot], par)) {
                if (ps < total && ps < s + (int)ns + P.pf_stages && s > 0 && mbar_test(full_bar(sv.full, last_slot, last_use), full_par(last_use))) {
                    long long now = clock64();
                    if (now >= next_ok) {
                        uint32_t pb;
                        const unsigned long long pa = stage_addr(pe, pl, ps, pb);
                        bulk_prefetch_l2(reinterpret_cast<const void *>(pa), pb);
                        next_ok = now + (((long long)pb * P.pace) >> 10);
                        ps++;
                        if (++pe == e_layer_end && pl + 1 < P.L) { pe = 1; pl++; }
                    }
                }
            }
        } else {
            mbar_wait(&sv.empty[slot], par, 1);
        }
        if (P.pace > 0) {
            long long now = clock64();
            while (now < next_ok) now = clock64();
            next_ok = now + (((long long)bytes * P.pace) >> 10);
        }
        uint64_t *fb = full_bar(sv.full, slot, use);
        mbar_arrive_expect_tx(fb, bytes);
        bulk_g2s(sv.ring + (size_t)slot * P.slot_bytes, reinterpret_cast<const void *>(src), bytes, fb, pol);
        last_slot = slot; last_use = use; par_of_last = par ^ 1u;  // its release completes the empty phase of parity (use & 1)
        if (++slot == ns) { slot = 0; par ^= 1u; use++; }
        *issued = s + 1;  // progress of the copy cursor, for the per-CTA trace
    }
}

// ------------------------------------------------------------------ consumer helpers
struct Cons {
    int tid, warp, lane, nt, nw;  // within the consumer group
};

// ring cursor of the consumer side: the next stage of the schedule
struct CState {
    RingPos pos;
    uint32_t gmod;  // schedule index of the next stage, mod NG (-> the group that owns it)
};
__device__ __forceinline__ void cons_advance(CState &cs, uint32_t n, uint32_t ns)
{
    ring_advance(cs.pos, n, ns);
    cs.gmod = (cs.gmod + n) % (uint32_t)NG;
}

__device__ __forceinline__ void cons_sync(const Cons &c) { named_bar_sync(CONS_BAR, c.nt); }

// Consume the `nst` stages of one weight phase.  The consumer warps form NG groups of GW warps;
// group g owns the stages whose schedule index is g (mod NG) and touches no barrier of the others,
// so a stage costs GW full-waits + GW empty-arrives, consecutive stages are in flight in different
// groups, and slots are handed back one after the other in ring order (the producer refills
// progressively, the phase tail is one stage of one group).
//   f32 / f16: warp cl of a group takes the 128-bit units u = 128 k + 32 cl + lane of every row of
//     the stage and keeps those units of the activation vector in registers for the whole phase
//     (KUT <= 4 units per lane; wider rows stream x from shared memory).
//     Lanes past the end of a row read a clamped (valid) unit against x = 0: no predication.
//     Lane-partial sums of four rows are reduced together (6 shuffles instead of 20) and the slot
//     is released as soon as its weights are in registers; partial sums go to plane cl of `res`,
//     the epilogue adds the GW planes in a fixed order.
//   q4_0: tensor cores, see consume_q4.
// Not inlined on purpose: own register allocation (x stays in registers), one copy for all phases.
struct ConsumeArgs {
    const PhaseW *ph;      // shared memory
    uint8_t *ring;
    const float *xs;
    float *res;
    uint64_t *full, *empty;
    long long *wait_cycles;  // optional trace accumulators (or null); warp-uniform
    long long *stamps;       // optional 8 clock stamps of this call (trace), or null
    int nrows, nst, slot_bytes, n_slots, slot0, use0, gmod0, warp, lane;  // use0: use count of slot0 (mod 4)
};

// four pending lane-partial sums -> four row results (see consume_phase)
struct Pending {
    float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
    int d0 = 0, d1 = 0, d2 = 0, d3 = 0, np = 0;
    __device__ __forceinline__ void flush(float *res, int lane)
    {
        const bool hi16 = lane & 16, hi8 = lane & 8;
        float k0 = hi16 ? p2 : p0, k1 = hi16 ? p3 : p1;
        k0 += __shfl_xor_sync(0xffffffffu, hi16 ? p0 : p2, 16);
        k1 += __shfl_xor_sync(0xffffffffu, hi16 ? p1 : p3, 16);
        float k = hi8 ? k1 : k0;
        k += __shfl_xor_sync(0xffffffffu, hi8 ? k0 : k1, 8);
        k += __shfl_xor_sync(0xffffffffu, k, 4);
        k += __shfl_xor_sync(0xffffffffu, k, 2);
        k += __shfl_xor_sync(0xffffffffu, k, 1);
        const int r = (lane >> 3) & 3;  // which of the four pending results this lane group holds
        const int d = r == 0 ? d0 : (r == 1 ? d1 : (r == 2 ? d2 : d3));
        if ((lane & 7) == 0 && r < np) res[d] = k;
        np = 0;
    }
    __device__ __forceinline__ void push(float v, int dst, float *res, int lane)
    {
        p3 = p2; p2 = p1; p1 = p0; p0 = v;
        d3 = d2; d2 = d1; d1 = d0; d0 = dst;
        if (++np == 4) flush(res, lane);
    }
};

// ring walk of one group through a phase: owned stages are NG apart (PROF: trace instrumentation)
template <bool PROF>
struct StageWalk {
    const ConsumeArgs &a;
    uint32_t slot, use;
    int s;  // phase-relative index of the group's next stage
    long long tc0 = 0;
    int nwaits = 0;
    __device__ __forceinline__ void stampc(int i) const
    {
        if constexpr (PROF)
            if (a.stamps && a.lane == 0) a.stamps[i] = clock64();
    }
    __device__ __forceinline__ StageWalk(const ConsumeArgs &a_) : a(a_)
    {
        stampc(1);
        int first = (a.warp / GW) - a.gmod0;
        if (first < 0) first += NG;
        s = first;
        slot = (uint32_t)a.slot0 + (uint32_t)first;
        use = (uint32_t)a.use0;
        if (slot >= (uint32_t)a.n_slots) { slot -= (uint32_t)a.n_slots; use++; }
    }
    __device__ __forceinline__ bool more() const { return s < a.nst; }
    __device__ __forceinline__ const uint8_t *wait()
    {
        if (PROF && a.wait_cycles) {
            const long long w0 = clock64();
            mbar_wait(full_bar(a.full, slot, use), full_par(use), 2);
            tc0 = clock64();
            if (a.lane == 0) a.wait_cycles[0] += tc0 - w0;
            if (nwaits++ == 0) stampc(2);
        } else {
            mbar_wait(full_bar(a.full, slot, use), full_par(use), 2);
        }
        return a.ring + (size_t)slot * a.slot_bytes;
    }
    __device__ __forceinline__ void release()
    {
        __syncwarp();
        if (PROF && a.wait_cycles && a.lane == 0) { a.wait_cycles[4] += clock64() - tc0; a.wait_cycles[8] += 1; }
        if (a.lane == 0) mbar_arrive(&a.empty[slot]);
        if (PROF && a.wait_cycles && nwaits == 1) stampc(3);
        s += NG;
        slot += NG;
        if (slot >= (uint32_t)a.n_slots) { slot -= (uint32_t)a.n_slots; use++; }
    }
};

template <int WT, int KUT, bool PROF>
__device__ __forceinline__ void consume_xreg(const ConsumeArgs &a)
{
    const PhaseW *ph = a.ph;
    const int nunits = ph->cols >> (WT == WT_F32 ? 2 : 3), rps = ph->rps, lane = a.lane, cl = a.warp % GW;
    const uint32_t rs = ph->rs;
    const float4 *x4 = reinterpret_cast<const float4 *>(a.xs);
    float *res = a.res + (size_t)cl * ph->rows_cap;
    XRegs<WT, KUT> x;
    uint32_t off[KUT];  // byte offsets of this lane's units inside a row (clamped to the row)
#pragma unroll
    for (int k = 0; k < KUT; k++) {
        const int u = k * (32 * GW) + cl * 32 + lane;
        const bool ok = u < nunits;
        off[k] = (uint32_t)min(u, nunits - 1) * 16u;
        if (WT == WT_F16) {
            x.v[2 * k] = ok ? x4[2 * u] : make_float4(0.f, 0.f, 0.f, 0.f);
            x.v[2 * k + 1] = ok ? x4[2 * u + 1] : make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            x.v[k] = ok ? x4[u] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    Pending pd;
    StageWalk<PROF> w(a);
    while (w.more()) {
        const uint8_t *sp = w.wait();
        const int base = w.s * rps, n = min(rps, a.nrows - base);
#pragma unroll 1
        for (int r = 0; r < n; r++) {
            const uint8_t *row = sp + (uint32_t)r * rs;
            // units in batches of four: four independent accumulator chains, <= 4 loads in flight
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
            for (int kb = 0; kb < KUT; kb += 4) {
                uint4 wv[4];
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (kb + k < KUT) wv[k] = *reinterpret_cast<const uint4 *>(row + off[kb + k]);
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (kb + k < KUT) {
                        if (WT == WT_F16) {
                            const float2 f0 = __half22float2(*reinterpret_cast<const __half2 *>(&wv[k].x));
                            const float2 f1 = __half22float2(*reinterpret_cast<const __half2 *>(&wv[k].y));
                            const float2 f2 = __half22float2(*reinterpret_cast<const __half2 *>(&wv[k].z));
                            const float2 f3 = __half22float2(*reinterpret_cast<const __half2 *>(&wv[k].w));
                            const float4 xa = x.v[2 * (kb + k)], xb = x.v[2 * (kb + k) + 1];
                            a0 = fmaf(f0.x, xa.x, a0); a1 = fmaf(f0.y, xa.y, a1);
                            a2 = fmaf(f1.x, xa.z, a2); a3 = fmaf(f1.y, xa.w, a3);
                            a0 = fmaf(f2.x, xb.x, a0); a1 = fmaf(f2.y, xb.y, a1);
                            a2 = fmaf(f3.x, xb.z, a2); a3 = fmaf(f3.y, xb.w, a3);
                        } else {
                            const float4 xa = x.v[kb + k];
                            a0 = fmaf(__uint_as_float(wv[k].x), xa.x, a0);
                            a1 = fmaf(__uint_as_float(wv[k].y), xa.y, a1);
                            a2 = fmaf(__uint_as_float(wv[k].z), xa.z, a2);
                            a3 = fmaf(__uint_as_float(wv[k].w), xa.w, a3);
                        }
                    }
            }
            pd.push((a0 + a1) + (a2 + a3), base + r, res, lane);
        }
        w.release();
    }
    if (pd.np) pd.flush(res, lane);
}

// rows wider than the register budget: the activation units come from shared memory
template <int WT, bool PROF>
__device__ __forceinline__ void consume_xsmem(const ConsumeArgs &a)
{
    constexpr int KB = 4;
    const PhaseW *ph = a.ph;
    const int nunits = ph->cols >> (WT == WT_F32 ? 2 : 3), rps = ph->rps, lane = a.lane, cl = a.warp % GW;
    const uint32_t rs = ph->rs;
    const float4 *x4 = reinterpret_cast<const float4 *>(a.xs);
    float *res = a.res + (size_t)cl * ph->rows_cap;
    Pending pd;
    StageWalk<PROF> w(a);
    while (w.more()) {
        const uint8_t *sp = w.wait();
        const int base = w.s * rps, n = min(rps, a.nrows - base);
#pragma unroll 1
        for (int r = 0; r < n; r++) {
            const uint8_t *row = sp + (uint32_t)r * rs;
            float acc = 0.f;
#pragma unroll 1
            for (int u0 = cl * 32 + lane; u0 < nunits; u0 += KB * 32 * GW) {
                XRegs<WT, KB> x;
                uint4 wv[KB];
#pragma unroll
                for (int k = 0; k < KB; k++) {
                    const int u = u0 + k * (32 * GW);
                    const bool ok = u < nunits;
                    const int uc = min(u, nunits - 1);
                    wv[k] = *reinterpret_cast<const uint4 *>(row + (uint32_t)uc * 16u);
                    if (WT == WT_F16) {
                        x.v[2 * k] = ok ? x4[2 * uc] : make_float4(0.f, 0.f, 0.f, 0.f);
                        x.v[2 * k + 1] = ok ? x4[2 * uc + 1] : make_float4(0.f, 0.f, 0.f, 0.f);
                    } else {
                        x.v[k] = ok ? x4[uc] : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
                acc += dot_units<WT, KB>(wv, x);
            }
            pd.push(acc, base + r, res, lane);
        }
        w.release();
    }
    if (pd.np) pd.flush(res, lane);
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
// Tiled weight format: common.cuh.  A stage holds the groups [g0, g1) of one row group of 16 rows;
// the warps of the consumer group take them round-robin and write per-warp partial sums to plane
// (stage within the row group) * GW + warp.
__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1)
{
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t uint4_word(const uint4 &v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }

template <bool PROF>
__device__ __forceinline__ void consume_q4(const ConsumeArgs &a)
{
    const PhaseW *ph = a.ph;
    const int ngrp = ph->ngrp, spg = ph->spg, lane = a.lane, g = lane >> 2, t = lane & 3, cl = a.warp % GW;
    // activations as f16 pairs x = hi + lo in B-fragment order, [half group of 4 blocks][hi | lo][block][t],
    // then the per-block offset corrections [hi | lo][block] (store_x4)
    const uint4 *xh4 = reinterpret_cast<const uint4 *>(a.xs);
    const float *C = a.xs + (size_t)ngrp * 256 + (size_t)(t >> 1) * ngrp * 8;
    StageWalk<PROF> w(a);
    while (w.more()) {
        const uint8_t *sp = w.wait();
        const int rgl = w.s / spg, si = w.s - rgl * spg;
        const int g0 = si * ngrp / spg, g1 = (si + 1) * ngrp / spg;
        float acc0 = 0.f, acc1 = 0.f;  // rows g and g + 8 of the row group
#pragma unroll 1
        for (int gi = g0 + cl; gi < g1; gi += GW) {
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
                const uint4 xb = xh4[((gi * 2 + hb) * 8 + g) * 4 + t];
                const uint2 sc = *reinterpret_cast<const uint2 *>(gp + 2048 + (g * 4 + 2 * hb + (t & 1)) * 8);
                const float2 cc = *reinterpret_cast<const float2 *>(C + gi * 8 + 4 * hb + 2 * (t & 1));
                // two accumulators (low / high nibbles): two independent mma chains of four
                float d[4] = {0.f, 0.f, 0.f, 0.f}, e[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint32_t wg = uint4_word(cg[hb], j), w8 = uint4_word(c8[hb], j);
                    const uint32_t m = ((g & 3) == j) ? 0xffffffffu : 0u;
                    const uint32_t wgs = wg >> 8, w8s = w8 >> 8;
                    mma16816(d, (wg & 0x000f000fu) | 0x64006400u, (w8 & 0x000f000fu) | 0x64006400u,
                             (wgs & 0x000f000fu) | 0x64006400u, (w8s & 0x000f000fu) | 0x64006400u, xb.x & m, xb.y & m);
                    mma16816(e, (wg & 0x00f000f0u) | 0x64006400u, (w8 & 0x00f000f0u) | 0x64006400u,
                             (wgs & 0x00f000f0u) | 0x64006400u, (w8s & 0x00f000f0u) | 0x64006400u, xb.z & m, xb.w & m);
                }
#pragma unroll
                for (int k = 0; k < 4; k++) d[k] += e[k];
                const float2 s01 = __half22float2(*reinterpret_cast<const __half2 *>(&sc.x));
                const float2 s23 = __half22float2(*reinterpret_cast<const __half2 *>(&sc.y));
                acc0 = fmaf(s01.x, d[0] - cc.x, acc0); acc0 = fmaf(s01.y, d[1] - cc.y, acc0);
                acc1 = fmaf(s23.x, d[2] - cc.x, acc1); acc1 = fmaf(s23.y, d[3] - cc.y, acc1);
            }
        }
        // sum over t: blocks 0,1 | 2,3 and the hi | lo parts
        acc0 += __shfl_xor_sync(0xffffffffu, acc0, 1); acc0 += __shfl_xor_sync(0xffffffffu, acc0, 2);
        acc1 += __shfl_xor_sync(0xffffffffu, acc1, 1); acc1 += __shfl_xor_sync(0xffffffffu, acc1, 2);
        if (t == 0) {
            float *res = a.res + (size_t)(si * GW + cl) * ph->rows_cap + rgl * 16 + g;
            res[0] = acc0;
            res[8] = acc1;
        }
        w.release();
    }
}

// Scalar parameters only (they travel in registers): a by-value struct would be written to and read
// back from local memory by every thread, four times per layer.
template <int WT, bool PROF>
__device__ __noinline__ void consume_phase(const CtaPlan *cp, Prof *pf, int ph, uint32_t cursor /* slot0 | use0 << 8 | gmod0 << 16 | trace << 24 */)
{
    extern __shared__ __align__(128) uint8_t smem[];
    ConsumeArgs a;
    a.ph = &cp->ph[ph]; a.ring = smem;
    a.xs = reinterpret_cast<const float *>(smem + cp->off_xs);
    a.res = reinterpret_cast<float *>(smem + cp->off_res);
    a.full = reinterpret_cast<uint64_t *>(smem + cp->off_full); a.empty = a.full + 2 * MAX_SLOTS;
    a.nrows = cp->r1[ph] - cp->r0[ph]; a.nst = cp->nst[ph]; a.slot_bytes = cp->slot_bytes; a.n_slots = cp->n_slots;
    a.slot0 = (int)(cursor & 255u); a.use0 = (int)((cursor >> 8) & 255u); a.gmod0 = (int)((cursor >> 16) & 255u);
    a.warp = (int)(threadIdx.x >> 5); a.lane = (int)(threadIdx.x & 31);
    a.wait_cycles = (PROF && (cursor >> 24)) ? &pf->twait[ph] : nullptr;
    a.stamps = a.wait_cycles ? &pf->twait[12 + 8 * ph] : nullptr;
    if (PROF && a.stamps && a.lane == 0) a.stamps[0] = clock64();
    if constexpr (WT != WT_Q4_0) {
        // units per lane per row, rounded up to an instantiated register budget
        const int ku = a.ph->ku;
        if (ku <= 2) consume_xreg<WT, 2, PROF>(a);
        else if (ku <= 4) consume_xreg<WT, 4, PROF>(a);
        else if (WT == WT_F16 && ku <= 6) consume_xreg<WT, WT == WT_F16 ? 6 : 2, PROF>(a);
        else if (WT == WT_F32 && ku <= 12) consume_xreg<WT, WT == WT_F32 ? 12 : 2, PROF>(a);
        else consume_xsmem<WT, PROF>(a);
    } else {
        consume_q4<PROF>(a);
    }
}

// ------------------------------------------------------------------ LL buffers
// One float per 64-bit word: low half = value bits, high half = epoch.  64-bit scalar accesses
// are single-copy atomic, so value and epoch always arrive together; relaxed gpu-scope accesses
// go to L2 (never a stale L1 line).  A buffer is rewritten one layer later at the earliest, and
// a CTA can only get there after it has seen every other CTA's output of the phases in between,
// i.e. after every reader of the old contents is done -- no write-after-read hazard.
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
// poll NV (<= 4) consecutive floats (i % NV == 0)
template <int NV>
__device__ __forceinline__ void ll_waitv(const unsigned long long *buf, int i, uint32_t ep, float (&o)[NV])
{
    if (NV == 4) {
        const float4 t = ll_wait4(buf, i, ep);
        o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
    } else if (NV == 2) {
        unsigned long long a, b;
        LLMF90_WD_DECL;
        do { ll_load2(buf + i, a, b); LLMF90_WD_CHECK(103, i, ep) } while (!(ll_ok(a, ep) && ll_ok(b, ep)));
        o[0] = ll_val(a); o[1] = ll_val(b);
    } else {
        unsigned long long a;
        LLMF90_WD_DECL;
        do { a = ll_load1(buf + i); LLMF90_WD_CHECK(104, i, ep) } while (!ll_ok(a, ep));
        o[0] = ll_val(a);
    }
}
// A thread's batch of PV float4 positions j = base + tid + k * nt of an LL vector: all requests of
// a polling round are issued before the first check (one L2 round trip per round).  Positions
// past the end are clamped to the last one (a harmless duplicate request).
template <int PV>
__device__ __forceinline__ void ll_gather(const unsigned long long *buf, int n4, int base, uint32_t ep,
                                          const Cons &c, float4 (&v)[PV])
{
    unsigned long long w[PV][4];
    int jj[PV];
#pragma unroll
    for (int k = 0; k < PV; k++) {
        const int j = base + c.tid + k * c.nt;
        jj[k] = min(j, n4 - 1);
    }
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
// the consumer-wide barrier that ends the prologue the warps of the group that owns it hand it back
// (the out-of-line pieces and the phase loop address shared memory through the plan's offsets instead
// of keeping a SmemView alive: whatever is live across a call is spilled around it)
__device__ __forceinline__ uint8_t *smem_base()
{
    extern __shared__ __align__(128) uint8_t smem[];
    return smem;
}
__device__ __forceinline__ uint64_t *plan_full(const CtaPlan *cp) { return reinterpret_cast<uint64_t *>(smem_base() + cp->off_full); }
__device__ __forceinline__ uint64_t *plan_empty(const CtaPlan *cp) { return plan_full(cp) + 2 * MAX_SLOTS; }
__device__ __forceinline__ Cons plan_cons(const CtaPlan *cp)
{
    Cons c;
    c.tid = (int)threadIdx.x; c.warp = c.tid >> 5; c.lane = c.tid & 31;
    c.nw = cp->n_cons_warps; c.nt = c.nw * 32;
    return c;
}
__device__ __forceinline__ const uint8_t *vec_stage_wait(const CtaPlan *cp, const RingPos &at)
{
    mbar_wait(full_bar(plan_full(cp), at.mod, at.div), full_par(at.div), 3);
    return smem_base() + (size_t)at.mod * cp->slot_bytes;
}
// call in stage order, after a cons_sync that follows the last read of the stage
__device__ __forceinline__ void vec_stage_release(const CtaPlan *cp, CState &cs)
{
    if ((threadIdx.x & 31) == 0 && (uint32_t)(threadIdx.x >> 5) / GW == cs.gmod) mbar_arrive(&plan_empty(cp)[cs.pos.mod]);
    cons_advance(cs, 1u, (uint32_t)cp->n_slots);
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
template <int WT>
__device__ __forceinline__ void store_x4(float *xs, int n, int j4, const float4 v, bool valid)
{
    if constexpr (WT != WT_Q4_0) {
        if (valid) reinterpret_cast<float4 *>(xs)[j4] = v;
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
__device__ __forceinline__ void q4_zero_tail(float *xs, int n, int tid, int nt)
{
    const int nblk = n >> 5, ngrp = q4t_groups(n);
    for (int b = nblk + tid; b < ngrp * 8; b += nt) {
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

// (out of line with scalar parameters: its registers do not add to the pressure of the phase loop)
template <int WT>
__device__ __noinline__ float gather_x(const CtaPlan *cp, const unsigned long long *src, int nsrc, uint32_t ep, int n,
                                       int norm, int wtype, const uint8_t *emb_row, const float *wn /* shared */)
{
    extern __shared__ __align__(128) uint8_t smem[];
    SmemView sv;
    sv.xs = reinterpret_cast<float *>(smem + cp->off_xs);
    sv.xres = reinterpret_cast<float *>(smem + cp->off_xres);
    sv.red = reinterpret_cast<float *>(smem + cp->off_red);
    Cons c;
    c.tid = (int)threadIdx.x; c.warp = c.tid >> 5; c.lane = c.tid & 31;
    c.nw = cp->n_cons_warps; c.nt = c.nw * 32;
    constexpr int PV = 2;
    const int n4 = n >> 2;
    const float4 *wn4 = reinterpret_cast<const float4 *>(wn);
    float4 *xr4 = reinterpret_cast<float4 *>(sv.xres);
    float ss = 0.f;
#pragma unroll 1
    for (int base = 0; base < n4; base += PV * c.nt) {
        float4 v[PV];
        if (emb_row) {
#pragma unroll
            for (int k = 0; k < PV; k++) {
                const int j = base + c.tid + k * c.nt;
                v[k] = emb_row4(emb_row, wtype, n, min(j, n4 - 1));
            }
        } else {
            ll_gather<PV>(src, n4, base, ep, c, v);
            if (norm) {
#pragma unroll
                for (int k = 0; k < PV; k++) {
                    const int j = base + c.tid + k * c.nt;
                    const float4 x = xr4[min(j, n4 - 1)];
                    v[k].x += x.x; v[k].y += x.y; v[k].z += x.z; v[k].w += x.w;
                }
#pragma unroll 1
                for (int r = 1; r < nsrc; r++) {
                    float4 t[PV];
                    ll_gather<PV>(src + (size_t)r * n, n4, base, ep, c, t);
#pragma unroll
                    for (int k = 0; k < PV; k++) { v[k].x += t[k].x; v[k].y += t[k].y; v[k].z += t[k].z; v[k].w += t[k].w; }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < PV; k++) {
            const int j = base + c.tid + k * c.nt;
            const bool valid = j < n4;
            float4 t = v[k];
            if (valid && norm) {
                xr4[j] = t;
                ss = fmaf(t.x, t.x, ss); ss = fmaf(t.y, t.y, ss); ss = fmaf(t.z, t.z, ss); ss = fmaf(t.w, t.w, ss);
                const float4 w = wn4[j];
                t.x *= w.x; t.y *= w.y; t.z *= w.z; t.w *= w.w;
            }
            store_x4<WT>(sv.xs, n, j, t, valid);
        }
    }
    if (WT == WT_Q4_0) q4_zero_tail(sv.xs, n, c.tid, c.nt);
    if (!norm) {
        cons_sync(c);
        return 1.f;
    }
    ss = warp_sum(ss);
    if (c.lane == 0) sv.red[c.warp] = ss;
    cons_sync(c);
    float tot = 0.f;
    for (int i = 0; i < c.nw; i++) tot += sv.red[i];
    return 1.0f / sqrtf(tot / (float)n + 1e-5f);
}

// ------------------------------------------------------------------ attention phase
// (head, split) items over the CTAs (llama2.f90:574-598); a split is a run of positions (<= 256
// up to 2048 positions of context).  Inside an item each consumer warp takes groups of 8
// positions; lane = 4 * (position in group) + (quarter of the head dimension):
//   scores : each lane dots its quarter of q_h with its quarter of one K row (q, K and V requests
//            are all issued up front: ONE L2 round trip per group), two shuffles finish the dot
//   softmax: online (running max / sum) over the 8 positions, three shuffles each
//   values : lane <-> head dimension (hs/32 consecutive dims), p_t broadcast by shuffle
// Positions past the end of a group read a clamped (valid) row with probability 0.
// Warp partials merge through shared memory.  With one split the normalised head output goes
// straight to P.att; otherwise {m, l, acc} partials go to P.att_part and the Wo prologue merges.
constexpr int ATT_PSTRIDE_PAD = 4;  // partial record = {m, l, -, -, acc[hs]}

template <int VEC>
__device__ __forceinline__ void load_vec(const float *p, float (&o)[VEC])
{
    if (VEC == 4) {
        const float4 t = __ldcg(reinterpret_cast<const float4 *>(p));
        o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
    } else if (VEC == 2) {
        const float2 t = __ldcg(reinterpret_cast<const float2 *>(p));
        o[0] = t.x; o[1] = t.y;
    } else {
        o[0] = __ldcg(p);
    }
}

template <int HS>
__device__ __noinline__ void attention_phase_t(const CtaPlan *cp, int n_splits, int layer, int pos, uint32_t ep)
{
    extern __shared__ __align__(128) uint8_t smem[];
    SmemView sv;
    sv.xs = reinterpret_cast<float *>(smem + cp->off_xs);
    Cons c;
    c.tid = (int)threadIdx.x; c.warp = c.tid >> 5; c.lane = c.tid & 31;
    c.nw = cp->n_cons_warps; c.nt = c.nw * 32;
    constexpr int hs = HS, vec = HS >> 5;  // HS in {32, 64, 128}
    constexpr int q4n = HS >> 4;           // float4 per lane of a quarter head: 2, 4, 8
    const CtaPlan::Att &T = cp->att;
    const int S = n_splits;
    const int items = T.H * S;
    const int npast = pos - 1;  // positions 0 .. pos-2 come from the cache (earlier launches)
    const int s_shift = 31 - __clz(S), kvm_shift = 31 - __clz(T.kv_mul);
    const bool kvm_pow2 = (T.kv_mul & (T.kv_mul - 1)) == 0;
    const int chunk = (((npast + S - 1) >> s_shift) + 7) & ~7;
    const int pstride = hs + ATT_PSTRIDE_PAD;
    const float rscale = 1.0f / sqrtf((float)hs);
    float *sc = sv.xs;  // [nw][pstride]  (xs is dead between weight phases)
    const float *kc = T.kc + (size_t)layer * T.seq * T.kv;
    const float *vc = T.vc + (size_t)layer * T.seq * T.kv;
    const int pl = c.lane >> 2, dq = c.lane & 3;
    const int rep = (int)blockIdx.x % T.ll_rep;  // the replica of the LL vectors this CTA polls
    const unsigned long long *ll_q = T.ll_q + (size_t)rep * T.att_dim, *ll_kv = T.ll_kv + (size_t)rep * 2 * T.kv;
#pragma unroll 1
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        // (S is a power of two; kv_mul is one in every common model: shifts, the division is a cold path)
        const int h = item >> s_shift, sp = item & (S - 1);
        const int g = kvm_pow2 ? h >> kvm_shift : h / T.kv_mul;
        const int t0 = sp * chunk, t1 = min(npast, t0 + chunk);
        float m = -INFINITY, l = 0.f, acc[vec];
#pragma unroll
        for (int i = 0; i < vec; i++) acc[i] = 0.f;
        // this lane's quarter of the query head (written by the QKV epilogues of this launch);
        // polled after the first group's K / V requests are in flight
        float4 qq[q4n];
        bool have_q = false;
#pragma unroll 1
        for (int tb = t0 + 8 * c.warp; tb < t1; tb += 8 * c.nw) {
            const int t = tb + pl;
            const bool valid = t < t1;
            const float4 *kr = reinterpret_cast<const float4 *>(kc + (size_t)min(t, t1 - 1) * T.kv + (size_t)g * hs +
                                                                dq * (hs >> 2));
            const float *vb = vc + (size_t)g * hs + c.lane * vec;
            float4 kk[q4n];
            float vv[8][vec];
#pragma unroll
            for (int i = 0; i < q4n; i++) kk[i] = __ldcg(kr + i);
#pragma unroll
            for (int u = 0; u < 8; u++) load_vec<vec>(vb + (size_t)min(tb + u, t1 - 1) * T.kv, vv[u]);
            if (!have_q) {
                ll_wait4n<q4n>(ll_q, h * hs + dq * (hs >> 2), ep, qq);
                have_q = true;
            }
            float sdot = 0.f;
#pragma unroll
            for (int i = 0; i < q4n; i++) {
                sdot = fmaf(qq[i].x, kk[i].x, sdot); sdot = fmaf(qq[i].y, kk[i].y, sdot);
                sdot = fmaf(qq[i].z, kk[i].z, sdot); sdot = fmaf(qq[i].w, kk[i].w, sdot);
            }
            sdot += __shfl_xor_sync(0xffffffffu, sdot, 1);
            sdot += __shfl_xor_sync(0xffffffffu, sdot, 2);
            sdot = valid ? sdot * rscale : -INFINITY;  // dot_product(q_t,k_t)/sqrt(head_size), :582
            float bm = sdot;
            bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 4));
            bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 8));
            bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 16));
            const float mn = fmaxf(m, bm);
            const float corr = expf(m - mn);
            const float p = valid ? expf(sdot - mn) : 0.f;
            float ps = p;  // every position is held by 4 lanes: sum over the position bits only
            ps += __shfl_xor_sync(0xffffffffu, ps, 4);
            ps += __shfl_xor_sync(0xffffffffu, ps, 8);
            ps += __shfl_xor_sync(0xffffffffu, ps, 16);
            l = fmaf(l, corr, ps);
            m = mn;
#pragma unroll
            for (int i = 0; i < vec; i++) acc[i] *= corr;
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const float pt = __shfl_sync(0xffffffffu, p, 4 * u);
#pragma unroll
                for (int i = 0; i < vec; i++) acc[i] = fmaf(pt, vv[u][i], acc[i]);
            }
        }
        if (sp == S - 1 && c.warp == 0) {
            // the current position: its key / value rows were produced in this launch (LL buffer)
            float4 kk[q4n];
            float vv[vec];
            if (!have_q) ll_wait4n<q4n>(ll_q, h * hs + dq * (hs >> 2), ep, qq);
            ll_wait4n<q4n>(ll_kv, g * hs + dq * (hs >> 2), ep, kk);
            ll_waitv<vec>(ll_kv, T.kv + g * hs + c.lane * vec, ep, vv);
            float sdot = 0.f;
#pragma unroll
            for (int i = 0; i < q4n; i++) {
                sdot = fmaf(qq[i].x, kk[i].x, sdot); sdot = fmaf(qq[i].y, kk[i].y, sdot);
                sdot = fmaf(qq[i].z, kk[i].z, sdot); sdot = fmaf(qq[i].w, kk[i].w, sdot);
            }
            sdot += __shfl_xor_sync(0xffffffffu, sdot, 1);
            sdot += __shfl_xor_sync(0xffffffffu, sdot, 2);
            sdot = sdot * rscale;  // identical in every lane (all lanes hold the same position)
            const float mn = fmaxf(m, sdot);
            const float corr = expf(m - mn), p = expf(sdot - mn);
            l = fmaf(l, corr, p);
            m = mn;
#pragma unroll
            for (int i = 0; i < vec; i++) acc[i] = fmaf(acc[i], corr, p * vv[i]);
        }
        float *mine = sc + (size_t)c.warp * pstride;
        if (c.lane == 0) { mine[0] = m; mine[1] = l; }
#pragma unroll
        for (int i = 0; i < vec; i++) mine[ATT_PSTRIDE_PAD + c.lane * vec + i] = acc[i];
        cons_sync(c);
        // thread = (head dimension d, LL replica): every thread publishes one word
        for (int w = c.tid; w < hs * (S == 1 ? T.ll_rep : 1); w += c.nt) {
            const int d = w & (hs - 1), rr = w / hs;
            float M = -INFINITY;
#pragma unroll 1
            for (int w = 0; w < c.nw; w++) M = fmaxf(M, sc[w * pstride]);
            float L = 0.f, A = 0.f;
#pragma unroll 1
            for (int w = 0; w < c.nw; w++) {
                const float mw = sc[w * pstride];
                if (mw > -INFINITY) {
                    const float e = __expf(mw - M);
                    L = fmaf(sc[w * pstride + 1], e, L);
                    A = fmaf(sc[w * pstride + ATT_PSTRIDE_PAD + d], e, A);
                }
            }
            if (S == 1) {
                ll_store(T.ll_att + (size_t)rr * T.att_dim, h * hs + d, A / L, ep);
            } else {
                unsigned long long *out = T.ll_part + (size_t)(h * S + sp) * pstride;
                ll_store(out, ATT_PSTRIDE_PAD + d, A, ep);
                if (d == 0) { ll_store(out, 0, M, ep); ll_store(out, 1, L, ep); }
            }
        }
        cons_sync(c);
    }
}

// xs = attention output (all heads), merging the position splits (n_splits > 1).  A thread owns one
// float4 of the output; the {m, l} pair and the float4 of all S partial records are requested
// together (one L2 round trip per polling round).
template <int WT>
__device__ __noinline__ void load_x_attn(const StreamParams &P, const CtaPlan *cp, uint32_t ep)
{
    SmemView sv;
    sv.xs = reinterpret_cast<float *>(smem_base() + cp->off_xs);
    const Cons c = plan_cons(cp);
    const int S = P.n_splits;
    const int hs = P.hs, pstride = hs + ATT_PSTRIDE_PAD, n4 = P.att_dim >> 2;
    const int hs_shift = hs == 64 ? 6 : (hs == 128 ? 7 : 5);
    for (int jj = c.tid; jj < ((n4 + 31) & ~31); jj += c.nt) {
        const bool valid = jj < n4;
        const int j = valid ? jj : n4 - 1;
        const int h = (4 * j) >> hs_shift, d = (4 * j) & (hs - 1);
        const unsigned long long *part = P.ll_part + (size_t)h * S * pstride;
        unsigned long long ml[8][2], av[8][4];
        bool ok;
        LLMF90_WD_DECL;
        do {
            LLMF90_WD_CHECK(107, j, ep)
#pragma unroll
            for (int s = 0; s < 8; s++)
                if (s < S) {
                    const unsigned long long *rec = part + (size_t)s * pstride;
                    ll_load2(rec, ml[s][0], ml[s][1]);
                    ll_load2(rec + ATT_PSTRIDE_PAD + d, av[s][0], av[s][1]);
                    ll_load2(rec + ATT_PSTRIDE_PAD + d + 2, av[s][2], av[s][3]);
                }
            ok = true;
#pragma unroll
            for (int s = 0; s < 8; s++)
                if (s < S)
                    ok = ok && ll_ok(ml[s][0], ep) && ll_ok(ml[s][1], ep) && ll_ok(av[s][0], ep) && ll_ok(av[s][1], ep) &&
                         ll_ok(av[s][2], ep) && ll_ok(av[s][3], ep);
        } while (!ok);
        float M = -INFINITY;
#pragma unroll
        for (int s = 0; s < 8; s++)
            if (s < S) M = fmaxf(M, ll_val(ml[s][0]));
        float den = 0.f;
        float4 num = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int s = 0; s < 8; s++)
            if (s < S && ll_val(ml[s][0]) > -INFINITY) {
                const float w = expf(ll_val(ml[s][0]) - M);
                den = fmaf(ll_val(ml[s][1]), w, den);
                num.x = fmaf(ll_val(av[s][0]), w, num.x); num.y = fmaf(ll_val(av[s][1]), w, num.y);
                num.z = fmaf(ll_val(av[s][2]), w, num.z); num.w = fmaf(ll_val(av[s][3]), w, num.w);
            }
        store_x4<WT>(sv.xs, P.att_dim, jj, make_float4(num.x / den, num.y / den, num.z / den, num.w / den), valid);
    }
    if (WT == WT_Q4_0) q4_zero_tail(sv.xs, P.att_dim, c.tid, c.nt);
    cons_sync(c);
}

// ------------------------------------------------------------------ profiling hooks (out of line)
__device__ __noinline__ void prof_lap(Prof *pf, int bucket)
{
    const long long now = clock64();
    pf->tacc[bucket] += now - pf->tmark;
    pf->tmark = now;
}
// per-CTA trace of one layer: globaltimer stamp, producer / consumer ring cursors and the number
// of landed stages at phase edge k
__device__ __noinline__ void prof_stamp(const StreamParams &P, const CtaPlan *cp, uint32_t pos_mod, uint32_t pos_div, Prof *pf, int k)
{
    SmemView sv;
    sv.full = plan_full(cp);
    CState cs;
    cs.pos.mod = pos_mod; cs.pos.div = pos_div;
    unsigned long long *row = P.trace + (size_t)blockIdx.x * 128;
    row[k] = globaltimer_ns();
    row[32 + k] = (unsigned long long)pf->prod_issued;
    RingPos at = cs.pos;
    int landed = 0;
    for (int i = 0; i < P.n_slots; i++) {
        landed += mbar_test(full_bar(sv.full, at.mod, at.div), full_par(at.div)) ? 1 : 0;
        ring_advance(at, 1u, (uint32_t)P.n_slots);
    }
    row[48 + k] = (unsigned long long)(cs.pos.div * (uint32_t)P.n_slots + cs.pos.mod) | ((unsigned long long)landed << 32);
}

// ------------------------------------------------------------------ token tail (once per launch)
// all-gathered logits must have landed on every rank before any rank's kernel ends (tp > 1), then
// maxloc(logits) (llama2.f90:388): first maximum wins at every reduction level
__device__ __noinline__ void token_tail(const StreamParams &P, const CtaPlan *cp, float best, int bidx, int pos)
{
    SmemView sv;
    sv.red = reinterpret_cast<float *>(smem_base() + cp->off_red);
    const Cons c = plan_cons(cp);
    const uint32_t epl = P.ep_base + (uint32_t)P.L + 1u;
    const int G = (int)gridDim.x;
    if (P.tp > 1) {
        // each CTA flags every rank once its rows are stored, CTA 0 of every rank collects the flags
        cons_sync(c);
        if (c.tid == 0) {
            asm volatile("fence.acq_rel.sys;" ::: "memory");
            for (int k = 0; k < P.tp; k++) ll_store_sys(P.done[k], P.rank * G + (int)blockIdx.x, 0.f, epl);
        }
        if (blockIdx.x == 0) {
            for (int i = c.tid; i < P.tp * G; i += c.nt) {
                float t[1];
                ll_waitv<1>(P.done[P.rank], i, epl, t);
            }
            asm volatile("fence.acq_rel.sys;" ::: "memory");
        }
    }
    if (!P.do_argmax) return;
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
        for (int k = 0; k < P.tp; k++) {
            unsigned long long *dst = P.amax[k] + (size_t)(P.rank * G + (int)blockIdx.x) * 2;
            ll_store_sys(dst, 0, best, epl);
            ll_store_sys(dst, 1, __int_as_float(bidx), epl);
        }
    }
    if (blockIdx.x == 0 && c.warp == 0) {
        // every rank reduces the same tp * grid records in the same order -> the same token
        best = -INFINITY; bidx = 0x7fffffff;
        for (int i = c.lane; i < P.tp * G; i += 32) {
            float rec[2];
            ll_waitv<2>(P.amax[P.rank], 2 * i, epl, rec);
            const float v = rec[0];
            const int ix = __float_as_int(rec[1]);
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

// ------------------------------------------------------------------ the kernel
// PROF = false is the production kernel; PROF = true adds the phase timers of CTA 0 and the optional
// per-CTA trace (the instrumentation alone costs 5-8 % of the token time: measured).
template <int WT, bool PROF>
__global__ void __launch_bounds__(416, 1)
stream_decode_kernel(const __grid_constant__ StreamParams P)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const SmemView sv = carve(smem, P);
    const int n_cons_warps = P.n_cons_warps;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    __shared__ CtaPlan cp;
    __shared__ float2 rope[64];
    __shared__ Prof pf;
    if (threadIdx.x == 0) pf.prod_issued = 0;
    if (threadIdx.x == 5) {
        cp.off_xs = (int)(reinterpret_cast<uint8_t *>(sv.xs) - smem);
        cp.off_res = (int)(reinterpret_cast<uint8_t *>(sv.res) - smem);
        cp.off_xres = (int)(reinterpret_cast<uint8_t *>(sv.xres) - smem);
        cp.off_red = (int)(reinterpret_cast<uint8_t *>(sv.red) - smem);
        cp.off_full = (int)(reinterpret_cast<uint8_t *>(sv.full) - smem);
        cp.slot_bytes = P.slot_bytes; cp.n_slots = P.n_slots; cp.n_cons_warps = P.n_cons_warps;
        cp.att.kc = P.kc; cp.att.vc = P.vc; cp.att.ll_q = P.ll_q; cp.att.ll_kv = P.ll_kv; cp.att.ll_att = P.ll_att;
        cp.att.ll_part = P.ll_part; cp.att.H = P.H; cp.att.seq = P.seq; cp.att.kv = P.kv; cp.att.kv_mul = P.kv_mul;
        cp.att.att_dim = P.att_dim; cp.att.ll_rep = P.ll_rep;
    }
    if (threadIdx.x < 5) {
        const int i = threadIdx.x;
        cp.ph[i] = P.ph[i];
        int r0, r1;
        cta_rows(P.ph[i], blockIdx.x, gridDim.x, r0, r1);
        cp.r0[i] = r0; cp.r1[i] = r1;
        cp.nst[i] = phase_stages(P.ph[i], r1 - r0);
    }
    {
        const uint4 *g = reinterpret_cast<const uint4 *>(P.sched + (size_t)blockIdx.x * P.sched_stride);
        uint4 *d = reinterpret_cast<uint4 *>(const_cast<SchedStage *>(sv.sched));
        for (int i = threadIdx.x; i < P.sched_stride; i += blockDim.x) d[i] = __ldg(g + i);
    }
    if (threadIdx.x == 32) {
        for (int i = 0; i < P.n_slots; i++) {
            mbar_init(&sv.full[i], 1);
            mbar_init(&sv.full[MAX_SLOTS + i], 1);
            mbar_init(&sv.empty[i], (uint32_t)GW);
        }
        fence_mbar_init();
    }
    const int token = P.token > 0 ? P.token : P.tokpos[0];
    const int pos = P.token > 0 ? P.pos : P.tokpos[1];
    // this position's RoPE row (a cold HBM read, issued first thing, used after the QKV phase)
    if (threadIdx.x < (P.hs >> 1)) rope[threadIdx.x] = P.rope_tab[(size_t)(pos - 1) * (P.hs >> 1) + threadIdx.x];
    __syncthreads();

    if (warp == n_cons_warps) {
        // ===================== producer warp =====================
        if (lane == 0) producer_loop(P, sv, &cp, token, &pf.prod_issued);
        return;
    }

    // ===================== consumer warps =====================
    const int tid = (int)threadIdx.x, nt = n_cons_warps * 32;
    CState cs;
    cs.pos.mod = 0; cs.pos.div = 0; cs.gmod = 0;
    // phase timers (CTA 0, thread 0): SM cycles per fine-grained bucket, see PH_* in kernels.cuh;
    // optional per-CTA trace of one layer (debug/profiling)
    const bool timer = PROF && (blockIdx.x == 0 && tid == 0);
    const bool tracing = PROF && P.trace != nullptr;
    const unsigned long long t_ns0 = timer ? globaltimer_ns() : 0ull;
    const long long t_c0 = timer ? clock64() : 0ll;
    if (timer) {
        for (int i = 0; i < PH_COUNT; i++) pf.tacc[i] = 0;
        pf.tmark = t_c0;
    }
    if (tracing && tid == 0)
        for (int i = 0; i < 44; i++) pf.twait[i] = 0;
#define LAP(b) do { if constexpr (PROF) { if (timer) prof_lap(&pf, (b)); } } while (0)
#define STAMP(l_, k_) do { if constexpr (PROF) { if (tracing && tid == 0 && (l_) == P.trace_layer) prof_stamp(P, &cp, cs.pos.mod, cs.pos.div, &pf, (k_)); } } while (0)
    const int half_mask = (P.hs >> 1) - 1;
    const uint32_t ns = (uint32_t)P.n_slots;
    const int nrep = P.ll_rep, rep = (int)blockIdx.x % nrep;  // LL vector replicas; the one this CTA polls
    const int hb_stride = (P.hid + 1) & ~1;
    float best = -INFINITY;  // running maxloc of this thread's logits (classifier epilogue)
    int bidx = 0x7fffffff;

    // One loop over the 4 L + 1 weight phases (q = 4 l + {0 QKV, 1 WO, 2 W13, 3 W2}; q = 4 L is the
    // classifier): prologue -> ring consumption -> epilogue (-> attention).  A single copy of
    // every piece serves all phases.
    const int nq = 4 * P.L + 1;
#pragma unroll 1
    for (int q = 0; q < nq; q++) {
        const int ph = q < 4 * P.L ? (q & 3) : 4, l = q >> 2;
        const uint32_t ep = P.ep_base + (uint32_t)l + 1u;  // epoch of everything layer l publishes
        const int tb = ph == 0 ? 0 : 2 + 3 * ph;  // timer bucket / trace stamp base of this phase
        const bool norm = !(ph & 1);
        const int r0 = cp.r0[ph], nr = cp.r1[ph] - cp.r0[ph];
        if (ph == 0) STAMP(l, 0);

        // ---- prologue: the activation vector of this phase, in shared memory
        float rscale = 1.f;
        if (norm) {
            // rmsnorm (llama2.f90:527, :608, :627); layer 0 starts from the embedding row (:520);
            // adds the tp partial outputs of the phase before (Wo of this layer / W2 of the previous one)
            const uint8_t *emb_row = nullptr;
            RingPos at = cs.pos;
            if (q == 0) {
                emb_row = vec_stage_wait(&cp, at);
                ring_advance(at, 1u, ns);
            }
            const float *wn = reinterpret_cast<const float *>(vec_stage_wait(&cp, at));
            rscale = gather_x<WT>(&cp, (ph == 2 ? P.part1[P.rank] : P.part2[P.rank]) + (size_t)rep * P.tp * P.emb, P.tp,
                                  ph == 2 ? ep : ep - 1u, P.emb, 1, P.wtype, emb_row, wn);
            if (q == 0) vec_stage_release(&cp, cs);
            vec_stage_release(&cp, cs);
        } else if (nr > 0) {
            // (a CTA without rows in a Wo / W2 phase does not need its input vector: skipping the poll
            // also keeps it from ever lagging behind on a buffer nobody waits for it to have read)
            if (ph == 1 && P.n_splits > 1) load_x_attn<WT>(P, &cp, ep);
            else gather_x<WT>(&cp, ph == 1 ? P.ll_att + (size_t)rep * P.att_dim : P.ll_hb + (size_t)rep * hb_stride, 1, ep,
                              ph == 1 ? P.att_dim : P.hid, 0, P.wtype, nullptr, nullptr);
        }
        LAP(tb);
        STAMP(l, ph == 0 ? 1 : 3 + 3 * ph);

        // ---- the mat-vec: consume this CTA's stages of the phase from the ring
        {
            const bool trace_me = tracing && warp == 0 && l == P.trace_layer && ph < 4;
            const uint32_t cursor = cs.pos.mod | ((cs.pos.div & 3u) << 8) | (cs.gmod << 16) | (trace_me ? 1u << 24 : 0u);
            if (trace_me && lane == 0) pf.twait[12 + 8 * ph + 7] = clock64();  // before the call
            consume_phase<WT, PROF>(&cp, &pf, ph, cursor);
            if (trace_me && lane == 0) pf.twait[12 + 8 * ph + 4] = clock64();  // after the return
            cons_advance(cs, (uint32_t)cp.nst[ph], ns);
        }
        STAMP(l, ph == 0 ? 2 : 4 + 3 * ph);
        named_bar_sync(CONS_BAR, nt);
        if (tracing && tid == 0 && l == P.trace_layer && ph < 4) pf.twait[12 + 8 * ph + 5] = clock64();
        LAP(tb + 1);

        // ---- epilogue.  A work item is (row pair, LL replica [, destination rank]): the planes of
        // partial sums are added in a fixed order, times the 1 / rms factor of the phase's rmsnorm,
        // and every thread publishes ONE 16-byte LL record -- relaxed gpu-scope stores are slow one
        // after the other from one thread, so they are spread over the threads instead.
        {
            const int cap = cp.ph[ph].rows_cap, rows_real = cp.ph[ph].rows_real;
            const int planes = (WT == WT_Q4_0) ? cp.ph[ph].spg * GW : GW;
            // (replica and rank counts are powers of two: shifts)
            const int nsub = ph == 4 ? 1 : ((ph & 1) ? nrep * P.tp : nrep), npairs = (nr + 1) >> 1;
            const int sub_shift = 31 - __clz(nsub);
            const float *res = reinterpret_cast<const float *>(smem_base() + cp.off_res);
#pragma unroll 1
            for (int w = tid; w < npairs * nsub; w += nt) {
                const int pair = w >> sub_shift, sub = w & (nsub - 1), i = 2 * pair;
                if (r0 + i >= rows_real) continue;  // padding rows of a tiled q4_0 matrix
                const bool two = i + 1 < nr && r0 + i + 1 < rows_real;
                float a = res[i], b = two ? res[i + 1] : 0.f;
#pragma unroll 4
                for (int p = 1; p < planes; p++) {
                    a += res[p * cap + i];
                    b += two ? res[p * cap + i + 1] : 0.f;
                }
                a *= rscale; b *= rscale;
                const int r = r0 + i;
                if (ph == 0) {
                    // RoPE on q and k (reference quirks Q1/Q2 are in the table), KV append (llama2.f90:543-565)
                    const bool isq = r < P.att_dim, isv = r >= P.att_dim + P.kv;
                    const int rk = isq ? r : (isv ? r - P.att_dim - P.kv : r - P.att_dim);
                    const float2 cs2 = isv ? make_float2(1.f, 0.f) : rope[(rk >> 1) & half_mask];
                    const float o0 = a * cs2.x - b * cs2.y, o1 = a * cs2.y + b * cs2.x;
                    if (!isq && sub == 0) {
                        // the cache row serves later launches, the LL copy this launch's attention
                        float *cache = (isv ? P.vc : P.kc) + ((size_t)l * P.seq + (pos - 1)) * P.kv;
                        *reinterpret_cast<float2 *>(cache + rk) = make_float2(o0, o1);
                    }
                    unsigned long long *dst = isq ? P.ll_q + (size_t)sub * P.att_dim
                                                  : P.ll_kv + (size_t)sub * 2 * P.kv + (isv ? P.kv : 0);
                    ll_store2(dst, rk, o0, o1, ep);
                } else if (ph == 2) {
                    // SwiGLU on the interleaved gate/up rows (llama2.f90:613-616)
                    ll_store(P.ll_hb + (size_t)sub * hb_stride, r >> 1, (a * (1.0f / (1.0f + expf(-a)))) * b, ep);
                } else if (ph == 4) {
                    // logits rows of this rank go to every rank's full logits buffer (all-gather)
                    const int gi = P.v_off + r;
                    for (int k = 0; k < P.tp; k++) {
                        P.logits[k][gi] = a;
                        if (two) P.logits[k][gi + 1] = b;
                    }
                    if (a > best) { best = a; bidx = gi; }
                    if (two && b > best) { best = b; bidx = gi + 1; }
                } else {
                    // Wo / W2 (llama2.f90:603-605, :618-620): publish this rank's partial sums to every
                    // rank; the residual add happens in the next norm prologue, on every CTA's copy of x
                    const int k = sub >> (31 - __clz(nrep)), rr = sub & (nrep - 1);  // destination rank, replica
                    unsigned long long *dst = (ph == 1 ? P.part1[k] : P.part2[k]) + ((size_t)rr * P.tp + P.rank) * P.emb;
                    ll_store_sys(dst, r, a, ep);  // (r may be odd here: no 16-byte store)
                    if (two) ll_store_sys(dst, r + 1, b, ep);
                }
            }
        }
        if (tracing && tid == 0 && l == P.trace_layer && ph < 4) pf.twait[12 + 8 * ph + 6] = clock64();
        if (ph == 4) break;
        LAP(tb + 2);
        STAMP(l, ph == 0 ? 3 : 5 + 3 * ph);

        if (ph == 0) {
            // ---- attention (llama2.f90:574-598)
            if (P.hs == 64) attention_phase_t<64>(&cp, P.n_splits, l, pos, ep);
            else if (P.hs == 128) attention_phase_t<128>(&cp, P.n_splits, l, pos, ep);
            else attention_phase_t<32>(&cp, P.n_splits, l, pos, ep);
            LAP(PH_ATT);
            STAMP(l, 4);
            STAMP(l, 5);
        }
        if (ph == 3 && tracing && tid == 0 && l == P.trace_layer)
            for (int i = 0; i < 44; i++)
                P.trace[(size_t)blockIdx.x * 128 + (i < 12 ? 16 + i : 64 + i - 12)] = (unsigned long long)pf.twait[i];
    }
    LAP(PH_CLS_MV);

    token_tail(P, &cp, best, bidx, pos);
    if (timer) {
        prof_lap(&pf, PH_ARGMAX);
        for (int i = 0; i < PH_COUNT; i++) P.phase_cycles[i] += (unsigned long long)pf.tacc[i];
        P.phase_cycles[PH_COUNT] += (unsigned long long)(clock64() - t_c0);
        P.phase_cycles[PH_COUNT + 1] += globaltimer_ns() - t_ns0;
    }
#undef LAP
#undef STAMP
}

// ------------------------------------------------------------------ host side
// stages a CTA can have in the layer section / after it (upper bounds over all CTAs)
static void sched_caps(const StreamParams &p, int *layer, int *post)
{
    int per_layer = 2, cls = 0;
    for (int i = 0; i < 5; i++) {
        const int st = phase_stages(p.ph[i], p.ph[i].rows_cap);
        if (i < 4) per_layer += st; else cls = st;
    }
    *layer = per_layer;
    *post = 1 + cls;
}

int plan_stream(StreamParams &p, int grid, int max_smem_optin, int target_slot_bytes, int max_slots,
                int cons_warps, StreamPlan *out)
{
    cons_warps = MAX_CONS_WARPS;  // NG groups of GW warps
    const bool tiled = p.ph[0].ngrp > 0;  // q4_0 in the tiled mma format
    unsigned int rs_max = (unsigned)p.emb * 4u;  // f32 vector stages
    if ((unsigned)row_stride_bytes(p.wtype, p.emb) > rs_max) rs_max = (unsigned)row_stride_bytes(p.wtype, p.emb);
    for (int i = 0; i < 5; i++) {
        PhaseW &w = p.ph[i];
        if (tiled) {
            // split a row group into stages of whole 8-block groups that fit the target slot
            w.spg = (int)((w.rs + (unsigned)target_slot_bytes - 1) / (unsigned)target_slot_bytes);
            const unsigned stage = (unsigned)((w.ngrp + w.spg - 1) / w.spg) * Q4T_GROUP_BYTES;
            if (stage > rs_max) rs_max = stage;
        } else {
            w.spg = 0;
            if (w.rs > rs_max) rs_max = w.rs;
        }
    }
    int slot = (!tiled && target_slot_bytes > (int)rs_max) ? target_slot_bytes : (int)rs_max;
    slot = (slot + 127) & ~127;
    int res_floats = 0;
    for (int i = 0; i < 5; i++) {
        PhaseW &w = p.ph[i];
        const int U = w.rows / w.unit;
        w.rows_cap = ((U + grid - 1) / grid) * w.unit;
        const int nunits = w.cols >> (p.wtype == WT_F32 ? 2 : 3);
        w.ku = tiled ? 0 : (nunits + 32 * GW - 1) / (32 * GW);
        const int planes = tiled ? w.spg * GW : GW;
        if (planes * w.rows_cap > res_floats) res_floats = planes * w.rows_cap;
    }
    int xs_floats = p.emb > p.hid ? p.emb : p.hid;
    if (tiled)  // f16 hi + lo activations in fragment order + two corrections per block (store_x4)
        for (int i = 0; i < 5; i++) xs_floats = xs_floats > p.ph[i].ngrp * 272 ? xs_floats : p.ph[i].ngrp * 272;
    // attention scratch ([consumer warps][hs+4] partials) also lives in xs
    const int att_scratch = MAX_SLOTS * (p.hs + 4);
    if (att_scratch > xs_floats) xs_floats = att_scratch;
    xs_floats = (xs_floats + 31) & ~31;
    res_floats = (res_floats + 31) & ~31;
    if (max_slots > MAX_SLOTS) max_slots = MAX_SLOTS;
    for (int i = 0; i < 5; i++) p.ph[i].rps = (!tiled && slot / (int)p.ph[i].rs > 0) ? slot / (int)p.ph[i].rs : 1;
    int cap_layer, cap_post;
    sched_caps(p, &cap_layer, &cap_post);
    const int sched_entries = 1 + cap_layer + cap_post + 1;
    int n_slots = max_slots;
    while (n_slots > 0 &&
           smem_bytes_for(n_slots, slot, xs_floats, res_floats, p.emb, sched_entries) > (size_t)max_smem_optin)
        n_slots--;
    if (n_slots < NG) return 1;
    out->n_slots = n_slots;
    out->slot_bytes = slot;
    out->n_cons_warps = cons_warps;
    out->threads = (cons_warps + 1) * 32;
    out->smem_bytes = (int)smem_bytes_for(n_slots, slot, xs_floats, res_floats, p.emb, sched_entries);
    out->grid = grid;
    out->xs_floats = xs_floats;
    out->res_floats = res_floats;
    return 0;
}

void build_schedule(StreamParams &p, int grid, SchedStage **out)
{
    // section sizes are per CTA (they follow from its row ranges: the kernel derives them from the same
    // cta_rows()); only the padded stride is shared
    int cap_layer, cap_post;
    sched_caps(p, &cap_layer, &cap_post);
    const int cap = 1 + cap_layer + cap_post + 1;
    SchedStage *tab = (SchedStage *)calloc((size_t)grid * cap, sizeof(SchedStage));
    for (int cta = 0; cta < grid; cta++) {
        SchedStage *t = tab + (size_t)cta * cap;
        int n = 0;
        bool start = true;  // the next stage pushed is the first of a phase (its vector stage, if it has one)
        auto vec = [&](const void *ptr, unsigned bytes, size_t layer_stride) {
            t[n].src = (unsigned long long)ptr; t[n].bytes = bytes;
            t[n].stride16 = (unsigned)(layer_stride >> 4) | (start ? SCHED_PHASE_START : 0u);
            start = false;
            n++;
        };
        auto rows = [&](int ph) {
            int r0, r1;
            cta_rows(p.ph[ph], cta, grid, r0, r1);
            const PhaseW &w = p.ph[ph];
            const size_t ls = ph < 4 ? (size_t)w.layer_stride : 0;
            if (ph == 1 || ph == 3) start = true;  // Wo / W2 have no vector stage: the phase starts with its rows
            if (w.spg > 0) {
                // tiled q4_0: spg stages per row group of 16, stage i = groups [i ngrp / spg, (i + 1) ngrp / spg)
                for (int rg = r0 >> 4; rg < (r1 >> 4); rg++)
                    for (int i = 0; i < w.spg; i++) {
                        const int g0 = i * w.ngrp / w.spg, g1 = (i + 1) * w.ngrp / w.spg;
                        vec(w.base + ((size_t)rg * w.ngrp + g0) * Q4T_GROUP_BYTES, (unsigned)(g1 - g0) * Q4T_GROUP_BYTES, ls);
                    }
                return;
            }
            for (int r = r0; r < r1; r += w.rps) {
                const int k = r1 - r < w.rps ? r1 - r : w.rps;
                vec(w.base + (size_t)r * w.rs, (unsigned)k * w.rs, ls);
            }
        };
        vec(p.emb_table, (unsigned)row_stride_bytes(p.wtype, p.emb), 0);  // row 0; the kernel adds (token - 1) rows
        start = true;
        vec(p.rms_att, (unsigned)p.emb * 4u, (size_t)p.emb * 4u);
        rows(0);
        rows(1);
        start = true;
        vec(p.rms_ffn, (unsigned)p.emb * 4u, (size_t)p.emb * 4u);
        rows(2);
        rows(3);
        start = true;
        vec(p.rms_final, (unsigned)p.emb * 4u, 0);
        rows(4);
    }
    p.sched_stride = cap;
    *out = tab;
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

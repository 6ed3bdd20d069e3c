// kernels.cuh -- host-side launch interface of the sm_100a kernels (ops.cu, stream.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace llmf90 {

// ---------------------------------------------------------------- granular operators (ops.cu)
// All pointers are device pointers; kernels run on `st`.  Each returns the launch error.
cudaError_t launch_rmsnorm(const float *x, const float *w, float *out, int n, cudaStream_t st);
cudaError_t launch_softmax(const float *x, float *p, int n, int s, cudaStream_t st);
// y[r] = (residual ? residual[r] : 0) + dot(W[r,:], x), W in device row format
cudaError_t launch_matvec(const uint8_t *W, int wtype, int rows, int cols, const float *x,
                          const float *residual, float *y, cudaStream_t st);
// y[r] = dot(dequant(W[r,:]), x), W = rows x (cols / 256) ggml Q6_K super-blocks of 210 bytes (file format, no re-layout)
cudaError_t launch_matvec_q6k(const uint8_t *W, int rows, int cols, const float *x, float *y, cudaStream_t st);
// standalone RoPE with on-the-fly trig (llama2.f90:543-559); pos is read from *pos_dev if given
cudaError_t launch_rope(float *q, float *k, int emb, int kv, int hs, int pos, cudaStream_t st);
// rope table: tab[(p*hs/2 + j)] = (cos, sin)((p+1) * 10000^-((2j+1)/hs)),  p = 0..seq-1
cudaError_t launch_rope_table(float2 *tab, int seq, int hs, cudaStream_t st);
// x = dequant(table row token-1); tokpos = {token, pos} on the device
cudaError_t launch_embed(const uint8_t *table, int wtype, int cols, const int *tokpos, float *x,
                         cudaStream_t st);
// rotate q (in place) and k with the table, append k,v at position pos-1 of this layer's cache
cudaError_t launch_rope_kv(float *qkv, int emb, int kv, int hs, const float2 *tab,
                           const int *tokpos, float *kc_layer, float *vc_layer, cudaStream_t st);
// GQA decode attention for one layer: out[h*hs + d] (llama2.f90:574-598)
cudaError_t launch_attention(const float *q, const float *kc_layer, const float *vc_layer,
                             const int *tokpos, float *out, int n_heads, int kv_mul, int hs, int kv,
                             int seq, cudaStream_t st);
// hb[i] = silu(h13[2i]) * h13[2i+1]   (rows interleaved gate/up at upload)
cudaError_t launch_swiglu(const float *h13, float *hb, int hid, cudaStream_t st);
// out[0] = 1-based index of the first maximum (maxloc, llama2.f90:388); also tokpos update
cudaError_t launch_argmax(const float *v, int n, int *out_token, cudaStream_t st);

// out[0] = 1-based index sampled from softmax(v / temperature) by the CDF walk against r (llama2.f90:390-391, :428-447)
cudaError_t launch_sample(const float *v, int n, float temperature, float r, int *out_token, cudaStream_t st);

// ---------------------------------------------------------------- upload helpers (ops.cu)
// dst row r (device format) = src row rowmap(r), columns [col0, col0+ncols) of a host-format
// matrix already copied to the device.  map_kind: 0 identity (+row0), 1 interleave halves
// (dst 2i = src i, dst 2i+1 = src half + i).
cudaError_t launch_repack(const uint8_t *src, int wtype, int src_cols, uint8_t *dst, int dst_rows,
                          int col0, int ncols, int map_kind, int row0, int half, cudaStream_t st);

// same for q4_0 into the tiled mma format of the fused kernel (common.cuh, Q4T_GROUP_BYTES)
cudaError_t launch_repack_q4_tiled(const uint8_t *src, int src_cols, uint8_t *dst, int dst_rows, int col0,
                                   int ncols, int map_kind, int row0, int half, cudaStream_t st);

// ---------------------------------------------------------------- fused streaming kernel (stream.cu)
// fine-grained phase timers of the streaming kernel (CTA 0); _PRO = prologue (activation
// vector load / rmsnorm / attention merge), _MV = ring consumption, _BAR = epilogue + grid barrier
enum {
    PH_QKV_PRO, PH_QKV_MV, PH_ROPE_BAR, PH_ATT, PH_ATT_BAR, PH_WO_PRO, PH_WO_MV, PH_WO_BAR,
    PH_W13_PRO, PH_W13_MV, PH_W13_BAR, PH_W2_PRO, PH_W2_MV, PH_W2_BAR, PH_CLS_PRO, PH_CLS_MV,
    PH_ARGMAX, PH_COUNT
};

struct PhaseW {
    const uint8_t *base;     // layer 0 base, device row format
    unsigned long long layer_stride;  // bytes between layers
    unsigned int rs;         // f32 / f16: bytes between rows; tiled q4_0: bytes of one 16-row group
    int rows, cols;
    int unit;                // rows are assigned to CTAs in multiples of `unit` (2: row pairs; 16: tiled q4_0)
    int rows_real;           // rows of the matrix (rows is padded to a multiple of `unit`)
    // A TILE is R consecutive rows (4 for f32 / f16, one 16-row group for tiled q4_0), consumed by ONE warp.
    // Its contraction range is cut into `nch` chunks; one chunk of one tile is one ring STAGE (R row
    // segments of f32 / f16 weights, or a run of whole 8-block groups of the q4_0 row group).
    int R, nch;
    int nu;                  // f32 / f16: 16-byte weight units per row; q4_0: 8-block groups per row group
    int cu;                  // units (groups) per chunk; the last chunk of a row has the remainder
    int rows_cap;            // max rows of any CTA in this phase
};

constexpr int MAX_TP = 8;

// One ring stage of a CTA's static weight-streaming schedule (built on the host once per engine,
// copied to shared memory at kernel start): the producer warp just walks its CTA's list
//   [embedding row][the stages of ONE layer][final norm vector + classifier stages],
// repeating the layer section L times.  A stage is `nseg` bulk copies of `seg_bytes` each, source
// segments `kind`-specific bytes apart (the rows of a tile), landing back to back in one ring slot.
struct SchedStage {
    unsigned long long src;  // device address of the first segment (layer 0)
    unsigned int seg_bytes;
    unsigned int meta;       // bits 0-7: nseg; bits 8-11: kind (SK_*); bit 31: first stage of a phase
};
constexpr unsigned int SCHED_PHASE_START = 0x80000000u;
// stage kinds: 0-4 = the five streamed matrices (QKV, WO, W13, W2, CLS), then the vector stages
enum { SK_RMS_ATT = 5, SK_RMS_FFN = 6, SK_RMS_FINAL = 7, SK_EMB_ROW = 8, SK_COUNT = 9 };

// All dimensions are THIS GPU's share under tensor parallelism (tp ranks): H / KVH / kv / hid / V /
// nqkv / att_dim are local, emb is the full residual width (the residual stream is replicated).
struct StreamParams {
    int emb, hid, L, H, KVH, V, seq, hs, kv, kv_mul, nqkv, wtype;
    int att_dim;             // H * hs: rows of Wq and contraction length of Wo on this rank
    int tp, rank;            // tensor-parallel size and rank (1, 0 on a single GPU)
    int v_off, v_total;      // this rank's first vocabulary row and the full vocabulary size
    PhaseW ph[5];            // 0 QKV, 1 WO, 2 W13 (interleaved), 3 W2, 4 CLS
    const uint8_t *emb_table;  // [V][rs_emb] device row format
    const float *rms_att, *rms_ffn, *rms_final;
    const float2 *rope_tab;  // [seq][hs/2]
    // Activations between phases travel in "LL" buffers: every float is one 64-bit word
    // {value, epoch} written and read atomically, so a reader that sees the expected epoch has
    // the value -- no grid barrier, no fence (see stream.cu).  Sizes in 64-bit words.
    // Wo / W2 outputs are PARTIAL sums over this rank's share of the contraction: each rank stores
    // its partial vector into every rank's buffer (peer stores over NVLink for the others) and
    // every CTA of every rank adds the tp partials to its copy of the residual stream -- the
    // all-reduce after Wo and after W2 is fused into the hand-over, one NVLink hop, no NCCL call.
    unsigned long long *part1[MAX_TP];  // rank k's [tp][emb] buffer of Wo partials (k = rank: local)
    unsigned long long *part2[MAX_TP];  // same for W2
    unsigned long long *ll_hb;    // [hid]   SwiGLU output
    unsigned long long *ll_q;     // [emb]   rotated query
    unsigned long long *ll_kv;    // [2 kv]  this position's rotated key and value rows
    unsigned long long *ll_att;   // [emb]   attention output (n_splits == 1)
    unsigned long long *ll_part;  // [H][MAX_SPLITS][hs + 4] split partials {m, l, -, -, acc[hs]}
    unsigned long long *amax[MAX_TP];   // rank k's [tp][grid][2] per-CTA {max logit bits, index}
    unsigned long long *done[MAX_TP];   // rank k's [tp][grid] "logits rows stored" flags (tp > 1)
    float *logits[MAX_TP];              // rank k's full [v_total] logits buffer (all-gathered)
    // Every CTA polls every LL vector, i.e. 148 CTAs hammer the same few L2 lines: each vector is kept
    // in ll_rep replicas (producers store to all of them, CTA b polls replica b % ll_rep) to spread
    // the hot spot.  Replica strides in 64-bit words: q / att: att_dim, hb: hid rounded up to even,
    // kv: 2 kv, part1 / part2: tp * emb.
    int ll_rep;
    unsigned int ep_base;         // epoch of layer l of this launch = ep_base + l + 1
    float *kc, *vc;          // [L][seq][kv]
    unsigned long long *phase_cycles;  // [PH_COUNT + 2] SM-cycle accumulators (+ total cycles, total ns)
    const int *tokpos;       // device {token, pos} (1-based); used when token < 0
    int token, pos;          // by-value inputs (token >= 1) -- no H2D copy needed
    int n_splits;            // attention position splits
    unsigned long long *trace;  // optional [grid][128] globaltimer stamps at the phase edges of layer `trace_layer` (or null)
    int trace_layer;
    const SchedStage *sched; // [grid][sched_stride] per-CTA stage lists (entry 0 = embedding row 0: + (token-1) * bytes)
    int sched_stride;        // entries per CTA (padded)
    int pace;                // producer pacing: SM cycles per KB issued (0 = unpaced)
    int pf_lead;             // L2 prefetch distance ahead of the ring cursor, in stages (0 = off)
    int do_argmax;           // fuse maxloc after the classifier and write tokpos = {argmax, pos+1}
    const int *forced;       // optional device array of forced next tokens (prompt), or null
    int *out_tokens;         // optional device array: out_tokens[pos-1] = chosen token
    int *err_flag;           // host-mapped error word: set when a tensor-parallel poll timed out (a peer is gone)
    // ring geometry
    int n_slots, slot_bytes;
    int xs_floats;
    int tile_warps[5];       // consumer warps per tile group, per phase (a divisor of 12)
};

struct StreamPlan {
    int n_slots, slot_bytes, threads, smem_bytes, grid;
    int xs_floats;
    int tile_warps[5];
};

// Decide tile / ring geometry for a model on `grid` CTAs (fills p.ph[i].R / nch / nu / cu / rows_cap);
// returns non-zero if it cannot fit.
int plan_stream(StreamParams &p, int grid, int max_smem_optin, int target_slot_bytes, int max_slots,
                StreamPlan *out);
// the per-CTA stage lists for `grid` CTAs: out has grid * (*stride) entries (call after plan_stream)
void build_schedule(StreamParams &p, int grid, SchedStage **out);  // fills p.sched_stride
cudaError_t prepare_stream_kernel(int wtype, int threads, int smem_bytes);
// prof: the instrumented kernel (phase timers of CTA 0, optional per-CTA trace) instead of the production one
cudaError_t launch_stream(const StreamParams &p, const StreamPlan &plan, bool prof, cudaStream_t st);
// f32 / f16: one layer of streamed matrix `phase` from plain rows (src) to the tile-major order the decode kernel
// streams (dst, same size; not in place); call after plan_stream
cudaError_t launch_tile_pass(const StreamParams &p, int phase, int grid, const uint8_t *src, uint8_t *dst, cudaStream_t st);

// ---------------------------------------------------------------- batched prompt pass (prefill.cu)
// The KV rows of up to prefill_max_positions() prompt positions per call, GEMMs on tcgen05 (see prefill.cu).
struct PrefillDims { int emb, hid, L, H, KVH, hs, kv, nqkv, seq, wtype; };
struct PrefillRun {
    const uint8_t *emb_table;            // device row format (plain rows)
    const float *rms_att, *rms_ffn;      // [L][emb]
    const float2 *rope_tab;              // [seq][hs/2]
    float *kc, *vc;                      // [L][seq][kv]
};
struct Prefill;
// geometry of one GEMM of the pass (host-side arithmetic only: also what llmf90_b200_prefill_plan reports)
struct PrefillGemmGeom {
    int rows, cols;          // weight rows N, contraction length K
    int planes;              // f16 planes of the weights: 1 (f16 storage) or 2 (hi + lo of f32 / q4_0)
    int m_tiles, k_chunks;   // tiles of 128 rows, chunks of 64 contraction elements (both padded with zeros)
    int chunks_per_split, n_splits;  // grid = m_tiles x n_splits; split s takes chunks [s * cps, min(k_chunks, (s + 1) * cps))
    int ppad, tmem_cols;     // positions padded to 16; TMEM columns allocated (power of two >= 32)
    int stages, stage_bytes, smem_bytes;
    unsigned long long weight_bytes;   // operand-order copy of one layer of this matrix
    unsigned long long partial_bytes;  // Y[n_splits][ppad][rows] f32
};
void prefill_gemm_geometry(int rows, int cols, int wtype, int n_sms, int n_pos, PrefillGemmGeom *g);
size_t prefill_weight_bytes(const PrefillDims &d);   // HBM the operand-order copy of the layer matrices takes
cudaError_t prefill_create(Prefill **out, const PrefillDims &d, int n_sms);
void prefill_destroy(Prefill *pf);
int prefill_max_positions();
int *prefill_token_buffer(Prefill *pf);              // device int[prefill_max_positions()]: the input tokens of a pass
// matrix 0 QKV, 1 Wo, 2 W1|W3, 3 W2 of `layer`, from the HOST-format matrix already copied to the device
cudaError_t prefill_pack_weights(Prefill *pf, int matrix, int layer, const uint8_t *src_host_format, cudaStream_t st);
// cache rows of positions pos0 .. pos0 + n - 1 (1-based) of every layer; *launches = kernels enqueued
cudaError_t prefill_run(Prefill *pf, const PrefillRun &r, int n, int pos0, cudaStream_t st, int *launches);
// the tcgen05 GEMM on its own: y[P][N] = x[P][K] . W^T, W in host format on the device (synchronises `st`)
cudaError_t prefill_gemm_op(const uint8_t *w_host_format, int wtype, int N, int K, const float *x, int P, float *y,
                            int n_sms, cudaStream_t st);

}  // namespace llmf90

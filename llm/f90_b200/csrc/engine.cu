// engine.cu -- the C ABI of libllmf90_b200.so (include/llmf90_b200.h): device state, weight
// upload / re-layout, the two forward drivers (fused streaming kernel; granular kernels in a
// CUDA graph) and the operator wrappers.  Host logic only -- every computation is a kernel in
// ops.cu / stream.cu.  There is deliberately no CPU fallback: if CUDA is unavailable every
// entry point fails with an error string.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../../include/llmf90_b200.h"
#include "kernels.cuh"

using namespace llmf90;

namespace {

thread_local std::string g_err;

int fail(const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return 1;
}

#define CK(call)                                                                             \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess)                                                               \
            return fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

constexpr int MAX_SPLITS = 8;

struct Engine {
    bool ready = false;
    llmf90_b200_config cfg{};
    int hs = 0, kv = 0, nqkv = 0, kv_mul = 0, hid_l = 0, att_dim = 0, v_l = 0;  // this rank's share
    // peer-visible buffer (see init) and the mapped buffers of the other tensor-parallel ranks
    uint8_t *d_shared = nullptr, *peer[MAX_TP] = {};
    size_t sh_part1 = 0, sh_part2 = 0, sh_amax = 0, sh_done = 0, sh_logits = 0, sh_bytes = 0;
    bool peers_ready = false;
    int n_sms = 0;
    cudaStream_t st = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
    // weights (device row format)
    uint8_t *d_emb = nullptr, *d_wqkv = nullptr, *d_wo = nullptr, *d_w13 = nullptr, *d_w2 = nullptr,
            *d_wcls = nullptr;
    float *d_rms_att = nullptr, *d_rms_ffn = nullptr, *d_rms_final = nullptr;
    float2 *d_rope = nullptr;
    // RunState + activations
    float *d_kc = nullptr, *d_vc = nullptr;
    float *d_x = nullptr, *d_xb = nullptr, *d_qkv = nullptr, *d_att = nullptr, *d_att_part = nullptr,
          *d_h13 = nullptr, *d_hb = nullptr;
    unsigned long long *d_times = nullptr;  // [PH_COUNT + 2]
    int *d_tokpos = nullptr, *d_forced = nullptr, *d_out_tokens = nullptr, *d_amax = nullptr;
    unsigned long long *d_ll = nullptr;  // all LL buffers of the fused kernel, one allocation
    SchedStage *d_sched = nullptr;       // per-CTA stage lists of the producer warps
    unsigned int launch_seq = 0;
    size_t ll_words = 0;
    float *h_logits = nullptr;  // pinned staging buffer (used when the caller's buffer cannot be registered)
    float *reg_logits = nullptr;  // the caller's logits array, page-locked in place so the D2H copy lands in it directly
    int *h_tokpos = nullptr;    // pinned
    int *h_err = nullptr, *d_err = nullptr;  // host-mapped error word the kernel sets when a tensor-parallel poll times out
    // drivers
    bool use_stream = true;
    bool cls_q6k = false;  // wcls is ggml Q6_K super-blocks (LLMF90_FLAG_CLS_Q6K): granular engine, Q6_K classifier kernel
    bool prof = false;  // instrumented fused kernel (LLMF90_FLAG_PROFILE / LLMF90_PROFILE=1)
    StreamParams sp{};
    StreamPlan plan{};
    cudaGraphExec_t graph = nullptr;
    Prefill *pf = nullptr;  // batched prompt pass (LLMF90_FLAG_PREFILL), prefill.cu
    int graph_kernels = 0;
    // stats
    uint64_t launches = 0, forwards = 0, weight_bytes = 0, active_bytes = 0;
    float last_ms = 0.f, loop_total_ms = 0.f, loop_after_first_ms = 0.f;
    float host_times[5] = {0, 0, 0, 0, 0};
};

Engine E;

template <typename T>
cudaError_t dalloc(T **p, size_t n)
{
    return cudaMalloc((void **)p, n * sizeof(T) > 0 ? n * sizeof(T) : 1);
}

void release_all()
{
    if (E.graph) cudaGraphExecDestroy(E.graph);
    prefill_destroy(E.pf);
    void *ptrs[] = {E.d_emb, E.d_wqkv, E.d_wo, E.d_w13, E.d_w2, E.d_wcls, E.d_rms_att, E.d_rms_ffn,
                    E.d_rms_final, E.d_rope, E.d_kc, E.d_vc, E.d_x, E.d_xb, E.d_qkv, E.d_att,
                    E.d_att_part, E.d_h13, E.d_hb, (void *)E.d_times, E.d_tokpos, E.d_forced,
                    E.d_out_tokens, E.d_amax, E.d_ll, E.d_sched};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    for (int k = 0; k < MAX_TP; k++)
        if (E.peer[k] && E.peer[k] != E.d_shared) cudaIpcCloseMemHandle(E.peer[k]);
    if (E.d_shared) cudaFree(E.d_shared);
    if (E.reg_logits) cudaHostUnregister(E.reg_logits);
    if (E.h_logits) cudaFreeHost(E.h_logits);
    if (E.h_tokpos) cudaFreeHost(E.h_tokpos);
    if (E.h_err) cudaFreeHost(E.h_err);
    if (E.ev0) cudaEventDestroy(E.ev0);
    if (E.ev1) cudaEventDestroy(E.ev1);
    if (E.ev2) cudaEventDestroy(E.ev2);
    if (E.st) cudaStreamDestroy(E.st);
    E = Engine{};
}

// Copy `src_rows` host-format rows to a staging buffer on the device, then re-lay them out into
// `dst` (device row format): dst row r = src row map(r), columns [col0, col0+ncols).
int upload_matrix(uint8_t *dst, const void *src_host, int wtype, int src_rows, int src_cols,
                  int dst_rows, int col0, int ncols, int map_kind, int row0, int half,
                  uint8_t *stage, size_t stage_bytes, bool tiled = false)
{
    const size_t src_bytes = (size_t)src_rows * host_row_bytes(wtype, src_cols);
    if (src_bytes > stage_bytes) return fail("internal: staging buffer too small");
    if (src_host)  // null: the staging buffer already holds this matrix (another row range of the same upload)
        CK(cudaMemcpyAsync(stage, src_host, src_bytes, cudaMemcpyHostToDevice, E.st));
    if (tiled)  // q4_0 matrices of the fused kernel: tiled mma format (common.cuh)
        CK(launch_repack_q4_tiled(stage, src_cols, dst, dst_rows, col0, ncols, map_kind, row0, half, E.st));
    else
        CK(launch_repack(stage, wtype, src_cols, dst, dst_rows, col0, ncols, map_kind, row0, half, E.st));
    CK(cudaStreamSynchronize(E.st));
    return 0;
}

// point the kernel parameters at every rank's shared buffer (own rank: local allocation)
void bind_peers()
{
    StreamParams &p = E.sp;
    for (int k = 0; k < E.cfg.tp_size; k++) {
        uint8_t *b = E.peer[k] ? E.peer[k] : E.d_shared;  // not yet connected: harmless placeholder
        p.part1[k] = reinterpret_cast<unsigned long long *>(b + E.sh_part1);
        p.part2[k] = reinterpret_cast<unsigned long long *>(b + E.sh_part2);
        p.amax[k] = reinterpret_cast<unsigned long long *>(b + E.sh_amax);
        p.done[k] = reinterpret_cast<unsigned long long *>(b + E.sh_done);
        p.logits[k] = reinterpret_cast<float *>(b + E.sh_logits);
    }
}

float *logits_dev() { return reinterpret_cast<float *>(E.d_shared + E.sh_logits); }

int n_splits_for(int pos)
{
    // a power of two (the kernel splits items with shifts).  Hard rule: runs of <= 256 positions (the score
    // buffer).  Then: an item whose cached positions fit ONE group per warp (11 warps x 512 / head_size
    // positions) has all its K rows in flight before the query arrives and its V rows before the softmax
    // statistics -- every further round is a serial L2 round trip on the layer's critical path -- so split
    // further while that is not the case and (head, split) items still map one-to-one onto CTAs.
    int s = 1;
    while (s < MAX_SPLITS && s * 256 < pos) s *= 2;
    static const int one_round = getenv("LLMF90_ATT_ONE_ROUND") ? atoi(getenv("LLMF90_ATT_ONE_ROUND")) : 1;
    const int per_round = 11 * (512 / E.hs);
    while (one_round && s < MAX_SPLITS && E.sp.H * s * 2 <= E.plan.grid && (pos + s - 1) / s > per_round) s *= 2;
    static const int min_items = getenv("LLMF90_ATT_ITEMS") ? atoi(getenv("LLMF90_ATT_ITEMS")) : 0;
    while (s < MAX_SPLITS && E.sp.H * s * 2 <= min_items && (pos - 1) / (2 * s) >= 16) s *= 2;
    return s;
}

int ensure_device(int dev)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail("no CUDA device available (%s); this library has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (dev < 0 || dev >= n) return fail("device %d out of range (have %d)", dev, n);
    CK(cudaSetDevice(dev));
    return 0;
}

// ------------------------------------------------------------------ granular forward
// One kernel per step of llama2.f90:520-636, captured once into a CUDA graph.  d_tokpos holds
// {token, pos}; kernels that need them read them from there, so the graph is position-agnostic.
int enqueue_granular(bool count)
{
    const auto &c = E.cfg;
    const int emb = c.emb_dim, hid = c.hidden_dim, L = c.n_layers, wt = c.wtype;
    const size_t rs_e = row_stride_bytes(wt, emb), rs_h = row_stride_bytes(wt, hid);
    int k = 0;
    CK(launch_embed(E.d_emb, wt, emb, E.d_tokpos, E.d_x, E.st)); k++;
    for (int l = 0; l < L; l++) {
        float *kc = E.d_kc + (size_t)l * c.seq_len * E.kv, *vc = E.d_vc + (size_t)l * c.seq_len * E.kv;
        CK(launch_rmsnorm(E.d_x, E.d_rms_att + (size_t)l * emb, E.d_xb, emb, E.st)); k++;
        CK(launch_matvec(E.d_wqkv + (size_t)l * E.nqkv * rs_e, wt, E.nqkv, emb, E.d_xb, nullptr,
                         E.d_qkv, E.st)); k++;
        CK(launch_rope_kv(E.d_qkv, emb, E.kv, E.hs, E.d_rope, E.d_tokpos, kc, vc, E.st)); k++;
        CK(launch_attention(E.d_qkv, kc, vc, E.d_tokpos, E.d_att, c.n_heads, E.kv_mul, E.hs, E.kv,
                            c.seq_len, E.st)); k++;
        CK(launch_matvec(E.d_wo + (size_t)l * emb * rs_e, wt, emb, emb, E.d_att, E.d_x, E.d_x, E.st)); k++;
        CK(launch_rmsnorm(E.d_x, E.d_rms_ffn + (size_t)l * emb, E.d_xb, emb, E.st)); k++;
        CK(launch_matvec(E.d_w13 + (size_t)l * 2 * hid * rs_e, wt, 2 * hid, emb, E.d_xb, nullptr,
                         E.d_h13, E.st)); k++;
        CK(launch_swiglu(E.d_h13, E.d_hb, hid, E.st)); k++;
        CK(launch_matvec(E.d_w2 + (size_t)l * emb * rs_h, wt, emb, hid, E.d_hb, E.d_x, E.d_x, E.st)); k++;
    }
    CK(launch_rmsnorm(E.d_x, E.d_rms_final, E.d_xb, emb, E.st)); k++;
    if (E.cls_q6k) CK(launch_matvec_q6k(E.d_wcls, c.vocab_size, emb, E.d_xb, logits_dev(), E.st));
    else CK(launch_matvec(E.d_wcls, wt, c.vocab_size, emb, E.d_xb, nullptr, logits_dev(), E.st));
    k++;
    if (count) E.graph_kernels = k;
    return 0;
}

int build_granular_graph()
{
    // warm-up pass outside capture so that every kernel's attributes are set before capturing
    E.h_tokpos[0] = 1; E.h_tokpos[1] = 1;
    CK(cudaMemcpyAsync(E.d_tokpos, E.h_tokpos, 8, cudaMemcpyHostToDevice, E.st));
    if (enqueue_granular(true)) return 1;
    CK(cudaStreamSynchronize(E.st));
    cudaGraph_t g;
    CK(cudaStreamBeginCapture(E.st, cudaStreamCaptureModeThreadLocal));
    int rc = enqueue_granular(false);
    cudaError_t e = cudaStreamEndCapture(E.st, &g);
    if (rc) return rc;
    if (e != cudaSuccess) return fail("graph capture failed: %s", cudaGetErrorString(e));
    CK(cudaGraphInstantiate(&E.graph, g, 0));
    CK(cudaGraphDestroy(g));
    // the warm-up pass wrote position 1 of the caches; restore the zeroed RunState
    const size_t cache = (size_t)E.cfg.n_layers * E.cfg.seq_len * E.kv;
    CK(cudaMemsetAsync(E.d_kc, 0, cache * 4, E.st));
    CK(cudaMemsetAsync(E.d_vc, 0, cache * 4, E.st));
    return 0;
}

__global__ void advance_kernel(int *tokpos, const int *amax, const int *forced, int *out_tokens)
{
    const int pos = tokpos[1];
    int next = amax[0];
    if (forced && forced[pos - 1] > 0) next = forced[pos - 1];
    if (out_tokens) out_tokens[pos - 1] = next;
    tokpos[0] = next;
    tokpos[1] = pos + 1;
}

// enqueue one forward on the stream.  token > 0: inputs by value; token < 0: read d_tokpos.
int enqueue_forward(int token, int pos, bool device_loop, const int *forced, int *out_tokens)
{
    if (E.use_stream) {
        StreamParams &p = E.sp;
        p.token = device_loop ? -1 : token;
        p.pos = pos;
        p.n_splits = n_splits_for(pos);
        p.do_argmax = device_loop ? 1 : 0;
        p.forced = forced;
        p.out_tokens = out_tokens;
        // The LL epoch (launch counter x (L + 1) + layer) is 32 bits and epoch 0 means "never written": before it
        // can wrap, drain the stream, zero every LL buffer and restart the counter.  Tensor-parallel ranks
        // launch in lock-step (same calls in the same order), so they all get here at the same launch.
        if ((unsigned long long)(E.launch_seq + 2) * (unsigned)(E.cfg.n_layers + 1) >= 0xffffffffull) {
            CK(cudaStreamSynchronize(E.st));
            CK(cudaMemsetAsync(E.d_ll, 0, E.ll_words * 8, E.st));
            CK(cudaMemsetAsync(E.d_shared, 0, E.sh_logits, E.st));  // partials, argmax records, flags (not the logits)
            CK(cudaStreamSynchronize(E.st));
            E.launch_seq = 0;
        }
        E.launch_seq++;
        p.ep_base = E.launch_seq * (unsigned)(E.cfg.n_layers + 1);
        CK(launch_stream(p, E.plan, E.prof || p.trace != nullptr, E.st));
        E.launches += 1;
    } else {
        if (!device_loop) {
            E.h_tokpos[0] = token; E.h_tokpos[1] = pos;
            CK(cudaMemcpyAsync(E.d_tokpos, E.h_tokpos, 8, cudaMemcpyHostToDevice, E.st));
        }
        CK(cudaGraphLaunch(E.graph, E.st));
        E.launches += E.graph_kernels;
        if (device_loop) {
            CK(launch_argmax(logits_dev(), E.cfg.vocab_size, E.d_amax, E.st));
            advance_kernel<<<1, 1, 0, E.st>>>(E.d_tokpos, E.d_amax, forced, out_tokens);
            CK(cudaGetLastError());
            E.launches += 2;
        }
    }
    E.forwards++;
    return 0;
}

// llmf90_b200_transformer page-locks the caller's logits array in place (one registration for the whole token
// loop).  A caller that frees that array leaves a stale registration behind, and a later cudaMemcpy whose host
// range overlaps only part of it fails with "invalid argument": every entry point that copies to or from OTHER
// caller memory drops the registration first (the next transformer call registers again).
void drop_logits_registration()
{
    if (E.reg_logits) { cudaHostUnregister(E.reg_logits); E.reg_logits = nullptr; }
}

// after a stream synchronisation: did a kernel give up waiting for a tensor-parallel peer?
int check_peers()
{
    if (E.h_err && *E.h_err) {
        *E.h_err = 0;
        return fail("a tensor-parallel rank stopped answering (a poll for its partial sums timed out): the forward "
                    "pass is invalid; every rank must make the same calls in the same order after tp_connect");
    }
    return 0;
}

int device_loop(int first_token, int pos0, int n, const int *d_forced, int *d_out, float *ms_after_first,
                float *ms_total)
{
    E.h_tokpos[0] = first_token; E.h_tokpos[1] = pos0;
    CK(cudaMemcpyAsync(E.d_tokpos, E.h_tokpos, 8, cudaMemcpyHostToDevice, E.st));
    CK(cudaEventRecord(E.ev0, E.st));
    for (int i = 0; i < n; i++) {
        if (enqueue_forward(-1, pos0 + i, true, d_forced, d_out)) return 1;
        if (i == 0) CK(cudaEventRecord(E.ev1, E.st));
    }
    CK(cudaEventRecord(E.ev2, E.st));
    CK(cudaStreamSynchronize(E.st));
    if (check_peers()) return 1;
    float a = 0, b = 0;
    CK(cudaEventElapsedTime(&a, E.ev0, E.ev2));
    CK(cudaEventElapsedTime(&b, E.ev1, E.ev2));
    E.loop_total_ms = a;
    E.loop_after_first_ms = b;
    if (ms_total) *ms_total = a;
    if (ms_after_first) *ms_after_first = b;
    return 0;
}

}  // namespace

// helpers of the operator wrappers
namespace {
struct TmpStream {
    cudaStream_t s = nullptr;
    std::vector<void *> bufs;
    ~TmpStream()
    {
        for (void *b : bufs) cudaFree(b);
        if (s) cudaStreamDestroy(s);
    }
};
int op_begin(TmpStream &t)
{
    if (ensure_device(E.ready ? E.cfg.device : 0)) return 1;
    CK(cudaStreamCreate(&t.s));
    return 0;
}
template <typename T>
int op_buf(TmpStream &t, T **p, size_t n, const void *host)
{
    CK(cudaMalloc((void **)p, std::max<size_t>(n * sizeof(T), 16)));
    t.bufs.push_back(*p);
    if (host) CK(cudaMemcpyAsync(*p, host, n * sizeof(T), cudaMemcpyHostToDevice, t.s));
    return 0;
}
// ---- what init and the dry planner share: the checks on a configuration ...
int check_config(const llmf90_b200_config &c, int *hs_out, int *tp_out, int *rank_out)
{
    if (c.emb_dim <= 0 || c.hidden_dim <= 0 || c.n_layers <= 0 || c.n_heads <= 0 || c.n_kv_heads <= 0 ||
        c.vocab_size <= 0 || c.seq_len <= 0)
        return fail("config: non-positive dimension");
    if (c.wtype < 0 || c.wtype > 2) return fail("config: unknown wtype %d", c.wtype);
    if (c.emb_dim % c.n_heads || c.n_heads % c.n_kv_heads) return fail("config: heads do not divide");
    const int hs = c.emb_dim / c.n_heads;
    if (hs != 32 && hs != 64 && hs != 128) return fail("config: head size %d not in {32,64,128}", hs);
    const int colmul = c.wtype == WT_Q4_0 ? 32 : (c.wtype == WT_F16 ? 8 : 4);
    if (c.emb_dim % colmul || c.hidden_dim % colmul)
        return fail("config: emb_dim/hidden_dim must be multiples of %d for this wtype", colmul);
    const int tp = c.tp_size <= 0 ? 1 : c.tp_size, rank = tp == 1 ? 0 : c.tp_rank;
    if (tp != 1 && tp != 2 && tp != 4 && tp != 8) return fail("config: tp_size %d not in {1,2,4,8}", tp);
    if (rank < 0 || rank >= tp) return fail("config: tp_rank %d out of range", rank);
    if (tp > 1) {
        if (c.flags & (LLMF90_FLAG_GRANULAR | LLMF90_FLAG_CLS_Q6K)) return fail("the granular forward (and with it a Q6_K classifier) is single-GPU only");
        // fewer KV heads than ranks: each KV head is replicated on tp / n_kv_heads ranks
        if (c.n_heads % tp || (c.n_kv_heads % tp && tp % c.n_kv_heads))
            return fail("config: n_heads %d must be a multiple of tp_size %d, n_kv_heads %d a multiple or a divisor",
                        c.n_heads, tp, c.n_kv_heads);
        if (c.hidden_dim % (tp * colmul) || (c.emb_dim / tp) % colmul || c.vocab_size % tp)
            return fail("config: hidden_dim / emb_dim / vocab_size do not split %d ways for this wtype", tp);
    }
    if ((c.flags & LLMF90_FLAG_CLS_Q6K) && (c.emb_dim % 256 || c.emb_dim > 12288))
        return fail("config: a Q6_K classifier needs emb_dim to be a multiple of 256 (and at most 12288)");
    *hs_out = hs; *tp_out = tp; *rank_out = rank;
    return 0;
}

// ... and the fused kernel's view of one rank's share of the model: dimensions and the five streamed
// matrices (bases[] = QKV, Wo, W13, W2, classifier in device memory -- or in the planner's virtual
// address space)
void stream_geometry(StreamParams &p, const llmf90_b200_config &c, int hs, int tp, int rank, bool tiled,
                     const uint8_t *const bases[5])
{
    const int emb = c.emb_dim, wt = c.wtype;
    const int Hl = c.n_heads / tp, KVHl = std::max(1, c.n_kv_heads / tp);
    const int att = Hl * hs, kvl = KVHl * hs, nqkv = att + 2 * kvl, hid = c.hidden_dim / tp, Vl = c.vocab_size / tp;
    p = StreamParams{};
    p.emb = emb; p.hid = hid; p.L = c.n_layers; p.H = Hl; p.KVH = KVHl; p.V = Vl;
    p.seq = c.seq_len; p.hs = hs; p.kv = kvl; p.kv_mul = Hl / KVHl /* local heads per local KV head */; p.nqkv = nqkv; p.wtype = wt;
    p.att_dim = att; p.tp = tp; p.rank = rank; p.v_off = rank * Vl; p.v_total = c.vocab_size;
    auto mk = [&](const uint8_t *base, int rows, int cols) {
        PhaseW w{};
        w.base = base; w.rows_real = rows; w.cols = cols;
        // rows go to CTAs in pairs (RoPE and SwiGLU pair rows 2i, 2i + 1; the LL stores are 16-byte pairs)
        w.rows = (rows + 1) & ~1; w.unit = 2;
        w.rs = (unsigned)row_stride_bytes(wt, cols);
        w.layer_stride = (unsigned long long)rows * w.rs;
        if (tiled) {
            // tiled q4_0: rows go to CTAs in row groups of 16; rs = bytes of one row group
            w.rows = (rows + 15) & ~15; w.unit = 16;
            w.rs = (unsigned)(q4t_groups(cols) * Q4T_GROUP_BYTES);
            w.layer_stride = (unsigned long long)q4t_matrix_bytes(rows, cols);
        }
        return w;
    };
    p.ph[0] = mk(bases[0], nqkv, emb);
    p.ph[1] = mk(bases[1], emb, att);
    p.ph[2] = mk(bases[2], 2 * hid, emb);
    p.ph[3] = mk(bases[3], emb, hid);
    p.ph[4] = mk(bases[4], Vl, emb);
}
// every CTA must own W13 rows (the LL hand-over's no-overwrite argument, stream.cu): tiny models run
// on fewer CTAs
int stream_grid(const StreamParams &p, int n_sms, bool tiled)
{
    return std::min(n_sms, tiled ? (2 * p.hid + 15) / 16 : p.hid);
}
// The ring takes what shared memory the activation vector leaves (the kernel has no local-memory traffic
// worth an L1): as many slots as fit, up to 16.
constexpr int STREAM_MAX_SLOTS = 16;
// Stage size the planner aims for: f32 / f16 one whole tile of the emb-column matrices (4 rows), between 16
// and 32 KB -- cutting a tile's columns into several stages costs 30 % (short inner loops: TinyLlama f32
// 1.28 ms per token with 16 KB stages against 0.97 with 32 KB), a slot larger than the tile wastes ring;
// tiled q4_0 16 groups of 16 rows x 256 columns (a whole row group of a 4096-column matrix: Llama-2-7B q4_0
// 2.10 ms per token with 8-group stages, 2.04 with 16).
inline int stream_target_slot(bool tiled, int wtype, int emb)
{
    if (tiled) return 16 * Q4T_GROUP_BYTES;
    if (wtype == WT_F16) {
        // tensor-core tiles of 8 rows, row segments of at most 4096 bytes (2048 columns)
        const int seg = std::min((int)row_stride_bytes(wtype, emb), 4096);
        return std::max(8 * seg, 8192);
    }
    const int tile = 4 * (int)row_stride_bytes(wtype, emb);
    return tile < 16384 ? 16384 : (tile > 32768 ? 32768 : tile);
}
constexpr int STREAM_STATIC_SMEM = 3072;  // the kernel's static shared memory (plan, RoPE row, timers)
}  // namespace

// ---- the batched prompt pass: positions pos0 .. pos0 + n - 1 of host tokens, in passes of prefill_max_positions()
namespace {
int prefill_positions(const int32_t *tokens, int n, int pos0, float *ms)
{
    drop_logits_registration();
    const int maxp = prefill_max_positions();
    const PrefillRun r{E.d_emb, E.d_rms_att, E.d_rms_ffn, E.d_rope, E.d_kc, E.d_vc};
    CK(cudaEventRecord(E.ev0, E.st));
    for (int off = 0; off < n; off += maxp) {
        const int m = std::min(maxp, n - off);
        CK(cudaMemcpyAsync(prefill_token_buffer(E.pf), tokens + off, (size_t)m * 4, cudaMemcpyHostToDevice, E.st));
        int k = 0;
        CK(prefill_run(E.pf, r, m, pos0 + off, E.st, &k));
        E.launches += (uint64_t)k;
    }
    CK(cudaEventRecord(E.ev1, E.st));
    CK(cudaStreamSynchronize(E.st));
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, E.ev0, E.ev1));
    E.host_times[3] += t;
    if (ms) *ms = t;
    return 0;
}
}  // namespace

// ====================================================================== C ABI
extern "C" {

const char *llmf90_b200_last_error(void) { return g_err.c_str(); }

int llmf90_b200_free(void)
{
    if (E.st) cudaStreamSynchronize(E.st);
    release_all();
    return 0;
}

int llmf90_b200_init(const llmf90_b200_config *cfg, const void *tok_emb, const float *rms_att,
                     const void *wqkv, const void *wo, const float *rms_ffn, const void *w13,
                     const void *w2, const float *rms_final, const void *wcls)
{
    if (!cfg) return fail("config is null");
    if (E.ready || E.st) llmf90_b200_free();
    const llmf90_b200_config c = *cfg;
    int hs, tp, rank;
    if (check_config(c, &hs, &tp, &rank)) return 1;
    if (!tok_emb || !rms_att || !wqkv || !wo || !rms_ffn || !w13 || !w2 || !rms_final || !wcls)
        return fail("null weight pointer");
    if (ensure_device(c.device)) return 1;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, c.device));
    if (prop.major < 10)
        return fail("device %d is sm_%d%d; this library is built for sm_100a (B200) only", c.device,
                    prop.major, prop.minor);

    E.cfg = c;
    E.cfg.tp_size = tp; E.cfg.tp_rank = rank;
    E.hs = hs; E.kv_mul = c.n_heads / c.n_kv_heads;
    E.n_sms = prop.multiProcessorCount;
    E.cls_q6k = (c.flags & LLMF90_FLAG_CLS_Q6K) != 0;
    E.use_stream = !(c.flags & LLMF90_FLAG_GRANULAR) && !E.cls_q6k;
    E.prof = (c.flags & LLMF90_FLAG_PROFILE) != 0;
    if (const char *s = getenv("LLMF90_PROFILE")) E.prof = E.prof || atoi(s) != 0;
    const int emb = c.emb_dim, L = c.n_layers, V = c.vocab_size, wt = c.wtype;
    const int hid_full = c.hidden_dim, kv_full = c.n_kv_heads * hs;
    // this rank's share: heads and their KV heads, FFN rows, vocabulary rows
    // (with fewer KV heads than ranks a rank's heads all map to ONE KV head, h / kv_mul, which it holds a
    // copy of: KV head index rank * KVH / tp)
    const int Hl = c.n_heads / tp, KVHl = std::max(1, c.n_kv_heads / tp);
    const int att = Hl * hs, kvl = KVHl * hs, nqkv = att + 2 * kvl, hid = hid_full / tp, Vl = V / tp;
    const int kv_row0 = (rank * c.n_kv_heads / tp) * hs;  // first Wk / Wv row of this rank's KV heads
    E.kv = kvl; E.nqkv = nqkv; E.hid_l = hid; E.att_dim = att; E.v_l = Vl;
    const size_t rs_e = row_stride_bytes(wt, emb), rs_a = row_stride_bytes(wt, att), rs_h = row_stride_bytes(wt, hid);
    // the five streamed matrices of a q4_0 model use the tiled mma format in the fused kernel (the
    // embedding table and the granular path keep plain rows)
    const bool tiled = E.use_stream && wt == WT_Q4_0;
    auto mbytes = [&](int rows, int cols) -> size_t {
        return tiled ? q4t_matrix_bytes(rows, cols) : (size_t)rows * row_stride_bytes(wt, cols);
    };

    CK(cudaStreamCreateWithFlags(&E.st, cudaStreamNonBlocking));
    CK(cudaEventCreate(&E.ev0)); CK(cudaEventCreate(&E.ev1)); CK(cudaEventCreate(&E.ev2));

    // ---- weights (device row format), this rank's shard
    CK(dalloc(&E.d_emb, (size_t)V * rs_e));
    CK(dalloc(&E.d_wqkv, (size_t)L * mbytes(nqkv, emb)));
    CK(dalloc(&E.d_wo, (size_t)L * mbytes(emb, att)));
    CK(dalloc(&E.d_w13, (size_t)L * mbytes(2 * hid, emb)));
    CK(dalloc(&E.d_w2, (size_t)L * mbytes(emb, hid)));
    const size_t cls_bytes = E.cls_q6k ? (size_t)Vl * (emb / 256) * 210 : (size_t)Vl * rs_e;  // classifier bytes in HBM
    CK(dalloc(&E.d_wcls, E.cls_q6k ? cls_bytes : mbytes(Vl, emb)));
    CK(dalloc(&E.d_rms_att, (size_t)L * emb));
    CK(dalloc(&E.d_rms_ffn, (size_t)L * emb));
    CK(dalloc(&E.d_rms_final, (size_t)emb));
    E.weight_bytes = (size_t)V * rs_e + cls_bytes +
                     (size_t)L * ((size_t)nqkv * rs_e + (size_t)emb * rs_a + (size_t)2 * hid * rs_e + (size_t)emb * rs_h) +
                     (size_t)(2 * L + 1) * emb * 4;
    // algorithmic bytes per token on this GPU, host row sizes
    const size_t hb_e = host_row_bytes(wt, emb), hb_hf = host_row_bytes(wt, hid_full);
    E.active_bytes = (size_t)L * ((size_t)(nqkv + 2 * hid) * hb_e + (size_t)emb * host_row_bytes(wt, att) +
                                  (size_t)emb * host_row_bytes(wt, hid) + 2 * (size_t)emb * 4) +
                     (E.cls_q6k ? cls_bytes : (size_t)Vl * hb_e) + (size_t)emb * 4 + hb_e;
    if (E.use_stream) {
        // the fused kernel's plan comes first: the f32 / f16 matrices are laid out in the order it streams them
        int coop = 0, smem_optin = 0;
        CK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c.device));
        CK(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, c.device));
        if (!coop) { release_all(); return fail("device does not support cooperative launch"); }
        const uint8_t *const bases[5] = {E.d_wqkv, E.d_wo, E.d_w13, E.d_w2, E.d_wcls};
        stream_geometry(E.sp, c, hs, tp, rank, tiled, bases);
        int target_slot = stream_target_slot(tiled, wt, emb), max_slots = STREAM_MAX_SLOTS;
        if (const char *s = getenv("LLMF90_SLOT_BYTES")) target_slot = atoi(s);
        if (const char *s = getenv("LLMF90_MAX_SLOTS")) max_slots = atoi(s);
        if (plan_stream(E.sp, stream_grid(E.sp, E.n_sms, tiled), smem_optin - STREAM_STATIC_SMEM, target_slot, max_slots, &E.plan)) {
            release_all();
            return fail("model rows do not fit the shared-memory ring (row stride too large)");
        }
    }
    if (c.flags & LLMF90_FLAG_PREFILL) {
        // a second copy of the four layer matrices, f16 planes in tensor-core operand order (prefill.cu)
        if (tp > 1) { release_all(); return fail("LLMF90_FLAG_PREFILL: the batched prompt pass is single-GPU"); }
        const PrefillDims pd{emb, hid_full, L, c.n_heads, c.n_kv_heads, hs, kv_full, emb + 2 * kv_full, c.seq_len, wt};
        cudaError_t e = prefill_create(&E.pf, pd, E.n_sms);
        if (e != cudaSuccess) { release_all(); return fail("prefill_create: %s", cudaGetErrorString(e)); }
        E.weight_bytes += prefill_weight_bytes(pd);
    }
    {
        const int nqkv_full = emb + 2 * kv_full;
        size_t stage_bytes = std::max({(size_t)V * hb_e, (size_t)2 * hid_full * hb_e, (size_t)emb * hb_hf,
                                       (size_t)nqkv_full * hb_e});
        uint8_t *stage = nullptr, *tmp = nullptr;
        CK(dalloc(&stage, stage_bytes));
        // f32 / f16 matrices of the fused kernel: plain rows -> tile-major (stream.cu: tile_pass_kernel), one layer
        // of one matrix at a time through a scratch copy
        const bool tile_major = E.use_stream && !tiled;
        if (tile_major) {
            size_t tb = 0;
            for (int i = 0; i < 5; i++) tb = std::max(tb, (size_t)E.sp.ph[i].layer_stride);
            CK(dalloc(&tmp, tb));
        }
        auto to_tiles = [&](int phase, uint8_t *layer) -> int {
            if (!tile_major) return 0;
            CK(cudaMemcpyAsync(tmp, layer, (size_t)E.sp.ph[phase].layer_stride, cudaMemcpyDeviceToDevice, E.st));
            CK(launch_tile_pass(E.sp, phase, E.plan.grid, tmp, layer, E.st));
            CK(cudaStreamSynchronize(E.st));
            return 0;
        };
        int rc = 0;
        rc |= upload_matrix(E.d_emb, tok_emb, wt, V, emb, V, 0, emb, 0, 0, 0, stage, stage_bytes);
        if (E.cls_q6k) {  // file-format super-blocks, consumed as they are
            if (cudaMemcpy(E.d_wcls, wcls, cls_bytes, cudaMemcpyHostToDevice) != cudaSuccess) rc |= fail("upload of the Q6_K classifier failed");
        } else
            rc |= upload_matrix(E.d_wcls, wcls, wt, V, emb, Vl, 0, emb, 0, rank * Vl, 0, stage, stage_bytes, tiled);
        if (!rc) rc |= to_tiles(4, E.d_wcls);
        for (int l = 0; l < L && !rc; l++) {
            const uint8_t *s_qkv = (const uint8_t *)wqkv + (size_t)l * nqkv_full * hb_e;
            const uint8_t *s_wo = (const uint8_t *)wo + (size_t)l * emb * hb_e;
            const uint8_t *s_w13 = (const uint8_t *)w13 + (size_t)l * 2 * hid_full * hb_e;
            const uint8_t *s_w2 = (const uint8_t *)w2 + (size_t)l * emb * hb_hf;
            uint8_t *d_qkv = E.d_wqkv + (size_t)l * mbytes(nqkv, emb);
            // Wq rows of this rank's heads | Wk rows | Wv rows of its KV heads (read_ggml.f90:272,286,300):
            // ONE host-to-device copy of the layer's fused matrix, three re-layouts out of the staging buffer
            // (att and kvl are multiples of 32, so the three pieces start on tile boundaries)
            rc |= upload_matrix(d_qkv, s_qkv, wt, nqkv_full, emb, att, 0, emb, 0, rank * att, 0, stage, stage_bytes, tiled);
            rc |= upload_matrix(d_qkv + mbytes(att, emb), nullptr, wt, nqkv_full, emb, kvl, 0, emb, 0,
                                emb + kv_row0, 0, stage, stage_bytes, tiled);
            rc |= upload_matrix(d_qkv + mbytes(att + kvl, emb), nullptr, wt, nqkv_full, emb, kvl, 0, emb, 0,
                                emb + kv_full + kv_row0, 0, stage, stage_bytes, tiled);
            if (!rc) rc |= to_tiles(0, d_qkv);
            auto pf_pack = [&](int m) -> int {  // the staging buffer holds the layer's host-format matrix m
                if (!E.pf || rc) return 0;
                CK(prefill_pack_weights(E.pf, m, l, stage, E.st));
                CK(cudaStreamSynchronize(E.st));
                return 0;
            };
            rc |= pf_pack(0);
            // Wo: all rows, the input columns of this rank's heads
            rc |= upload_matrix(E.d_wo + (size_t)l * mbytes(emb, att), s_wo, wt, emb, emb, emb, rank * att, att, 0, 0, 0,
                                stage, stage_bytes, tiled);
            if (!rc) rc |= to_tiles(1, E.d_wo + (size_t)l * mbytes(emb, att));
            rc |= pf_pack(1);
            // gate/up rows of this rank's FFN slice, interleaved: row 2i = W1 row i, row 2i+1 = W3 row i
            rc |= upload_matrix(E.d_w13 + (size_t)l * mbytes(2 * hid, emb), s_w13, wt, 2 * hid_full, emb, 2 * hid, 0,
                                emb, 1, rank * hid, hid_full, stage, stage_bytes, tiled);
            if (!rc) rc |= to_tiles(2, E.d_w13 + (size_t)l * mbytes(2 * hid, emb));
            rc |= pf_pack(2);
            // W2: all rows, the input columns of this rank's FFN slice
            rc |= upload_matrix(E.d_w2 + (size_t)l * mbytes(emb, hid), s_w2, wt, emb, hid_full, emb, rank * hid, hid, 0, 0,
                                0, stage, stage_bytes, tiled);
            if (!rc) rc |= to_tiles(3, E.d_w2 + (size_t)l * mbytes(emb, hid));
            rc |= pf_pack(3);
        }
        cudaFree(stage);
        if (tmp) cudaFree(tmp);
        if (rc) { release_all(); return 1; }
    }
    CK(cudaMemcpy(E.d_rms_att, rms_att, (size_t)L * emb * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(E.d_rms_ffn, rms_ffn, (size_t)L * emb * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(E.d_rms_final, rms_final, (size_t)emb * 4, cudaMemcpyHostToDevice));

    // ---- RunState, activations
    const size_t cache = (size_t)L * c.seq_len * kvl;
    CK(dalloc(&E.d_kc, cache)); CK(dalloc(&E.d_vc, cache));
    CK(cudaMemsetAsync(E.d_kc, 0, cache * 4, E.st)); CK(cudaMemsetAsync(E.d_vc, 0, cache * 4, E.st));
    CK(dalloc(&E.d_rope, (size_t)c.seq_len * (hs / 2)));
    CK(launch_rope_table(E.d_rope, c.seq_len, hs, E.st));
    CK(dalloc(&E.d_x, (size_t)emb)); CK(dalloc(&E.d_xb, (size_t)emb)); CK(dalloc(&E.d_qkv, (size_t)nqkv));
    CK(dalloc(&E.d_att, (size_t)emb));
    CK(dalloc(&E.d_h13, (size_t)2 * hid)); CK(dalloc(&E.d_hb, (size_t)hid));
    CK(dalloc(&E.d_times, (size_t)(PH_COUNT + 2)));
    CK(cudaMemsetAsync(E.d_times, 0, (PH_COUNT + 2) * 8, E.st));
    CK(dalloc(&E.d_tokpos, (size_t)2)); CK(dalloc(&E.d_forced, (size_t)c.seq_len));
    CK(dalloc(&E.d_out_tokens, (size_t)c.seq_len)); CK(dalloc(&E.d_amax, (size_t)2 * 1024));
    E.launch_seq = 0;
    CK(cudaMallocHost((void **)&E.h_logits, (size_t)V * 4));
    CK(cudaMallocHost((void **)&E.h_tokpos, 64));
    CK(cudaHostAlloc((void **)&E.h_err, 64, cudaHostAllocMapped));
    *E.h_err = 0;
    CK(cudaHostGetDevicePointer((void **)&E.d_err, E.h_err, 0));

    if (E.use_stream) {
        StreamParams &p = E.sp;
        // cycles per KB of the HBM-generating cursor: 46 = an SM's fair share of the measured HBM bandwidth (22.3
        // B/cycle x 148 SMs x 1.965 GHz = 6.5 TB/s).  Asking for more only queues requests in front of the hand-over
        // traffic: TinyLlama f32 0.906 / 0.878 / 0.871 / 0.868 / 0.872 ms per token at 38 / 42 / 46 / 50 / 54, the
        // other configurations within 0.5 % of each other.
        p.pace = 46;
        if (const char *s = getenv("LLMF90_PACE")) p.pace = std::max(0, atoi(s));
        // L2 prefetch distance: ~48 MB over the 148 SMs (a third of L2) of stages ahead of the ring
        p.pf_lead = std::max(1, (int)((48ull << 20) / ((size_t)E.n_sms * E.plan.slot_bytes)));
        if (const char *s = getenv("LLMF90_PF_LEAD")) p.pf_lead = std::max(0, atoi(s));
        p.emb_table = E.d_emb;
        p.rms_att = E.d_rms_att; p.rms_ffn = E.d_rms_ffn; p.rms_final = E.d_rms_final;
        p.rope_tab = E.d_rope;
        {
            // rank-private LL buffers (64-bit words), zero-filled: epoch 0 is never expected
            // replicas multiply the peer stores of a tensor-parallel run (rows x tp x replicas over NVLink):
            // 4 on one GPU, 2 at tp 2, 1 from tp 4 on
            int rep = tp >= 4 ? 1 : (tp == 2 ? 2 : 4);
            if (const char *s = getenv("LLMF90_LL_REP")) rep = std::max(1, std::min(16, atoi(s)));
            while (rep & (rep - 1)) rep &= rep - 1;  // a power of two (the kernel splits work items with shifts)
            p.ll_rep = rep;
            const size_t n_part = (size_t)Hl * MAX_SPLITS * (hs + 4);
            const size_t words = (size_t)rep * (2 * (size_t)att + ((hid + 1) & ~1) + 2 * (size_t)kvl) + n_part + 64;
            E.ll_words = words;
            CK(dalloc(&E.d_ll, words));
            CK(cudaMemsetAsync(E.d_ll, 0, words * 8, E.st));
            unsigned long long *w = E.d_ll;
            p.ll_q = w; w += (size_t)rep * att;
            p.ll_att = w; w += (size_t)rep * att;
            p.ll_hb = w; w += (size_t)rep * ((hid + 1) & ~1);
            p.ll_kv = w; w += (size_t)rep * 2 * kvl;
            p.ll_part = w;
        }
        {
            // buffers the other ranks write into (one allocation, exported over CUDA IPC when tp > 1):
            // Wo / W2 partials, argmax records, "logits stored" flags, the all-gathered logits
            const size_t G = (size_t)E.plan.grid;
            E.sh_part1 = 0;
            E.sh_part2 = E.sh_part1 + (size_t)p.ll_rep * tp * emb * 8;
            E.sh_amax = E.sh_part2 + (size_t)p.ll_rep * tp * emb * 8;
            E.sh_done = E.sh_amax + (size_t)tp * G * 2 * 8;
            E.sh_logits = E.sh_done + (((size_t)tp * G * 8 + 15) & ~(size_t)15);
            E.sh_bytes = E.sh_logits + (size_t)V * 4;
            CK(cudaMalloc((void **)&E.d_shared, E.sh_bytes));
            CK(cudaMemsetAsync(E.d_shared, 0, E.sh_bytes, E.st));
            for (int k = 0; k < MAX_TP; k++) E.peer[k] = nullptr;
            E.peer[rank] = E.d_shared;
            E.peers_ready = (tp == 1);
            bind_peers();
        }
        p.kc = E.d_kc; p.vc = E.d_vc;
        p.err_flag = E.d_err;
        p.phase_cycles = E.d_times; p.tokpos = E.d_tokpos;
        p.n_slots = E.plan.n_slots; p.slot_bytes = E.plan.slot_bytes;
        p.xs_floats = E.plan.xs_floats;
        for (int i = 0; i < 5; i++) p.tile_warps[i] = E.plan.tile_warps[i];
        if (const char *s = getenv("LLMF90_TILE_WARPS")) {
            // one value for all phases, or five comma-separated ones (QKV, Wo, W13, W2, classifier)
            int g[5], n = sscanf(s, "%d,%d,%d,%d,%d", &g[0], &g[1], &g[2], &g[3], &g[4]);
            for (int i = 0; i < 5 && n >= 1; i++) {
                const int v = g[n == 5 ? i : 0];
                if (v >= 1 && v <= 12 && 12 % v == 0) p.tile_warps[i] = v;
            }
        }
        {
            SchedStage *h = nullptr;
            build_schedule(p, E.plan.grid, &h);
            const size_t bytes = (size_t)E.plan.grid * p.sched_stride * sizeof(SchedStage);
            cudaError_t e = cudaMalloc((void **)&E.d_sched, bytes);
            if (e == cudaSuccess) e = cudaMemcpy(E.d_sched, h, bytes, cudaMemcpyHostToDevice);
            free(h);
            CK(e);
            p.sched = E.d_sched;
        }
        CK(prepare_stream_kernel(wt, E.plan.threads, E.plan.smem_bytes));
    } else {
        CK(dalloc(&E.d_att_part, (size_t)c.n_heads * MAX_SPLITS * (hs + 4)));
        CK(cudaMalloc((void **)&E.d_shared, (size_t)V * 4));
        E.sh_logits = 0;
        E.peers_ready = true;
        if (build_granular_graph()) { release_all(); return 1; }
    }
    CK(cudaStreamSynchronize(E.st));
    E.ready = true;
    return 0;
}

int llmf90_b200_transformer(int32_t token, int32_t pos, float *logits)
{
    if (!E.ready) return fail("llmf90_b200_transformer: engine not initialised");
    if (!E.peers_ready) return fail("tensor-parallel engine: call llmf90_b200_tp_connect first");
    if (token < 1 || token > E.cfg.vocab_size) return fail("token %d out of range 1..%d", token, E.cfg.vocab_size);
    if (pos < 1 || pos > E.cfg.seq_len) return fail("pos %d out of range 1..%d", pos, E.cfg.seq_len);
    if (!logits) return fail("logits is null");
    // The host loop passes the same logits array every token (llama2.f90:380): page-lock it in place once, so
    // the device-to-host copy lands in the caller's memory directly (no staging copy, no second memcpy).
    const size_t lbytes = (size_t)E.cfg.vocab_size * 4;
    if (logits != E.reg_logits) {
        if (E.reg_logits) { cudaHostUnregister(E.reg_logits); E.reg_logits = nullptr; }
        if (cudaHostRegister(logits, lbytes, cudaHostRegisterDefault) == cudaSuccess) E.reg_logits = logits;
        else cudaGetLastError();  // not registrable (e.g. read-only mapping): stage through the pinned buffer
    }
    CK(cudaEventRecord(E.ev0, E.st));
    if (enqueue_forward(token, pos, false, nullptr, nullptr)) return 1;
    CK(cudaEventRecord(E.ev1, E.st));
    float *dst = E.reg_logits == logits ? logits : E.h_logits;
    CK(cudaMemcpyAsync(dst, logits_dev(), lbytes, cudaMemcpyDeviceToHost, E.st));
    CK(cudaStreamSynchronize(E.st));
    if (check_peers()) return 1;
    if (dst != logits) memcpy(logits, E.h_logits, lbytes);
    CK(cudaEventElapsedTime(&E.last_ms, E.ev0, E.ev1));
    if (!E.use_stream || !E.prof) E.host_times[3] += E.last_ms;  // no per-phase timers: whole forward in bucket 4
    return 0;
}

int llmf90_b200_transformer_sample(int32_t token, int32_t pos, float temperature, float r, int32_t *next_token)
{
    if (!E.ready) return fail("llmf90_b200_transformer_sample: engine not initialised");
    if (!E.peers_ready) return fail("tensor-parallel engine: call llmf90_b200_tp_connect first");
    if (token < 1 || token > E.cfg.vocab_size) return fail("token %d out of range 1..%d", token, E.cfg.vocab_size);
    if (pos < 1 || pos > E.cfg.seq_len) return fail("pos %d out of range 1..%d", pos, E.cfg.seq_len);
    if (!next_token || !(temperature >= 0.f)) return fail("transformer_sample: bad argument");
    CK(cudaEventRecord(E.ev0, E.st));
    if (enqueue_forward(token, pos, false, nullptr, nullptr)) return 1;
    CK(cudaEventRecord(E.ev1, E.st));
    // the pick is made next to the logits: maxloc for temperature 0 (llama2.f90:388), else softmax(logits / T) and
    // the CDF walk against r (:390-391, :428-447); 4 bytes come back instead of the vocabulary's logits
    if (temperature == 0.f) CK(launch_argmax(logits_dev(), E.cfg.vocab_size, E.d_amax, E.st));
    else CK(launch_sample(logits_dev(), E.cfg.vocab_size, temperature, r, E.d_amax, E.st));
    E.launches += 1;
    CK(cudaMemcpyAsync(E.h_tokpos + 8, E.d_amax, 4, cudaMemcpyDeviceToHost, E.st));
    CK(cudaStreamSynchronize(E.st));
    if (check_peers()) return 1;
    *next_token = E.h_tokpos[8];
    CK(cudaEventElapsedTime(&E.last_ms, E.ev0, E.ev1));
    if (!E.use_stream || !E.prof) E.host_times[3] += E.last_ms;
    return 0;
}

int llmf90_b200_debug_trace(int32_t token, int32_t pos, int32_t layer, uint64_t *out, int32_t n_ctas)
{
    if (!E.ready || !E.use_stream) return fail("debug_trace: needs the fused streaming engine");
    if (!E.peers_ready) return fail("tensor-parallel engine: call llmf90_b200_tp_connect first");
    if (!out || n_ctas < E.plan.grid) return fail("debug_trace: buffer must hold %d x 128 entries", E.plan.grid);
    drop_logits_registration();
    unsigned long long *d = nullptr;
    CK(cudaMalloc((void **)&d, (size_t)E.plan.grid * 128 * 8));
    CK(cudaMemset(d, 0, (size_t)E.plan.grid * 128 * 8));
    E.sp.trace = d; E.sp.trace_layer = layer;
    int rc = enqueue_forward(token, pos, false, nullptr, nullptr);
    E.sp.trace = nullptr;
    if (!rc) {
        cudaError_t e = cudaStreamSynchronize(E.st);
        if (e == cudaSuccess) e = cudaMemcpy(out, d, (size_t)E.plan.grid * 128 * 8, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = fail("debug_trace: %s", cudaGetErrorString(e));
    }
    cudaFree(d);
    return rc;
}

int llmf90_b200_phase_times(float *ms, int32_t n)
{
    if (!E.ready) return fail("engine not initialised");
    if (!ms || n < 1) return fail("phase_times: bad argument");
    unsigned long long c[PH_COUNT + 2];
    drop_logits_registration();
    CK(cudaStreamSynchronize(E.st));
    CK(cudaMemcpy(c, E.d_times, sizeof c, cudaMemcpyDeviceToHost));
    // cycles -> ms with the kernel's own (globaltimer ns / clock64 cycles) ratio
    const double ns_per_cycle = c[PH_COUNT] ? (double)c[PH_COUNT + 1] / (double)c[PH_COUNT] : 0.0;
    for (int i = 0; i < n; i++) ms[i] = i < PH_COUNT ? (float)((double)c[i] * ns_per_cycle * 1e-6) : 0.f;
    return 0;
}

int llmf90_b200_times(float t[5])
{
    if (!E.ready) return fail("engine not initialised");
    float ph[PH_COUNT];
    if (llmf90_b200_phase_times(ph, PH_COUNT)) return 1;
    // the reference's five buckets (llama2.f90:526-638): 1 rmsnorm+QKV, 2 RoPE+KV append,
    // 3 attention, 4 Wo+FFN, 5 final norm+classifier
    float d[5] = {ph[PH_QKV_PRO] + ph[PH_QKV_MV], ph[PH_ROPE_BAR], ph[PH_ATT] + ph[PH_ATT_BAR], 0.f,
                  ph[PH_CLS_PRO] + ph[PH_CLS_MV] + ph[PH_ARGMAX]};
    for (int i = PH_WO_PRO; i <= PH_W2_BAR; i++) d[3] += ph[i];
    for (int i = 0; i < 5; i++) t[i] = d[i] + E.host_times[i];
    return 0;
}

int llmf90_b200_reset(void)
{
    if (!E.ready) return fail("engine not initialised");
    const size_t cache = (size_t)E.cfg.n_layers * E.cfg.seq_len * E.kv;
    CK(cudaMemsetAsync(E.d_kc, 0, cache * 4, E.st));
    CK(cudaMemsetAsync(E.d_vc, 0, cache * 4, E.st));
    CK(cudaMemsetAsync(E.d_times, 0, (PH_COUNT + 2) * 8, E.st));
    CK(cudaStreamSynchronize(E.st));
    for (float &h : E.host_times) h = 0;
    E.launches = 0; E.forwards = 0;
    return 0;
}

int llmf90_b200_generate_greedy(const int32_t *prompt_tokens, int32_t n_prompt, int32_t n,
                                int32_t *out_tokens, float *elapsed_ms)
{
    if (!E.ready) return fail("engine not initialised");
    if (!E.peers_ready) return fail("tensor-parallel engine: call llmf90_b200_tp_connect first");
    if (n < 1 || n > E.cfg.seq_len) return fail("n %d out of range 1..%d", n, E.cfg.seq_len);
    if (n_prompt < 0 || (n_prompt > 0 && !prompt_tokens)) return fail("bad prompt");
    drop_logits_registration();
    std::vector<int> forced(E.cfg.seq_len, 0);
    for (int i = 0; i < n_prompt && i < n; i++) {
        if (prompt_tokens[i] < 1 || prompt_tokens[i] > E.cfg.vocab_size) return fail("prompt token out of range");
        forced[i] = prompt_tokens[i];
    }
    CK(cudaMemcpyAsync(E.d_forced, forced.data(), (size_t)E.cfg.seq_len * 4, cudaMemcpyHostToDevice, E.st));
    CK(cudaStreamSynchronize(E.st));
    // With the batched prompt pass the forced positions (inputs BOS, prompt[0 .. m-2]; llama2.f90:376-385) are one
    // pass over the weights; the loop then starts at the first position whose logits are used.  At least one
    // position is left to the loop.
    const int m = E.pf ? std::min(n_prompt, n - 1) : 0;
    if (m >= 1) {
        std::vector<int32_t> in(m);
        in[0] = 2;  // BOS
        for (int i = 1; i < m; i++) in[i] = prompt_tokens[i - 1];
        float pf_ms = 0.f, loop_ms = 0.f;
        if (prefill_positions(in.data(), m, 1, &pf_ms)) return 1;
        if (device_loop(prompt_tokens[m - 1], m + 1, n - m, E.d_forced, E.d_out_tokens, nullptr, &loop_ms)) return 1;
        CK(cudaMemcpy(out_tokens, E.d_out_tokens, (size_t)n * 4, cudaMemcpyDeviceToHost));
        for (int i = 0; i < m; i++) out_tokens[i] = prompt_tokens[i];
        E.loop_total_ms = pf_ms + loop_ms;  // the stats describe the whole generation: the prompt pass and the loop
        E.loop_after_first_ms = pf_ms + loop_ms;
        if (elapsed_ms) *elapsed_ms = pf_ms + loop_ms;  // every position is inside: there is no separate first token to leave out
        return 0;
    }
    float after_first = 0.f;
    if (device_loop(2 /* BOS, llama2.f90:376 */, 1, n, E.d_forced, E.d_out_tokens, &after_first, nullptr))
        return 1;
    CK(cudaMemcpy(out_tokens, E.d_out_tokens, (size_t)n * 4, cudaMemcpyDeviceToHost));
    if (elapsed_ms) *elapsed_ms = after_first;
    return 0;
}

// ---------------------------------------------------------------- batched prompt pass (prefill.cu)

int llmf90_b200_prefill(const int32_t *tokens, int32_t n_tokens, int32_t pos0)
{
    if (!E.ready) return fail("llmf90_b200_prefill: engine not initialised");
    if (!E.pf) return fail("llmf90_b200_prefill: initialise the engine with LLMF90_FLAG_PREFILL");
    if (!tokens || n_tokens < 1) return fail("prefill: no tokens");
    if (pos0 < 1 || pos0 + n_tokens - 1 > E.cfg.seq_len)
        return fail("prefill: positions %d..%d out of range 1..%d", pos0, pos0 + n_tokens - 1, E.cfg.seq_len);
    if ((size_t)(pos0 + n_tokens) * 4 > 200 * 1024) return fail("prefill: more than 51199 positions of attention scores");
    for (int i = 0; i < n_tokens; i++)
        if (tokens[i] < 1 || tokens[i] > E.cfg.vocab_size) return fail("prefill: token %d out of range 1..%d", tokens[i], E.cfg.vocab_size);
    return prefill_positions(tokens, n_tokens, pos0, nullptr);
}

int llmf90_b200_debug_read_kv(int32_t layer, int32_t pos, float *k, float *v)
{
    if (!E.ready) return fail("engine not initialised");
    if (layer < 0 || layer >= E.cfg.n_layers || pos < 1 || pos > E.cfg.seq_len || !k || !v) return fail("read_kv: bad argument");
    drop_logits_registration();
    CK(cudaStreamSynchronize(E.st));
    const size_t off = ((size_t)layer * E.cfg.seq_len + (size_t)(pos - 1)) * E.kv;
    CK(cudaMemcpy(k, E.d_kc + off, (size_t)E.kv * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(v, E.d_vc + off, (size_t)E.kv * 4, cudaMemcpyDeviceToHost));
    return 0;
}

int llmf90_b200_bench_device_loop(int32_t first_token, int32_t pos0, int32_t n_steps, float *elapsed_ms)
{
    if (!E.ready) return fail("engine not initialised");
    if (!E.peers_ready) return fail("tensor-parallel engine: call llmf90_b200_tp_connect first");
    if (pos0 < 1 || n_steps < 1 || pos0 + n_steps - 1 > E.cfg.seq_len) return fail("bad position range");
    if (first_token < 1 || first_token > E.cfg.vocab_size) return fail("bad token");
    float total = 0.f;
    if (device_loop(first_token, pos0, n_steps, nullptr, nullptr, nullptr, &total)) return 1;
    if (elapsed_ms) *elapsed_ms = total;
    return 0;
}

int llmf90_b200_get_stats(llmf90_b200_stats *out)
{
    if (!out) return fail("null");
    memset(out, 0, sizeof *out);
    if (!E.ready) return fail("engine not initialised");
    out->kernel_launches = E.launches;
    out->forward_calls = E.forwards;
    out->weight_bytes_device = E.weight_bytes;
    out->active_bytes_per_token = E.active_bytes;
    out->last_forward_ms = E.last_ms;
    out->n_sms = E.n_sms;
    out->last_loop_total_ms = E.loop_total_ms;
    out->last_loop_after_first_ms = E.loop_after_first_ms;
    if (E.use_stream) {
        out->stream_slots = E.plan.n_slots; out->stream_slot_bytes = E.plan.slot_bytes;
        out->stream_smem_bytes = E.plan.smem_bytes; out->stream_threads = E.plan.threads;
    }
    return 0;
}

// ---------------------------------------------------------------- the planner, without a device
int llmf90_b200_plan(const llmf90_b200_config *cfg, int32_t n_sms, int32_t smem_optin, llmf90_b200_plan_info *info,
                     llmf90_b200_sched_stage *sched, int64_t sched_entries)
{
    if (!cfg || !info) return fail("plan: null argument");
    if (n_sms <= 0 || smem_optin <= STREAM_STATIC_SMEM) return fail("plan: bad device description");
    int hs, tp, rank;
    if (check_config(*cfg, &hs, &tp, &rank)) return 1;
    if (cfg->flags & (LLMF90_FLAG_GRANULAR | LLMF90_FLAG_CLS_Q6K)) return fail("plan: the granular forward has no ring plan");
    const bool tiled = cfg->wtype == WT_Q4_0;
    const uint8_t *bases[5];
    for (int i = 0; i < 5; i++) bases[i] = reinterpret_cast<const uint8_t *>(LLMF90_PLAN_VBASE(i));
    StreamParams p;
    stream_geometry(p, *cfg, hs, tp, rank, tiled, bases);
    p.emb_table = reinterpret_cast<const uint8_t *>(LLMF90_PLAN_VBASE(5));
    p.rms_att = reinterpret_cast<const float *>(LLMF90_PLAN_VBASE(6));
    p.rms_ffn = reinterpret_cast<const float *>(LLMF90_PLAN_VBASE(7));
    p.rms_final = reinterpret_cast<const float *>(LLMF90_PLAN_VBASE(8));
    StreamPlan plan{};
    if (plan_stream(p, stream_grid(p, n_sms, tiled), smem_optin - STREAM_STATIC_SMEM, stream_target_slot(tiled, cfg->wtype, cfg->emb_dim),
                    STREAM_MAX_SLOTS, &plan))
        return fail("model rows do not fit the shared-memory ring (row stride too large)");
    SchedStage *h = nullptr;
    build_schedule(p, plan.grid, &h);
    // the report lists BULK COPIES: a stage of nseg row segments becomes nseg entries
    const unsigned long long lstride[SK_COUNT] = {p.ph[0].layer_stride, p.ph[1].layer_stride, p.ph[2].layer_stride,
                                                  p.ph[3].layer_stride, 0ull, (unsigned long long)p.emb * 4u,
                                                  (unsigned long long)p.emb * 4u, 0ull, 0ull};
    int copies_max = 0;
    for (int cta = 0; cta < plan.grid; cta++) {
        int n = 0;
        for (int i = 0; i < p.sched_stride; i++) n += (int)(h[(size_t)cta * p.sched_stride + i].meta & 0xffu);
        copies_max = std::max(copies_max, n);
    }
    memset(info, 0, sizeof *info);
    info->grid = plan.grid; info->threads = plan.threads; info->n_slots = plan.n_slots;
    info->slot_bytes = plan.slot_bytes; info->smem_bytes = plan.smem_bytes + STREAM_STATIC_SMEM;
    info->sched_stride = copies_max; info->n_layers = p.L;
    for (int i = 0; i < 5; i++) {
        info->rows[i] = p.ph[i].rows_real; info->cols[i] = p.ph[i].cols;
        info->matrix_bytes[i] = tiled ? (uint64_t)q4t_matrix_bytes(p.ph[i].rows_real, p.ph[i].cols)
                                      : (uint64_t)p.ph[i].rows_real * p.ph[i].rs;
        info->tile_rows[i] = p.ph[i].R; info->tile_chunks[i] = p.ph[i].nch; info->tile_warps[i] = plan.tile_warps[i];
    }
    info->vector_bytes = (uint64_t)p.emb * 4u;
    info->emb_row_bytes = (uint64_t)row_stride_bytes(p.wtype, p.emb);
    const int64_t need = (int64_t)plan.grid * copies_max;
    int rc = 0;
    if (sched) {
        if (sched_entries < need) rc = fail("plan: schedule buffer too small (%lld entries needed)", (long long)need);
        else {
            memset(sched, 0, (size_t)need * sizeof *sched);
            for (int cta = 0; cta < plan.grid; cta++) {
                llmf90_b200_sched_stage *o = sched + (size_t)cta * copies_max;
                uint32_t stage = 0;
                for (int i = 0; i < p.sched_stride; i++) {
                    const SchedStage &e = h[(size_t)cta * p.sched_stride + i];
                    const unsigned nseg = e.meta & 0xffu, kind = (e.meta >> 8) & 0xfu;
                    if (!nseg) continue;
                    const unsigned sstride = kind < 5 ? p.ph[kind].rs : 0u;
                    for (unsigned k = 0; k < nseg; k++, o++) {
                        o->src = e.src + (uint64_t)k * sstride; o->bytes = e.seg_bytes;
                        o->layer_stride16 = (uint32_t)(lstride[kind] >> 4);
                        o->phase_start = (k == 0 && (e.meta & SCHED_PHASE_START)) ? 1u : 0u;
                        o->stage = stage;
                    }
                    stage++;
                }
            }
        }
    }
    free(h);
    return rc;
}

int llmf90_b200_prefill_plan(const llmf90_b200_config *cfg, int32_t n_sms, int32_t n_pos, llmf90_b200_prefill_gemm out[4])
{
    if (!cfg || !out) return fail("prefill_plan: null argument");
    if (n_sms <= 0 || n_pos < 1 || n_pos > prefill_max_positions()) return fail("prefill_plan: n_sms > 0 and n_pos in 1..%d", prefill_max_positions());
    int hs, tp, rank;
    if (check_config(*cfg, &hs, &tp, &rank)) return 1;
    if (tp > 1) return fail("prefill_plan: the batched prompt pass is single-GPU");
    const int emb = cfg->emb_dim, hid = cfg->hidden_dim, kv = cfg->n_kv_heads * hs;
    const int N[4] = {emb + 2 * kv, emb, 2 * hid, emb}, K[4] = {emb, emb, emb, hid};
    for (int i = 0; i < 4; i++) {
        PrefillGemmGeom g;
        prefill_gemm_geometry(N[i], K[i], cfg->wtype, n_sms, n_pos, &g);
        out[i] = {g.rows, g.cols, g.planes, g.m_tiles, g.k_chunks, g.chunks_per_split, g.n_splits, g.ppad, g.tmem_cols,
                  g.stages, g.stage_bytes, g.smem_bytes, g.weight_bytes, g.partial_bytes};
    }
    return 0;
}

// ---------------------------------------------------------------- operator wrappers

int llmf90_b200_matvec(const void *w, int32_t wtype, int32_t rows, int32_t cols, const float *x, float *y)
{
    if (!w || !x || !y || rows <= 0 || cols <= 0) return fail("matvec: bad argument");
    if (wtype == LLMF90_WTYPE_Q6_K) {  // ggml super-blocks straight from the file
        if (cols % 256 || cols > 12288) return fail("matvec: Q6_K needs cols to be a multiple of 256, at most 12288");
        TmpStream t;
        if (op_begin(t)) return 1;
        uint8_t *d_w; float *d_x, *d_y;
        if (op_buf(t, &d_w, (size_t)rows * (cols / 256) * 210, w)) return 1;
        if (op_buf(t, &d_x, (size_t)cols, x)) return 1;
        if (op_buf(t, &d_y, (size_t)rows, nullptr)) return 1;
        CK(launch_matvec_q6k(d_w, rows, cols, d_x, d_y, t.s));
        CK(cudaMemcpyAsync(y, d_y, (size_t)rows * 4, cudaMemcpyDeviceToHost, t.s));
        CK(cudaStreamSynchronize(t.s));
        return 0;
    }
    if (wtype < 0 || wtype > 2) return fail("matvec: unknown wtype");
    const int colmul = wtype == WT_Q4_0 ? 32 : (wtype == WT_F16 ? 8 : 4);
    if (cols % colmul) return fail("matvec: cols must be a multiple of %d", colmul);
    TmpStream t;
    if (op_begin(t)) return 1;
    uint8_t *d_src, *d_w; float *d_x, *d_y;
    if (op_buf(t, &d_src, (size_t)rows * host_row_bytes(wtype, cols), w)) return 1;
    if (op_buf(t, &d_w, (size_t)rows * row_stride_bytes(wtype, cols), nullptr)) return 1;
    if (op_buf(t, &d_x, (size_t)cols, x)) return 1;
    if (op_buf(t, &d_y, (size_t)rows, nullptr)) return 1;
    CK(launch_repack(d_src, wtype, cols, d_w, rows, 0, cols, 0, 0, 0, t.s));
    CK(launch_matvec(d_w, wtype, rows, cols, d_x, nullptr, d_y, t.s));
    CK(cudaMemcpyAsync(y, d_y, (size_t)rows * 4, cudaMemcpyDeviceToHost, t.s));
    CK(cudaStreamSynchronize(t.s));
    return 0;
}

int llmf90_b200_matmul(const void *w, int32_t wtype, int32_t rows, int32_t cols, const float *x, int32_t n_pos, float *y)
{
    if (!w || !x || !y || rows <= 0 || cols <= 0) return fail("matmul: bad argument");
    if (wtype < 0 || wtype > 2) return fail("matmul: unknown wtype");
    if (n_pos < 1 || n_pos > prefill_max_positions()) return fail("matmul: n_pos must be 1..%d", prefill_max_positions());
    const int colmul = wtype == WT_Q4_0 ? 32 : (wtype == WT_F16 ? 8 : 4);
    if (cols % colmul) return fail("matmul: cols must be a multiple of %d", colmul);
    TmpStream t;
    if (op_begin(t)) return 1;
    int dev = 0, n_sms = 148;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev));
    uint8_t *d_w; float *d_x, *d_y;
    if (op_buf(t, &d_w, (size_t)rows * host_row_bytes(wtype, cols), w)) return 1;
    if (op_buf(t, &d_x, (size_t)n_pos * cols, x)) return 1;
    if (op_buf(t, &d_y, (size_t)n_pos * rows, nullptr)) return 1;
    CK(prefill_gemm_op(d_w, wtype, rows, cols, d_x, n_pos, d_y, n_sms, t.s));
    CK(cudaMemcpyAsync(y, d_y, (size_t)n_pos * rows * 4, cudaMemcpyDeviceToHost, t.s));
    CK(cudaStreamSynchronize(t.s));
    return 0;
}

int llmf90_b200_rmsnorm(const float *x, const float *w, int32_t n, float *out)
{
    if (!x || !w || !out || n <= 0) return fail("rmsnorm: bad argument");
    TmpStream t;
    if (op_begin(t)) return 1;
    float *d_x, *d_w, *d_o;
    if (op_buf(t, &d_x, (size_t)n, x) || op_buf(t, &d_w, (size_t)n, w) || op_buf(t, &d_o, (size_t)n, nullptr)) return 1;
    CK(launch_rmsnorm(d_x, d_w, d_o, n, t.s));
    CK(cudaMemcpyAsync(out, d_o, (size_t)n * 4, cudaMemcpyDeviceToHost, t.s));
    CK(cudaStreamSynchronize(t.s));
    return 0;
}

int llmf90_b200_softmax(const float *x, int32_t n, int32_t s, float *p)
{
    if (!x || !p || n <= 0 || s <= 0 || s > n) return fail("softmax: bad argument");
    TmpStream t;
    if (op_begin(t)) return 1;
    float *d_x, *d_p;
    if (op_buf(t, &d_x, (size_t)n, x) || op_buf(t, &d_p, (size_t)n, nullptr)) return 1;
    CK(launch_softmax(d_x, d_p, n, s, t.s));
    CK(cudaMemcpyAsync(p, d_p, (size_t)n * 4, cudaMemcpyDeviceToHost, t.s));
    CK(cudaStreamSynchronize(t.s));
    return 0;
}

int llmf90_b200_rope(float *q, float *k, int32_t emb, int32_t kv, int32_t head_size, int32_t pos)
{
    if (!q || !k || emb <= 0 || kv <= 0 || kv > emb || head_size <= 0 || (head_size & 1) || emb % head_size ||
        kv % head_size || pos < 1)
        return fail("rope: bad argument");
    TmpStream t;
    if (op_begin(t)) return 1;
    float *d_q, *d_k;
    if (op_buf(t, &d_q, (size_t)emb, q) || op_buf(t, &d_k, (size_t)kv, k)) return 1;
    CK(launch_rope(d_q, d_k, emb, kv, head_size, pos, t.s));
    CK(cudaMemcpyAsync(q, d_q, (size_t)emb * 4, cudaMemcpyDeviceToHost, t.s));
    CK(cudaMemcpyAsync(k, d_k, (size_t)kv * 4, cudaMemcpyDeviceToHost, t.s));
    CK(cudaStreamSynchronize(t.s));
    return 0;
}

int llmf90_b200_tp_export(void *handle64)
{
    if (!E.ready || !E.use_stream) return fail("tp_export: engine not initialised");
    if (!handle64) return fail("tp_export: null");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, E.d_shared));
    memcpy(handle64, &h, 64);
    return 0;
}

int llmf90_b200_tp_connect(const void *handles, int32_t n)
{
    if (!E.ready || !E.use_stream) return fail("tp_connect: engine not initialised");
    if (!handles || n != E.cfg.tp_size) return fail("tp_connect: expected %d handles", E.cfg.tp_size);
    for (int k = 0; k < n; k++) {
        if (k == E.cfg.tp_rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const uint8_t *)handles + (size_t)k * 64, 64);
        void *ptr = nullptr;
        CK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        E.peer[k] = (uint8_t *)ptr;
    }
    bind_peers();
    // Every rank restarts its launch counter (the LL epochs derive from it) and clears its hand-over buffers
    // here: launches made before the connection (warm-ups, a debug trace on one rank) no longer matter, the
    // ranks are in step from this call on as long as they make the same calls in the same order.
    CK(cudaStreamSynchronize(E.st));
    CK(cudaMemsetAsync(E.d_ll, 0, E.ll_words * 8, E.st));
    CK(cudaMemsetAsync(E.d_shared, 0, E.sh_logits, E.st));
    CK(cudaStreamSynchronize(E.st));
    E.launch_seq = 0;
    E.peers_ready = true;
    return 0;
}

}  // extern "C"

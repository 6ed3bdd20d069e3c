/*
 * llmf90_b200.h -- C ABI of the B200-native decode engine that replaces the
 * forward pass of rbitr/llm.f90.
 *
 * Drop-in boundary: the single call site
 *     logits = transformer(token, pos, s, weights)        (llama2.f90:380)
 * and the inner subroutines it uses (rmsnorm llama2.f90:450-457, softmax
 * :468-478, the inline mat-vec loops :529-531/:603-605/:610-612/:618-620/
 * :634-636, RoPE :543-559).  Everything above that call (CLI, GGUF loader,
 * tokenizer, sampler, token loop) stays on the host.  The reference has no FFI
 * today; these are the entry points an ISO_C_BINDING interface block binds
 * (fortran/llmf90_b200_iface.f90), following the author's own
 * bind(C) convention (load.f90:123-152: scalar `value` args, c_float/c_int).
 *
 * Conventions
 *  - plain C types only; every function returns 0 on success, non-zero on error;
 *    llmf90_b200_last_error() then describes it (the reference's convention is
 *    `print *, msg; stop`, read_ggml.f90:122-125 -- the host prints and stops).
 *  - `token` and `pos` are 1-based, exactly what the Fortran loop passes
 *    (llama2.f90:376-380); BOS = 2.
 *  - weight pointers are the base addresses of the TransformerWeights
 *    allocatables (weight_module.f90:13-26).  Column-major Fortran == row-major C
 *    with reversed index order: wqkv(emb, emb+2kv, L) is [L][emb+2kv][emb] etc.
 *    The library copies (and re-lays-out / shards) them to the device in init and
 *    never frees or mutates host memory.
 *  - the engine is a process-wide singleton (the reference is a single-model
 *    program); the caller is single-threaded; calls are synchronous -- on return
 *    of llmf90_b200_transformer the logits are in host memory (llama2.f90:388-391
 *    reads them immediately).
 *  - RunState (weight_module.f90:33-40: att, key_cache, value_cache, times) lives
 *    on the device and is owned by the library.
 *  - there is NO CPU fallback: without a CUDA device every compute entry point
 *    fails with an error.
 */
#ifndef LLMF90_B200_H
#define LLMF90_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* storage of the 2-D weight tensors (ggml tensor type ids, read_ggml.f90:613-635) */
#define LLMF90_WTYPE_F32  0   /* what the reference's master branch runs                */
#define LLMF90_WTYPE_F16  1   /* ggml type 1 (the optimize16 branch's storage)          */
#define LLMF90_WTYPE_Q4_0 2   /* ggml type 2: 18-byte blocks of 32 (four_bit_dev branch) */

/* flags */
#define LLMF90_FLAG_GRANULAR 1u  /* run the forward as separate kernels (rmsnorm, matvec, rope,
                                    attention ...) instead of the fused weight-streaming kernel */
#define LLMF90_FLAG_PROFILE  2u  /* fused kernel with per-phase timers (llmf90_b200_phase_times, the five
                                    reference buckets of llmf90_b200_times); costs 5-8 % of the token
                                    time, so it is off by default: all forward time then goes to bucket 4.
                                    Also switched on by the environment variable LLMF90_PROFILE=1 */
#define LLMF90_FLAG_PREFILL  4u  /* keep a second copy of the layer matrices in tensor-core operand order
                                    (f16 hi / lo planes) so that llmf90_b200_prefill can run the prompt
                                    positions as one batched tcgen05 pass; single-GPU */

#define LLMF90_FLAG_CLS_Q6K  8u  /* wcls points at ggml Q6_K super-blocks (type 14: 210 bytes per 256 weights, as in the
                                    file) instead of `wtype` rows: the output.weight of stock llama.cpp q4_0 GGUFs,
                                    which the reference's type switch stops at (read_ggml.f90:613-635).  The fused
                                    kernel streams ONE storage type, so such a model runs on the granular engine
                                    (the flag implies LLMF90_FLAG_GRANULAR); single-GPU; emb_dim % 256 == 0 */
#define LLMF90_WTYPE_Q6_K 14  /* accepted by llmf90_b200_matvec only */

/* mirror of `type Config` (weight_module.f90:28-31) + dtype and placement */
typedef struct llmf90_b200_config {
    int32_t emb_dim;      /* llama2.f90:102 */
    int32_t hidden_dim;   /* :103 */
    int32_t n_layers;     /* :104 */
    int32_t n_heads;      /* :105 */
    int32_t n_kv_heads;   /* :106 */
    int32_t vocab_size;   /* :107 */
    int32_t seq_len;      /* :108 -- size of the KV cache in positions */
    int32_t wtype;        /* LLMF90_WTYPE_* */
    int32_t device;       /* CUDA device ordinal */
    int32_t tp_rank;      /* tensor-parallel rank of this process (0 when tp_size == 1) */
    int32_t tp_size;      /* number of GPUs the heads / FFN rows are sharded over (1,2,4,8) */
    uint32_t flags;       /* LLMF90_FLAG_* */
} llmf90_b200_config;

/* replaces the allocation + load of `weights` and `s` (llama2.f90:151, :311-319).
 * Norm vectors are always f32; the other tensors are `wtype` rows. */
int llmf90_b200_init(const llmf90_b200_config *cfg,
                     const void *token_embedding_table, /* [V][emb]            */
                     const float *rms_att_weight,       /* [L][emb]            */
                     const void *wqkv,                  /* [L][emb+2kv][emb]   */
                     const void *wo,                    /* [L][emb][emb]       */
                     const float *rms_ffn_weight,       /* [L][emb]            */
                     const void *w13,                   /* [L][2*hid][emb]     */
                     const void *w2,                    /* [L][emb][hid]       */
                     const float *rms_final_weight,     /* [emb]               */
                     const void *wcls);                 /* [V][emb]            */

/* replaces `function transformer(token,pos,s,w) result(logits)` (llama2.f90:480-640).
 * logits[vocab_size] is host memory. */
int llmf90_b200_transformer(int32_t token, int32_t pos, float *logits);

/* transformer() and the pick of the next token in one call, the pick made on the device: maxloc for
 * temperature == 0 (llama2.f90:388), otherwise softmax(logits / temperature) and the CDF walk against the uniform
 * number r in [0, 1) that the caller draws (llama2.f90:390-391, :428-447; the reference draws it with
 * random_number).  Returns the 1-based token in *next_token; the logits stay in HBM (4 bytes come back instead of
 * vocab_size floats).  The probabilities are summed in f32 like the reference's, in chunks instead of one chain:
 * the pick equals the sequential walk's unless r lies within rounding (~1e-6) of a CDF boundary. */
int llmf90_b200_transformer_sample(int32_t token, int32_t pos, float temperature, float r, int32_t *next_token);

/* s%times(1:5) in milliseconds, accumulated since init/reset (llama2.f90:526-638, :407-410).
 * With LLMF90_FLAG_PROFILE (or LLMF90_PROFILE=1) the five buckets come from the kernel's phase
 * timers; without it (the default, faster kernel) and on the granular path the whole forward pass
 * is reported in bucket 4 and the others are 0. */
int llmf90_b200_times(float t[5]);

/* profiling aid: the fused kernel's fine-grained phase timers in ms (17 buckets per layer loop:
 * qkv prologue, qkv mat-vec, rope+barrier, attention, barrier, wo prologue, wo mat-vec, barrier,
 * w13 prologue, w13 mat-vec, swiglu+barrier, w2 prologue, w2 mat-vec, barrier, cls prologue,
 * cls mat-vec, argmax), measured on CTA 0, accumulated since init/reset. */
#define LLMF90_N_PHASES 17
int llmf90_b200_phase_times(float *ms, int32_t n);

/* profiling aid: run one forward with the instrumented kernel and record, for every CTA, globaltimer
 * stamps (ns) taken by its thread 0 at the phase edges of `layer`: entry 0 layer start, 1 QKV prologue
 * done, 2 QKV tiles of warp 0 done, 4 attention done, 6 / 7 the same for Wo, 9 / 10 for W13, 12 / 13 for
 * W2 (other entries 0); out is [n_ctas][128]. */
int llmf90_b200_debug_trace(int32_t token, int32_t pos, int32_t layer, uint64_t *out, int32_t n_ctas);

/* zero the KV cache and the timers (the state llama2.f90:316-319 initialises) */
int llmf90_b200_reset(void);

int llmf90_b200_free(void);

const char *llmf90_b200_last_error(void);

/* The whole generation loop of llama2.f90:376-402 for temperature == 0 kept on the device
 * (argmax fused after the classifier, next token never leaves HBM).  out_tokens[i] is the
 * 1-based token chosen after position i+1 (forced prompt token or greedy pick); elapsed_ms
 * is device time from after the first token to the end (llama2.f90:399-406). */
int llmf90_b200_generate_greedy(const int32_t *prompt_tokens, int32_t n_prompt, int32_t n,
                                int32_t *out_tokens, float *elapsed_ms);

/* The forced prompt positions of llama2.f90:379-385 as ONE batched pass (needs LLMF90_FLAG_PREFILL): after the
 * call the KV cache holds the key / value rows of positions pos0 .. pos0 + n_tokens - 1 (1-based) for the input
 * tokens tokens[0 .. n_tokens) -- what n_tokens calls of llmf90_b200_transformer would have left there; their
 * logits, which the reference discards (llama2.f90:383-385), are not computed.  The caller continues with
 * llmf90_b200_transformer(next_token, pos0 + n_tokens, logits).  llmf90_b200_generate_greedy does this itself for
 * its prompt when the flag is set.  Prompts longer than 128 positions are taken in passes of 128. */
int llmf90_b200_prefill(const int32_t *tokens, int32_t n_tokens, int32_t pos0);

/* test aid: the key and value rows (n_kv_heads * head_size floats each; this rank's share under tensor
 * parallelism) of cache position pos (1-based) in `layer` (0-based) */
int llmf90_b200_debug_read_kv(int32_t layer, int32_t pos, float *k, float *v);

/* ---- the inner subroutines as separately callable operators (host pointers in/out) ---- */
/* y(ix) = dot_product(x, w(:,ix)), ix = 1..rows; w is `rows` rows of `cols` weights of `wtype` */
int llmf90_b200_matvec(const void *w, int32_t wtype, int32_t rows, int32_t cols, const float *x,
                       float *y);
/* the same dot products for n_pos (1..128) activation vectors at once -- the GEMM of the batched prompt pass
 * (tcgen05 tensor cores, f16 hi + lo operand planes, f32 accumulation): y[p][ix] = dot_product(x[p][:], w(:,ix));
 * x is [n_pos][cols], y is [n_pos][rows] */
int llmf90_b200_matmul(const void *w, int32_t wtype, int32_t rows, int32_t cols, const float *x, int32_t n_pos,
                       float *y);
/* llama2.f90:450-457 */
int llmf90_b200_rmsnorm(const float *x, const float *w, int32_t n, float *out);
/* llama2.f90:468-478: softmax over x(1:s), zeros in p(s+1:n) */
int llmf90_b200_softmax(const float *x, int32_t n, int32_t s, float *p);
/* llama2.f90:543-559, in place on q(1:emb) and k(1:kv); pos is 1-based */
int llmf90_b200_rope(float *q, float *k, int32_t emb, int32_t kv, int32_t head_size, int32_t pos);

/* ---- tensor parallelism plumbing (one process per GPU) ----
 * Every rank calls llmf90_b200_init with the FULL host weights and its tp_rank / tp_size: the
 * library uploads only the rank's shard (its heads' Wq rows, their KV heads' Wk / Wv rows, the
 * matching Wo columns, its FFN rows of W1 / W3 and columns of W2, its vocabulary rows).
 * n_heads must be a multiple of tp_size; n_kv_heads a multiple, or a divisor (each KV head is then
 * replicated on tp_size / n_kv_heads ranks: TinyLlama's 4 KV heads on 8 GPUs).  The
 * all-reduce after Wo and after W2 is fused into the decode kernel: every rank stores its partial
 * vector straight into every other rank's hand-over buffer over NVLink.  For that the ranks
 * exchange one 64-byte CUDA IPC handle each (any host transport: MPI, torch.distributed, a file):
 *   llmf90_b200_tp_export(h)          -> this rank's handle
 *   llmf90_b200_tp_connect(all, n)    <- the n = tp_size handles in rank order
 * After connect every rank calls llmf90_b200_transformer / _generate_greedy with the same
 * arguments in the same order; each returns the full, identical logits. */
int llmf90_b200_tp_export(void *handle64);
int llmf90_b200_tp_connect(const void *handles, int32_t n);

/* ---- introspection used by the benchmark harness ---- */
typedef struct llmf90_b200_stats {
    uint64_t kernel_launches;     /* kernels of this library launched since init/reset */
    uint64_t forward_calls;
    uint64_t weight_bytes_device; /* bytes of weights resident in HBM on this GPU      */
    uint64_t active_bytes_per_token; /* weight bytes one token streams on this GPU      */
    float last_forward_ms;        /* CUDA-event time of the most recent forward kernel(s) */
    int32_t n_sms;
    int32_t stream_slots, stream_slot_bytes, stream_smem_bytes, stream_threads;
    float last_loop_total_ms;       /* device loop (generate_greedy / bench_device_loop): all positions */
    float last_loop_after_first_ms; /* same, from after the first token (llama2.f90:399-406)            */
} llmf90_b200_stats;
int llmf90_b200_get_stats(llmf90_b200_stats *out);

/* ---- the fused kernel's plan for a configuration, computed WITHOUT a device ----
 * What init decides before the first launch: the grid, the shared-memory ring, and for every CTA
 * the list of bulk copies one token takes -- [embedding row][one layer: rms_att, its QKV rows, its Wo
 * rows, rms_ffn, its W13 rows, its W2 rows][rms_final, its classifier rows]; the kernel walks the layer
 * section n_layers times adding layer_stride16 * 16 bytes per layer.  A ring stage is one chunk of the
 * contraction range of one tile (tile_rows rows, owned by one group of tile_warps consumer warps) and ONE
 * bulk copy: the f32 / f16 matrices are stored tile-major on the device (the chunk's segments of the
 * tile's rows are contiguous), tiled q4_0 in whole 8-block groups.  Sources are in
 * a virtual address space: region k starts at LLMF90_PLAN_VBASE(k), k = 0..4 the five streamed
 * matrices (QKV, Wo, W13, W2, classifier) of this rank's shard, 5 the embedding table, 6 / 7 / 8 the
 * rms_att / rms_ffn / rms_final vectors.  The CPU test-suite uses it to check, for full-size models
 * and every tensor-parallel split, that the stages are 16-byte aligned, fit a ring slot and cover every
 * matrix exactly once (tests/test_plan.py).  n_sms / smem_optin describe the device (B200: 148 SMs,
 * 232448 bytes of opt-in shared memory per block). */
#define LLMF90_PLAN_VBASE(k) (((uint64_t)(k) + 1u) << 40)
typedef struct llmf90_b200_plan_info {
    int32_t grid, threads;            /* CTAs (one per SM) and threads per CTA                     */
    int32_t n_slots, slot_bytes;      /* the ring                                                   */
    int32_t smem_bytes;               /* dynamic + static shared memory per CTA                     */
    int32_t sched_stride;             /* schedule entries reserved per CTA (unused ones: bytes = 0) */
    int32_t n_layers;
    int32_t rows[5], cols[5];         /* this rank's share of the five matrices                     */
    int32_t tile_rows[5];             /* rows of a tile (the unit one group of consumer warps owns) */
    int32_t tile_chunks[5];           /* ring stages a tile's contraction range is cut into         */
    int32_t tile_warps[5];            /* consumer warps that share one tile (a divisor of 12)       */
    int32_t reserved0;
    uint64_t matrix_bytes[5];         /* device bytes of one layer of each                          */
    uint64_t vector_bytes;            /* one rmsnorm weight vector                                  */
    uint64_t emb_row_bytes;           /* one row of the embedding table                             */
} llmf90_b200_plan_info;
typedef struct llmf90_b200_sched_stage {   /* one bulk copy (cp.async.bulk) */
    uint64_t src;                     /* virtual source address of the copy (layer 0)               */
    uint32_t bytes;                   /* 0 = unused entry                                           */
    uint32_t layer_stride16;          /* bytes / 16 to add per layer                                */
    uint32_t phase_start;             /* 1 on the first copy of a phase                             */
    uint32_t stage;                   /* ring stage of this CTA the copy belongs to (copies of one    */
                                      /* stage land back to back in one ring slot)                  */
} llmf90_b200_sched_stage;
/* sched may be NULL (info only); otherwise it receives grid * sched_stride entries, CTA-major. */
int llmf90_b200_plan(const llmf90_b200_config *cfg, int32_t n_sms, int32_t smem_optin,
                     llmf90_b200_plan_info *info, llmf90_b200_sched_stage *sched, int64_t sched_entries);

/* ---- the batched prompt pass's plan for a configuration, computed WITHOUT a device ----
 * One entry per GEMM of a layer (0 QKV, 1 Wo, 2 W1|W3, 3 W2) for a pass over n_pos positions (1..128): the
 * grid (m_tiles x n_splits CTAs: 128 weight rows x a range of 64-element contraction chunks), the TMEM columns
 * of the accumulator, the shared-memory pipeline, and the bytes of the operand-order weight copy and of the
 * K-split partial products.  The CPU test-suite checks with it that the splits cover every chunk exactly once,
 * that the pipeline fits a B200 SM (227 KB, 512 TMEM columns) and what the flag costs in HBM (tests/test_plan.py). */
typedef struct llmf90_b200_prefill_gemm {
    int32_t rows, cols, planes;
    int32_t m_tiles, k_chunks, chunks_per_split, n_splits;
    int32_t ppad, tmem_cols;
    int32_t stages, stage_bytes, smem_bytes;
    uint64_t weight_bytes;   /* per layer */
    uint64_t partial_bytes;
} llmf90_b200_prefill_gemm;
int llmf90_b200_prefill_plan(const llmf90_b200_config *cfg, int32_t n_sms, int32_t n_pos, llmf90_b200_prefill_gemm out[4]);

/* device-resident variant used for the HBM-resident measurement: runs `n_steps` forwards for
 * positions pos0..pos0+n_steps-1 feeding each one the greedy pick of the previous one, without
 * any host<->device traffic; returns the device time in ms. */
int llmf90_b200_bench_device_loop(int32_t first_token, int32_t pos0, int32_t n_steps,
                                  float *elapsed_ms);

#ifdef __cplusplus
}
#endif
#endif /* LLMF90_B200_H */

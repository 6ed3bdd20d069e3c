/*
 * llmf90_host.h -- C ABI of libllmf90_host.so, the C++ mirror of the reference's host program
 * (GGUF loader read_ggml.f90:53-511, tokenizer llama2.f90:643-724, sampler llama2.f90:387-447).
 * It exists because no Fortran compiler is available where this repo is built: the `llm` binary
 * and the tests drive the CUDA library through it.  Not part of the drop-in boundary.
 */
#ifndef LLMF90_HOST_H
#define LLMF90_HOST_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct llmf90_host_model llmf90_host_model;

typedef struct llmf90_host_config {
    int32_t emb_dim, hidden_dim, n_layers, n_heads, n_kv_heads, vocab_size, seq_len, wtype;
    int32_t cls_wtype; /* storage of tensor 8 (wcls): wtype, or 14 = ggml Q6_K blocks (stock llama.cpp q4_0 files) */
} llmf90_host_config;

/* NULL on failure; llmf90_host_last_error() says why.  verbose: 1 = the reference's -v listing,
 * 0 = only the unconditional "data offset" line (read_ggml.f90:196), -1 = silent */
llmf90_host_model *llmf90_host_load(const char *gguf_path, int32_t verbose);
/* the legacy `--ak` packed f32 model file (llama2.f90:158-294); no vocabulary: add one with
 * llmf90_host_load_tokenizer (the reference's -s) */
llmf90_host_model *llmf90_host_load_ak(const char *path, int32_t verbose);
void llmf90_host_free(llmf90_host_model *m);
const char *llmf90_host_last_error(void);
int llmf90_host_get_config(const llmf90_host_model *m, llmf90_host_config *out);
uint64_t llmf90_host_data_offset(const llmf90_host_model *m);
/* which: 0 token_embedding_table, 1 rms_att, 2 wqkv, 3 wo, 4 rms_ffn, 5 w13, 6 w2, 7 rms_final, 8 wcls */
const void *llmf90_host_tensor(const llmf90_host_model *m, int32_t which, uint64_t *nbytes);
/* vocabulary entry i (0-based): returns its byte length, copies at most cap bytes */
int32_t llmf90_host_vocab(const llmf90_host_model *m, int32_t i, char *buf, int32_t cap, float *score);
int llmf90_host_load_tokenizer(llmf90_host_model *m, const char *path);
/* bpe_encode: returns the number of tokens (1-based ids in out), or -1 */
int32_t llmf90_host_encode(const llmf90_host_model *m, const char *text, int32_t text_len, int32_t *out, int32_t cap);
int32_t llmf90_host_argmax(const float *logits, int32_t n);
int32_t llmf90_host_sample(const float *logits, int32_t n, float temperature, float r);

#ifdef __cplusplus
}
#endif
#endif

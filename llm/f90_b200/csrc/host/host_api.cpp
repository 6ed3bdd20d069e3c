// host_api.cpp -- C ABI over the host mirror (include/llmf90_host.h)
#include <cstring>
#include <stdexcept>

#include "../../../../include/llmf90_host.h"
#include "host.hpp"

struct llmf90_host_model {
    llmhost::Model m;
};

namespace {
thread_local std::string g_err;
}

extern "C" {

const char *llmf90_host_last_error(void) { return g_err.c_str(); }

llmf90_host_model *llmf90_host_load(const char *path, int32_t verbose)
{
    try {
        auto *h = new llmf90_host_model;
        h->m = llmhost::load_gguf(path ? path : "", verbose > 0, verbose >= 0);
        return h;
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}

llmf90_host_model *llmf90_host_load_ak(const char *path, int32_t verbose)
{
    try {
        auto *h = new llmf90_host_model;
        h->m = llmhost::load_ak(path ? path : "", verbose > 0);
        return h;
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}

void llmf90_host_free(llmf90_host_model *m) { delete m; }

int llmf90_host_get_config(const llmf90_host_model *m, llmf90_host_config *out)
{
    if (!m || !out) return 1;
    const auto &c = m->m.cfg;
    *out = {c.emb_dim, c.hidden_dim, c.n_layers, c.n_heads, c.n_kv_heads, c.vocab_size, c.seq_len, c.wtype, c.cls_wtype};
    return 0;
}

uint64_t llmf90_host_data_offset(const llmf90_host_model *m) { return m ? m->m.data_offset : 0; }

const void *llmf90_host_tensor(const llmf90_host_model *m, int32_t which, uint64_t *nbytes)
{
    if (!m) return nullptr;
    const auto &w = m->m.w;
    const void *p = nullptr;
    uint64_t n = 0;
    switch (which) {
    case 0: p = w.token_embedding_table.data(); n = w.token_embedding_table.size(); break;
    case 1: p = w.rms_att_weight.data(); n = w.rms_att_weight.size() * 4; break;
    case 2: p = w.wqkv.data(); n = w.wqkv.size(); break;
    case 3: p = w.wo.data(); n = w.wo.size(); break;
    case 4: p = w.rms_ffn_weight.data(); n = w.rms_ffn_weight.size() * 4; break;
    case 5: p = w.w13.data(); n = w.w13.size(); break;
    case 6: p = w.w2.data(); n = w.w2.size(); break;
    case 7: p = w.rms_final_weight.data(); n = w.rms_final_weight.size() * 4; break;
    case 8: p = w.wcls.data(); n = w.wcls.size(); break;
    default: return nullptr;
    }
    if (nbytes) *nbytes = n;
    return p;
}

int32_t llmf90_host_vocab(const llmf90_host_model *m, int32_t i, char *buf, int32_t cap, float *score)
{
    if (!m || i < 0 || i >= (int)m->m.vocab.tokens.size()) return -1;
    const std::string &t = m->m.vocab.tokens[i];
    if (buf && cap > 0) memcpy(buf, t.data(), std::min<size_t>(cap, t.size()));
    if (score) *score = m->m.vocab.scores[i];
    return (int32_t)t.size();
}

int llmf90_host_load_tokenizer(llmf90_host_model *m, const char *path)
{
    if (!m) return 1;
    try {
        llmhost::load_tokenizer_bin(path ? path : "", m->m.cfg.vocab_size, m->m.vocab);
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return 1;
    }
}

int32_t llmf90_host_encode(const llmf90_host_model *m, const char *text, int32_t text_len, int32_t *out, int32_t cap)
{
    if (!m || (!text && text_len > 0)) return -1;
    try {
        const auto toks = llmhost::bpe_encode(m->m.vocab, std::string(text ? text : "", (size_t)text_len));
        if ((int)toks.size() > cap) { g_err = "token buffer too small"; return -1; }
        for (size_t i = 0; i < toks.size(); i++) out[i] = toks[i];
        return (int32_t)toks.size();
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

int32_t llmf90_host_argmax(const float *logits, int32_t n) { return llmhost::argmax1(logits, n); }

int32_t llmf90_host_sample(const float *logits, int32_t n, float temperature, float r)
{
    std::vector<float> scratch;
    return llmhost::sample_cdf(logits, n, temperature, r, scratch);
}

}  // extern "C"

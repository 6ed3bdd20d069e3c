// ak_loader.cpp -- the reference's legacy `--ak` model file (llama2.f90:158-294): a packed f32 dump in
// the llama2.c order, WITHOUT the RoPE tables llama2.c carries:
//     7 x int32   emb_dim, hidden_dim, n_layers, n_heads, n_kv_heads, vocab_size, seq_len   (:164, order :163)
//     f32 token_embedding_table[V][emb]                                                      (:189-190)
//     f32 rms_att[L][emb]                                                                     (:196-197)
//     f32 Wq[L][emb][emb], Wk[L][kv][emb], Wv[L][kv][emb]                                     (:207-229)
//     f32 Wo[L][emb][emb]                                                                     (:235-238)
//     f32 rms_ffn[L][emb]                                                                     (:244-245)
//     f32 W1[L][hid][emb], W2[L][emb][hid], W3[L][hid][emb]                                   (:251-276)
//     f32 rms_final[emb], wcls[V][emb]                                                        (:282-292)
// The reference reads the header into a dummy and keeps its compile-time dimensions (:164); this
// mirror takes the dimensions from the header, like the GGUF path does.  A negative vocab_size
// (llama2.c's "unshared classifier" marker, see the commented code at :181-186) is accepted as
// its absolute value; the classifier is always read, as in the reference (shared_weights = .false.).
// The file has no vocabulary: pair it with `-s tokenizer.bin` (llama2.f90:321-356).
#include <cstdio>
#include <cstring>
#include <stdexcept>

#include "host.hpp"

namespace llmhost {

namespace {
struct File {
    FILE *f;
    explicit File(const std::string &p) : f(fopen(p.c_str(), "rb")) {
        if (!f) throw std::runtime_error("cannot open model file " + p);
    }
    ~File() { fclose(f); }
    void read(void *dst, size_t bytes, const char *what) {
        if (fread(dst, 1, bytes, f) != bytes) throw std::runtime_error(std::string("unexpected end of file reading ") + what);
    }
};
}  // namespace

Model load_ak(const std::string &path, bool verbose)
{
    File in(path);
    int32_t h[7];
    in.read(h, sizeof h, "the header");
    Model m;
    ModelConfig &c = m.cfg;
    c.emb_dim = h[0]; c.hidden_dim = h[1]; c.n_layers = h[2]; c.n_heads = h[3]; c.n_kv_heads = h[4];
    c.vocab_size = h[5] < 0 ? -h[5] : h[5]; c.seq_len = h[6]; c.wtype = 0;
    if (c.emb_dim <= 0 || c.hidden_dim <= 0 || c.n_layers <= 0 || c.n_heads <= 0 || c.n_kv_heads <= 0 ||
        c.vocab_size <= 0 || c.seq_len <= 0 || c.emb_dim % c.n_heads || c.n_heads % c.n_kv_heads)
        throw std::runtime_error("--ak header: implausible dimensions");
    const size_t emb = c.emb_dim, hid = c.hidden_dim, L = c.n_layers, V = c.vocab_size;
    const size_t kv = (size_t)c.n_kv_heads * (emb / c.n_heads), nqkv = emb + 2 * kv;
    if (verbose) {  // the reference's -v listing (:169-177)
        printf(" Embedding dimension:  %d\n Hidden dimension:  %d\n Layers:  %d\n Heads:  %d\n kv Heads:  %d\n"
               " Vocabulary Size:  %d\n Sequence Length:  %d\n Head Size:  %zu\n kv Head Size:  %zu\n",
               c.emb_dim, c.hidden_dim, c.n_layers, c.n_heads, c.n_kv_heads, c.vocab_size, c.seq_len, emb / c.n_heads, kv);
    }
    Weights &w = m.w;
    auto say = [&](const char *what, size_t n) { if (verbose) printf(" loaded %s: %zu\n", what, n); };
    w.token_embedding_table.resize(V * emb * 4);
    in.read(w.token_embedding_table.data(), w.token_embedding_table.size(), "token_embedding_table");
    say("embedding weights", V * emb);
    w.rms_att_weight.resize(L * emb);
    in.read(w.rms_att_weight.data(), L * emb * 4, "rms_att_weight");
    say("rms att weights", L * emb);
    // fused wqkv(emb, emb + 2 kv, L): per layer rows Wq | Wk | Wv (weight_module.f90:15, llama2.f90:206-229)
    w.wqkv.resize(L * nqkv * emb * 4);
    auto piece = [&](std::vector<uint8_t> &dst, size_t layer_rows, size_t row0, size_t rows, size_t cols, const char *what) {
        for (size_t l = 0; l < L; l++) in.read(dst.data() + ((l * layer_rows + row0) * cols) * 4, rows * cols * 4, what);
        say(what, L * rows * cols);
    };
    piece(w.wqkv, nqkv, 0, emb, emb, "wq weights");
    piece(w.wqkv, nqkv, emb, kv, emb, "wk weights");
    piece(w.wqkv, nqkv, emb + kv, kv, emb, "wv weights");
    w.wo.resize(L * emb * emb * 4);
    piece(w.wo, emb, 0, emb, emb, "wo weights");
    w.rms_ffn_weight.resize(L * emb);
    in.read(w.rms_ffn_weight.data(), L * emb * 4, "rms_ffn_weight");
    say("rms ffn  weights", L * emb);
    // fused w13(emb, 2 hid, L): rows W1 | W3, with W2 stored between them in the file (:250-276)
    w.w13.resize(L * 2 * hid * emb * 4);
    w.w2.resize(L * emb * hid * 4);
    piece(w.w13, 2 * hid, 0, hid, emb, "w1 weights");
    piece(w.w2, emb, 0, emb, hid, "w2 weights");
    piece(w.w13, 2 * hid, hid, hid, emb, "w3 weights");
    w.rms_final_weight.resize(emb);
    in.read(w.rms_final_weight.data(), emb * 4, "rms_final_weight");
    say("rms_final weights", emb);
    w.wcls.resize(V * emb * 4);
    in.read(w.wcls.data(), w.wcls.size(), "wcls");
    say("wcls weights", V * emb);
    // no vocabulary in this format: placeholders until -s tokenizer.bin replaces them
    m.vocab.tokens.assign(V, std::string());
    m.vocab.scores.assign(V, 0.f);
    m.vocab.build_index();
    m.arch = "llama";
    m.name = "ak";
    return m;
}

}  // namespace llmhost

// gguf_loader.cpp -- GGUF reader producing the fused weight_module layout.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <map>
#include <stdexcept>

#include "host.hpp"

namespace llmhost {

size_t row_bytes(int wtype, int n)
{
    if (wtype == 0) return (size_t)n * 4;
    if (wtype == 1) return (size_t)n * 2;
    if (wtype == 14) return (size_t)n / 256 * 210;  // Q6_K super-blocks
    return (size_t)n / 32 * 18;
}

namespace {

constexpr uint32_t GGUF_MAGIC = 1179993927u;  // "GGUF" little endian (read_ggml.f90:122)

struct Reader {
    std::ifstream f;
    std::string path;
    explicit Reader(const std::string &p) : f(p, std::ios::binary), path(p)
    {
        if (!f) throw std::runtime_error("cannot open model file " + p);
    }
    template <typename T>
    T get()
    {
        T v;
        f.read(reinterpret_cast<char *>(&v), sizeof v);
        if (!f) throw std::runtime_error("unexpected end of file in " + path);
        return v;
    }
    std::string str()
    {
        const uint64_t n = get<uint64_t>();
        if (n > (1u << 26)) throw std::runtime_error("GGUF string too long");
        std::string s(n, '\0');
        f.read(s.data(), (std::streamsize)n);
        if (!f) throw std::runtime_error("unexpected end of file in " + path);
        return s;
    }
    void skip(uint64_t n) { f.seekg((std::streamoff)n, std::ios::cur); }
    uint64_t tell() { return (uint64_t)f.tellg(); }
};

// GGUF KV value types 0..12; element sizes of the fixed-size ones (read_ggml.f90:663-685 handles 4,5,6,8,9)
size_t scalar_size(uint32_t t)
{
    switch (t) {
    case 0: case 1: case 7: return 1;        // u8, i8, bool
    case 2: case 3: return 2;                // u16, i16
    case 4: case 5: case 6: return 4;        // u32, i32, f32
    case 10: case 11: case 12: return 8;     // u64, i64, f64
    default: return 0;
    }
}

struct TensorInfo {
    std::vector<uint64_t> dims;  // innermost (contraction) dimension first
    uint32_t type = 0;
    uint64_t offset = 0;
};

}  // namespace

void Vocab::build_index()
{
    index.clear();
    index.reserve(tokens.size() * 2);
    for (int i = 0; i < (int)tokens.size(); i++) index.emplace(tokens[i], i);  // emplace keeps the first
}

int Vocab::lookup(const std::string &s) const
{
    auto it = index.find(s);
    return it == index.end() ? -1 : it->second;
}

Model load_gguf(const std::string &path, bool verbose, bool print_offset)
{
    Reader r(path);
    Model m;
    if (r.get<uint32_t>() != GGUF_MAGIC) throw std::runtime_error("GGUF magic number not found in " + path);
    m.gguf_version = (int)r.get<uint32_t>();
    if (m.gguf_version < 2 || m.gguf_version > 3)
        throw std::runtime_error("unsupported GGUF version " + std::to_string(m.gguf_version));
    const uint64_t n_tensors = r.get<uint64_t>(), n_kv = r.get<uint64_t>();
    if (verbose) printf(" GGUF version %d, %llu tensors, %llu key-value pairs\n", m.gguf_version,
                        (unsigned long long)n_tensors, (unsigned long long)n_kv);

    std::map<std::string, double> num;  // every numeric scalar KV
    uint32_t alignment = 32;            // read_ggml.f90:104
    for (uint64_t i = 0; i < n_kv; i++) {
        const std::string key = r.str();
        const uint32_t t = r.get<uint32_t>();
        if (t == 8) {
            const std::string v = r.str();
            if (key == "general.architecture") m.arch = v;
            if (key == "general.name") m.name = v;
            if (verbose) printf(" %s = %s\n", key.c_str(), v.size() < 80 ? v.c_str() : "(long string)");
        } else if (t == 9) {
            const uint32_t et = r.get<uint32_t>();
            const uint64_t n = r.get<uint64_t>();
            if (key == "tokenizer.ggml.tokens" && et == 8) {
                m.vocab.tokens.resize(n);
                for (auto &s : m.vocab.tokens) s = r.str();
            } else if (key == "tokenizer.ggml.scores" && et == 6) {
                m.vocab.scores.resize(n);
                r.f.read(reinterpret_cast<char *>(m.vocab.scores.data()), (std::streamsize)(n * 4));
            } else if (et == 8) {
                for (uint64_t k = 0; k < n; k++) r.str();
            } else if (scalar_size(et)) {
                r.skip(n * scalar_size(et));
            } else {
                throw std::runtime_error("GGUF: unsupported array element type " + std::to_string(et) + " for " + key);
            }
            if (verbose) printf(" %s = array[%llu]\n", key.c_str(), (unsigned long long)n);
        } else if (scalar_size(t)) {
            double v = 0;
            switch (t) {
            case 0: v = r.get<uint8_t>(); break;
            case 1: v = r.get<int8_t>(); break;
            case 2: v = r.get<uint16_t>(); break;
            case 3: v = r.get<int16_t>(); break;
            case 4: v = r.get<uint32_t>(); break;
            case 5: v = r.get<int32_t>(); break;
            case 6: v = r.get<float>(); break;
            case 7: v = r.get<uint8_t>(); break;
            case 10: v = (double)r.get<uint64_t>(); break;
            case 11: v = (double)r.get<int64_t>(); break;
            default: v = r.get<double>(); break;
            }
            num[key] = v;
            if (key == "general.alignment") {
                // (a zero or non-power-of-two alignment would divide by zero / misplace the tensor data below)
                if (!(v >= 1 && v <= 65536) || ((uint32_t)v & ((uint32_t)v - 1)))
                    throw std::runtime_error("GGUF: general.alignment " + std::to_string(v) + " is not a power of two in 1..65536");
                alignment = (uint32_t)v;
            }
            if (verbose) printf(" %s = %g\n", key.c_str(), v);
        } else {
            throw std::runtime_error("GGUF: unsupported value type " + std::to_string(t) + " for key " + key);
        }
    }

    std::map<std::string, TensorInfo> tensors;
    for (uint64_t i = 0; i < n_tensors; i++) {
        const std::string name = r.str();
        TensorInfo ti;
        const uint32_t nd = r.get<uint32_t>();
        if (nd > 4) throw std::runtime_error("GGUF: tensor " + name + " has too many dimensions");
        for (uint32_t d = 0; d < nd; d++) ti.dims.push_back(r.get<uint64_t>());
        ti.type = r.get<uint32_t>();
        ti.offset = r.get<uint64_t>();
        tensors[name] = ti;
    }
    // tensor data starts at the next multiple of the alignment (read_ggml.f90:176-196)
    uint64_t pos = r.tell();
    m.data_offset = (pos + alignment - 1) / alignment * alignment;
    // the reference prints this unconditionally (:196); library callers may silence it
    if (print_offset) printf(" data offset %llu\n", (unsigned long long)m.data_offset);

    auto need = [&](const std::string &name) -> const TensorInfo & {
        auto it = tensors.find(name);
        if (it == tensors.end()) throw std::runtime_error("key not found: " + name);  // read_ggml.f90:571
        return it->second;
    };
    auto key_u = [&](const std::string &k, int fallback) {
        auto it = num.find(k);
        return it == num.end() ? fallback : (int)it->second;
    };

    // ---- dimensions: from the llama.* keys, cross-checked against the tensor shapes
    ModelConfig &c = m.cfg;
    const TensorInfo &te = need("token_embd.weight");
    if (te.dims.size() != 2) throw std::runtime_error("token_embd.weight is not 2-D");
    c.emb_dim = (int)te.dims[0];
    c.vocab_size = (int)te.dims[1];
    c.wtype = (int)te.type;
    if (c.wtype < 0 || c.wtype > 2)
        throw std::runtime_error("Type not supported: ggml tensor type " + std::to_string(te.type) +
                                 " (f32, f16 and q4_0 are)");  // read_ggml.f90:633-635
    c.n_layers = key_u("llama.block_count", 0);
    if (c.n_layers <= 0) {
        while (tensors.count("blk." + std::to_string(c.n_layers) + ".attn_q.weight")) c.n_layers++;
    }
    if (c.n_layers <= 0) throw std::runtime_error("key not found: blk.0.attn_q.weight");
    c.hidden_dim = (int)need("blk.0.ffn_gate.weight").dims.at(1);
    c.n_heads = key_u("llama.attention.head_count", 0);
    if (c.n_heads <= 0) throw std::runtime_error("key not found: llama.attention.head_count");
    c.n_kv_heads = key_u("llama.attention.head_count_kv", c.n_heads);
    c.seq_len = key_u("llama.context_length", 2048);
    if (key_u("llama.embedding_length", c.emb_dim) != c.emb_dim ||
        key_u("llama.feed_forward_length", c.hidden_dim) != c.hidden_dim)
        throw std::runtime_error("GGUF: llama.* dimensions disagree with the tensor shapes");
    if (c.emb_dim % c.n_heads || c.n_heads % c.n_kv_heads) throw std::runtime_error("GGUF: head counts do not divide");
    const int hs = c.emb_dim / c.n_heads, kv = c.n_kv_heads * hs, e = c.emb_dim, h = c.hidden_dim, L = c.n_layers,
              V = c.vocab_size, wt = c.wtype;
    if (wt == 2 && (e % 32 || h % 32)) throw std::runtime_error("q4_0 needs dimensions that are multiples of 32");

    // ---- tensors -> fused arrays (read_ggml.f90:238-410)
    auto read_into = [&](const std::string &name, uint8_t *dst, int rows, int cols, int type) {
        const TensorInfo &ti = need(name);
        if ((int)ti.type != type)
            throw std::runtime_error("tensor " + name + " has ggml type " + std::to_string(ti.type) + ", expected " +
                                     std::to_string(type));
        const uint64_t r_have = ti.dims.size() > 1 ? ti.dims[1] : 1;
        if ((int)ti.dims[0] != cols || (int)r_have != rows)
            throw std::runtime_error("tensor " + name + " has an unexpected shape");
        r.f.seekg((std::streamoff)(m.data_offset + ti.offset));
        const size_t nbytes = type == 0 && ti.dims.size() == 1 ? (size_t)cols * 4 : (size_t)rows * row_bytes(type, cols);
        r.f.read(reinterpret_cast<char *>(dst), (std::streamsize)nbytes);
        if (!r.f) throw std::runtime_error("unexpected end of file reading " + name);
        if (verbose) printf(" %s %d x %d\n", name.c_str(), cols, rows);
    };
    const size_t rb_e = row_bytes(wt, e), rb_h = row_bytes(wt, h);
    const int nqkv = e + 2 * kv;
    Weights &w = m.w;
    w.token_embedding_table.resize((size_t)V * rb_e);
    // output.weight: the model's type, or Q6_K -- llama.cpp's q4_0 files keep the classifier in ggml type 14, which
    // the reference's type switch (read_ggml.f90:613-635) stops at; SURVEY.md 8f1
    c.cls_wtype = (int)need("output.weight").type;
    if (c.cls_wtype != wt && c.cls_wtype != 14)
        throw std::runtime_error("Type not supported: output.weight has ggml tensor type " + std::to_string(c.cls_wtype) +
                                 " (the model's type and Q6_K are)");
    if (c.cls_wtype == 14 && e % 256) throw std::runtime_error("Q6_K output.weight needs an embedding length that is a multiple of 256");
    w.wcls.resize((size_t)V * row_bytes(c.cls_wtype, e));
    w.wqkv.resize((size_t)L * nqkv * rb_e);
    w.wo.resize((size_t)L * e * rb_e);
    w.w13.resize((size_t)L * 2 * h * rb_e);
    w.w2.resize((size_t)L * e * rb_h);
    w.rms_att_weight.resize((size_t)L * e);
    w.rms_ffn_weight.resize((size_t)L * e);
    w.rms_final_weight.resize(e);
    read_into("token_embd.weight", w.token_embedding_table.data(), V, e, wt);
    for (int l = 0; l < L; l++) {
        const std::string p = "blk." + std::to_string(l) + ".";
        uint8_t *qkv = w.wqkv.data() + (size_t)l * nqkv * rb_e;
        read_into(p + "attn_norm.weight", reinterpret_cast<uint8_t *>(w.rms_att_weight.data() + (size_t)l * e), 1, e, 0);
        read_into(p + "attn_q.weight", qkv, e, e, wt);                          // rows 0 .. emb-1        (:272)
        read_into(p + "attn_k.weight", qkv + (size_t)e * rb_e, kv, e, wt);       // next kv rows           (:286)
        read_into(p + "attn_v.weight", qkv + (size_t)(e + kv) * rb_e, kv, e, wt);  // last kv rows           (:300)
        read_into(p + "attn_output.weight", w.wo.data() + (size_t)l * e * rb_e, e, e, wt);
        read_into(p + "ffn_norm.weight", reinterpret_cast<uint8_t *>(w.rms_ffn_weight.data() + (size_t)l * e), 1, e, 0);
        uint8_t *w13 = w.w13.data() + (size_t)l * 2 * h * rb_e;
        read_into(p + "ffn_gate.weight", w13, h, e, wt);                         // rows 0 .. hid-1 = W1   (:347)
        read_into(p + "ffn_up.weight", w13 + (size_t)h * rb_e, h, e, wt);        // rows hid .. 2hid-1 = W3 (:376)
        read_into(p + "ffn_down.weight", w.w2.data() + (size_t)l * e * rb_h, e, h, wt);
    }
    read_into("output_norm.weight", reinterpret_cast<uint8_t *>(w.rms_final_weight.data()), 1, e, 0);
    read_into("output.weight", w.wcls.data(), V, e, c.cls_wtype);  // a separate classifier is required (:406)

    // ---- vocabulary: a leading U+2581 becomes one space (read_ggml.f90:483-503)
    if ((int)m.vocab.tokens.size() != V) throw std::runtime_error("key not found: tokenizer.ggml.tokens");
    if ((int)m.vocab.scores.size() != V) m.vocab.scores.assign(V, 0.f);
    for (auto &t : m.vocab.tokens)
        if (t.size() >= 3 && (uint8_t)t[0] == 0xE2 && (uint8_t)t[1] == 0x96 && (uint8_t)t[2] == 0x81)
            t = " " + t.substr(3);
    m.vocab.build_index();
    return m;
}

void load_tokenizer_bin(const std::string &path, int vocab_size, Vocab &out)
{
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot open tokenizer file " + path);
    int32_t max_len = 0;
    f.read(reinterpret_cast<char *>(&max_len), 4);
    out.tokens.assign(vocab_size, "");
    out.scores.assign(vocab_size, 0.f);
    for (int i = 0; i < vocab_size; i++) {
        float score;
        int32_t len;
        f.read(reinterpret_cast<char *>(&score), 4);
        f.read(reinterpret_cast<char *>(&len), 4);
        if (!f || len < 0 || len > (1 << 16)) throw std::runtime_error("malformed tokenizer file " + path);
        out.tokens[i].resize(len);
        f.read(out.tokens[i].data(), len);
        out.scores[i] = score;
    }
    if (!f) throw std::runtime_error("unexpected end of tokenizer file " + path);
    out.build_index();
}

}  // namespace llmhost

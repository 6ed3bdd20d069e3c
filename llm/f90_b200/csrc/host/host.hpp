// host.hpp -- C++ mirror of the reference's HOST program (everything above the transformer()
// call in llama2.f90): GGUF loader, tokenizer, sampler.  No Fortran compiler exists in this
// image, so this is the caller of the C ABI that is actually built, tested and timed; the
// Fortran-side binding a maintainer would add is shown in INTEGRATION.md / fortran/.
//
// Behaviour follows the reference (citations into /root/reference), the structure does not:
// hash-map vocabulary lookup instead of a linear scan, one generic KV reader instead of a
// per-type cascade, dimensions taken from the file instead of compile-time parameters.
#pragma once
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

namespace llmhost {

struct ModelConfig {  // type Config (weight_module.f90:28-31) + storage type
    int emb_dim = 0, hidden_dim = 0, n_layers = 0, n_heads = 0, n_kv_heads = 0, vocab_size = 0, seq_len = 0;
    int wtype = 0;  // 0 f32, 1 f16, 2 q4_0 (ggml tensor type ids)
    int cls_wtype = 0;  // storage of wcls: wtype, or 14 (ggml Q6_K: the output.weight of stock llama.cpp q4_0 files)
};

// TransformerWeights (weight_module.f90:13-26) in the C view of the Fortran layout; 2-D tensors
// are raw bytes of `wtype` rows, norm vectors f32.
struct Weights {
    std::vector<uint8_t> token_embedding_table, wqkv, wo, w13, w2, wcls;
    std::vector<float> rms_att_weight, rms_ffn_weight, rms_final_weight;
};

struct Vocab {
    std::vector<std::string> tokens;  // true byte strings (leading U+2581 already rewritten to ' ')
    std::vector<float> scores;
    std::unordered_map<std::string, int> index;  // first occurrence wins, like the reference's scan
    void build_index();
    int lookup(const std::string &s) const;  // 0-based id or -1 (llama2.f90:643-655)
};

struct Model {
    ModelConfig cfg;
    Weights w;
    Vocab vocab;
    uint64_t data_offset = 0;
    int gguf_version = 0;
    std::string arch, name;
};

size_t row_bytes(int wtype, int n);

// load_ggml (read_ggml.f90:53-511) extended per SURVEY.md 8f: tensor types 0/1/2, every GGUF KV
// value type, dimensions from the llama.* keys.  Throws std::runtime_error with the reference's
// style of message ("key not found", "GGUF magic", ...).
Model load_gguf(const std::string &path, bool verbose, bool print_offset = true);
// legacy `--ak` packed f32 model file (llama2.f90:158-294; ak_loader.cpp); carries no vocabulary
Model load_ak(const std::string &path, bool verbose);
// legacy `-s tokenizer.bin` (llama2.f90:321-356): i32 max_len, then per token f32 score, i32 len, bytes
void load_tokenizer_bin(const std::string &path, int vocab_size, Vocab &out);

// bpe_encode (llama2.f90:658-724): one token per input BYTE, then repeatedly merge the adjacent pair
// whose concatenation is a vocabulary entry with the highest score.  Returns 1-based ids like the
// reference; a byte with no single-byte entry throws (the reference indexes vocab(-1)).
std::vector<int> bpe_encode(const Vocab &v, const std::string &text);

// maxloc (first maximum, 1-based; llama2.f90:388)
int argmax1(const float *logits, int n);
// softmax(logits / T) then the CDF walk against r in [0,1) (llama2.f90:390-391, :428-447); 1-based
int sample_cdf(const float *logits, int n, float temperature, float r, std::vector<float> &scratch);

}  // namespace llmhost

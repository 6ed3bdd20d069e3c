// llm_main.cpp -- the `./llm -m <gguf>` command (program llama2, llama2.f90:87-410) with the forward
// pass behind the C ABI of libllmf90_b200.so.  Same flags, same token loop, same report lines.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>

#include "../../../../include/llmf90_b200.h"
#include "host.hpp"

namespace {

struct Args {  // type args (llama2.f90:7-14), defaults :26-32
    float temperature = 0.f;
    std::string model_file = "stories15M.bin", prompt, tokenizer;
    bool verbose = false, ak = false;
    int n = 256;
    int device = 0, granular = 0;
};

[[noreturn]] void die(const std::string &msg)
{
    printf(" %s\n", msg.c_str());  // the reference prints and stops (read_ggml.f90:122-125)
    exit(1);
}

Args parse_args(int argc, char **argv)
{
    Args a;
    for (int i = 1; i < argc;) {
        const std::string f = argv[i];
        auto val = [&]() -> std::string {
            if (i + 1 >= argc) die("Missing value for option: " + f);
            return argv[i + 1];
        };
        if (f == "-m" || f == "--model") { a.model_file = val(); i += 2; }
        else if (f == "-p" || f == "--prompt") { a.prompt = val(); i += 2; }
        else if (f == "-s" || f == "--tokenizer") { a.tokenizer = val(); i += 2; }
        else if (f == "-t" || f == "--temperature") { a.temperature = (float)atof(val().c_str()); i += 2; }
        else if (f == "-n" || f == "--num_tokens") { a.n = atoi(val().c_str()); i += 2; }
        else if (f == "-v" || f == "--verbose") { a.verbose = true; i += 1; }
        else if (f == "--ak") { a.ak = true; i += 1; }
        else if (f == "--device") { a.device = atoi(val().c_str()); i += 2; }       // extension
        else if (f == "--granular") { a.granular = 1; i += 1; }                     // extension
        else die("Unrecognized option: " + f);                                      // llama2.f90:74-75
    }
    return a;
}

}  // namespace

int main(int argc, char **argv)
{
    const Args a = parse_args(argc, argv);
    if (a.ak && a.tokenizer.empty()) die("--ak model files carry no vocabulary: pass -s tokenizer.bin");

    llmhost::Model m;
    try {
        m = a.ak ? llmhost::load_ak(a.model_file, a.verbose) : llmhost::load_gguf(a.model_file, a.verbose);
        if (!a.tokenizer.empty()) llmhost::load_tokenizer_bin(a.tokenizer, m.cfg.vocab_size, m.vocab);
    } catch (const std::exception &e) {
        die(e.what());
    }
    if (a.verbose) printf(" Loaded weights\n");

    int seq_len = m.cfg.seq_len;
    if (a.n <= seq_len) seq_len = a.n;  // llama2.f90:363-368
    else printf(" %d greater than maxinum squence length\n set to %d\n", a.n, seq_len);

    llmf90_b200_config cfg{};
    cfg.emb_dim = m.cfg.emb_dim; cfg.hidden_dim = m.cfg.hidden_dim; cfg.n_layers = m.cfg.n_layers;
    cfg.n_heads = m.cfg.n_heads; cfg.n_kv_heads = m.cfg.n_kv_heads; cfg.vocab_size = m.cfg.vocab_size;
    cfg.seq_len = m.cfg.seq_len; cfg.wtype = m.cfg.wtype; cfg.device = a.device; cfg.tp_rank = 0; cfg.tp_size = 1;
    cfg.flags = a.granular ? LLMF90_FLAG_GRANULAR : 0;
    if (llmf90_b200_init(&cfg, m.w.token_embedding_table.data(), m.w.rms_att_weight.data(), m.w.wqkv.data(),
                         m.w.wo.data(), m.w.rms_ffn_weight.data(), m.w.w13.data(), m.w.w2.data(),
                         m.w.rms_final_weight.data(), m.w.wcls.data()))
        die(llmf90_b200_last_error());

    std::vector<int> prompt_tokens;
    try {
        prompt_tokens = llmhost::bpe_encode(m.vocab, a.prompt);
    } catch (const std::exception &e) {
        die(e.what());
    }

    std::vector<float> logits(m.cfg.vocab_size), scratch;
    std::mt19937 rng(std::random_device{}());  // the reference never seeds random_number (llama2.f90:433)
    std::uniform_real_distribution<float> uni(0.f, 1.f);
    using clk = std::chrono::steady_clock;
    clk::time_point t_start{};
    bool started = false;
    int token = 2;  // <s>, 1-based (llama2.f90:376)
    for (int pos = 1; pos <= seq_len; pos++) {
        if (llmf90_b200_transformer(token, pos, logits.data())) die(llmf90_b200_last_error());
        if (pos <= (int)prompt_tokens.size()) token = prompt_tokens[pos - 1];
        else if (a.temperature == 0.f) token = llmhost::argmax1(logits.data(), m.cfg.vocab_size);
        else token = llmhost::sample_cdf(logits.data(), m.cfg.vocab_size, a.temperature, uni(rng), scratch);
        const std::string &piece = m.vocab.tokens[token - 1];
        fwrite(piece.data(), 1, piece.size(), stdout);
        fflush(stdout);
        if (!started) { t_start = clk::now(); started = true; }  // start after the first token (:399-401)
    }
    const double ms = std::chrono::duration<double, std::milli>(clk::now() - t_start).count();
    printf("\n Inference time:  %g  seconds\n", ms / 1000.0);
    printf(" %g tokens/second\n", 1000.0 * (seq_len - 1) / ms);
    printf(" Timings\n");
    float t[5] = {0, 0, 0, 0, 0};
    llmf90_b200_times(t);
    for (int l = 0; l < 5; l++) printf(" %d %g\n", l + 1, t[l] / seq_len);
    llmf90_b200_free();
    return 0;
}

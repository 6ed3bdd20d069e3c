// llm_main.cpp -- the `./llm -m <gguf>` command (program llama2, llama2.f90:87-410) with the forward
// pass behind the C ABI of libllmf90_b200.so.  Same flags, same token loop, same report lines.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "../../../../include/llmf90_b200.h"
#include "host.hpp"

namespace {

struct Args {  // type args (llama2.f90:7-14), defaults :26-32
    float temperature = 0.f;
    std::string model_file = "stories15M.bin", prompt, tokenizer;
    bool verbose = false, ak = false;
    int n = 256;
    int device = -1, granular = 0;
    bool host_sampler = false;  // extension: copy the logits back and pick on the host, like the reference
    bool prefill = false;       // extension: the prompt positions as one batched tensor-core pass (single GPU)
    // extension: tensor parallelism, one `llm` process per GPU started with the same flags plus its
    // --tp-rank; the ranks meet in --tp-dir (a fresh directory on a shared file system)
    int tp_size = 1, tp_rank = 0;
    std::string tp_dir;
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
        else if (f == "--host-sampler") { a.host_sampler = true; i += 1; }          // extension
        else if (f == "--prefill") { a.prefill = true; i += 1; }                    // extension
        else if (f == "--tp-size") { a.tp_size = atoi(val().c_str()); i += 2; }     // extension
        else if (f == "--tp-rank") { a.tp_rank = atoi(val().c_str()); i += 2; }     // extension
        else if (f == "--tp-dir") { a.tp_dir = val(); i += 2; }                     // extension
        else die("Unrecognized option: " + f);                                      // llama2.f90:74-75
    }
    return a;
}

// ---- tensor-parallel rendezvous through a directory: every rank publishes its 64-byte CUDA IPC handle
// (llmf90_b200_tp_export) plus a random seed as <dir>/rank<r>.tp (written to a temporary name, then
// renamed: a reader never sees half a file), waits for the other ranks' files and connects
// (llmf90_b200_tp_connect).  Rank 0's seed drives every rank's sampler: all ranks must feed the same
// token, so they must draw the same random numbers.
constexpr size_t TP_REC = 64 + 8;

std::string tp_file(const Args &a, int r) { return a.tp_dir + "/rank" + std::to_string(r) + ".tp"; }

uint64_t tp_rendezvous(const Args &a, uint64_t my_seed)
{
    unsigned char rec[TP_REC];
    if (llmf90_b200_tp_export(rec)) die(llmf90_b200_last_error());
    memcpy(rec + 64, &my_seed, 8);
    const std::string mine = tp_file(a, a.tp_rank), tmp = mine + ".tmp";
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f || fwrite(rec, 1, TP_REC, f) != TP_REC || fclose(f) != 0 || rename(tmp.c_str(), mine.c_str()) != 0)
        die("cannot write " + mine);
    std::vector<unsigned char> all((size_t)a.tp_size * 64);
    uint64_t seed0 = my_seed;
    const auto deadline = std::chrono::steady_clock::now() + std::chrono::seconds(600);
    for (int r = 0; r < a.tp_size; r++) {
        unsigned char got[TP_REC];
        for (;;) {
            FILE *g = fopen(tp_file(a, r).c_str(), "rb");
            const size_t n = g ? fread(got, 1, TP_REC, g) : 0;
            if (g) fclose(g);
            if (n == TP_REC) break;
            if (std::chrono::steady_clock::now() > deadline) die("timed out waiting for " + tp_file(a, r));
            std::this_thread::sleep_for(std::chrono::milliseconds(20));
        }
        memcpy(all.data() + (size_t)r * 64, got, 64);
        if (r == 0) memcpy(&seed0, got + 64, 8);
    }
    if (llmf90_b200_tp_connect(all.data(), a.tp_size)) die(llmf90_b200_last_error());
    return seed0;
}

}  // namespace

int main(int argc, char **argv)
{
    const Args a = parse_args(argc, argv);
    if (a.ak && a.tokenizer.empty()) die("--ak model files carry no vocabulary: pass -s tokenizer.bin");
    if (a.tp_size != 1 && a.tp_size != 2 && a.tp_size != 4 && a.tp_size != 8) die("--tp-size must be 1, 2, 4 or 8");
    if (a.tp_rank < 0 || a.tp_rank >= a.tp_size) die("--tp-rank out of range");
    if (a.tp_size > 1 && a.tp_dir.empty()) die("--tp-size > 1 needs --tp-dir <fresh directory shared by the ranks>");
    const bool talk = a.tp_rank == 0;  // one rank prints; all ranks compute the same tokens

    llmhost::Model m;
    try {
        m = a.ak ? llmhost::load_ak(a.model_file, a.verbose && talk)
                 : llmhost::load_gguf(a.model_file, a.verbose && talk, /*print_offset=*/talk);
        if (!a.tokenizer.empty()) llmhost::load_tokenizer_bin(a.tokenizer, m.cfg.vocab_size, m.vocab);
    } catch (const std::exception &e) {
        die(e.what());
    }
    if (a.verbose && talk) printf(" Loaded weights\n");

    int seq_len = m.cfg.seq_len;
    if (a.n <= seq_len) seq_len = a.n;  // llama2.f90:363-368
    else if (talk) printf(" %d greater than maxinum squence length\n set to %d\n", a.n, seq_len);

    llmf90_b200_config cfg{};
    cfg.emb_dim = m.cfg.emb_dim; cfg.hidden_dim = m.cfg.hidden_dim; cfg.n_layers = m.cfg.n_layers;
    cfg.n_heads = m.cfg.n_heads; cfg.n_kv_heads = m.cfg.n_kv_heads; cfg.vocab_size = m.cfg.vocab_size;
    cfg.seq_len = m.cfg.seq_len; cfg.wtype = m.cfg.wtype; cfg.device = a.device >= 0 ? a.device : a.tp_rank;
    cfg.tp_rank = a.tp_rank; cfg.tp_size = a.tp_size;
    cfg.flags = a.granular ? LLMF90_FLAG_GRANULAR : 0;
    if (m.cfg.cls_wtype == 14) cfg.flags |= LLMF90_FLAG_CLS_Q6K;  // stock llama.cpp q4_0 file: Q6_K classifier
    if (a.prefill && a.tp_size == 1) cfg.flags |= LLMF90_FLAG_PREFILL;
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
    uint64_t seed = ((uint64_t)std::random_device{}() << 32) ^ std::random_device{}();  // the reference never seeds random_number (llama2.f90:433)
    if (a.tp_size > 1) seed = tp_rendezvous(a, seed);
    std::mt19937_64 rng(seed);
    std::uniform_real_distribution<float> uni(0.f, 1.f);
    using clk = std::chrono::steady_clock;
    clk::time_point t_start{};
    bool started = false;
    int token = 2;  // <s>, 1-based (llama2.f90:376)
    int first_pos = 1, timed_positions = seq_len - 1;  // the reference's clock starts after the first token (:399-401)
    if ((cfg.flags & LLMF90_FLAG_PREFILL) && !prompt_tokens.empty() && seq_len > 1) {
        t_start = clk::now(); started = true; timed_positions = seq_len;  // here every position is inside the clock
        // The forced positions 1..np (inputs <s>, prompt[0..np-2]) only leave KV rows behind -- their picks are
        // overwritten by the prompt (llama2.f90:383-385) -- so they are one batched pass; the loop starts at np + 1.
        const int np = std::min((int)prompt_tokens.size(), seq_len - 1);
        std::vector<int32_t> in(np);
        in[0] = 2;
        for (int i = 1; i < np; i++) in[i] = prompt_tokens[i - 1];
        if (llmf90_b200_prefill(in.data(), np, 1)) die(llmf90_b200_last_error());
        for (int i = 0; i < np && talk; i++) {  // the reference prints the forced tokens as it goes (:395)
            const std::string &piece = m.vocab.tokens[prompt_tokens[i] - 1];
            fwrite(piece.data(), 1, piece.size(), stdout);
        }
        fflush(stdout);
        token = prompt_tokens[np - 1];
        first_pos = np + 1;
    }
    for (int pos = first_pos; pos <= seq_len; pos++) {
        if (a.host_sampler || a.tp_size > 1) {  // (tensor-parallel runs keep the path their tests cover: logits on every rank, host pick)
            // the reference's own shape: logits to the host, pick there (llama2.f90:380-392)
            if (llmf90_b200_transformer(token, pos, logits.data())) die(llmf90_b200_last_error());
            if (pos <= (int)prompt_tokens.size()) token = prompt_tokens[pos - 1];
            else if (a.temperature == 0.f) token = llmhost::argmax1(logits.data(), m.cfg.vocab_size);
            else token = llmhost::sample_cdf(logits.data(), m.cfg.vocab_size, a.temperature, uni(rng), scratch);
        } else {
            // default: maxloc / softmax(logits / T) + CDF walk next to the logits on the device; only the token
            // comes back.  The uniform number is drawn here, once per sampled position, like random_number (:433).
            const bool forced = pos <= (int)prompt_tokens.size();
            const float r = (!forced && a.temperature != 0.f) ? uni(rng) : 0.f;
            int32_t next = 0;
            if (llmf90_b200_transformer_sample(token, pos, a.temperature, r, &next)) die(llmf90_b200_last_error());
            token = forced ? prompt_tokens[pos - 1] : next;
        }
        if (talk) {
            const std::string &piece = m.vocab.tokens[token - 1];
            fwrite(piece.data(), 1, piece.size(), stdout);
            fflush(stdout);
        }
        if (!started) { t_start = clk::now(); started = true; }  // start after the first token (:399-401)
    }
    const double ms = std::chrono::duration<double, std::milli>(clk::now() - t_start).count();
    if (talk) {
        printf("\n Inference time:  %g  seconds\n", ms / 1000.0);
        printf(" %g tokens/second\n", 1000.0 * timed_positions / ms);
        printf(" Timings\n");
        float t[5] = {0, 0, 0, 0, 0};
        llmf90_b200_times(t);
        for (int l = 0; l < 5; l++) printf(" %d %g\n", l + 1, t[l] / seq_len);
    }
    llmf90_b200_free();
    if (a.tp_size > 1) remove(tp_file(a, a.tp_rank).c_str());
    return 0;
}

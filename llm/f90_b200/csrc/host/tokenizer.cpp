// tokenizer.cpp -- bpe_encode and the two ways the host picks a token (llama2.f90:387-392, :658-724)
#include <cmath>
#include <stdexcept>

#include "host.hpp"

namespace llmhost {

std::vector<int> bpe_encode(const Vocab &v, const std::string &text)
{
    std::vector<int> toks;  // 0-based while merging
    toks.reserve(text.size());
    for (char ch : text) {
        const int id = v.lookup(std::string(1, ch));
        if (id < 0)
            throw std::runtime_error("prompt byte " + std::to_string((unsigned char)ch) +
                                     " has no single-byte vocabulary entry");
        toks.push_back(id);
    }
    while (toks.size() >= 2) {
        float best_score = -1e10f;
        int best_i = -1, best_id = -1;
        for (size_t i = 0; i + 1 < toks.size(); i++) {
            const int id = v.lookup(v.tokens[toks[i]] + v.tokens[toks[i + 1]]);
            if (id >= 0 && v.scores[id] > best_score) {  // strict: the first of equal scores wins (:699)
                best_score = v.scores[id];
                best_i = (int)i;
                best_id = id;
            }
        }
        if (best_i < 0) break;
        toks[best_i] = best_id;
        toks.erase(toks.begin() + best_i + 1);
    }
    for (int &t : toks) t += 1;  // the reference's ids are 1-based
    return toks;
}

int argmax1(const float *logits, int n)
{
    int best = 0;
    for (int i = 1; i < n; i++)
        if (logits[i] > logits[best]) best = i;
    return best + 1;
}

int sample_cdf(const float *logits, int n, float temperature, float r, std::vector<float> &p)
{
    p.resize(n);
    float mx = -INFINITY;
    for (int i = 0; i < n; i++) {
        p[i] = logits[i] / temperature;
        mx = std::fmax(mx, p[i]);
    }
    float sum = 0.f;
    for (int i = 0; i < n; i++) {
        p[i] = std::exp(p[i] - mx);
        sum += p[i];
    }
    float cdf = 0.f;
    for (int i = 0; i < n; i++) {
        cdf += p[i] / sum;
        if (r < cdf) return i + 1;
    }
    return n;  // fallback to the last index (:444)
}

}  // namespace llmhost

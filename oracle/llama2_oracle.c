/*
 * llama2_oracle.c -- CPU restatement of the llm.f90 decode hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker and the CPU
 * baseline ("C restatement of llama2.f90"); nothing under llm/f90_b200/ (the
 * product) may call, link or import it.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or
 * known-answer fixtures for this path, and no Fortran compiler exists in this
 * image, so the reference binary cannot be run here either.  What pins this
 * file instead is (a) the float64 numpy restatement in oracle/oracle_np.py,
 * (b) a cross-check of the canonical (quirk-free) mode against the Hugging
 * Face Llama implementation (tests/golden/make_hf_golden.py) and (c) desk
 * review against the cited reference lines.
 *
 * Every function cites the reference lines (into /root/reference) it follows.
 * All arithmetic is IEEE binary32 like the reference's `wp = kind(1.0)`
 * (weight_module.f90:4).  Indices across this API are 1-based for `token` and
 * `pos`, exactly like the Fortran caller (llama2.f90:376-380).
 *
 * Weight layout = weight_module.f90:13-26 seen from C (column-major Fortran
 * arrays are row-major C arrays with the index order reversed):
 *   token_embedding_table(emb,V)      -> [V][emb]
 *   rms_att_weight(emb,L)             -> [L][emb]
 *   wqkv(emb, emb+2kv, L)             -> [L][emb+2kv][emb]  rows: Wq | Wk | Wv
 *   wo(emb,emb,L)                     -> [L][emb][emb]
 *   rms_ffn_weight(emb,L)             -> [L][emb]
 *   w13(emb, 2hid, L)                 -> [L][2hid][emb]     rows: W1(gate) | W3(up)
 *   w2(hid, emb, L)                   -> [L][emb][hid]
 *   rms_final_weight(emb)             -> [emb]
 *   wcls(emb,V)                       -> [V][emb]
 * For wtype 1 (f16) a "row" of n weights is n little-endian binary16 values;
 * for wtype 2 (q4_0) it is n/32 ggml blocks of 18 bytes (f16 scale d, then 16
 * bytes: low nibbles = elements 0..15, high nibbles = elements 16..31, value
 * d*(q-8)).  Both are dequantised exactly to f32 element by element and then
 * go through the same f32 arithmetic as wtype 0.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_F32 0
#define ORACLE_F16 1
#define ORACLE_Q4_0 2
#define ORACLE_Q6_K 14 /* ggml type 14: the output.weight of stock llama.cpp q4_0 files (classifier only) */

typedef struct {
    int emb_dim, hidden_dim, n_layers, n_heads, n_kv_heads, vocab_size, seq_len;
    int wtype;      /* 0 f32, 1 f16, 2 q4_0 (2-D tensors only; norm vectors are always f32) */
    int canonical;  /* 0 = reference behaviour (quirks Q1,Q2 on); 1 = canonical llama RoPE
                       (exponent 2j/hs, 0-based angle) -- used only to cross-check against HF */
    int n_threads;  /* 1 = like the reference (single core, README.md:21) */
} oracle_cfg;

typedef struct {
    oracle_cfg c;
    const void *tok_emb, *wqkv, *wo, *w13, *w2, *wcls;
    const float *rms_att, *rms_ffn, *rms_final;
    /* RunState, weight_module.f90:33-40 */
    float *att;         /* [H][seq] */
    float *key_cache;   /* [L][seq][kv] */
    float *value_cache; /* [L][seq][kv] */
    double times[5];    /* ms per bucket, llama2.f90:526-638 */
    /* scratch */
    float *x, *xb, *qkv, *hb13, *row;
    int cls_type;       /* storage of wcls: c.wtype, or ORACLE_Q6_K (oracle_set_cls_type) */
} oracle_model;

static float f16_table[65536];
static int f16_table_ready = 0;

static float half_bits_to_float(uint16_t h)
{
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1fu;
    uint32_t man = h & 0x3ffu;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else { /* subnormal: renormalise */
            int e = -1;
            do { man <<= 1; e++; } while (!(man & 0x400u));
            man &= 0x3ffu;
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | (man << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7f800000u | (man << 13);
    } else {
        bits = sign | ((exp + 127 - 15) << 23) | (man << 13);
    }
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

static void f16_init(void)
{
    if (f16_table_ready) return;
    for (uint32_t i = 0; i < 65536; i++) f16_table[i] = half_bits_to_float((uint16_t)i);
    f16_table_ready = 1;
}

size_t oracle_row_bytes(int wtype, int n)
{
    if (wtype == ORACLE_F32) return (size_t)n * 4;
    if (wtype == ORACLE_F16) return (size_t)n * 2;
    if (wtype == ORACLE_Q6_K) return (size_t)(n / 256) * 210;
    return (size_t)(n / 32) * 18;
}

/* Exact dequantisation of one row of n weights to f32 (q4_0 per
 * the public ggml block format, cross-checked against gguf.quants in tests). */
void oracle_dequant_row(const void *src, int wtype, int n, float *dst)
{
    if (wtype == ORACLE_F32) {
        memcpy(dst, src, (size_t)n * 4);
    } else if (wtype == ORACLE_F16) {
        f16_init();
        const uint16_t *h = (const uint16_t *)src;
        for (int i = 0; i < n; i++) dst[i] = f16_table[h[i]];
    } else if (wtype == ORACLE_Q6_K) {
        /* ggml block_q6_K, 256 weights in 210 bytes: ql[128] low 4 bits, qh[64] high 2 bits, 16 int8
         * sub-block scales, f16 super-scale d; weight = d * scale * (q - 32) with q the 6-bit value
         * (public ggml format: ggml-quants.c dequantize_row_q6_K; checked against gguf.quants in
         * tests/test_oracle.py).  d * scale * q has at most 11 + 7 + 6 = 24 significant bits: exact in f32. */
        f16_init();
        const uint8_t *b = (const uint8_t *)src;
        for (int blk = 0; blk < n / 256; blk++, b += 210) {
            const uint8_t *ql = b, *qh = b + 128;
            const int8_t *sc = (const int8_t *)(b + 192);
            uint16_t dh;
            memcpy(&dh, b + 208, 2);
            const float d = f16_table[dh];
            for (int half = 0; half < 2; half++, dst += 128, ql += 64, qh += 32, sc += 8) {
                for (int l = 0; l < 32; l++) {
                    const int is = l / 16;
                    const int q1 = (int)((ql[l] & 0xF) | (((qh[l] >> 0) & 3) << 4)) - 32;
                    const int q2 = (int)((ql[l + 32] & 0xF) | (((qh[l] >> 2) & 3) << 4)) - 32;
                    const int q3 = (int)((ql[l] >> 4) | (((qh[l] >> 4) & 3) << 4)) - 32;
                    const int q4 = (int)((ql[l + 32] >> 4) | (((qh[l] >> 6) & 3) << 4)) - 32;
                    dst[l] = d * (float)sc[is] * (float)q1;
                    dst[l + 32] = d * (float)sc[is + 2] * (float)q2;
                    dst[l + 64] = d * (float)sc[is + 4] * (float)q3;
                    dst[l + 96] = d * (float)sc[is + 6] * (float)q4;
                }
            }
        }
    } else {
        f16_init();
        const uint8_t *b = (const uint8_t *)src;
        for (int blk = 0; blk < n / 32; blk++, b += 18, dst += 32) {
            uint16_t dh;
            memcpy(&dh, b, 2);
            const float d = f16_table[dh];
            for (int j = 0; j < 16; j++) {
                dst[j] = d * (float)((int)(b[2 + j] & 0x0f) - 8);
                dst[j + 16] = d * (float)((int)(b[2 + j] >> 4) - 8);
            }
        }
    }
}

static double now_ms(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

/* dot_product(a, b) over n f32 values -- the Fortran intrinsic used at
 * llama2.f90:454,530,582,604,611,619,635.  Plain loop; the compiler is free to
 * vectorise it under -ffast-math exactly as gfortran is (Makefile:7). */
static inline float dotf(const float *a, const float *b, int n)
{
    float s = 0.0f;
    for (int i = 0; i < n; i++) s += a[i] * b[i];
    return s;
}

/* llama2.f90:450-457  rmsnorm: xn = sqrt(dot(x,x)/n + 1e-5); xr = x*w/xn */
void oracle_rmsnorm(const float *x, const float *w, int n, float *out)
{
    const float xn = sqrtf(dotf(x, x, n) / (float)n + 1e-5f);
    for (int i = 0; i < n; i++) out[i] = x[i] * w[i] / xn;
}

/* llama2.f90:468-478  softmax over the first s entries of x(1:n); zeros after */
void oracle_softmax(const float *x, int n, int s, float *p)
{
    float mx = x[0];
    for (int i = 1; i < s; i++) if (x[i] > mx) mx = x[i];
    float sum = 0.0f;
    for (int i = 0; i < s; i++) { p[i] = expf(x[i] - mx); sum += p[i]; }
    for (int i = 0; i < s; i++) p[i] = p[i] / sum;
    for (int i = s; i < n; i++) p[i] = 0.0f;
}

/* The inline mat-vec loops: y(ix) = dot_product(x, w(:,ix))  for ix = 1..rows
 * (llama2.f90:529-531, 603-605, 610-612, 618-620, 634-636).  `scratch` holds
 * one dequantised row when wtype != f32. */
static void matvec_rows(const void *w, int wtype, int rows, int cols, const float *x, float *y,
                        int n_threads)
{
    const size_t rb = oracle_row_bytes(wtype, cols);
    (void)n_threads;
#ifdef _OPENMP
#pragma omp parallel num_threads(n_threads) if (n_threads > 1)
#endif
    {
        float *tmp = NULL;
        if (wtype != ORACLE_F32) tmp = (float *)malloc((size_t)cols * 4);
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
        for (int r = 0; r < rows; r++) {
            const char *src = (const char *)w + (size_t)r * rb;
            if (wtype == ORACLE_F32) {
                y[r] = dotf(x, (const float *)src, cols);
            } else {
                oracle_dequant_row(src, wtype, cols, tmp);
                y[r] = dotf(x, tmp, cols);
            }
        }
        free(tmp);
    }
}

void oracle_matvec(const void *w, int wtype, int rows, int cols, const float *x, float *y)
{
    matvec_rows(w, wtype, rows, cols, x, y, 1);
}

/* llama2.f90:543-559  RoPE on interleaved pairs.
 * Q1: the loop index i is 1-based and odd, head_dim = mod(i, head_size), so the
 *     exponent is (2j+1)/hs, not the canonical 2j/hs.
 * Q2: the angle is pos*freq with the 1-based pos.
 * k is rotated for pairs with i < kv_head_size, i.e. all of k. */
void oracle_rope(float *q, float *k, int emb, int kv, int head_size, int pos, int canonical)
{
    for (int i = 1; i <= emb; i += 2) {
        const int head_dim = canonical ? ((i - 1) % head_size) : (i % head_size);
        const float freq = 1.0f / powf(10000.0f, (float)head_dim / (float)head_size);
        const float rval = (float)(canonical ? pos - 1 : pos) * freq;
        const float fcr = cosf(rval), fci = sinf(rval);
        const float q0 = q[i - 1], q1 = q[i];
        q[i - 1] = q0 * fcr - q1 * fci;
        q[i] = q0 * fci + q1 * fcr;
        if (i < kv) {
            const float k0 = k[i - 1], k1 = k[i];
            k[i - 1] = k0 * fcr - k1 * fci;
            k[i] = k0 * fci + k1 * fcr;
        }
    }
}

oracle_model *oracle_create(const oracle_cfg *cfg, const void *tok_emb, const float *rms_att,
                            const void *wqkv, const void *wo, const float *rms_ffn,
                            const void *w13, const void *w2, const float *rms_final,
                            const void *wcls)
{
    oracle_model *m = (oracle_model *)calloc(1, sizeof(oracle_model));
    m->c = *cfg;
    if (m->c.n_threads < 1) m->c.n_threads = 1;
    m->tok_emb = tok_emb; m->rms_att = rms_att; m->wqkv = wqkv; m->wo = wo;
    m->rms_ffn = rms_ffn; m->w13 = w13; m->w2 = w2; m->rms_final = rms_final; m->wcls = wcls;
    m->cls_type = cfg->wtype;
    const int hs = cfg->emb_dim / cfg->n_heads;
    const int kv = cfg->n_kv_heads * hs;
    /* llama2.f90:311-319: caches allocated at seq_len and zeroed */
    m->att = (float *)calloc((size_t)cfg->n_heads * cfg->seq_len, 4);
    m->key_cache = (float *)calloc((size_t)cfg->n_layers * cfg->seq_len * kv, 4);
    m->value_cache = (float *)calloc((size_t)cfg->n_layers * cfg->seq_len * kv, 4);
    m->x = (float *)malloc((size_t)cfg->emb_dim * 4);
    m->xb = (float *)malloc((size_t)cfg->emb_dim * 4);
    m->qkv = (float *)malloc((size_t)(cfg->emb_dim + 2 * kv) * 4);
    m->hb13 = (float *)malloc((size_t)2 * cfg->hidden_dim * 4);
    const int mx = cfg->emb_dim > cfg->hidden_dim ? cfg->emb_dim : cfg->hidden_dim;
    m->row = (float *)malloc((size_t)mx * 4);
    f16_init();
    return m;
}

/* the classifier tensor is stored in another ggml type than the rest (Q6_K output.weight of a q4_0 file) */
void oracle_set_cls_type(oracle_model *m, int type) { m->cls_type = type; }

void oracle_free(oracle_model *m)
{
    if (!m) return;
    free(m->att); free(m->key_cache); free(m->value_cache);
    free(m->x); free(m->xb); free(m->qkv); free(m->hb13); free(m->row);
    free(m);
}

void oracle_reset(oracle_model *m)
{
    const int hs = m->c.emb_dim / m->c.n_heads;
    const int kv = m->c.n_kv_heads * hs;
    memset(m->att, 0, (size_t)m->c.n_heads * m->c.seq_len * 4);
    memset(m->key_cache, 0, (size_t)m->c.n_layers * m->c.seq_len * kv * 4);
    memset(m->value_cache, 0, (size_t)m->c.n_layers * m->c.seq_len * kv * 4);
    for (int i = 0; i < 5; i++) m->times[i] = 0;
}

void oracle_times(const oracle_model *m, double t[5])
{
    for (int i = 0; i < 5; i++) t[i] = m->times[i];
}

/* llama2.f90:480-640  transformer(token, pos, s, w) -> logits(vocab) */
int oracle_transformer(oracle_model *m, int token, int pos, float *logits)
{
    const oracle_cfg *c = &m->c;
    const int emb = c->emb_dim, hid = c->hidden_dim, L = c->n_layers, H = c->n_heads;
    const int V = c->vocab_size, seq = c->seq_len, wt = c->wtype, nt = c->n_threads;
    const int hs = emb / H;                 /* :515 */
    const int kv = c->n_kv_heads * hs;      /* :154 */
    const int kv_mul = H / c->n_kv_heads;   /* :572 */
    const int nqkv = emb + 2 * kv;
    if (token < 1 || token > V || pos < 1 || pos > seq) return 1;
    const size_t rb_emb = oracle_row_bytes(wt, emb), rb_hid = oracle_row_bytes(wt, hid);
    float *x = m->x, *xb = m->xb, *qkv = m->qkv, *hb13 = m->hb13;
    double t;

    /* :520  x = token_embedding_table(:,token) */
    oracle_dequant_row((const char *)m->tok_emb + (size_t)(token - 1) * rb_emb, wt, emb, x);

    for (int l = 0; l < L; l++) {
        /* :527-531  rmsnorm + fused QKV mat-vec */
        t = now_ms();
        oracle_rmsnorm(x, m->rms_att + (size_t)l * emb, emb, xb);
        matvec_rows((const char *)m->wqkv + (size_t)l * nqkv * rb_emb, wt, nqkv, emb, xb, qkv, nt);
        float *q = qkv, *k = qkv + emb, *v = qkv + emb + kv; /* :533-535 */
        m->times[0] += now_ms() - t;

        /* :543-559 */
        t = now_ms();
        oracle_rope(q, k, emb, kv, hs, pos, c->canonical);
        m->times[1] += now_ms() - t;

        /* :564-565  cache k and v for this position */
        float *kc = m->key_cache + ((size_t)l * seq + (pos - 1)) * kv;
        float *vc = m->value_cache + ((size_t)l * seq + (pos - 1)) * kv;
        memcpy(kc, k, (size_t)kv * 4);
        memcpy(vc, v, (size_t)kv * 4);

        /* :574-598  attention.  Q3: effective kv head of query head h is h/kv_mul
         * (hs contiguous floats from the slice's lower bound), SURVEY.md 8a. */
        t = now_ms();
        const float sqrt_hs = sqrtf((float)hs);
        for (int h = 0; h < H; h++) {
            const float *q_t = q + (size_t)h * hs;
            const int kvh = h / kv_mul;
            float *att = m->att + (size_t)h * seq;
            for (int tt = 0; tt < pos; tt++) {
                const float *k_t = m->key_cache + ((size_t)l * seq + tt) * kv + (size_t)kvh * hs;
                att[tt] = dotf(q_t, k_t, hs) / sqrt_hs; /* :582 */
            }
            oracle_softmax(att, seq, pos, att); /* :586 */
            float *xbh = xb + (size_t)h * hs;
            for (int i = 0; i < hs; i++) xbh[i] = 0.0f;
            for (int tt = 0; tt < pos; tt++) {
                const float *v_t = m->value_cache + ((size_t)l * seq + tt) * kv + (size_t)kvh * hs;
                const float a = att[tt];
                for (int i = 0; i < hs; i++) xbh[i] += a * v_t[i]; /* :593 */
            }
        }
        m->times[2] += now_ms() - t;

        /* :603-620  Wo + residual, FFN */
        t = now_ms();
        matvec_rows((const char *)m->wo + (size_t)l * emb * rb_emb, wt, emb, emb, xb, m->row, nt);
        for (int i = 0; i < emb; i++) x[i] += m->row[i];
        oracle_rmsnorm(x, m->rms_ffn + (size_t)l * emb, emb, xb);
        matvec_rows((const char *)m->w13 + (size_t)l * 2 * hid * rb_emb, wt, 2 * hid, emb, xb, hb13, nt);
        float *hb = hb13, *hb2 = hb13 + hid;
        for (int i = 0; i < hid; i++) hb[i] = hb[i] * (1.0f / (1.0f + expf(-hb[i]))); /* :615 */
        for (int i = 0; i < hid; i++) hb[i] = hb[i] * hb2[i];                          /* :616 */
        matvec_rows((const char *)m->w2 + (size_t)l * emb * rb_hid, wt, emb, hid, hb, m->row, nt);
        for (int i = 0; i < emb; i++) x[i] += m->row[i];
        m->times[3] += now_ms() - t;
    }

    /* :627-636  final norm + classifier */
    t = now_ms();
    oracle_rmsnorm(x, m->rms_final, emb, x);
    matvec_rows(m->wcls, m->cls_type, V, emb, x, logits, nt);
    m->times[4] += now_ms() - t;
    return 0;
}

/* maxloc(logits, DIM=1): first maximum wins; returns a 1-based index (llama2.f90:388) */
int oracle_argmax1(const float *v, int n)
{
    int best = 0;
    for (int i = 1; i < n; i++) if (v[i] > v[best]) best = i;
    return best + 1;
}

/* llama2.f90:376-402 with temperature == 0: BOS = 2, prompt tokens forced, greedy
 * afterwards.  out_tokens[pos-1] = the token chosen AFTER the forward at pos.
 * logits_out (optional) receives all n logit vectors.  Returns elapsed ms from
 * after the first token to the end, like :399-406. */
double oracle_generate(oracle_model *m, const int *prompt_tokens, int n_prompt, int n,
                       int *out_tokens, float *logits_out)
{
    const int V = m->c.vocab_size;
    float *logits = (float *)malloc((size_t)V * 4);
    int token = 2;
    double t_start = 0, t_end;
    for (int pos = 1; pos <= n; pos++) {
        oracle_transformer(m, token, pos, logits);
        if (logits_out) memcpy(logits_out + (size_t)(pos - 1) * V, logits, (size_t)V * 4);
        if (pos <= n_prompt) token = prompt_tokens[pos - 1];
        else token = oracle_argmax1(logits, V);
        out_tokens[pos - 1] = token;
        if (t_start == 0) t_start = now_ms();
    }
    t_end = now_ms();
    free(logits);
    return t_end - t_start;
}

int oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

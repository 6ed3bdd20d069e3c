"""The oracle against everything that can pin it without the (uncompilable) Fortran binary:
the float64 numpy restatement, the Hugging Face golden vectors (canonical mode), the gguf
package's Q4_0 codec and hand-derived values for the reference's quirks."""
import math
import os

import numpy as np
import pytest

from conftest import rel_err
from llm.f90_b200 import fixtures as fx
from llm.f90_b200.layout import Config, TINY, SMALL, F32, F16, Q4_0, row_bytes, active_weight_bytes, \
    TINYLLAMA, LLAMA2_7B
from oracle import oracle_c as oc
from oracle.oracle_np import OracleNP

GOLD_DIR = os.path.join(os.path.dirname(__file__), "golden")
GOLD = os.path.join(GOLD_DIR, "hf_tiny_logits.npz")
# the head geometries of tests/golden/make_hf_golden.py (CASES): TINY; TinyLlama's 8 query heads per KV
# head (quirk Q3's case); Llama-2-7B's multi-head attention with head size 128
GOLD_SHAPES = {
    "tiny": TINY,
    "gqa8": dict(emb_dim=256, hidden_dim=704, n_layers=2, n_heads=8, n_kv_heads=1, vocab_size=512, seq_len=64),
    "mha128": dict(emb_dim=256, hidden_dim=704, n_layers=2, n_heads=2, n_kv_heads=2, vocab_size=512, seq_len=64),
}


@pytest.mark.parametrize("name", list(GOLD_SHAPES))
def test_hf_golden_canonical_mode(name):
    """Structure pin: canonical-RoPE oracle == Hugging Face LlamaForCausalLM (f32) to 5e-6."""
    g = np.load(os.path.join(GOLD_DIR, f"hf_{name}_logits.npz"))
    cfg = Config(**GOLD_SHAPES[name], wtype=F32)
    w = fx.synth_weights(cfg, int(g["seed"]))
    o = oc.Oracle(w, canonical=True)
    n = OracleNP(w, canonical=True)
    for i, t in enumerate(g["tokens_0based"]):
        lg = o.transformer(int(t) + 1, i + 1)
        assert rel_err(lg, g["logits"][i]) < 5e-6
        assert rel_err(n.transformer(int(t) + 1, i + 1), g["logits"][i]) < 5e-6


def test_reference_mode_differs_from_canonical():
    """The quirks are real: reference-mode logits must NOT equal the canonical ones after pos 1."""
    g = np.load(GOLD)
    cfg = Config(**TINY, wtype=F32)
    w = fx.synth_weights(cfg, int(g["seed"]))
    o = oc.Oracle(w)
    errs = [rel_err(o.transformer(int(t) + 1, i + 1), g["logits"][i]) for i, t in enumerate(g["tokens_0based"])]
    assert max(errs[1:]) > 1e-3


@pytest.mark.parametrize("wt", [F32, F16, Q4_0])
@pytest.mark.parametrize("shape", [TINY, SMALL])
def test_c_oracle_matches_float64(wt, shape):
    cfg = Config(**shape, wtype=wt)
    w = fx.synth_weights(cfg, seed=7)
    prompt = [5, 6, 7, 8, 9]
    n = 14
    toks, lg, _ = oc.Oracle(w).generate(prompt, n, want_logits=True)
    toks64, lg64 = OracleNP(w).generate(prompt, n)
    assert (toks == toks64).all()
    assert rel_err(lg, lg64) < 5e-6


def test_oracle_threads_do_not_change_results():
    cfg = Config(**SMALL, wtype=F32)
    w = fx.synth_weights(cfg, seed=1)
    a = oc.Oracle(w, n_threads=1).generate([3, 4], 8, want_logits=True)
    b = oc.Oracle(w, n_threads=4).generate([3, 4], 8, want_logits=True)
    assert (a[0] == b[0]).all() and np.array_equal(a[1], b[1])


def test_rmsnorm_formula():
    rng = np.random.default_rng(0)
    x = rng.standard_normal(257).astype(np.float32)
    w = rng.standard_normal(257).astype(np.float32)
    ref = x.astype(np.float64) * w / math.sqrt(float(x.astype(np.float64) @ x) / 257 + 1e-5)
    assert rel_err(oc.rmsnorm(x, w), ref) < 1e-6


def test_softmax_prefix_and_zero_tail():
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(64) * 5).astype(np.float32)
    p = oc.softmax(x, 10)
    e = np.exp(x[:10].astype(np.float64) - x[:10].max())
    assert rel_err(p[:10], e / e.sum()) < 1e-6
    assert (p[10:] == 0).all()
    assert abs(float(p.sum()) - 1) < 1e-6


def test_rope_quirks_q1_q2():
    """Q1: exponent (2j+1)/hs; Q2: angle = pos*freq with the 1-based pos (llama2.f90:543-548)."""
    hs, emb, kv = 8, 16, 8
    q = np.arange(1, emb + 1, dtype=np.float32)
    k = np.arange(1, kv + 1, dtype=np.float32) * 0.5
    pos = 3
    q2, k2 = oc.rope(q, k, hs, pos)
    for i in range(0, emb, 2):
        j = (i // 2) % (hs // 2)
        ang = pos * 10000.0 ** (-(2 * j + 1) / hs)
        c, s = math.cos(ang), math.sin(ang)
        assert abs(q2[i] - (q[i] * c - q[i + 1] * s)) < 1e-5
        assert abs(q2[i + 1] - (q[i] * s + q[i + 1] * c)) < 1e-5
        if i < kv:
            assert abs(k2[i] - (k[i] * c - k[i + 1] * s)) < 1e-5
    qc, _ = oc.rope(q, k, hs, 1, canonical=True)  # canonical position 0 -> identity
    assert np.allclose(qc, q)


def test_gqa_head_mapping_q3():
    """Q3: query head h reads kv head h // kv_mul.  Make kv heads distinguishable via Wv."""
    cfg = Config(emb_dim=32, hidden_dim=32, n_layers=1, n_heads=4, n_kv_heads=2, vocab_size=32, seq_len=8)
    t = fx.synth_tensors(cfg, 0)
    w = fx.fuse_tensors(cfg, t)
    a = OracleNP(w).transformer(3, 1)
    b = oc.Oracle(w).transformer(3, 1)
    assert rel_err(b, a) < 5e-6


def test_q4_0_codec_matches_gguf_package():
    gguf = pytest.importorskip("gguf")
    rng = np.random.default_rng(5)
    w = rng.standard_normal((6, 128)).astype(np.float32)
    ours = fx.quantize_q4_0(w)
    theirs = gguf.quants.quantize(w, gguf.GGMLQuantizationType.Q4_0)
    assert np.array_equal(ours, theirs.reshape(ours.shape))
    deq = gguf.quants.dequantize(theirs, gguf.GGMLQuantizationType.Q4_0)
    assert np.array_equal(fx.dequantize_q4_0(ours, 128), deq.reshape(6, 128))
    for r in range(6):
        assert np.array_equal(oc.dequant_row(ours[r], Q4_0, 128), deq.reshape(6, 128)[r])


def test_f16_dequant_exact():
    rng = np.random.default_rng(2)
    h = rng.standard_normal(64).astype(np.float16)
    assert np.array_equal(oc.dequant_row(h, F16, 64), h.astype(np.float32))


@pytest.mark.parametrize("wt", [F32, F16, Q4_0])
def test_matvec_operator(wt):
    rng = np.random.default_rng(3)
    rows, cols = 37, 96
    wf = rng.standard_normal((rows, cols)).astype(np.float32)
    enc = fx.encode_matrix(wf, wt)
    x = rng.standard_normal(cols).astype(np.float32)
    ref = fx.decode_matrix(enc, wt, cols).astype(np.float64) @ x
    assert rel_err(oc.matvec(enc, wt, rows, cols, x), ref) < 1e-6


def test_argmax_first_maximum_wins():
    v = np.array([1, 5, 5, 2, 5], np.float32)
    assert oc.lib().oracle_argmax1(oc._fp(v), 5) == 2  # maxloc -> 1-based first max (llama2.f90:388)


def test_roofline_numerators_match_baseline_md():
    """BASELINE.md section 2 / SURVEY.md 8a: algorithmic bytes per token."""
    assert active_weight_bytes(Config(**TINYLLAMA, wtype=F32)) == 4_138_057_728
    assert active_weight_bytes(Config(**TINYLLAMA, wtype=F16)) == 2_069_213_184
    assert active_weight_bytes(Config(**LLAMA2_7B, wtype=Q4_0)) == 3_717_548_288
    assert active_weight_bytes(Config(**LLAMA2_7B, wtype=F16)) == 13_215_227_904
    assert row_bytes(Q4_0, 4096) == 4096 // 32 * 18


def test_forced_prompt_then_greedy():
    """llama2.f90:376-393: BOS=2 fed first, prompt tokens forced, then maxloc."""
    cfg = Config(**TINY, wtype=F32)
    w = fx.synth_weights(cfg, seed=9)
    o = oc.Oracle(w)
    toks, lg, _ = o.generate([40, 41, 42], 8, want_logits=True)
    assert list(toks[:3]) == [40, 41, 42]
    assert all(int(np.argmax(lg[i])) + 1 == toks[i] for i in range(3, 8))
    o2 = oc.Oracle(w)
    assert np.array_equal(o2.transformer(2, 1), lg[0])
    assert np.array_equal(o2.transformer(40, 2), lg[1])


# ------------------------------------------------------------------ Q6_K classifier
def test_q6_k_codec_matches_gguf_quants():
    """The oracle's and the fixtures' Q6_K dequantisation against the `gguf` package's (a third implementation of
    the public ggml block format), bit for bit -- on quantised rows and on arbitrary block bytes."""
    from gguf import quants, GGMLQuantizationType as T
    from llm.f90_b200.layout import Q6_K
    rng = np.random.default_rng(6)
    x = (rng.standard_normal((9, 768)) * 0.05).astype(np.float32)
    q = fx.quantize_q6_k(x)
    ref = quants.dequantize(q, T.Q6_K)
    assert np.array_equal(fx.dequantize_q6_k(q, 768), ref)
    assert np.array_equal(np.stack([oc.dequant_row(q[i], Q6_K, 768) for i in range(9)]), ref)
    assert np.abs(ref - x).max() < 0.03 * np.abs(x).max()  # it is a 6-bit quantiser
    raw = rng.integers(0, 256, (6, 630), dtype=np.uint8)
    for b in range(3):  # finite super-scales
        raw[:, 210 * b + 208:210 * b + 210] = np.array([0.37 * (b + 1) * (-1) ** b], np.float16).view(np.uint8)
    ref = quants.dequantize(raw, T.Q6_K)
    assert np.array_equal(fx.dequantize_q6_k(raw, 768), ref)
    assert np.array_equal(np.stack([oc.dequant_row(raw[i], Q6_K, 768) for i in range(6)]), ref)


def test_oracle_q6_k_classifier_equals_the_dequantised_f32_classifier():
    """Q6_K dequantises exactly to f32, so a model with a Q6_K output.weight must give the very logits of the same
    model with that tensor stored as the dequantised f32 values."""
    from llm.f90_b200.layout import Q6_K
    cfg = Config(emb_dim=256, hidden_dim=352, n_layers=2, n_heads=4, n_kv_heads=2, vocab_size=300, seq_len=32, wtype=F32)
    t = fx.synth_tensors(cfg, 4)
    wq = fx.fuse_tensors(cfg, t, cls_wtype=Q6_K)
    t2 = dict(t)
    t2["output.weight"] = fx.dequantize_q6_k(wq.wcls.reshape(cfg.vocab_size, -1), cfg.emb_dim)
    wf = fx.fuse_tensors(cfg, t2)
    a, b = oc.Oracle(wq), oc.Oracle(wf)
    for pos, tok in enumerate([2, 17, 250], 1):
        assert np.array_equal(a.transformer(tok, pos), b.transformer(tok, pos))


def test_q6_k_golden_vectors():
    """The committed gguf.quants vectors (tests/golden/make_q6k_golden.py): no third-party package needed at test time."""
    import os
    from llm.f90_b200.layout import Q6_K
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "q6k_golden.npz"))
    blocks, values = g["blocks"], g["values"]
    n = values.shape[1]
    assert np.array_equal(fx.dequantize_q6_k(blocks, n), values)
    assert np.array_equal(np.stack([oc.dequant_row(np.ascontiguousarray(blocks[i]), Q6_K, n) for i in range(len(blocks))]), values)

"""Parity tests proper: the sm_100a kernels, called through the C ABI exactly as the Fortran
host would call them, against the CPU oracle on the same seeded inputs.

Tolerances are the north star's: logits within 1e-4 relative (max|diff| / max|ref|) for f32 and
f16 storage (f32 arithmetic on exactly dequantised weights), 1e-2 for q4_0; greedy token ids
identical.  The f32 budget covers summation-order noise only: the reference itself is built
with -ffast-math (Makefile:7), so neither side has a canonical order.
"""
import numpy as np
import pytest

from conftest import rel_err
from llm.f90_b200 import capi, fixtures as fx
from llm.f90_b200.layout import Config, TINY, SMALL, F32, F16, Q4_0
from oracle import oracle_c as oc

pytestmark = pytest.mark.gpu

TOL = {F32: 1e-4, F16: 1e-4, Q4_0: 1e-2}
# a mid-size shape with TinyLlama's head geometry (hs 64, kv_mul 8) and a multi-slot ring
MID = dict(emb_dim=1024, hidden_dim=2816, n_layers=4, n_heads=16, n_kv_heads=2, vocab_size=4096, seq_len=512)
# llama-2-7B head geometry (hs 128, kv_mul 1)
MHA = dict(emb_dim=512, hidden_dim=1376, n_layers=2, n_heads=4, n_kv_heads=4, vocab_size=1024, seq_len=128)


@pytest.fixture(scope="module", autouse=True)
def _built(built):
    capi.load()


# ------------------------------------------------------------------ operators
@pytest.mark.parametrize("n", [4, 64, 2048, 4096, 11008, 333])
def test_rmsnorm(n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n).astype(np.float32)
    w = (1 + 0.1 * rng.standard_normal(n)).astype(np.float32)
    assert rel_err(capi.rmsnorm(x, w), oc.rmsnorm(x, w)) < 1e-6


@pytest.mark.parametrize("n,s", [(1, 1), (64, 1), (64, 64), (2048, 97), (2048, 2048), (5000, 4999)])
def test_softmax(n, s):
    rng = np.random.default_rng(n + s)
    x = (4 * rng.standard_normal(n)).astype(np.float32)
    p, ref = capi.softmax(x, s), oc.softmax(x, s)
    assert rel_err(p, ref) < 1e-6
    assert (p[s:] == 0).all()


@pytest.mark.parametrize("emb,kv,hs,pos", [(128, 64, 32, 1), (2048, 256, 64, 97), (4096, 4096, 128, 2048),
                                            (512, 128, 64, 1000)])
def test_rope(emb, kv, hs, pos):
    rng = np.random.default_rng(pos)
    q = rng.standard_normal(emb).astype(np.float32)
    k = rng.standard_normal(kv).astype(np.float32)
    q1, k1 = capi.rope(q, k, hs, pos)
    q2, k2 = oc.rope(q, k, hs, pos)
    # the angle pos*freq reaches ~2000 rad: f32 sin/cos argument rounding is 2000*6e-8 ~ 1e-4 abs
    # in the worst case on BOTH sides; compare with the f64 truth instead of each other there
    assert np.abs(q1 - q2).max() < 2e-4 * max(1.0, pos / 100)
    assert np.abs(k1 - k2).max() < 2e-4 * max(1.0, pos / 100)


@pytest.mark.parametrize("wt", [F32, F16, Q4_0])
@pytest.mark.parametrize("rows,cols", [(1, 32), (7, 64), (37, 96), (2560, 2048), (2048, 5632), (1000, 4096),
                                        (333, 11008)])
def test_matvec(wt, rows, cols):
    rng = np.random.default_rng(rows * 7 + cols)
    wf = (rng.standard_normal((rows, cols)) / np.sqrt(cols)).astype(np.float32)
    enc = fx.encode_matrix(wf, wt)
    x = rng.standard_normal(cols).astype(np.float32)
    ref = fx.decode_matrix(enc, wt, cols).astype(np.float64) @ x.astype(np.float64)
    got = capi.matvec(enc, wt, rows, cols, x)
    assert rel_err(got, ref) < (2e-5 if wt == Q4_0 else 2e-6)
    assert rel_err(oc.matvec(enc, wt, rows, cols, x), ref) < 2e-6


def test_operator_errors():
    x = np.ones(8, np.float32)
    with pytest.raises(capi.EngineError):
        capi.softmax(x, 9)
    with pytest.raises(capi.EngineError):
        capi.matvec(np.ones((2, 6), np.float32), F32, 2, 6, np.ones(6, np.float32))  # cols % 4
    with pytest.raises(capi.EngineError):
        capi.rope(x, x, 3, 1)  # odd head size


# ------------------------------------------------------------------ the forward
def run_both(cfg, seed, prompt, n, granular):
    w = fx.synth_weights(cfg, seed)
    ref_toks, ref_lg, _ = oc.Oracle(w).generate(prompt, n, want_logits=True)
    with capi.Engine(w, granular=granular) as eng:
        toks, lg = capi.host_generate(eng, prompt, n, want_logits=True)
        st = eng.stats()
    return ref_toks, ref_lg, toks, lg, st


@pytest.mark.parametrize("granular", [False, True], ids=["stream", "granular"])
@pytest.mark.parametrize("wt", [F32, F16, Q4_0], ids=["f32", "f16", "q4_0"])
@pytest.mark.parametrize("shape", [TINY, SMALL, MHA], ids=["tiny", "small", "mha"])
def test_transformer_matches_oracle(shape, wt, granular):
    cfg = Config(**shape, wtype=wt)
    ref_toks, ref_lg, toks, lg, st = run_both(cfg, 11, [21, 22, 23, 24, 25], 24, granular)
    errs = [rel_err(lg[i], ref_lg[i]) for i in range(len(lg))]
    assert max(errs) < TOL[wt], errs
    assert (toks == ref_toks).all()
    assert st["kernel_launches"] >= 24


@pytest.mark.parametrize("wt", [F32, Q4_0], ids=["f32", "q4_0"])
def test_transformer_mid_shape_long(wt):
    """TinyLlama head geometry, 300 positions: crosses the attention split threshold (256) and uses
    several 32-position blocks per warp."""
    cfg = Config(**MID, wtype=wt)
    ref_toks, ref_lg, toks, lg, _ = run_both(cfg, 5, [100, 200, 300], 300, False)
    errs = [rel_err(lg[i], ref_lg[i]) for i in range(len(lg))]
    assert max(errs) < TOL[wt], (int(np.argmax(errs)), max(errs))
    assert (toks == ref_toks).all()


@pytest.mark.parametrize("granular", [False, True], ids=["stream", "granular"])
@pytest.mark.parametrize("wt", [F32, F16, Q4_0], ids=["f32", "f16", "q4_0"])
def test_device_greedy_loop_matches_host_loop(wt, granular):
    cfg = Config(**SMALL, wtype=wt)
    w = fx.synth_weights(cfg, 2)
    prompt, n = [7, 8, 9], 40
    ref_toks, _, _ = oc.Oracle(w).generate(prompt, n)
    with capi.Engine(w, granular=granular) as eng:
        toks, ms = eng.generate_greedy(prompt, n)
        assert (toks == ref_toks).all()
        assert ms > 0
        eng.reset()
        toks2, _ = capi.host_generate(eng, prompt, n)
        assert (toks2 == ref_toks).all()


def test_reset_restores_initial_state():
    cfg = Config(**TINY, wtype=F32)
    w = fx.synth_weights(cfg, 4)
    with capi.Engine(w) as eng:
        a = eng.transformer(2, 1).copy()
        eng.transformer(9, 2)
        eng.reset()
        b = eng.transformer(2, 1)
        assert np.array_equal(a, b)
        t = eng.times()
        assert t.shape == (5,) and (t >= 0).all() and t.sum() > 0


def test_full_context_position():
    """last position of the cache (pos == seq_len) and the range checks around it"""
    cfg = Config(**TINY, wtype=F32)
    w = fx.synth_weights(cfg, 6)
    o = oc.Oracle(w)
    with capi.Engine(w) as eng:
        tok = 2
        for pos in range(1, cfg.seq_len + 1):
            ref = o.transformer(tok, pos)
            got = eng.transformer(tok, pos)
            assert rel_err(got, ref) < 1e-4
            tok = int(np.argmax(ref)) + 1
        with pytest.raises(capi.EngineError):
            eng.transformer(2, cfg.seq_len + 1)
        with pytest.raises(capi.EngineError):
            eng.transformer(0, 1)
        with pytest.raises(capi.EngineError):
            eng.transformer(cfg.vocab_size + 1, 1)


def test_init_rejects_bad_configs():
    cfg = Config(**TINY, wtype=F32)
    w = fx.synth_weights(cfg, 0)
    L = capi.load()
    cc = capi.CConfig(cfg.emb_dim, cfg.hidden_dim, cfg.n_layers, 5, cfg.n_kv_heads, cfg.vocab_size, cfg.seq_len,
                      0, 0, 0, 1, 0)
    import ctypes as C
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    args = [ptr(getattr(w, f)) for f in w.FIELDS]
    assert L.llmf90_b200_init(C.byref(cc), *args) != 0
    assert b"heads" in L.llmf90_b200_last_error()
    cc.n_heads, cc.wtype = cfg.n_heads, 7
    assert L.llmf90_b200_init(C.byref(cc), *args) != 0


def test_linearity_of_classifier_property():
    """size-independent property: logits are linear in wcls -> doubling wcls doubles logits."""
    cfg = Config(**SMALL, wtype=F32)
    w = fx.synth_weights(cfg, 8)
    with capi.Engine(w) as eng:
        a = eng.transformer(2, 1).copy()
    w.wcls *= 2
    with capi.Engine(w) as eng:
        b = eng.transformer(2, 1)
    assert np.allclose(b, 2 * a, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("wt", [F32, Q4_0], ids=["f32", "q4_0"])
def test_cli_end_to_end(tmp_path, wt):
    """`./llm -m <gguf> -n N -p prompt` (llama2.f90:87-410 mirrored in C++): GGUF loader -> tokenizer ->
    C ABI -> printed tokens, against the oracle driven by the same file and prompt."""
    import subprocess
    from llm.f90_b200 import hostapi
    cfg = Config(**SMALL, wtype=wt)
    p = str(tmp_path / "m.gguf")
    w = fx.write_synth_gguf(p, cfg, seed=12)
    m = hostapi.HostModel(p)
    vocab, _ = m.vocab()
    prompt = "the cat sat"
    ptoks = m.encode(prompt)
    m.close()
    n = 24
    ref_toks, _, _ = oc.Oracle(w).generate(ptoks, n)
    want = b"".join(vocab[t - 1] for t in ref_toks)
    r = subprocess.run([hostapi.LLM_BIN, "-m", p, "-n", str(n), "-p", prompt, "-t", "0"], capture_output=True, timeout=300)
    assert r.returncode == 0, r.stdout[-400:] + r.stderr[-400:]
    out = r.stdout
    assert b"data offset" in out and b"tokens/second" in out and b"Timings" in out
    text = out.split(b"\n Inference time:")[0].split(b"\n", 1)[1]  # after the loader's "data offset" line
    assert text == want


def test_cli_ak_packed_model(tmp_path):
    """`./llm --ak -m model.bin -s tokenizer.bin` (llama2.f90:158-294, :321-356): the legacy packed f32 file
    and the separate tokenizer file drive the same forward pass as the GGUF of the same weights."""
    import subprocess
    from llm.f90_b200 import hostapi
    cfg = Config(**SMALL, wtype=F32)
    t = fx.synth_tensors(cfg, seed=12)
    w = fx.fuse_tensors(cfg, t)
    p, tb = str(tmp_path / "m.bin"), str(tmp_path / "tok.bin")
    fx.write_ak(p, cfg, t)
    vocab, scores = fx.synth_vocab(cfg.vocab_size)
    # tokenizer.bin entries are used as they are (llama2.f90:321-356): write them the way load_ggml leaves
    # GGUF tokens, with a leading U+2581 already turned into a space (read_ggml.f90:483-503)
    vocab = [b" " + t[3:] if t.startswith(b"\xe2\x96\x81") else t for t in vocab]
    fx.write_tokenizer_bin(tb, vocab, scores)
    m = hostapi.HostModel(p, ak=True)
    m.load_tokenizer(tb)
    prompt = "the cat sat"
    ptoks = m.encode(prompt)
    m.close()
    n = 20
    ref_toks, _, _ = oc.Oracle(w).generate(ptoks, n)
    want = b"".join(vocab[t - 1] for t in ref_toks)
    r = subprocess.run([hostapi.LLM_BIN, "--ak", "-m", p, "-s", tb, "-n", str(n), "-p", prompt, "-t", "0"],
                       capture_output=True, timeout=300)
    assert r.returncode == 0, r.stdout[-400:] + r.stderr[-400:]
    assert r.stdout.split(b"\n Inference time:")[0] == want
    # without -s there is no vocabulary: refuse like the reference's other fatal conditions (print + stop)
    r = subprocess.run([hostapi.LLM_BIN, "--ak", "-m", p, "-n", "4"], capture_output=True, timeout=60)
    assert r.returncode != 0 and b"tokenizer" in r.stdout


def test_profile_flag_selects_the_instrumented_kernel():
    """LLMF90_FLAG_PROFILE: same logits bit for bit, per-phase timers filled; without it every forward goes
    to the reference's bucket 4 (llama2.f90:407-410 prints five buckets)."""
    cfg = Config(**SMALL, wtype=F16)
    w = fx.synth_weights(cfg, 9)
    with capi.Engine(w) as eng:
        a = [eng.transformer(5 + i, 1 + i).copy() for i in range(6)]
        assert sum(eng.phase_times().values()) == 0.0
        t = eng.times()
        assert t[3] > 0 and t[0] == 0 and t[4] == 0
    with capi.Engine(w, profile=True) as eng:
        b = [eng.transformer(5 + i, 1 + i).copy() for i in range(6)]
        ph = eng.phase_times()
        assert ph["w13_mv"] > 0 and ph["qkv_pro"] > 0 and ph["cls_mv"] > 0
        t = eng.times()
        # bucket 2 (RoPE + KV append, llama2.f90:543-565) is part of the QKV tiles' epilogue here: it reads 0,
        # like the reference's own transcript (README.md:77-82: 26.67 / 0.0 / 0.0 / 192.0 / 17.33)
        assert (t[[0, 2, 3, 4]] > 0).all() and t[1] >= 0
    assert all(np.array_equal(x, y) for x, y in zip(a, b))

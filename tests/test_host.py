"""The C++ host mirror (GGUF loader, tokenizer, sampler) against the reference's documented
behaviour on synthetic GGUF files."""
import os
import struct
import subprocess

import numpy as np
import pytest

from llm.f90_b200 import fixtures as fx, hostapi
from llm.f90_b200.layout import Config, TINY, SMALL, F32, F16, Q4_0


@pytest.fixture(scope="module", autouse=True)
def _built(built):
    hostapi.load()


def py_bpe(vocab, scores, text: bytes):
    """Independent restatement of llama2.f90:658-724 (linear first-match lookup, strict best score)."""
    def lookup(s):
        for i, t in enumerate(vocab):
            if t == s:
                return i
        return -1
    toks = [lookup(bytes([b])) for b in text]
    assert all(t >= 0 for t in toks)
    while len(toks) >= 2:
        best, best_i, best_id = -1e10, -1, -1
        for i in range(len(toks) - 1):
            ind = lookup(vocab[toks[i]] + vocab[toks[i + 1]])
            if ind >= 0 and scores[ind] > best:
                best, best_i, best_id = scores[ind], i, ind
        if best_i < 0:
            break
        toks[best_i:best_i + 2] = [best_id]
    return [t + 1 for t in toks]


@pytest.mark.parametrize("wt", [F32, F16, Q4_0], ids=["f32", "f16", "q4_0"])
@pytest.mark.parametrize("alignment", [32, 64])
def test_loader_rebuilds_the_fused_layout(tmp_path, wt, alignment):
    cfg = Config(**TINY, wtype=wt)
    p = str(tmp_path / "m.gguf")
    want = fx.write_synth_gguf(p, cfg, seed=3, alignment=alignment)
    m = hostapi.HostModel(p)
    assert m.cfg == cfg
    assert m.data_offset % alignment == 0
    got = m.weights()
    for f in got.FIELDS:
        assert np.array_equal(getattr(got, f).view(np.uint8).ravel(), getattr(want, f).view(np.uint8).ravel()), f
    m.close()


def test_vocab_rewrites_leading_u2581_only(tmp_path):
    cfg = Config(**TINY)
    p = str(tmp_path / "m.gguf")
    fx.write_synth_gguf(p, cfg, seed=0)
    raw, _ = fx.synth_vocab(cfg.vocab_size)
    m = hostapi.HostModel(p)
    toks, scores = m.vocab()
    sp = "▁".encode()
    for r, t in zip(raw, toks):
        assert t == (b" " + r[3:] if r.startswith(sp) else r)
    assert scores[5] == -5.0
    m.close()


def test_bpe_encode_matches_the_reference_algorithm(tmp_path):
    cfg = Config(**SMALL)
    p = str(tmp_path / "m.gguf")
    fx.write_synth_gguf(p, cfg, seed=0)
    m = hostapi.HostModel(p)
    vocab, scores = m.vocab()
    for text in (b"", b"a", b"the rain in spain", b"I stopped posting on knitting forums because",
                 b"  double  spaces", b"tttttt", b"hello, world! 123"):
        assert m.encode(text) == py_bpe(vocab, scores, text), text
    assert m.encode(b"ab") and all(1 <= t <= cfg.vocab_size for t in m.encode(b"some text here"))
    with pytest.raises(hostapi.HostError, match="single-byte"):
        m.encode(b"caf\xc3\xa9")  # bytes without a vocabulary entry: the reference indexes vocab(-1)
    m.close()


def test_loader_errors_follow_the_reference(tmp_path):
    cfg = Config(**TINY)
    t = fx.synth_tensors(cfg, 0)
    p = str(tmp_path / "bad.gguf")
    with open(p, "wb") as f:
        f.write(b"NOPE" + b"\0" * 64)
    with pytest.raises(hostapi.HostError, match="magic"):
        hostapi.HostModel(p)
    t2 = dict(t)
    del t2["output.weight"]  # tied-embedding models fail: a separate classifier is required (read_ggml.f90:406)
    fx.write_gguf(p, cfg, t2)
    with pytest.raises(hostapi.HostError, match="key not found: output.weight"):
        hostapi.HostModel(p)
    with pytest.raises(hostapi.HostError, match="cannot open"):
        hostapi.HostModel(str(tmp_path / "missing.gguf"))


def test_loader_accepts_every_gguf_kv_type(tmp_path):
    """The reference stops on KV types 0-3, 7, 10-12 (read_ggml.f90:682-685); real llama.cpp files
    contain them, so the mirror reads them all."""
    cfg = Config(**TINY)
    p = str(tmp_path / "m.gguf")
    fx.write_synth_gguf(p, cfg, seed=1)
    raw = open(p, "rb").read()
    n_kv = struct.unpack_from("<Q", raw, 16)[0]
    def kv(key, t, payload):
        return struct.pack("<Q", len(key)) + key + struct.pack("<I", t) + payload
    extra = (kv(b"x.u8", 0, b"\x07") + kv(b"x.i8", 1, b"\xff") + kv(b"x.u16", 2, struct.pack("<H", 9)) +
             kv(b"x.i16", 3, struct.pack("<h", -9)) + kv(b"x.bool", 7, b"\x01") + kv(b"x.u64", 10, struct.pack("<Q", 5)) +
             kv(b"x.i64", 11, struct.pack("<q", -5)) + kv(b"x.f64", 12, struct.pack("<d", 2.5)) +
             kv(b"x.arr", 9, struct.pack("<IQ", 2, 3) + struct.pack("<3H", 1, 2, 3)))
    # splice the extra pairs right after the header; the metadata block grows, so re-pad the data offset
    body = raw[24:]
    old = hostapi.HostModel(p)
    off = old.data_offset
    old.close()
    meta = struct.pack("<IIQQ", *struct.unpack_from("<II", raw, 0), struct.unpack_from("<Q", raw, 8)[0], n_kv + 9) + extra + raw[24:off]
    meta = meta.rstrip(b"\0")
    meta += b"\0" * ((-len(meta)) % 32)
    q = str(tmp_path / "m2.gguf")
    with open(q, "wb") as f:
        f.write(meta + raw[off:])
    a, b = hostapi.HostModel(p), hostapi.HostModel(q)
    assert a.cfg == b.cfg
    assert np.array_equal(a.weights().wqkv, b.weights().wqkv)
    a.close(); b.close()


def test_tokenizer_bin_override(tmp_path):
    cfg = Config(**TINY)
    p = str(tmp_path / "m.gguf")
    fx.write_synth_gguf(p, cfg, seed=0)
    tb = str(tmp_path / "tok.bin")
    with open(tb, "wb") as f:
        f.write(struct.pack("<i", 8))
        for i in range(cfg.vocab_size):
            s = b"%c" % (33 + i) if i < 90 else b"tok%d" % i
            f.write(struct.pack("<fi", float(-i), len(s)) + s)
    m = hostapi.HostModel(p)
    m.load_tokenizer(tb)
    toks, scores = m.vocab()
    assert toks[0] == b"!" and toks[100] == b"tok100" and scores[3] == -3.0
    assert m.encode(b"!\"") == [1, 2]
    m.close()


def test_ak_packed_file_loads_into_the_fused_layout(tmp_path):
    """--ak (llama2.f90:158-294): same TransformerWeights as the GGUF path; vocabulary from -s."""
    cfg = Config(**TINY)
    t = fx.synth_tensors(cfg, seed=4)
    p = str(tmp_path / "m.bin")
    fx.write_ak(p, cfg, t)
    m = hostapi.HostModel(p, ak=True)
    assert m.cfg == cfg
    got, ref = m.weights(), fx.fuse_tensors(cfg, t)
    for f in ref.FIELDS:
        assert np.array_equal(getattr(got, f).reshape(-1), getattr(ref, f).reshape(-1)), f
    toks, scores = fx.synth_vocab(cfg.vocab_size)
    tb = str(tmp_path / "tok.bin")
    fx.write_tokenizer_bin(tb, toks, scores)
    m.load_tokenizer(tb)
    assert m.vocab()[0] == toks
    m.close()
    # truncated file: the reference would hit end-of-file in a stream read; we say so
    with open(p, "r+b") as f:
        f.truncate(1000)
    with pytest.raises(hostapi.HostError, match="end of file"):
        hostapi.HostModel(p, ak=True)


def test_argmax_and_sampler():
    lg = np.array([0.5, 3.0, 3.0, -1.0], np.float32)
    assert hostapi.argmax(lg) == 2  # first maximum, 1-based
    p = np.exp(lg.astype(np.float64) / 0.7)
    p /= p.sum()
    cdf = np.cumsum(p)
    for r in (0.0, 0.01, 0.3, 0.6, 0.95, 0.999999):
        assert hostapi.sample(lg, 0.7, r) == int(np.searchsorted(cdf, r, side="right")) + 1
    assert hostapi.sample(lg, 0.7, 1.0) == 4  # fallback to the last index


def test_cli_rejects_unknown_flags_like_the_reference():
    r = subprocess.run([hostapi.LLM_BIN, "--bogus"], capture_output=True, text=True)
    assert r.returncode != 0 and "Unrecognized option: --bogus" in r.stdout


def test_cli_tensor_parallel_flags_are_checked_before_anything_is_loaded():
    """`llm --tp-size N --tp-rank r --tp-dir d` (extension: one process per GPU, rendezvous through a
    directory); bad combinations stop with a message, like every other bad flag."""
    run = lambda *a: subprocess.run([hostapi.LLM_BIN, *a], capture_output=True, text=True)
    r = run("--tp-size", "3")
    assert r.returncode != 0 and "--tp-size must be 1, 2, 4 or 8" in r.stdout
    r = run("--tp-size", "2")
    assert r.returncode != 0 and "needs --tp-dir" in r.stdout
    r = run("--tp-size", "2", "--tp-rank", "2", "--tp-dir", "/tmp")
    assert r.returncode != 0 and "--tp-rank out of range" in r.stdout


def test_bench_prompt_tokenises():
    toks = hostapi.encode_with_synth_vocab("I stopped posting on knitting forums because", 32000)
    assert 5 < len(toks) < 44 and all(1 <= t <= 32000 for t in toks)


_FUZZ_CHILD = r'''
import sys
import numpy as np
sys.path.insert(0, sys.argv[3])
from llm.f90_b200 import hostapi
data = open(sys.argv[1], "rb").read()
rng = np.random.default_rng(int(sys.argv[2]))
head = min(len(data) - 8, 6000)  # header, metadata and tensor infos of the tiny model
blobs = [data[:c] for c in list(range(0, 48)) + [int(c) for c in rng.integers(48, len(data), 40)]]
for _ in range(60):                # one wrong byte
    b = bytearray(data); b[int(rng.integers(0, head))] = int(rng.integers(0, 256)); blobs.append(bytes(b))
for _ in range(40):                # an absurd 64-bit count / length / offset
    b = bytearray(data); i = int(rng.integers(0, head)); b[i:i + 8] = (0xFFFFFFFFFFFFFFF0).to_bytes(8, "little")
    blobs.append(bytes(b))
ok = err = 0
for blob in blobs:
    open(sys.argv[1] + ".case", "wb").write(blob)
    try:
        hostapi.HostModel(sys.argv[1] + ".case").close()
        ok += 1
    except hostapi.HostError:
        err += 1
print("FUZZ", len(blobs), ok, err)
'''


@pytest.mark.parametrize("alignment", [0, 24])
def test_loader_rejects_a_bad_alignment(tmp_path, alignment):
    """general.alignment is divided by (read_ggml.f90:176-196): zero would be a SIGFPE, a value that is not a
    power of two misplaces the tensor data -- an error, not a crash"""
    cfg = Config(**TINY)
    p = str(tmp_path / "align.gguf")
    fx.write_gguf(p, cfg, fx.synth_tensors(cfg, 0), alignment=64)
    raw = bytearray(open(p, "rb").read())
    key = b"general.alignment"
    i = raw.index(key) + len(key) + 4  # key bytes, then the u32 value type, then the value
    assert int.from_bytes(raw[i:i + 4], "little") == 64
    raw[i:i + 4] = alignment.to_bytes(4, "little")
    open(p, "wb").write(bytes(raw))
    with pytest.raises(hostapi.HostError, match="alignment"):
        hostapi.HostModel(p)


def test_loader_survives_truncated_and_corrupted_files(tmp_path):
    """The reference prints a message and stops on a bad file (read_ggml.f90:122-125); the mirror must
    do the same -- an error, never a crash or an unbounded allocation -- whatever the bytes are.  Runs in
    a child process so that a crash would show up as a failed test, not a dead test run."""
    p = str(tmp_path / "good.gguf")
    fx.write_synth_gguf(p, Config(**TINY, wtype=F32), 0)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([os.sys.executable, "-c", _FUZZ_CHILD, p, "7", root], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    n, ok, err = (int(x) for x in r.stdout.split("FUZZ")[1].split())
    assert ok + err == n and err >= 48  # every truncation fails cleanly; a flipped weight byte may still load


def test_loader_accepts_a_q6_k_output_weight(tmp_path):
    """Stock llama.cpp q4_0 files keep output.weight in Q6_K (ggml type 14), where the reference's type switch
    stops (read_ggml.f90:613-635): the loader hands the super-blocks through unchanged and says so."""
    from llm.f90_b200.layout import Q6_K, Q4_0
    cfg = Config(emb_dim=256, hidden_dim=352, n_layers=2, n_heads=4, n_kv_heads=2, vocab_size=300, seq_len=32, wtype=Q4_0)
    p = str(tmp_path / "m.gguf")
    want = fx.write_synth_gguf(p, cfg, seed=3, cls_wtype=Q6_K)
    m = hostapi.HostModel(p)
    assert m.cfg == cfg and m.cls_wtype == Q6_K
    got = m.weights()
    assert got.cls_wtype == Q6_K and got.wcls.nbytes == cfg.vocab_size * 210
    for f in got.FIELDS:
        assert np.array_equal(getattr(got, f).view(np.uint8).ravel(), getattr(want, f).view(np.uint8).ravel()), f
    m.close()
    # any other classifier type is still refused with the reference's message
    bad = str(tmp_path / "bad.gguf")
    fx.write_synth_gguf(bad, Config(**{**cfg.asdict(), "wtype": F32}), seed=3, cls_wtype=Q4_0)
    with pytest.raises(hostapi.HostError, match="Type not supported"):
        hostapi.HostModel(bad)

"""Q6_K output.weight: the classifier of stock llama.cpp q4_0 files, which the reference's loader
stops at (read_ggml.f90:406, :613-635).  The Q6_K mat-vec kernel against float64 on exactly dequantised weights, a
q4_0 model with a Q6_K classifier against the C oracle (extended with the exact Q6_K dequantisation, pinned to
gguf.quants in tests/test_oracle.py), and the same file through `llm -m`."""
import subprocess

import numpy as np
import pytest

from conftest import rel_err
from llm.f90_b200 import capi, fixtures as fx, hostapi
from llm.f90_b200.layout import Config, F32, Q4_0, Q6_K
from oracle import oracle_c as oc

pytestmark = pytest.mark.gpu

Q6 = dict(emb_dim=512, hidden_dim=1408, n_layers=3, n_heads=8, n_kv_heads=2, vocab_size=2048, seq_len=256)


@pytest.fixture(scope="module", autouse=True)
def _built(built):
    capi.load()


@pytest.mark.parametrize("rows,cols", [(1, 256), (37, 512), (1000, 2048), (333, 4096)])
def test_matvec_q6_k(rows, cols):
    rng = np.random.default_rng(rows + cols)
    enc = fx.quantize_q6_k((rng.standard_normal((rows, cols)) / np.sqrt(cols)).astype(np.float32))
    x = rng.standard_normal(cols).astype(np.float32)
    ref = fx.dequantize_q6_k(enc, cols).astype(np.float64) @ x.astype(np.float64)
    assert rel_err(capi.matvec(enc, Q6_K, rows, cols, x), ref) < 2e-6
    assert rel_err(oc.matvec(enc, Q6_K, rows, cols, x), ref) < 2e-6


@pytest.mark.parametrize("wt", [Q4_0, F32], ids=["q4_0", "f32"])
def test_model_with_q6_k_classifier_matches_oracle(wt):
    cfg = Config(**Q6, wtype=wt)
    w = fx.fuse_tensors(cfg, fx.synth_tensors(cfg, 11), cls_wtype=Q6_K)
    prompt, n = [21, 22, 23, 24, 25], 24
    ref_toks, ref_lg, _ = oc.Oracle(w).generate(prompt, n, want_logits=True)
    with capi.Engine(w) as eng:
        toks, lg = capi.host_generate(eng, prompt, n, want_logits=True)
    assert max(rel_err(lg[i], ref_lg[i]) for i in range(n)) < (1e-2 if wt == Q4_0 else 1e-4)
    assert (toks == ref_toks).all()


def test_cli_loads_a_q4_0_file_with_q6_k_output_weight(tmp_path):
    cfg = Config(**Q6, wtype=Q4_0)
    p = str(tmp_path / "m.gguf")
    w = fx.write_synth_gguf(p, cfg, seed=12, cls_wtype=Q6_K)
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
    text = r.stdout.split(b"\n Inference time:")[0].split(b"\n", 1)[1]
    assert text == want

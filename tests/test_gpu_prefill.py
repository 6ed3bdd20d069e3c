"""Batched prompt pass: the tcgen05 GEMM against float64, and the KV cache / next-position logits
after llmf90_b200_prefill(P tokens) against P calls of llmf90_b200_transformer (llama2.f90:379-385), for the three
weight storages.  The cases of one test run in their own process under a timeout (tests/prefill_worker.py): a
tensor-core kernel that waits on a barrier forever must not take the pytest session with it.

Tolerances: the GEMM keeps f16 hi + lo operand planes and f32 accumulation, so it is held to 2e-5 relative
against float64 on exactly dequantised weights (measured 3e-7 .. 2e-6); KV rows and logits to the 1e-4 (f32 / f16)
and 1e-2 (q4_0: the decode path's own q4_0 arithmetic differs from the GEMM's) the per-token path is held to
against the oracle (measured 5e-7 and 2e-5).
"""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
F32, F16, Q4_0 = 0, 1, 2
WT = [F32, F16, Q4_0]
IDS = ["f32", "f16", "q4_0"]
TOL = {F32: 1e-4, F16: 1e-4, Q4_0: 1e-2}


def run_cases(cases, timeout=600):
    cmd = [sys.executable, os.path.join(HERE, "prefill_worker.py"), "multi"] + cases
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    except subprocess.TimeoutExpired:
        pytest.fail("timed out (a kernel of the batched pass did not finish): " + "; ".join(cases))
    assert r.returncode == 0, r.stderr[-2000:]
    out = [json.loads(ln) for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert [o["case"] for o in out] == cases, r.stdout[-2000:]
    for o in out:
        assert "error" not in o, o
    return out


@pytest.mark.parametrize("wt", WT, ids=IDS)
def test_matmul_tcgen05(wt):
    """rows / cols / positions: one tile, ragged everything, TinyLlama's QKV and W2 shapes, a full 128-position pass."""
    shapes = [(128, 64, 16), (37, 96, 5), (2560, 2048, 16), (2048, 5632, 33), (1000, 4096, 128)]
    if wt == Q4_0:
        shapes[1] = (37, 96, 5)  # 96 = 3 q4_0 blocks
    for o in run_cases([f"matmul {wt} {r} {c} {p}" for r, c, p in shapes]):
        assert o["finite"] and o["rel_err"] < 2e-5, o


@pytest.mark.parametrize("wt", WT, ids=IDS)
def test_prefill_matches_per_token_path(wt):
    """tiny: contraction lengths that are not multiples of 64; small / mha: GQA and MHA head geometries, a prompt that
    is not a multiple of 16; mid: more than 128 positions (two passes, the second attending to the first's cache rows)."""
    for o in run_cases([f"prefill tiny {wt} 5", f"prefill small {wt} 16", f"prefill mha {wt} 37", f"prefill mid {wt} 130"]):
        assert o["k_err"] < TOL[wt] and o["v_err"] < TOL[wt] and o["logit_err"] < TOL[wt], o
        assert o["argmax_same"], o


def test_generate_greedy_with_prefill_matches_oracle():
    for o in run_cases([f"greedy small {wt} 9 40" for wt in WT] + ["greedy small 0 60 40"]):  # prompt longer than the run
        assert o["same"], o


def test_cli_with_prefill_prints_the_oracles_text(tmp_path):
    """`llm --prefill`: the prompt as one batched pass, then the token loop -- the same text as the oracle driven by
    the same file and prompt (and as `llm` without the flag)."""
    import numpy as np
    from llm.f90_b200 import fixtures as fx, hostapi
    from llm.f90_b200.layout import Config, SMALL
    from oracle import oracle_c as oc
    cfg = Config(**SMALL, wtype=Q4_0)
    p = str(tmp_path / "m.gguf")
    w = fx.write_synth_gguf(p, cfg, seed=12)
    m = hostapi.HostModel(p)
    vocab, _ = m.vocab()
    prompt = "the cat sat on the mat and"
    ptoks = m.encode(prompt)
    m.close()
    n = 32
    ref_toks, _, _ = oc.Oracle(w).generate(ptoks, n)
    want = b"".join(vocab[t - 1] for t in ref_toks)
    for extra in (["--prefill"], ["--prefill", "--host-sampler"]):
        r = subprocess.run([hostapi.LLM_BIN, "-m", p, "-n", str(n), "-p", prompt, "-t", "0"] + extra, capture_output=True, timeout=300)
        assert r.returncode == 0, r.stdout[-400:] + r.stderr[-400:]
        text = r.stdout.split(b"\n Inference time:")[0].split(b"\n", 1)[1]
        assert text == want, extra

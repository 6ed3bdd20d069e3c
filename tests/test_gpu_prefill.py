"""Batched prompt pass: the tcgen05 GEMM against float64, and the KV cache / next-position logits
after llmf90_b200_prefill(P tokens) against P calls of llmf90_b200_transformer (llama2.f90:379-385), for the three
weight storages.  Every case runs in its own process under a timeout (tests/prefill_worker.py).

Tolerances: the GEMM keeps f16 hi + lo operand planes and f32 accumulation, so it is held to 2e-5 relative
against float64 on exactly dequantised weights; KV rows and logits to the 1e-4 (f32 / f16) and 1e-2 (q4_0: the
decode path's own q4_0 arithmetic differs from the GEMM's) the per-token path is held to against the oracle.
"""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
F32, F16, Q4_0 = 0, 1, 2
TOL = {F32: 1e-4, F16: 1e-4, Q4_0: 1e-2}


def run_case(*args, timeout=300):
    cmd = [sys.executable, os.path.join(HERE, "prefill_worker.py")] + [str(a) for a in args]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    except subprocess.TimeoutExpired:
        pytest.fail("timed out (a kernel of the batched pass did not finish): " + " ".join(cmd[2:]))
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("wt", [F32, F16, Q4_0], ids=["f32", "f16", "q4_0"])
@pytest.mark.parametrize("rows,cols,n_pos", [(128, 64, 16), (37, 96, 5), (2560, 2048, 16), (2048, 5632, 33), (1000, 4096, 128)])
def test_matmul_tcgen05(wt, rows, cols, n_pos):
    out = run_case("matmul", wt, rows, cols, n_pos)
    assert out["finite"]
    assert out["rel_err"] < 2e-5, out


@pytest.mark.parametrize("wt", [F32, F16, Q4_0], ids=["f32", "f16", "q4_0"])
@pytest.mark.parametrize("shape,n_prompt", [("tiny", 5), ("small", 16), ("mha", 37), ("mid", 130)])
def test_prefill_matches_per_token_path(shape, wt, n_prompt):
    out = run_case("prefill", shape, wt, n_prompt)
    assert out["k_err"] < TOL[wt] and out["v_err"] < TOL[wt], out
    assert out["logit_err"] < TOL[wt], out
    assert out["argmax_same"]


@pytest.mark.parametrize("wt", [F32, F16, Q4_0], ids=["f32", "f16", "q4_0"])
def test_generate_greedy_with_prefill_matches_oracle(wt):
    out = run_case("greedy", "small", wt, 9, 40)
    assert out["same"], out

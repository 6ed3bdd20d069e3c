"""Parity at BASELINE.json's full sizes, all against the C oracle (oracle/llama2_oracle.c).

The oracle dequantises f16 / q4_0 rows one at a time (llama2_oracle.c: row_dot), so Llama-2-7B q4_0 costs
3.7 GB of host memory and runs at a few positions per second on all host cores -- the full-size models,
the whole 128-position benchmark workload and the long-context attention splits are all checked
against it directly.  Tolerances are the north star's: 1e-4 relative for f32 / f16 storage, 1e-2 for
q4_0, greedy token ids identical."""
import os

import numpy as np
import pytest

from conftest import rel_err
from llm.f90_b200 import capi, fixtures as fx
from llm.f90_b200.layout import Config, TINYLLAMA, LLAMA2_7B, F32, F16, Q4_0
from oracle import oracle_c as oc

pytestmark = pytest.mark.gpu
PROMPT = [7, 1200, 31000, 45]
TOL = {F32: 1e-4, F16: 1e-4, Q4_0: 1e-2}
CORES = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def bench_prompt(cfg):
    """the benchmark's prompt (bench.py: README.md:42 through the mirrored BPE)"""
    import bench
    return bench.prompt_tokens(cfg)


def first_flip(toks, ref_toks, ref_lg):
    """diagnostic for a greedy-id mismatch: position and the oracle's top-2 logit gap there"""
    bad = np.nonzero(np.asarray(toks) != np.asarray(ref_toks))[0]
    if len(bad) == 0:
        return None
    i = int(bad[0])
    top = np.sort(ref_lg[i])[-2:]
    return i, float(top[1] - top[0]), float(np.abs(ref_lg[i]).max())


@pytest.mark.parametrize("wt", [F32, F16], ids=["f32", "f16"])
def test_tinyllama_matches_oracle_at_full_size(built, wt):
    cfg = Config(**TINYLLAMA, wtype=wt)
    w = fx.synth_weights_tiled(cfg, 0)
    n = 6
    ref_toks, ref_lg, _ = oc.Oracle(w, n_threads=CORES).generate(PROMPT, n, want_logits=True)
    with capi.Engine(w) as eng:
        toks, lg = capi.host_generate(eng, PROMPT, n, want_logits=True)
        errs = [rel_err(lg[i], ref_lg[i]) for i in range(n)]
        assert max(errs) < TOL[wt], errs
        assert (toks == ref_toks).all()
        eng.reset()
        dev_toks, _ = eng.generate_greedy(PROMPT, n)
        assert (dev_toks == ref_toks).all()


@pytest.mark.parametrize("wt", [Q4_0, F16], ids=["7b-q4_0", "7b-f16"])
def test_llama2_7b_matches_oracle_at_full_size(built, wt):
    """the second headline configuration (Llama-2-7B q4_0) and the multi-GPU one (7B f16), fused kernel vs
    the C oracle on all host cores (llama2.f90:480-640 on exactly dequantised weights)"""
    cfg = Config(**LLAMA2_7B, wtype=wt)
    w = fx.synth_weights_tiled(cfg, 0)
    n = 5
    ref_toks, ref_lg, _ = oc.Oracle(w, n_threads=CORES).generate(PROMPT, n, want_logits=True)
    with capi.Engine(w) as eng:
        toks, lg = capi.host_generate(eng, PROMPT, n, want_logits=True)
        eng.reset()
        dev_toks, _ = eng.generate_greedy(PROMPT, n)
    errs = [rel_err(lg[i], ref_lg[i]) for i in range(n)]
    assert np.isfinite(lg).all()
    assert max(errs) < TOL[wt], errs
    assert (toks == ref_toks).all(), first_flip(toks, ref_toks, ref_lg)
    assert (dev_toks == ref_toks).all()


@pytest.mark.parametrize("shape,wt", [(TINYLLAMA, F32), (LLAMA2_7B, Q4_0)], ids=["tinyllama-f32", "7b-q4_0"])
def test_whole_bench_workload_matches_oracle(built, shape, wt):
    """bench.py's workload itself: 128 positions, the benchmark prompt, greedy ids identical to the
    oracle's over the whole run (device loop and host loop), logits within tolerance at every position
    (positions 1, 64 and 128 are reported)"""
    cfg = Config(**shape, wtype=wt)
    w = fx.synth_weights_tiled(cfg, 0)
    prompt = bench_prompt(cfg)
    n = 128
    ref_toks, ref_lg, _ = oc.Oracle(w, n_threads=CORES).generate(prompt, n, want_logits=True)
    with capi.Engine(w) as eng:
        dev_toks, _ = eng.generate_greedy(prompt, n)
        eng.reset()
        toks, lg = capi.host_generate(eng, prompt, n, want_logits=True)
    errs = np.array([rel_err(lg[i], ref_lg[i]) for i in range(n)])
    print(f"rel err at positions 1/64/128: {errs[0]:.2e} {errs[63]:.2e} {errs[127]:.2e}; max {errs.max():.2e}")
    assert (dev_toks == ref_toks).all(), first_flip(dev_toks, ref_toks, ref_lg)
    assert (toks == ref_toks).all(), first_flip(toks, ref_toks, ref_lg)
    assert errs.max() < TOL[wt], (int(errs.argmax()), float(errs.max()))


# the attention phase splits the positions of a head over 1 / 2 / 4 / 8 work items (<= 256 / 512 / 1024 /
# 2048 positions, engine.cu: n_splits_for): run past 1024 positions so that every split count, the
# partial-record merge and the reference's seq_len of 2048 region (llama2.f90:108) are exercised
LONG_GQA = dict(emb_dim=512, hidden_dim=1408, n_layers=2, n_heads=8, n_kv_heads=1, vocab_size=512, seq_len=1280)
LONG_MHA = dict(emb_dim=512, hidden_dim=1376, n_layers=2, n_heads=4, n_kv_heads=4, vocab_size=512, seq_len=1280)


@pytest.mark.parametrize("shape,wt,n", [(LONG_GQA, F32, 1200), (LONG_GQA, Q4_0, 1100), (LONG_MHA, F16, 1100)],
                         ids=["gqa8-hs64-f32", "gqa8-hs64-q4_0", "mha-hs128-f16"])
def test_attention_splits_4_and_8_match_oracle(built, shape, wt, n):
    cfg = Config(**shape, wtype=wt)
    w = fx.synth_weights(cfg, 21)
    # teacher-forced: every position is fed the oracle's token, so one near-tie cannot fork the two runs
    # and every position's logits are comparable
    ref_toks, ref_lg, _ = oc.Oracle(w, n_threads=CORES).generate([9, 8, 7], n, want_logits=True)
    forced = [int(t) for t in ref_toks[:-1]]
    with capi.Engine(w) as eng:
        toks, lg = capi.host_generate(eng, forced, n, want_logits=True)
        eng.reset()
        dev_toks, _ = eng.generate_greedy(forced, n)
    errs = np.array([rel_err(lg[i], ref_lg[i]) for i in range(n)])
    for lo, hi in ((0, 256), (256, 512), (512, 1024), (1024, n)):
        print(f"positions {lo + 1}..{hi}: max rel err {errs[lo:hi].max():.2e}")
    assert errs.max() < TOL[wt], (int(errs.argmax()), float(errs.max()))
    # ids: the last position is free-running in both, the others are forced
    assert (toks == ref_toks).all() and (dev_toks == ref_toks).all()

"""Runs ONE case of the batched prompt pass (tcgen05 GEMM / prefill) in its own process and prints one JSON
line.  tests/test_gpu_prefill.py starts it under a timeout: a tensor-core kernel that waits on a barrier
forever must not take the whole pytest session with it.

    python tests/prefill_worker.py matmul  <wtype> <rows> <cols> <n_pos>
    python tests/prefill_worker.py prefill <shape> <wtype> <n_prompt>
    python tests/prefill_worker.py greedy  <shape> <wtype> <n_prompt> <n>
    python tests/prefill_worker.py q6k_matvec <rows> <cols> | q6k_model <wtype> | sample <shape> <wtype>
    python tests/prefill_worker.py multi "<case> <args>" "<case> <args>" ...     (several cases, one process)

Also what the hardware check scripts tools/gpu_r02b.sh .. use directly (LLMF90_WORKER_NO_BUILD=1 skips the build check).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from llm.f90_b200 import capi, fixtures as fx  # noqa: E402
from llm.f90_b200.layout import Config, TINY, SMALL, TINYLLAMA  # noqa: E402

MID = dict(emb_dim=1024, hidden_dim=2816, n_layers=4, n_heads=16, n_kv_heads=2, vocab_size=4096, seq_len=512)
MHA = dict(emb_dim=512, hidden_dim=1376, n_layers=2, n_heads=4, n_kv_heads=4, vocab_size=1024, seq_len=128)
SHAPES = {"tiny": TINY, "small": SMALL, "mid": MID, "mha": MHA, "tinyllama": TINYLLAMA}


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def case_matmul(wt, rows, cols, n_pos):
    rng = np.random.default_rng(rows * 7 + cols + n_pos)
    wf = (rng.standard_normal((rows, cols)) / np.sqrt(cols)).astype(np.float32)
    enc = fx.encode_matrix(wf, wt)
    x = rng.standard_normal((n_pos, cols)).astype(np.float32)
    ref = x.astype(np.float64) @ fx.decode_matrix(enc, wt, cols).astype(np.float64).T
    got = capi.matmul(enc, wt, rows, cols, x)
    return {"rel_err": rel(got, ref), "finite": bool(np.isfinite(got).all())}


def prompt_for(cfg, n_prompt, seed=3):
    rng = np.random.default_rng(seed)
    return [int(t) for t in rng.integers(3, cfg.vocab_size, n_prompt)]


def case_prefill(shape, wt, n_prompt):
    """KV rows of every layer and the logits of the next position: batched pass vs the per-token path."""
    cfg = Config(**SHAPES[shape], wtype=wt)
    if shape in ("tiny", "small", "mha"):
        w = fx.synth_weights(cfg, 11)
    else:  # large shapes: the quick generators of the benchmark
        w = fx.synth_weights_tiled(cfg, 0) if shape == "tinyllama" else fx.synth_weights_fast(cfg, 11)
    toks = [2] + prompt_for(cfg, n_prompt)  # inputs of positions 1 .. n_prompt + 1
    with capi.Engine(w, prefill=True) as eng:
        for p in range(n_prompt):
            eng.transformer(toks[p], p + 1)
        kv_ref = [[eng.read_kv(l, p + 1) for p in range(n_prompt)] for l in range(cfg.n_layers)]
        lg_ref = eng.transformer(toks[n_prompt], n_prompt + 1).copy()
        eng.reset()
        eng.prefill(toks[:n_prompt], 1)
        kv = [[eng.read_kv(l, p + 1) for p in range(n_prompt)] for l in range(cfg.n_layers)]
        lg = eng.transformer(toks[n_prompt], n_prompt + 1).copy()
        launches = eng.stats()["kernel_launches"]
    k_err = max(rel(np.stack([a[0] for a in kv[l]]), np.stack([a[0] for a in kv_ref[l]])) for l in range(cfg.n_layers))
    v_err = max(rel(np.stack([a[1] for a in kv[l]]), np.stack([a[1] for a in kv_ref[l]])) for l in range(cfg.n_layers))
    return {"k_err": k_err, "v_err": v_err, "logit_err": rel(lg, lg_ref), "argmax_same": bool(lg.argmax() == lg_ref.argmax()),
            "launches": int(launches)}


def case_greedy(shape, wt, n_prompt, n):
    """generate_greedy with the batched prompt pass against the CPU oracle's token ids."""
    from oracle import oracle_c as oc
    cfg = Config(**SHAPES[shape], wtype=wt)
    w = fx.synth_weights(cfg, 2)
    prompt = prompt_for(cfg, n_prompt, 5)
    ref, _, _ = oc.Oracle(w).generate(prompt, n)
    with capi.Engine(w, prefill=True) as eng:
        toks, ms = eng.generate_greedy(prompt, n)
    return {"same": bool((np.asarray(toks) == np.asarray(ref)).all()), "n_diff": int((np.asarray(toks) != np.asarray(ref)).sum()),
            "ms": float(ms)}


def case_q6k_matvec(rows, cols):
    from llm.f90_b200.layout import Q6_K
    rng = np.random.default_rng(rows + cols)
    enc = fx.quantize_q6_k((rng.standard_normal((rows, cols)) / np.sqrt(cols)).astype(np.float32))
    x = rng.standard_normal(cols).astype(np.float32)
    ref = fx.dequantize_q6_k(enc, cols).astype(np.float64) @ x.astype(np.float64)
    return {"rel_err": rel(capi.matvec(enc, Q6_K, rows, cols, x), ref)}


def case_q6k_model(wt):
    from oracle import oracle_c as oc
    from llm.f90_b200.layout import Q6_K
    cfg = Config(**SMALL, wtype=wt)
    w = fx.fuse_tensors(cfg, fx.synth_tensors(cfg, 11), cls_wtype=Q6_K)
    prompt, n = [21, 22, 23, 24, 25], 24
    ref_toks, ref_lg, _ = oc.Oracle(w).generate(prompt, n, want_logits=True)
    with capi.Engine(w) as eng:
        toks, lg = capi.host_generate(eng, prompt, n, want_logits=True)
    return {"logit_err": max(rel(lg[i], ref_lg[i]) for i in range(n)), "same": bool((toks == ref_toks).all())}


def case_sample(shape, wt):
    """transformer_sample against the sequential CDF walk of the host mirror on the same logits."""
    from llm.f90_b200 import hostapi
    cfg = Config(**SHAPES[shape], wtype=wt)
    w = fx.synth_weights(cfg, 7)
    rng = np.random.default_rng(1)
    bad, near, n = [], 0, 0
    with capi.Engine(w) as eng:
        tok = 2
        for pos in range(1, 25):
            lg = eng.transformer(tok, pos).copy()
            for T in (0.0, 0.7, 1.3):
                r = float(rng.random())
                got = eng.transformer_sample(tok, pos, T, r)   # same position again: the cache row is rewritten with the same values
                want = hostapi.argmax(lg) if T == 0 else hostapi.sample(lg, T, r)
                n += 1
                if got != want:
                    z = lg.astype(np.float64) / T
                    p = np.exp(z - z.max()); cdf = np.cumsum(p / p.sum())
                    if np.abs(cdf - r).min() < 1e-5:
                        near += 1
                    else:
                        bad.append((pos, T, r, int(got), int(want)))
            tok = int(lg.argmax()) + 1
    return {"n": n, "mismatch_near_boundary": near, "bad": bad[:5], "ok": not bad}


def main(argv):
    kind = argv[0]
    if kind == "multi":  # several cases in one process: "kind args..." strings
        for spec in argv[1:]:
            try:
                main(spec.split())
            except Exception as e:  # keep going: the later cases may still tell something
                print(json.dumps({"case": spec, "error": repr(e)[:400]}), flush=True)
        return
    if kind == "q6k_matvec":
        out = case_q6k_matvec(int(argv[1]), int(argv[2]))
    elif kind == "q6k_model":
        out = case_q6k_model(int(argv[1]))
    elif kind == "sample":
        out = case_sample(argv[1], int(argv[2]))
    elif kind == "matmul":
        out = case_matmul(int(argv[1]), int(argv[2]), int(argv[3]), int(argv[4]))
    elif kind == "prefill":
        out = case_prefill(argv[1], int(argv[2]), int(argv[3]))
    elif kind == "greedy":
        out = case_greedy(argv[1], int(argv[2]), int(argv[3]), int(argv[4]))
    else:
        raise SystemExit("unknown case " + kind)
    out["case"] = " ".join(argv)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    if not os.environ.get("LLMF90_WORKER_NO_BUILD"):
        import __graft_entry__ as ge
        ge.build()
    main(sys.argv[1:])

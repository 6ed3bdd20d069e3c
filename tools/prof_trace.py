"""Per-CTA phase-edge trace of one layer of the fused kernel (profiling aid; instrumented kernel).
usage: python tools/prof_trace.py <tinyllama|llama2-7b> <f32|f16|q4_0> [layer] [pos]"""
import sys
import numpy as np
sys.path.insert(0, '.')
from llm.f90_b200 import capi, fixtures as fx
from llm.f90_b200.layout import Config, TINYLLAMA, LLAMA2_7B, WTYPE_BY_NAME
model, wt = sys.argv[1], sys.argv[2]
layer = int(sys.argv[3]) if len(sys.argv) > 3 else 10
npos = int(sys.argv[4]) if len(sys.argv) > 4 else 64
cfg = Config(**(TINYLLAMA if model == 'tinyllama' else LLAMA2_7B), wtype=WTYPE_BY_NAME[wt])
w = fx.synth_weights_tiled(cfg, 0)
eng = capi.Engine(w)
toks, _ = eng.generate_greedy([5, 6, 7], npos)
tr = eng.debug_trace(int(toks[-1]), npos + 1, layer).astype(np.int64)
# thread 0 of every CTA stamps globaltimer (ns) when IT passes an edge: after a prologue (the activation
# vector is complete: includes the wait for the slowest publisher), after its own warp's tiles of a phase
edges = [(0, "layer start"), (1, "qkv prologue done"), (2, "qkv tiles done (warp 0)"), (4, "attention done"),
         (6, "wo prologue done"), (7, "wo tiles done (warp 0)"), (9, "w13 prologue done"),
         (10, "w13 tiles done (warp 0)"), (12, "w2 prologue done"), (13, "w2 tiles done (warp 0)")]
t0 = tr[:, 0].min()
print("edge                         min     p50     max   (us since the first CTA entered the layer)")
prev = None
for k, name in edges:
    col = (tr[:, k] - t0) / 1000.0
    d = "" if prev is None else f"   step p50 {np.median((tr[:, k] - tr[:, prev]) / 1000.0):6.2f}  min {((tr[:, k] - tr[:, prev]) / 1000.0).min():6.2f}  max {((tr[:, k] - tr[:, prev]) / 1000.0).max():6.2f}"
    print(f"{k:2d} {name:24s} {col.min():7.2f} {np.median(col):7.2f} {col.max():7.2f}{d}")
    prev = k
np.save("gpurun_out/trace.npy", tr)
# warp 0's first tile of each mat-vec phase: SM clocks relative to entering run_tiles
print("warp 0, first tile (cycles since run_tiles entry, p50 / max over CTAs): first stage landed, mat-vec done, reduced, "
      "published | all own tiles done | at entry: ring lead / prefetch lead (stages ahead of the phase's first) p50 | "
      "when warp 0 is done: ring / prefetch cursor - phase start, stages of the phase")
info = eng.stats()
for ph, name in enumerate(["qkv", "wo", "w13", "w2"]):
    b = tr[:, 16 + 8 * ph: 24 + 8 * ph].copy()
    ok = b[:, 1] > 0  # CTAs whose warp 0 had a tile
    if not ok.any():
        continue
    end_ring = (b[ok, 3] >> 32) & 0xffff
    end_pf = (b[ok, 3] >> 48) & 0xffff
    b[:, 3] &= 0xffffffff
    e = b[ok, 0]
    lo = e & 0xffffffff
    cols = [b[ok, k] - e for k in (1, 2, 4, 7)]
    red = (b[ok, 3] - lo) & 0xffffffff
    lead = (b[ok, 5] & 0xffffffff) - b[ok, 6]
    pfl = (b[ok, 5] >> 32) - b[ok, 6]
    print(f"   {name:4s}", " ".join(f"{int(np.median(c)):6d}/{int(c.max()):6d}" for c in cols), f"(reduced {int(np.median(red))})",
          f"| lead {int(np.median(lead))} / pf {int(np.median(pfl))} | end ring {int(np.median(end_ring))} pf {int(np.median(end_pf))}")
# attention phase of the CTAs that had an item: cycles since entering attention_phase_t (thread 0 = warp 0)
a = tr[:, 48:56]
ok = a[:, 6] > 0
if ok.any():
    e = a[ok, 0]
    names = {1: "K/V loads issued, poll starts", 2: "q (+ current k, v) arrived", 3: "positions done", 5: "merge barrier passed",
             4: "(probe build: first merge pass done)", 6: "published", 7: "left"}
    print(f"attention CTAs ({int(ok.sum())}): cycles since entry, p50 / max:",
          "; ".join(f"{n} {int(np.median(a[ok, k] - e))}/{int((a[ok, k] - e).max())}" for k, n in names.items() if (a[ok, k] > 0).any()))
# the activation-vector prologue (gather_x) of W13 and W2: cycles since entry (thread 0)
for k, name in ((72, "w13 prologue"), (80, "w2 prologue")):
    g = tr[:, k:k + 5]
    ok = g[:, 4] > 0
    if ok.any():
        e = g[ok, 0]
        print(f"{name}: first batch polled {int(np.median(g[ok, 1] - e))}/{int((g[ok, 1] - e).max())}; barrier A passed "
              f"{int(np.median(g[ok, 2] - e))}; vector stored {int(np.median(g[ok, 3] - e))}; barrier B passed "
              f"{int(np.median(g[ok, 4] - e))}/{int((g[ok, 4] - e).max())}   (p50[/max] over {int(ok.sum())} CTAs)")
# probe build (LLMF90_BUILD_PROBE=1): the attention phase run a second time on complete inputs, warm code
a2 = tr[:, 56:64]
ok = a2[:, 7] > 0
if ok.any():
    e = a2[ok, 0]
    print("attention, second (warm) call: cycles since entry p50/max:",
          "; ".join(f"[{k}] {int(np.median(a2[ok, k] - e))}/{int((a2[ok, k] - e).max())}" for k in (1, 2, 3, 5, 6, 7) if (a2[ok, k] > 0).any()))

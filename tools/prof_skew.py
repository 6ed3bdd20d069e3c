"""Is the per-CTA skew of the weight phases systematic (same SMs slow every time)?
usage: python tools/prof_skew.py <model> <wtype>"""
import sys
import numpy as np
sys.path.insert(0, '.')
from llm.f90_b200 import capi, fixtures as fx
from llm.f90_b200.layout import Config, TINYLLAMA, LLAMA2_7B, WTYPE_BY_NAME
model, wt = sys.argv[1], sys.argv[2]
cfg = Config(**(TINYLLAMA if model == 'tinyllama' else LLAMA2_7B), wtype=WTYPE_BY_NAME[wt])
w = fx.synth_weights_tiled(cfg, 0)
eng = capi.Engine(w)
toks, _ = eng.generate_greedy([5, 6, 7], 64)
ends = []
for rep, layer in enumerate([5, 10, 10, 15]):
    tr = eng.debug_trace(int(toks[-1]), 65 + rep, layer).astype(np.int64)
    t = (tr[:, :15] - tr[:, :1].min()) / 1000.0
    ends.append(t)
    print(f"layer {layer}: edge times min/p50/max")
    for k in (1, 2, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14):
        print(f"   edge {k:2d}: {t[:, k].min():7.2f} {np.median(t[:, k]):7.2f} {t[:, k].max():7.2f}")
w13 = [e[:, 10] - e[:, 9] for e in ends]
w2 = [e[:, 13] - e[:, 12] for e in ends]
print("corr of per-CTA W13 durations between runs:", np.round(np.corrcoef(np.array(w13)), 2).tolist())
print("corr of per-CTA W2 durations between runs:", np.round(np.corrcoef(np.array(w2)), 2).tolist())
order = np.argsort(np.mean(w13, axis=0))
print("slowest CTAs (W13):", order[-12:].tolist(), "fastest:", order[:12].tolist())
print("mean W13 duration by CTA id block of 16:", [round(float(np.mean(np.mean(w13, axis=0)[i:i+16])), 2) for i in range(0, 148, 16)])
np.save("gpurun_out/skew.npy", np.array(ends))

"""Per-CTA phase trace of one layer of the fused kernel (profiling aid).
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
t = tr[:, :15] - tr[:, :1].min()
names = ["start", "qkv_pro", "qkv_mv", "bar", "att", "bar", "wo_pro", "wo_mv", "bar", "w13_pro", "w13_mv", "bar",
         "w2_pro", "w2_mv", "bar"]
print("edge            min     p50     max   (us since the first CTA entered the layer)")
for k, nme in enumerate(names):
    col = t[:, k] / 1000.0
    print(f"{k:2d} {nme:8s} {col.min():8.2f}{np.median(col):8.2f}{col.max():8.2f}")
d = np.diff(t, axis=1) / 1000.0
print("per-CTA durations (us): min p50 max")
for k in range(14):
    print(f"   {names[k+1]:8s} {d[:, k].min():7.2f}{np.median(d[:, k]):7.2f}{d[:, k].max():7.2f}")
wait = tr[:, 16:20]
print("warp-0 wait cycles in mv phases (qkv, wo, w13, w2): p50", np.median(wait, axis=0), "max", wait.max(axis=0))
print("warp-0 compute cycles: p50", np.median(tr[:, 20:24], axis=0), "max", tr[:, 20:24].max(axis=0))
print("warp-0 stages: p50", np.median(tr[:, 24:28], axis=0), "max", tr[:, 24:28].max(axis=0))
landed = (tr[:, 48:63] >> 32).astype(np.int64)
tr[:, 48:63] &= 0xffffffff
lead = tr[:, 32:47] - tr[:, 48:63]
print("stages already landed in the ring at each edge: min p50 max")
for k, nme in enumerate(names):
    print(f"{k:2d} {nme:8s} {landed[:, k].min():4d} {int(np.median(landed[:, k])):4d} {landed[:, k].max():4d}")
print("producer lead (stages issued - stages consumed) at each edge: min p50 max")
for k, nme in enumerate(names):
    print(f"{k:2d} {nme:8s} {lead[:, k].min():4d} {int(np.median(lead[:, k])):4d} {lead[:, k].max():4d}   issued p50 {int(np.median(tr[:, 32 + k]))} consumed p50 {int(np.median(tr[:, 48 + k]))}")
st = tr[:, 64:96].reshape(-1, 4, 8).astype(np.int64)
print("warp-0 consume stamps (cycles since 'before call', p50 over CTAs): entry, walk-ctor, 1st wait done, 1st stage released, returned, after cons_sync, after epilogue")
for phi, nme in enumerate(["qkv", "wo", "w13", "w2"]):
    b = st[:, phi, 7]
    print(f"   {nme:4s}", [int(np.median(st[:, phi, i] - b)) for i in (0, 1, 2, 3, 4, 5, 6)])
np.save("gpurun_out/trace.npy", tr)

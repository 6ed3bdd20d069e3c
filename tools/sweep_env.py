"""ms/token of the device greedy loop for several values of one engine environment knob
(one process, the engine is re-initialised per value).
usage: python tools/sweep_env.py <tinyllama|llama2-7b> <f32|f16|q4_0> <ENV_NAME> v1 v2 ...
       python tools/sweep_env.py <model> <wtype> MULTI A=1,B=2 A=3,B=4 ...   (several knobs per point)"""
import os
import sys
sys.path.insert(0, '.')
from llm.f90_b200 import capi, fixtures as fx
from llm.f90_b200.layout import Config, TINYLLAMA, LLAMA2_7B, WTYPE_BY_NAME
model, wt, name = sys.argv[1], sys.argv[2], sys.argv[3]
cfg = Config(**(TINYLLAMA if model == 'tinyllama' else LLAMA2_7B), wtype=WTYPE_BY_NAME[wt])
w = fx.synth_weights_tiled(cfg, 0)
prompt = [5, 6, 7, 8, 9, 10, 11, 12, 13]
ref = None
touched = set()
for v in sys.argv[4:]:
    for k in touched:  # every point starts from the caller's environment
        os.environ.pop(k, None)
    touched.clear()
    if name == "MULTI":
        for kv in v.split(","):
            k, x = kv.split("=")
            os.environ[k] = x
            touched.add(k)
    else:
        os.environ[name] = v
        touched.add(name)
    eng = capi.Engine(w)
    for _ in range(3):
        eng.generate_greedy(prompt, 128)
    ms = []
    for _ in range(3):
        toks, _ = eng.generate_greedy(prompt, 128)
        ms.append(eng.stats()["last_loop_total_ms"] / 128)
    ref = toks if ref is None else ref
    print(f"{model} {wt} {name}={v}: ms/token {min(ms):.4f} (median {sorted(ms)[1]:.4f}) tokens_equal={bool((toks == ref).all())}", flush=True)
    eng.close()

import sys, json, numpy as np
sys.path.insert(0, '.')
from llm.f90_b200 import capi, fixtures as fx
from llm.f90_b200.layout import Config, TINYLLAMA, LLAMA2_7B, WTYPE_BY_NAME
model, wt = sys.argv[1], sys.argv[2]
cfg = Config(**(TINYLLAMA if model == 'tinyllama' else LLAMA2_7B), wtype=WTYPE_BY_NAME[wt])
w = fx.synth_weights_tiled(cfg, 0)
eng = capi.Engine(w, profile='--noprof' not in sys.argv)
prompt = [5, 6, 7, 8, 9, 10, 11, 12, 13]
for _ in range(2): eng.generate_greedy(prompt, 128)
eng.reset()
toks, af = eng.generate_greedy(prompt, 128)
st = eng.stats()
ph = eng.phase_times()
tot = sum(ph.values())
print(json.dumps({"model": model, "wtype": wt, "ms_per_token": st["last_loop_total_ms"]/128, "phase_ms_per_token": {k: v/128 for k, v in ph.items()}, "sum": tot/128, "stats": st}))

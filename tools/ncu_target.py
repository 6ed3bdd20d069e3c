"""Small driver for ncu captures: build an engine on synthetic weights and run a few decode steps.
usage: python tools/ncu_target.py <tinyllama|llama2-7b> <f32|f16|q4_0> [n_positions]"""
import sys
sys.path.insert(0, '.')
from llm.f90_b200 import capi, fixtures as fx
from llm.f90_b200.layout import Config, TINYLLAMA, LLAMA2_7B, WTYPE_BY_NAME
model, wt = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 24
cfg = Config(**(TINYLLAMA if model == 'tinyllama' else LLAMA2_7B), wtype=WTYPE_BY_NAME[wt])
w = fx.synth_weights_tiled(cfg, 0)
eng = capi.Engine(w, granular='--granular' in sys.argv)
toks, ms = eng.generate_greedy([5, 6, 7, 8, 9], n)
print("tokens", toks[:8], "ms after first", ms, eng.stats())
eng.close()

import sys, numpy as np
sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
from test_gpu_parity import run_both, rel_err, MID
from llm.f90_b200.layout import Config, Q4_0
cfg = Config(**MID, wtype=Q4_0)
ref_toks, ref_lg, toks, lg, _ = run_both(cfg, 5, [100, 200, 300], 300, False)
errs = np.array([rel_err(lg[i], ref_lg[i]) for i in range(len(lg))])
print("max err", errs.max(), "argmax", errs.argmax())
bad = np.where(errs > 5e-3)[0]
print("bad positions", bad[:40], len(bad))
print("errs 250..300", np.round(errs[250:300], 5))
print("tokens equal", (toks == ref_toks).all(), np.where(toks != ref_toks)[0][:10])

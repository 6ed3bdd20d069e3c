"""bench.py's CPU-side contract, checked without a GPU: the reference arm's JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GPU_ARM_CONFIG_KEYS = {"workload", "emb_dim", "hidden_dim", "n_layers", "n_heads", "n_kv_heads", "vocab_size",
                       "weight_storage", "positions_per_step", "prompt_tokens", "parallelism", "arithmetic"}


def test_reference_arm_line(built):
    """`bench.py --impl reference`: the 1-thread figure is `value` (the reference is a single-threaded program), the
    all-core figure sits beside it, the config carries the keys of the GPU arm and names the position sample, and
    the e2e object repeats the value with zero copy bytes."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--model", "small",
                        "--steps", "1", "--warmup", "0", "--cpu-sample-pos", "24"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "tokens/sec decode (128-tok gen)"
    assert line["unit"] == "tokens/s" and line["higher_is_better"] is True and line["dtype"] == "f32"
    assert GPU_ARM_CONFIG_KEYS <= set(line["config"])
    assert line["config"]["positions_per_step_sampled"] == 24
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["value"] == line["value"] > 0
    assert cb["all_cores"]["cores"] >= 1 and cb["all_cores"]["value"] > 0
    assert "positions 1..24" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}

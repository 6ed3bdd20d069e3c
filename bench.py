#!/usr/bin/env python
"""bench.py -- tokens/s of a 128-position greedy decode (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--model tinyllama|llama2-7b|small]
                    [--wtype f32|f16|q4_0] [--impl ours|reference]

One "step" = one 128-position greedy generation (BOS + forced prompt + greedy picks,
llama2.f90:376-402) on synthetic weights of the named architecture.

  value        tokens/s with everything resident in HBM: the device-side greedy loop
               (llmf90_b200_generate_greedy; one kernel launch per token, next token never leaves
               the device), timed with CUDA events on the engine's stream.
  e2e          the same 128 positions through the reference-facing call
               llmf90_b200_transformer(token, pos, logits) with HOST logits every token and the
               host picking the next token (what the Fortran loop does), wall clock between
               device synchronisations.
  roofline     the decode kernel (one launch = one token): algorithmic bytes per token
                / mean launch duration vs MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline the C restatement of llama2.f90 (oracle/, 1 thread like the reference) on a bounded
               sample of the same workload, on this box's host cores.

--impl reference times that CPU restatement (the Fortran binary cannot be built: no Fortran compiler
in the image) on a bounded sample of the same workload: `value` is the 1-thread figure -- the reference is a
single-threaded program -- and cpu_baseline.all_cores the same code with OpenMP on every host core.  Under
torchrun only rank 0 runs it.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_POS = 128
PROMPT = "I stopped posting on knitting forums because"  # README.md:42
METRIC = "tokens/sec decode (128-tok gen)"
# the arithmetic of the mat-vecs per weight storage type (activations, accumulators and everything else are f32)
DTYPE = {"f32": "f32", "f16": "f16", "q4_0": "q4_0"}
ARITHMETIC = {"f32": "f32 FMA (weights f32)",
              "f16": "f16 weights x (hi + lo f16 planes of the f32 activations) on mma.sync, f32 accumulate",
              "q4_0": "q4_0 nibbles -> f16 x (hi + lo f16 planes of the activations) on mma.sync, f32 scales / accumulate"}


def model_config(name: str, wtype: str):
    from llm.f90_b200.layout import Config, TINYLLAMA, LLAMA2_7B, SMALL, WTYPE_BY_NAME
    dims = {"tinyllama": TINYLLAMA, "llama2-7b": LLAMA2_7B, "small": SMALL}[name]
    return Config(**dims, wtype=WTYPE_BY_NAME[wtype])


def prompt_tokens(cfg) -> list[int]:
    """Tokenise PROMPT with the synthetic vocabulary (byte-level + merges, llama2.f90:658-724).
    Falls back to a fixed id sequence when the host tokenizer is not built."""
    try:
        from llm.f90_b200 import hostapi
        return hostapi.encode_with_synth_vocab(PROMPT, cfg.vocab_size)
    except Exception:
        rng = np.random.default_rng(42)
        return [int(t) for t in rng.integers(4, cfg.vocab_size, size=9)]


# ------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock and throttle reasons during the timed region (NVML; nvidia-smi fallback)."""

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thr = None
        self.how = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.how = "nvml"
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {}
        if nv:
            for n in ("HwSlowdown", "HwThermalSlowdown", "SwThermalSlowdown", "SwPowerCap", "HwPowerBrakeSlowdown"):
                v = getattr(nv, "nvmlClocksEventReason" + n, None) or getattr(nv, "nvmlClocksThrottleReason" + n, None)
                if v is not None:
                    names[v] = n
        while not self._stop.is_set():
            try:
                if nv:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                    try:
                        r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    except Exception:
                        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for bit, n in names.items():
                        if r & bit:
                            self.reasons.add(n)
                else:
                    out = subprocess.run(["nvidia-smi", f"--id={self.index}",
                                          "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
                                          "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                                          "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits"],
                                         capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                    self.samples.append(int(out[0]))
                    self.max_mhz = int(out[1])
                    for n, v in zip(("HwSlowdown", "HwThermalSlowdown", "SwThermalSlowdown", "SwPowerCap"), out[2:]):
                        if v.strip().lower().startswith("active"):
                            self.reasons.add(n)
                    self.how = "nvidia-smi"
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._thr = threading.Thread(target=self._loop, daemon=True)
        self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._thr.join(timeout=5)

    def summary(self) -> dict:
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s), "how": self.how}


# ------------------------------------------------------------------ NVLink counters
def nvlink_kib(index: int):
    """(tx, rx) data KiB this GPU has moved over NVLink so far (NVML throughput counters, summed over the links),
    or None when the driver does not expose them.  Read before and after the timed region of a tensor-parallel run:
    the only NVLink traffic in between is the decode kernel's own peer stores (partial Wo / W2 vectors, logits rows,
    argmax records) -- no NCCL call is inside it."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        out = []
        for fid in (pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX):
            tot, seen = 0, False
            try:  # scope UINT_MAX = all links
                v = pynvml.nvmlDeviceGetFieldValues(h, [(fid, 0xFFFFFFFF)])[0]
                if v.nvmlReturn == 0:
                    tot, seen = int(v.value.ullVal), True
            except Exception:
                pass
            if not seen:
                for link in range(18):
                    try:
                        v = pynvml.nvmlDeviceGetFieldValues(h, [(fid, link)])[0]
                        if v.nvmlReturn == 0:
                            tot += int(v.value.ullVal)
                            seen = True
                    except Exception:
                        break
            if not seen:
                raise RuntimeError("NVML NVLink throughput fields not supported")
            out.append(tot)
        return tuple(out)
    except Exception:
        pass
    try:  # the same counters through the CLI: "Link 3: Data Tx: 1234 KiB"
        import re
        txt = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(index)], capture_output=True, text=True,
                             timeout=10).stdout
        tx = [int(x) for x in re.findall(r"Data Tx:\s*(\d+)\s*KiB", txt)]
        rx = [int(x) for x in re.findall(r"Data Rx:\s*(\d+)\s*KiB", txt)]
        return (sum(tx), sum(rx)) if tx and rx else None
    except Exception:
        return None


# ------------------------------------------------------------------ CPU legs
def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_oracle_run(weights, prompt, n_pos: int, threads: int, arch: str = "native"):
    """tokens/s of the C restatement by the reference's formula (n-1)/(t_end - t_after_first)."""
    from oracle import oracle_c
    try:
        o = oracle_c.Oracle(weights, n_threads=threads, arch=arch)
    except Exception:
        o = oracle_c.Oracle(weights, n_threads=threads)
    t0 = time.perf_counter()
    _, _, ms_after_first = o.generate(prompt, n_pos)
    wall = time.perf_counter() - t0
    o.close()
    return (n_pos - 1) / (ms_after_first / 1000.0), wall


def peaks() -> tuple[float, str]:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(model: str, wtype: str):
    """dram bytes per launch of the decode kernel from the committed ncu --set full capture."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p)).get(f"{model}-{wtype}")
    except Exception:
        return None


def tp1_same_workload(model: str, wtype: str, steps: int, local_rank: int, timeout_s: int = 420):
    """tokens/s of the SAME workload on ONE GPU, measured by a child `bench.py --gpus 1` on this
    rank's GPU while the other ranks wait: the N > 1 line runs Llama-2-7B f16 but the N = 1 default
    is TinyLlama f32, so this is the figure tensor-parallel scaling should be read against.
    Returns (value, note); value None when the child did not produce a line."""
    env = {k: v for k, v in os.environ.items()
           if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "LOCAL_WORLD_SIZE", "GROUP_RANK", "ROLE_RANK",
                        "MASTER_ADDR", "MASTER_PORT") and not k.startswith("TORCHELASTIC")}
    vis = [d for d in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if d]
    env["CUDA_VISIBLE_DEVICES"] = vis[local_rank] if local_rank < len(vis) else str(local_rank)
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", "1", "--model", model, "--wtype", wtype,
           "--steps", str(max(1, min(steps, 5))), "--warmup", "3", "--no-cpu-baseline", "--no-tp1",
           "--no-prefill"]  # like the tensor-parallel run: every position one decode launch
    try:
        out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=timeout_s).stdout
        for ln in reversed(out.strip().splitlines()):
            if ln.startswith("{"):
                d = json.loads(ln)
                if "value" in d:
                    return float(d["value"]), "child bench.py --gpus 1, same workload, same GPU as rank 0"
                return None, str(d.get("error", "no value in child line"))[:200]
        return None, "child printed no JSON line"
    except Exception as e:  # a missing figure must not cost the tensor-parallel line itself
        return None, f"{type(e).__name__}: {e}"[:200]


# ------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default=None)
    ap.add_argument("--wtype", default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-prefill", action="store_true",
                    help="N = 1: run the prompt positions through the per-token kernel too (default: one batched tcgen05 pass)")
    ap.add_argument("--cpu-sample-pos", type=int, default=0, help="positions of the CPU sample (0 = auto)")
    ap.add_argument("--no-tp1", action="store_true", help="N > 1: skip the one-GPU run of the same workload")
    ap.add_argument("--force-tp1", action="store_true", help="N = 1: run the child measurement anyway (test hook)")
    a = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.gpus != world and world > 1:
        a.gpus = world
    # workload: configs[1] at N=1 (TinyLlama-1.1B f32, the configuration the metric is quoted on);
    # N > 1 runs configs[4], Llama-2-7B f16 row-parallel (the configuration BASELINE.json names for 8 GPUs)
    model = a.model or ("tinyllama" if a.gpus == 1 else "llama2-7b")
    wtype = a.wtype or ("f32" if a.gpus == 1 else "f16")
    a.warmup = max(a.warmup, 3) if a.impl == "ours" else a.warmup

    from llm.f90_b200 import fixtures as fx
    from llm.f90_b200.layout import active_weight_bytes
    cfg = model_config(model, wtype)
    workload = f"{model}-{wtype} {N_POS}-position greedy decode, synthetic weights (seed 0)"

    def config_of(parallelism: str, n_prompt: int) -> dict:
        return {"workload": workload, "emb_dim": cfg.emb_dim, "hidden_dim": cfg.hidden_dim,
                "n_layers": cfg.n_layers, "n_heads": cfg.n_heads, "n_kv_heads": cfg.n_kv_heads,
                "vocab_size": cfg.vocab_size, "weight_storage": wtype, "positions_per_step": N_POS,
                "prompt_tokens": n_prompt, "parallelism": parallelism,
                "arithmetic": ARITHMETIC[wtype]}

    if a.impl == "reference":
        # The reference is a single-threaded program (llama2.f90:93 has `use omp_lib` commented out; README.md:97
        # "single thread 32-bit operation"): `value` is the 1-thread figure the north star names; the same C
        # restatement with OpenMP over rows on every host core is reported beside it.
        if rank != 0:
            return 0
        cores = host_cores()
        w = fx.synth_weights_tiled(cfg, 0)
        prompt = prompt_tokens(cfg)
        n_pos = a.cpu_sample_pos or {"tinyllama": 12, "llama2-7b": 4, "small": N_POS}[model]
        for _ in range(a.warmup):
            cpu_oracle_run(w, prompt, min(n_pos, 3), 1)
        tps, walls = [], []
        for _ in range(a.steps):
            t, wall = cpu_oracle_run(w, prompt, n_pos, 1)
            tps.append(t)
            walls.append(wall)
        val = float(len(tps) * (n_pos - 1) / sum((n_pos - 1) / t for t in tps))
        all_tps, _ = cpu_oracle_run(w, prompt, n_pos, cores)
        sample = (f"positions 1..{n_pos} of the {N_POS}-position workload per step (the sample: a CPU pass over all "
                  f"{N_POS} would take minutes), tokens/s by the reference's formula (n-1)/(t_end-t_after_first)")
        cfgd = config_of("cpu, 1 thread", len(prompt))
        cfgd["positions_per_step_sampled"] = n_pos
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "tokens/s", "n_gpus": a.gpus,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1000.0 * sum(walls) / len(walls),
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": DTYPE[wtype],
                "data": "synthetic", "config": cfgd,
                "cpu_baseline": {"value": val, "unit": "tokens/s", "cores": 1, "kind": "port", "sample": sample,
                                 "all_cores": {"value": all_tps, "unit": "tokens/s", "cores": cores,
                                               "how": "same restatement, OpenMP over rows, one step"},
                                 "note": "C restatement of llama2.f90 (oracle/); the Fortran binary cannot be built "
                                         "in this image (no Fortran compiler)"},
                "e2e": {"value": val, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------- our arm
    import torch
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device; this engine has no CPU fallback"}))
        return 2
    from llm.f90_b200 import capi
    capi.load()
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        # gloo carries the IPC handles (objects), nccl the barriers / timing reductions
        dist.init_process_group("cpu:gloo,cuda:nccl", device_id=torch.device("cuda", local_rank))
    dev = local_rank

    w = fx.synth_weights_tiled(cfg, 0)
    prompt = prompt_tokens(cfg)
    # one GPU: the forced prompt positions are one batched pass on the tensor cores (llmf90_b200_prefill; their
    # logits are dead values in the reference, llama2.f90:383-385); tensor-parallel runs take them token by token
    use_prefill = world == 1 and not a.no_prefill
    eng = capi.make_engine(w, device=dev, tp_rank=rank, tp_size=world, prefill=use_prefill)
    act_bytes = active_weight_bytes(cfg)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- value: device-resident loop
    for _ in range(a.warmup):
        eng.generate_greedy(prompt, N_POS)
    barrier()
    l0 = eng.stats()["kernel_launches"]
    dev_ms, after_first = [], []
    nvl0 = nvlink_kib(dev) if world > 1 else None
    t0 = time.perf_counter()
    with ClockSampler(dev) as clk:
        for _ in range(a.steps):
            toks, af = eng.generate_greedy(prompt, N_POS)
            st = eng.stats()
            dev_ms.append(st["last_loop_total_ms"])
            after_first.append(af)
        barrier()
    wall_ms = (time.perf_counter() - t0) * 1000.0
    nvl1 = nvlink_kib(dev) if world > 1 else None
    launches = eng.stats()["kernel_launches"] - l0
    tot_ms = float(sum(dev_ms))
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([tot_ms], device=f"cuda:{dev}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tot_ms = float(t.item())
    value = a.steps * N_POS / (tot_ms / 1000.0)
    ref_formula = a.steps * (N_POS - 1) / (sum(after_first) / 1000.0)
    # the decode kernel on its own for the roofline: N_POS launches of it (positions 1..N_POS fed with their own
    # picks), CUDA events on the engine's stream inside the library; with the prompt pass on, `value` is not
    # N_POS launches of one kernel any more
    decode_ms = []
    if use_prefill:
        for _ in range(max(1, min(a.steps, 5))):
            eng.reset()
            decode_ms.append(eng.bench_device_loop(2, 1, N_POS))
        decode_launch_ms = float(sum(decode_ms)) / (len(decode_ms) * N_POS)
    else:
        decode_launch_ms = tot_ms / max(1, a.steps * N_POS)  # every position of the timed region is one decode launch

    # ---- e2e: the reference-facing per-token call with host logits
    for _ in range(2):
        eng.reset()
        capi.host_generate(eng, prompt, N_POS, prefill=use_prefill)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(a.steps, 5))
    for _ in range(e2e_steps):
        toks_h, _ = capi.host_generate(eng, prompt, N_POS, prefill=use_prefill)
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([e2e_s], device=f"cuda:{dev}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = e2e_steps * N_POS / e2e_s
    tokens_agree = bool((np.asarray(toks_h) == np.asarray(toks)).all())
    # the same host-driven loop with the pick made on the device (llmf90_b200_transformer_sample, temperature 0):
    # 4 bytes come back per position instead of the logits.  Reported beside e2e, not as e2e.
    pick = None
    if world == 1:
        eng.reset()
        capi.host_generate_device_pick(eng, prompt, N_POS, prefill=use_prefill)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            toks_p = capi.host_generate_device_pick(eng, prompt, N_POS, prefill=use_prefill)
        barrier()
        pick = (e2e_steps * N_POS / (time.perf_counter() - t0), bool((np.asarray(toks_p) == np.asarray(toks)).all()))

    if rank != 0:
        barrier()
        eng.close()
        return 0

    peak, peak_src = peaks()
    st = eng.stats()
    per_launch_ms = decode_launch_ms
    n_pf = min(len(prompt), N_POS - 1) if use_prefill else 0
    per_gpu_bytes = st["active_bytes_per_token"]
    achieved = per_gpu_bytes / (per_launch_ms / 1000.0) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": tot_ms / a.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": DTYPE[wtype], "data": "synthetic",
        "config": {**config_of(f"tp{world}", len(prompt)),
                   "l2": f"inputs larger than L2: {act_bytes / 1e6:.0f} MB of weights streamed per token vs 126 MB L2",
                   "tokens_per_s_reference_formula": ref_formula,
                   "prompt_pass": (f"positions 1..{n_pf} (BOS + forced prompt) in one batched tcgen05 pass "
                                   f"(llmf90_b200_prefill), positions {n_pf + 1}..{N_POS} one decode launch each"
                                   if n_pf else "every position one decode launch"),
                   "decode_only_tokens_per_s": 1000.0 / decode_launch_ms,
                   "wall_ms_timed_region": wall_ms},
        "clocks": clk.summary(),
        "e2e": {"value": e2e, "unit": "tokens/s", "h2d_bytes_per_step": 8 * (N_POS - n_pf) + 4 * n_pf,
                "d2h_bytes_per_step": 4 * cfg.vocab_size * (N_POS - n_pf), "steps": e2e_steps,
                "api": ("llmf90_b200_prefill(prompt positions) + " if n_pf else "") +
                       "llmf90_b200_transformer(token,pos,logits) per position, host argmax",
                "tokens_match_device_loop": tokens_agree},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic(model, wtype), "peak_source": peak_src,
                     "kernel": "stream_decode_kernel (1 launch = 1 token)" if st["stream_slots"] else "granular graph",
                     "algorithmic_bytes_per_launch": int(per_gpu_bytes), "launch_ms": per_launch_ms,
                     "timed": (f"{len(decode_ms)} x {N_POS} launches of the decode kernel alone (llmf90_b200_bench_device_loop), "
                               "CUDA events on the engine's stream") if decode_ms else
                              f"the {a.steps} x {N_POS} decode launches of the timed region, CUDA events on the engine's stream"},
    }
    if pick:
        line["e2e"]["device_pick"] = {"value": pick[0], "unit": "tokens/s",
                                      "api": "llmf90_b200_transformer_sample(token,pos,0,r,&next) per position: maxloc next to "
                                             "the logits, 4 bytes back",
                                      "d2h_bytes_per_step": 4 * (N_POS - n_pf), "tokens_match_device_loop": pick[1]}
    if model == "tinyllama" and wtype == "f32":
        # the one throughput figure the reference publishes; another workload
        # (-n 96, sampled, unspecified CPU), so it is quoted, not divided into vs_baseline
        line["config"]["reference_published"] = {"value": 4.217, "unit": "tokens/s",
                                                 "source": "README.md:76: TinyLlama f32, -n 96 -t 0.9, one thread, "
                                                           "unspecified Intel CPU"}
    if not a.no_cpu_baseline and world == 1:
        n_pos = a.cpu_sample_pos or {"tinyllama": 64, "llama2-7b": 6, "small": N_POS}[model]
        tps, wall = cpu_oracle_run(w, prompt, n_pos, 1)
        line["cpu_baseline"] = {"value": tps, "unit": "tokens/s", "cores": 1, "kind": "port",
                                "sample": f"first {n_pos} of the {N_POS} positions, 1 thread, {wall:.1f} s wall; "
                                          f"C restatement of llama2.f90 (no Fortran compiler in the image); "
                                          f"box has {host_cores()} host cores"}
    if world > 1:
        # what the fused all-reduce should move per token and GPU: twice per layer the rank's partial vector (emb LL
        # words of 8 bytes) to each of the other ranks, times the hand-over replicas; + its logits rows to each peer
        rep = 2 if world == 2 else 1
        expect = (2 * cfg.n_layers * cfg.emb_dim * 8 * rep + (cfg.vocab_size // world) * 4) * (world - 1)
        nv = {"expected_tx_bytes_per_token_per_gpu": int(expect),
              "how": "NVML NVLINK_THROUGHPUT_DATA_TX / _RX of rank 0's GPU before / after the timed region (device loop)"}
        if nvl0 and nvl1:
            toks_timed = a.steps * N_POS
            nv["tx_bytes_per_token"] = (nvl1[0] - nvl0[0]) * 1024.0 / toks_timed
            nv["rx_bytes_per_token"] = (nvl1[1] - nvl0[1]) * 1024.0 / toks_timed
        else:
            nv["tx_bytes_per_token"] = nv["rx_bytes_per_token"] = None
            nv["unavailable"] = ("the driver does not expose the NVLink throughput counters here (NVML field values "
                                 "unsupported, `nvidia-smi nvlink -gt d` prints N/A)")
        line["nvlink"] = nv
    if (world > 1 and not a.no_tp1) or a.force_tp1:
        v1, note = tp1_same_workload(model, wtype, a.steps, local_rank)
        line["config"]["tp1_same_workload"] = {"value": v1, "unit": "tokens/s", "how": note}
    if world > 1:
        barrier()
    eng.close()
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())

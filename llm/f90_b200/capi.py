"""ctypes binding of libllmf90_b200.so (include/llmf90_b200.h).

Plumbing for the tests and the benchmark: it passes numpy host buffers straight through the
C ABI, exactly as the Fortran host would pass its allocatables.  If the library is missing it
raises -- there is no Python or CPU fallback for any operator.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .layout import Config, Weights

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libllmf90_b200.so")

FLAG_GRANULAR = 1
FLAG_PROFILE = 2  # fused kernel with per-phase timers (slower; phase_times / debug_trace tools)
FLAG_CLS_Q6K = 8  # wcls is ggml Q6_K super-blocks (stock llama.cpp q4_0 files); implies the granular engine
FLAG_PREFILL = 4  # second copy of the layer matrices in tensor-core operand order: batched prompt pass

EXPORTS = [
    "llmf90_b200_init", "llmf90_b200_transformer", "llmf90_b200_times", "llmf90_b200_reset",
    "llmf90_b200_free", "llmf90_b200_last_error", "llmf90_b200_generate_greedy",
    "llmf90_b200_matvec", "llmf90_b200_rmsnorm", "llmf90_b200_softmax", "llmf90_b200_rope",
    "llmf90_b200_tp_export", "llmf90_b200_tp_connect", "llmf90_b200_get_stats",
    "llmf90_b200_bench_device_loop", "llmf90_b200_phase_times", "llmf90_b200_debug_trace",
    "llmf90_b200_plan", "llmf90_b200_prefill", "llmf90_b200_debug_read_kv", "llmf90_b200_matmul",
    "llmf90_b200_transformer_sample", "llmf90_b200_prefill_plan",
]


class CConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("emb_dim", "hidden_dim", "n_layers", "n_heads", "n_kv_heads",
                                         "vocab_size", "seq_len", "wtype", "device", "tp_rank",
                                         "tp_size")] + [("flags", C.c_uint32)]


class CStats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("forward_calls", C.c_uint64),
                ("weight_bytes_device", C.c_uint64), ("active_bytes_per_token", C.c_uint64),
                ("last_forward_ms", C.c_float), ("n_sms", C.c_int32), ("stream_slots", C.c_int32),
                ("stream_slot_bytes", C.c_int32), ("stream_smem_bytes", C.c_int32),
                ("stream_threads", C.c_int32), ("last_loop_total_ms", C.c_float),
                ("last_loop_after_first_ms", C.c_float)]


class CPlanInfo(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("grid", "threads", "n_slots", "slot_bytes", "smem_bytes", "sched_stride",
                                         "n_layers")] + \
               [("rows", C.c_int32 * 5), ("cols", C.c_int32 * 5), ("tile_rows", C.c_int32 * 5),
                ("tile_chunks", C.c_int32 * 5), ("tile_warps", C.c_int32 * 5), ("reserved0", C.c_int32),
                ("matrix_bytes", C.c_uint64 * 5),
                ("vector_bytes", C.c_uint64), ("emb_row_bytes", C.c_uint64)]


class CPrefillGemm(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("rows", "cols", "planes", "m_tiles", "k_chunks", "chunks_per_split", "n_splits",
                                         "ppad", "tmem_cols", "stages", "stage_bytes", "smem_bytes")] + \
               [("weight_bytes", C.c_uint64), ("partial_bytes", C.c_uint64)]


SCHED_DTYPE = np.dtype([("src", np.uint64), ("bytes", np.uint32), ("layer_stride16", np.uint32),
                        ("phase_start", np.uint32), ("stage", np.uint32)])
B200_SMS, B200_SMEM_OPTIN = 148, 232448


def plan_vbase(k: int) -> int:
    """LLMF90_PLAN_VBASE(k): start of region k of the planner's virtual address space."""
    return (k + 1) << 40


class EngineError(RuntimeError):
    pass


_lib = None


def load() -> C.CDLL:
    """Load the shared library (building it is __graft_entry__.build()'s job)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineError(f"{LIB_PATH} is missing: run `python -m llm.f90_b200.build` "
                          "(there is no fallback implementation)")
    L = C.CDLL(LIB_PATH)
    vp, fp, ip = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int32)
    L.llmf90_b200_init.argtypes = [C.POINTER(CConfig)] + [vp] * 9
    L.llmf90_b200_transformer.argtypes = [C.c_int32, C.c_int32, fp]
    L.llmf90_b200_times.argtypes = [fp]
    L.llmf90_b200_phase_times.argtypes = [fp, C.c_int32]
    L.llmf90_b200_debug_trace.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_uint64), C.c_int32]
    L.llmf90_b200_last_error.restype = C.c_char_p
    L.llmf90_b200_generate_greedy.argtypes = [ip, C.c_int32, C.c_int32, ip, fp]
    L.llmf90_b200_matvec.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, fp, fp]
    L.llmf90_b200_rmsnorm.argtypes = [fp, fp, C.c_int32, fp]
    L.llmf90_b200_softmax.argtypes = [fp, C.c_int32, C.c_int32, fp]
    L.llmf90_b200_rope.argtypes = [fp, fp, C.c_int32, C.c_int32, C.c_int32, C.c_int32]
    L.llmf90_b200_tp_export.argtypes = [vp]
    L.llmf90_b200_tp_connect.argtypes = [vp, C.c_int32]
    L.llmf90_b200_get_stats.argtypes = [C.POINTER(CStats)]
    L.llmf90_b200_bench_device_loop.argtypes = [C.c_int32, C.c_int32, C.c_int32, fp]
    L.llmf90_b200_transformer_sample.argtypes = [C.c_int32, C.c_int32, C.c_float, C.c_float, ip]
    L.llmf90_b200_prefill_plan.argtypes = [C.POINTER(CConfig), C.c_int32, C.c_int32, C.POINTER(CPrefillGemm)]
    L.llmf90_b200_prefill.argtypes = [ip, C.c_int32, C.c_int32]
    L.llmf90_b200_debug_read_kv.argtypes = [C.c_int32, C.c_int32, fp, fp]
    L.llmf90_b200_matmul.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, fp, C.c_int32, fp]
    L.llmf90_b200_plan.argtypes = [C.POINTER(CConfig), C.c_int32, C.c_int32, C.POINTER(CPlanInfo), vp, C.c_int64]
    for name in EXPORTS:
        if name != "llmf90_b200_last_error":
            getattr(L, name).restype = C.c_int
    _lib = L
    return L


def _check(rc: int) -> None:
    if rc != 0:
        raise EngineError(load().llmf90_b200_last_error().decode(errors="replace"))


def _fp(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a: np.ndarray):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_int32))


class Engine:
    """The process-wide engine singleton behind the C ABI."""

    def __init__(self, weights: Weights, device: int = 0, granular: bool = False, tp_rank: int = 0,
                 tp_size: int = 1, profile: bool = False, prefill: bool = False):
        self.L = load()
        c = weights.cfg
        self.cfg = c
        self.tp_rank, self.tp_size = tp_rank, tp_size
        cc = CConfig(c.emb_dim, c.hidden_dim, c.n_layers, c.n_heads, c.n_kv_heads, c.vocab_size,
                     c.seq_len, c.wtype, device, tp_rank, tp_size,
                     (FLAG_GRANULAR if granular else 0) | (FLAG_PROFILE if profile else 0) |
                     (FLAG_PREFILL if prefill else 0) |
                     (FLAG_CLS_Q6K if getattr(weights, "cls_wtype", c.wtype) == 14 else 0))
        ptr = lambda a: a.ctypes.data_as(C.c_void_p)
        _check(self.L.llmf90_b200_init(C.byref(cc), ptr(weights.token_embedding_table),
                                       ptr(weights.rms_att_weight), ptr(weights.wqkv), ptr(weights.wo),
                                       ptr(weights.rms_ffn_weight), ptr(weights.w13), ptr(weights.w2),
                                       ptr(weights.rms_final_weight), ptr(weights.wcls)))
        self._logits = np.empty(c.vocab_size, np.float32)
        self.open = True

    def tp_export(self) -> bytes:
        """This rank's 64-byte CUDA IPC handle of the buffers the other ranks write into."""
        buf = C.create_string_buffer(64)
        _check(self.L.llmf90_b200_tp_export(buf))
        return buf.raw

    def tp_connect(self, handles: list[bytes]) -> None:
        """Map the other ranks' buffers (handles in rank order, one per rank)."""
        assert len(handles) == self.tp_size and all(len(h) == 64 for h in handles)
        blob = C.create_string_buffer(b"".join(handles), 64 * len(handles))
        _check(self.L.llmf90_b200_tp_connect(blob, len(handles)))

    def transformer(self, token: int, pos: int, out: np.ndarray | None = None) -> np.ndarray:
        """logits = transformer(token, pos) with 1-based token/pos (llama2.f90:380)."""
        if out is None:
            # one buffer for the engine's lifetime: the library page-locks the array it is given in place, and a
            # fresh array per call would mean one registration per token (and stale ones over freed memory)
            _check(self.L.llmf90_b200_transformer(token, pos, _fp(self._logits)))
            return self._logits.copy()
        _check(self.L.llmf90_b200_transformer(token, pos, _fp(out)))
        return out

    def transformer_sample(self, token: int, pos: int, temperature: float, r: float) -> int:
        """transformer() + the pick of the next token on the device (maxloc, or softmax / T + CDF walk against r)."""
        nxt = C.c_int32(0)
        _check(self.L.llmf90_b200_transformer_sample(token, pos, temperature, r, C.byref(nxt)))
        return nxt.value

    def prefill(self, tokens, pos0: int = 1) -> None:
        """KV rows of positions pos0.. for the input tokens, as one batched tcgen05 pass (llama2.f90:379-385)."""
        t = np.ascontiguousarray(tokens, np.int32)
        _check(self.L.llmf90_b200_prefill(_ip(t), len(t), pos0))

    def read_kv(self, layer: int, pos: int):
        """The key and value rows of cache position pos (1-based) in `layer` (0-based)."""
        n = self.cfg.n_kv_heads * (self.cfg.emb_dim // self.cfg.n_heads) // max(1, min(self.tp_size, self.cfg.n_kv_heads))
        k, v = np.empty(n, np.float32), np.empty(n, np.float32)
        _check(self.L.llmf90_b200_debug_read_kv(layer, pos, _fp(k), _fp(v)))
        return k, v

    def generate_greedy(self, prompt_tokens, n: int):
        pt = np.ascontiguousarray(prompt_tokens, np.int32)
        out = np.empty(n, np.int32)
        ms = C.c_float(0)
        _check(self.L.llmf90_b200_generate_greedy(_ip(pt) if len(pt) else None, len(pt), n, _ip(out),
                                                  C.byref(ms)))
        return out, ms.value

    def bench_device_loop(self, first_token: int, pos0: int, n_steps: int) -> float:
        ms = C.c_float(0)
        _check(self.L.llmf90_b200_bench_device_loop(first_token, pos0, n_steps, C.byref(ms)))
        return ms.value

    def times(self) -> np.ndarray:
        t = np.zeros(5, np.float32)
        _check(self.L.llmf90_b200_times(_fp(t)))
        return t

    PHASES = ("qkv_pro", "qkv_mv", "rope_bar", "att", "att_bar", "wo_pro", "wo_mv", "wo_bar", "w13_pro",
              "w13_mv", "w13_bar", "w2_pro", "w2_mv", "w2_bar", "cls_pro", "cls_mv", "argmax")

    def phase_times(self) -> dict:
        t = np.zeros(len(self.PHASES), np.float32)
        _check(self.L.llmf90_b200_phase_times(_fp(t), len(t)))
        return dict(zip(self.PHASES, (float(v) for v in t)))

    def debug_trace(self, token: int, pos: int, layer: int) -> np.ndarray:
        n = self.stats()["n_sms"]
        out = np.zeros((n, 128), np.uint64)
        _check(self.L.llmf90_b200_debug_trace(token, pos, layer, out.ctypes.data_as(C.POINTER(C.c_uint64)), n))
        return out

    def reset(self) -> None:
        _check(self.L.llmf90_b200_reset())

    def stats(self) -> dict:
        s = CStats()
        _check(self.L.llmf90_b200_get_stats(C.byref(s)))
        return {n: getattr(s, n) for n, _ in CStats._fields_}

    def close(self) -> None:
        if getattr(self, "open", False):
            self.L.llmf90_b200_free()
            self.open = False

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def host_generate(engine: Engine, prompt_tokens, n: int, want_logits: bool = False, prefill: bool = False):
    """The reference's token loop (llama2.f90:376-393, temperature 0) driving the C ABI one
    token at a time with host logits -- what the Fortran host does.  prefill=True: the forced prompt
    positions (inputs BOS, prompt[:-1]; their logits are never looked at, llama2.f90:383-385) go through
    llmf90_b200_prefill as one batched pass and the loop starts at the first position whose logits are used."""
    V = engine.cfg.vocab_size
    toks = np.empty(n, np.int32)
    lg_all = np.full((n, V), np.nan, np.float32) if want_logits else None
    buf = np.empty(V, np.float32)
    token = 2
    m = min(len(prompt_tokens), n - 1) if prefill else 0
    if m >= 1:
        engine.prefill([2] + [int(t) for t in prompt_tokens[:m - 1]], 1)
        toks[:m] = prompt_tokens[:m]
        token = int(prompt_tokens[m - 1])
    for pos in range(m + 1, n + 1):
        engine.transformer(token, pos, buf)
        if want_logits:
            lg_all[pos - 1] = buf
        token = int(prompt_tokens[pos - 1]) if pos <= len(prompt_tokens) else int(np.argmax(buf)) + 1
        toks[pos - 1] = token
    return toks, lg_all


def host_generate_device_pick(engine: Engine, prompt_tokens, n: int, temperature: float = 0.0, rng=None,
                              prefill: bool = False) -> np.ndarray:
    """The same token loop with the pick made next to the logits (llmf90_b200_transformer_sample): one call per
    position, the token comes back instead of the logits.  rng draws the uniform number of a sampled position, in
    the host, like random_number at llama2.f90:433."""
    toks = np.empty(n, np.int32)
    token = 2
    m = min(len(prompt_tokens), n - 1) if prefill else 0
    if m >= 1:
        engine.prefill([2] + [int(t) for t in prompt_tokens[:m - 1]], 1)
        toks[:m] = prompt_tokens[:m]
        token = int(prompt_tokens[m - 1])
    for pos in range(m + 1, n + 1):
        forced = pos <= len(prompt_tokens)
        r = float(rng.random()) if (rng is not None and temperature != 0 and not forced) else 0.0
        nxt = engine.transformer_sample(token, pos, temperature, r)
        token = int(prompt_tokens[pos - 1]) if forced else nxt
        toks[pos - 1] = token
    return toks


def plan(cfg: Config, tp_rank: int = 0, tp_size: int = 1, n_sms: int = B200_SMS,
         smem_optin: int = B200_SMEM_OPTIN):
    """llmf90_b200_plan: the fused kernel's grid / ring / per-CTA stage lists for a configuration,
    computed on the host (no device needed).  Returns (info dict, stages[grid, sched_stride])."""
    L = load()
    cc = CConfig(cfg.emb_dim, cfg.hidden_dim, cfg.n_layers, cfg.n_heads, cfg.n_kv_heads, cfg.vocab_size,
                 cfg.seq_len, cfg.wtype, 0, tp_rank, tp_size, 0)
    info = CPlanInfo()
    _check(L.llmf90_b200_plan(C.byref(cc), n_sms, smem_optin, C.byref(info), None, 0))
    st = np.zeros((info.grid, info.sched_stride), SCHED_DTYPE)
    _check(L.llmf90_b200_plan(C.byref(cc), n_sms, smem_optin, C.byref(info), st.ctypes.data_as(C.c_void_p), st.size))
    d = {}
    for n, _ in CPlanInfo._fields_:
        v = getattr(info, n)
        d[n] = list(v) if hasattr(v, "__len__") else v
    return d, st


# ---- operators
def matvec(w: np.ndarray, wtype: int, rows: int, cols: int, x: np.ndarray) -> np.ndarray:
    y = np.empty(rows, np.float32)
    w = np.ascontiguousarray(w)
    _check(load().llmf90_b200_matvec(w.ctypes.data_as(C.c_void_p), wtype, rows, cols, _fp(x), _fp(y)))
    return y


def prefill_plan(cfg: Config, n_pos: int, n_sms: int = B200_SMS) -> list[dict]:
    """The four GEMMs of a batched prompt pass over n_pos positions, computed without a device."""
    cc = CConfig(cfg.emb_dim, cfg.hidden_dim, cfg.n_layers, cfg.n_heads, cfg.n_kv_heads, cfg.vocab_size, cfg.seq_len,
                 cfg.wtype, 0, 0, 1, FLAG_PREFILL)
    out = (CPrefillGemm * 4)()
    _check(load().llmf90_b200_prefill_plan(C.byref(cc), n_sms, n_pos, out))
    return [{n: int(getattr(g, n)) for n, _ in CPrefillGemm._fields_} for g in out]


def matmul(w: np.ndarray, wtype: int, rows: int, cols: int, x: np.ndarray) -> np.ndarray:
    """y[p] = W x[p] for the rows of x at once: the tcgen05 GEMM of the batched prompt pass."""
    x = np.ascontiguousarray(x, np.float32)
    y = np.empty((x.shape[0], rows), np.float32)
    w = np.ascontiguousarray(w)
    _check(load().llmf90_b200_matmul(w.ctypes.data_as(C.c_void_p), wtype, rows, cols, _fp(x), x.shape[0], _fp(y)))
    return y


def rmsnorm(x: np.ndarray, w: np.ndarray) -> np.ndarray:
    out = np.empty_like(x)
    _check(load().llmf90_b200_rmsnorm(_fp(x), _fp(w), len(x), _fp(out)))
    return out


def softmax(x: np.ndarray, s: int) -> np.ndarray:
    out = np.empty_like(x)
    _check(load().llmf90_b200_softmax(_fp(x), len(x), s, _fp(out)))
    return out


def rope(q: np.ndarray, k: np.ndarray, head_size: int, pos: int):
    q, k = q.copy(), k.copy()
    _check(load().llmf90_b200_rope(_fp(q), _fp(k), len(q), len(k), head_size, pos))
    return q, k


def make_engine(weights: Weights, device: int = 0, tp_rank: int = 0, tp_size: int = 1,
                granular: bool = False, prefill: bool = False) -> Engine:
    """Single GPU, or one tensor-parallel rank of `tp_size` (one process per GPU).  For tp_size > 1
    torch.distributed must be initialised: it carries the 64-byte IPC handles, nothing else -- the
    all-reduces of the forward run inside the decode kernel over NVLink peer stores."""
    eng = Engine(weights, device=device, granular=granular, tp_rank=tp_rank, tp_size=tp_size, prefill=prefill)
    if tp_size > 1:
        import torch.distributed as dist
        handles: list = [None] * tp_size
        dist.all_gather_object(handles, eng.tp_export())
        eng.tp_connect(handles)
        dist.barrier()
    return eng

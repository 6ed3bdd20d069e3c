"""ctypes binding of oracle/_build/liboracle.so (the C restatement of llama2.f90:450-640).

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (no reference golden vectors exist; see the
header of llama2_oracle.c).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class OracleCfg(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("emb_dim", "hidden_dim", "n_layers", "n_heads", "n_kv_heads",
                                       "vocab_size", "seq_len", "wtype", "canonical", "n_threads")]


def build(arch: str = "x86-64-v3", force: bool = False) -> str:
    out = "_build/liboracle.so" if arch == "x86-64-v3" else f"_build/liboracle_{arch}.so"
    path = os.path.join(HERE, out)
    src = os.path.join(HERE, "llama2_oracle.c")
    if force or not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, f"ARCH={arch}", f"OUT={out}"], check=True,
                       stdout=subprocess.DEVNULL)
    return path


_libs: dict[str, C.CDLL] = {}


def lib(arch: str = "x86-64-v3") -> C.CDLL:
    if arch in _libs:
        return _libs[arch]
    L = C.CDLL(build(arch))
    vp, fp, ip = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int)
    L.oracle_create.restype = vp
    L.oracle_create.argtypes = [C.POINTER(OracleCfg)] + [vp] * 9
    L.oracle_free.argtypes = [vp]
    L.oracle_set_cls_type.argtypes = [vp, C.c_int]
    L.oracle_reset.argtypes = [vp]
    L.oracle_transformer.restype = C.c_int
    L.oracle_transformer.argtypes = [vp, C.c_int, C.c_int, fp]
    L.oracle_generate.restype = C.c_double
    L.oracle_generate.argtypes = [vp, ip, C.c_int, C.c_int, ip, fp]
    L.oracle_times.argtypes = [vp, C.POINTER(C.c_double)]
    L.oracle_rmsnorm.argtypes = [fp, fp, C.c_int, fp]
    L.oracle_softmax.argtypes = [fp, C.c_int, C.c_int, fp]
    L.oracle_matvec.argtypes = [vp, C.c_int, C.c_int, C.c_int, fp, fp]
    L.oracle_rope.argtypes = [fp, fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.oracle_dequant_row.argtypes = [vp, C.c_int, C.c_int, fp]
    L.oracle_argmax1.restype = C.c_int
    L.oracle_argmax1.argtypes = [fp, C.c_int]
    L.oracle_max_threads.restype = C.c_int
    _libs[arch] = L
    return L


def _fp(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a: np.ndarray):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_int))


class Oracle:
    """One model instance; borrows the numpy weight arrays (keeps references alive)."""

    def __init__(self, weights, canonical: bool = False, n_threads: int = 1, arch: str = "x86-64-v3"):
        self.L = lib(arch)
        self.w = weights
        c = weights.cfg
        self.cfg = c
        oc = OracleCfg(c.emb_dim, c.hidden_dim, c.n_layers, c.n_heads, c.n_kv_heads, c.vocab_size,
                       c.seq_len, c.wtype, int(canonical), n_threads)
        ptr = lambda a: a.ctypes.data_as(C.c_void_p)
        self.h = self.L.oracle_create(C.byref(oc), ptr(weights.token_embedding_table),
                                      ptr(weights.rms_att_weight), ptr(weights.wqkv), ptr(weights.wo),
                                      ptr(weights.rms_ffn_weight), ptr(weights.w13), ptr(weights.w2),
                                      ptr(weights.rms_final_weight), ptr(weights.wcls))
        if getattr(weights, "cls_wtype", c.wtype) != c.wtype:  # Q6_K output.weight of a q4_0 file
            self.L.oracle_set_cls_type(self.h, weights.cls_wtype)

    def close(self):
        if self.h:
            self.L.oracle_free(self.h)
            self.h = None

    __del__ = close

    def reset(self):
        self.L.oracle_reset(self.h)

    def transformer(self, token: int, pos: int) -> np.ndarray:
        """token, pos are 1-based (llama2.f90:380)."""
        out = np.empty(self.cfg.vocab_size, np.float32)
        rc = self.L.oracle_transformer(self.h, token, pos, _fp(out))
        if rc:
            raise ValueError(f"oracle_transformer: bad token/pos {token}/{pos}")
        return out

    def generate(self, prompt_tokens, n: int, want_logits: bool = False):
        """Greedy run of n positions.  Returns (tokens[n] 1-based, logits[n,V] or None, ms)."""
        pt = np.ascontiguousarray(prompt_tokens, np.int32)
        out = np.empty(n, np.int32)
        lg = np.empty((n, self.cfg.vocab_size), np.float32) if want_logits else None
        ms = self.L.oracle_generate(self.h, _ip(pt) if len(pt) else None, len(pt), n, _ip(out),
                                    _fp(lg) if want_logits else None)
        return out, lg, ms

    def times(self) -> np.ndarray:
        t = (C.c_double * 5)()
        self.L.oracle_times(self.h, t)
        return np.array(list(t))


def rmsnorm(x, w):
    out = np.empty_like(x)
    lib().oracle_rmsnorm(_fp(x), _fp(w), len(x), _fp(out))
    return out


def softmax(x, s):
    out = np.empty_like(x)
    lib().oracle_softmax(_fp(x), len(x), s, _fp(out))
    return out


def matvec(w, wtype, rows, cols, x):
    y = np.empty(rows, np.float32)
    lib().oracle_matvec(w.ctypes.data_as(C.c_void_p), wtype, rows, cols, _fp(x), _fp(y))
    return y


def rope(q, k, head_size, pos, canonical=False):
    q, k = q.copy(), k.copy()
    lib().oracle_rope(_fp(q), _fp(k), len(q), len(k), head_size, pos, int(canonical))
    return q, k


def dequant_row(row, wtype, n):
    out = np.empty(n, np.float32)
    lib().oracle_dequant_row(row.ctypes.data_as(C.c_void_p), wtype, n, _fp(out))
    return out

"""Array layout of the drop-in boundary (the reference's weight_module.f90:13-40).

The Fortran derived type ``TransformerWeights`` holds column-major allocatables.
Seen from C / numpy (row-major) every array has its index order reversed, so a
weight "row" (one output feature, contiguous contraction index) is the unit of
storage:

    token_embedding_table(emb, V)   -> [V][emb]
    rms_att_weight(emb, L)          -> [L][emb]            (always f32)
    wqkv(emb, emb+2kv, L)           -> [L][emb+2kv][emb]   rows Wq | Wk | Wv   (read_ggml.f90:272,286,300)
    wo(emb, emb, L)                 -> [L][emb][emb]
    rms_ffn_weight(emb, L)          -> [L][emb]            (always f32)
    w13(emb, 2hid, L)               -> [L][2hid][emb]      rows W1(gate) | W3(up) (read_ggml.f90:347,376)
    w2(hid, emb, L)                 -> [L][emb][hid]
    rms_final_weight(emb)           -> [emb]               (always f32)
    wcls(emb, V)                    -> [V][emb]

``wtype`` selects the storage of the 2-D tensors: 0 = f32 (what the reference's
master branch runs), 1 = f16 (ggml type 1), 2 = q4_0 (ggml type 2: 18-byte blocks
of 32 weights).  For f16 / q4_0 the arrays are raw bytes with the same row order.
"""
from __future__ import annotations

from dataclasses import dataclass, asdict

import numpy as np

F32, F16, Q4_0 = 0, 1, 2
WTYPE_NAMES = {F32: "f32", F16: "f16", Q4_0: "q4_0"}
WTYPE_BY_NAME = {v: k for k, v in WTYPE_NAMES.items()}
GGML_TYPE = {F32: 0, F16: 1, Q4_0: 2, 14: 14}  # ggml tensor type ids used in GGUF tensor infos (14 = Q6_K)
QK4_0 = 32  # load.f90:8 (qk4)
Q4_0_BLOCK_BYTES = 18
Q6_K = 14  # ggml type 14: only as the storage of wcls (the output.weight of stock llama.cpp q4_0 files)
QK_K, Q6_K_BLOCK_BYTES = 256, 210


@dataclass(frozen=True)
class Config:
    """Mirror of ``type Config`` (weight_module.f90:28-31) plus the weight dtype."""

    emb_dim: int
    hidden_dim: int
    n_layers: int
    n_heads: int
    n_kv_heads: int
    vocab_size: int
    seq_len: int
    wtype: int = F32

    @property
    def head_size(self) -> int:  # llama2.f90:153
        return self.emb_dim // self.n_heads

    @property
    def kv_head_size(self) -> int:  # llama2.f90:154 (n_kv_heads * head_size)
        return self.n_kv_heads * self.head_size

    @property
    def n_qkv(self) -> int:
        return self.emb_dim + 2 * self.kv_head_size

    def asdict(self):
        return asdict(self)

    def validate(self) -> None:
        assert self.emb_dim % self.n_heads == 0
        assert self.n_heads % self.n_kv_heads == 0
        assert self.head_size % 2 == 0
        assert self.wtype in WTYPE_NAMES
        if self.wtype == Q4_0:
            assert self.emb_dim % QK4_0 == 0 and self.hidden_dim % QK4_0 == 0
        if self.wtype == F16:
            assert self.emb_dim % 8 == 0 and self.hidden_dim % 8 == 0
        assert self.emb_dim % 4 == 0 and self.hidden_dim % 4 == 0


def row_bytes(wtype: int, n: int) -> int:
    if wtype == F32:
        return 4 * n
    if wtype == F16:
        return 2 * n
    if wtype == Q6_K:
        assert n % QK_K == 0
        return n // QK_K * Q6_K_BLOCK_BYTES
    assert n % QK4_0 == 0
    return n // QK4_0 * Q4_0_BLOCK_BYTES


# Named configurations from BASELINE.json / SURVEY.md 8 (dims only; weights are synthetic).
TINYLLAMA = dict(emb_dim=2048, hidden_dim=5632, n_layers=22, n_heads=32, n_kv_heads=4,
                 vocab_size=32000, seq_len=2048)  # llama2.f90:102-108
LLAMA2_7B = dict(emb_dim=4096, hidden_dim=11008, n_layers=32, n_heads=32, n_kv_heads=32,
                 vocab_size=32000, seq_len=2048)
TINY = dict(emb_dim=128, hidden_dim=352, n_layers=2, n_heads=4, n_kv_heads=2,
            vocab_size=512, seq_len=64)  # CI-sized
SMALL = dict(emb_dim=512, hidden_dim=1408, n_layers=3, n_heads=8, n_kv_heads=2,
             vocab_size=2048, seq_len=256)


def active_weight_bytes(cfg: Config) -> int:
    """Algorithmic bytes per decoded token: every 2-D weight once,
    the f32 norm vectors, one embedding row."""
    e, h, L, V, kv = cfg.emb_dim, cfg.hidden_dim, cfg.n_layers, cfg.vocab_size, cfg.kv_head_size
    rb_e, rb_h = row_bytes(cfg.wtype, e), row_bytes(cfg.wtype, h)
    per_layer = (e + 2 * kv) * rb_e + e * rb_e + 2 * h * rb_e + e * rb_h + 2 * e * 4
    return L * per_layer + V * rb_e + e * 4 + rb_e


class Weights:
    """``TransformerWeights`` as numpy arrays in the C view of the Fortran layout."""

    FIELDS = ("token_embedding_table", "rms_att_weight", "wqkv", "wo", "rms_ffn_weight",
              "w13", "w2", "rms_final_weight", "wcls")

    def __init__(self, cfg: Config, cls_wtype: int | None = None, **arrays: np.ndarray):
        self.cfg = cfg
        # storage of wcls: the model's wtype, or Q6_K (ggml type 14) for a stock llama.cpp q4_0 file
        self.cls_wtype = cfg.wtype if cls_wtype is None else cls_wtype
        for f in self.FIELDS:
            a = np.ascontiguousarray(arrays[f])
            setattr(self, f, a)
        self.check()

    def check(self) -> None:
        c = self.cfg
        e, h, L, V = c.emb_dim, c.hidden_dim, c.n_layers, c.vocab_size
        rb_e, rb_h = row_bytes(c.wtype, e), row_bytes(c.wtype, h)
        expect = {
            "token_embedding_table": V * rb_e,
            "wqkv": L * c.n_qkv * rb_e,
            "wo": L * e * rb_e,
            "w13": L * 2 * h * rb_e,
            "w2": L * e * rb_h,
            "wcls": V * row_bytes(self.cls_wtype, e),
            "rms_att_weight": L * e * 4,
            "rms_ffn_weight": L * e * 4,
            "rms_final_weight": e * 4,
        }
        for f, nbytes in expect.items():
            a = getattr(self, f)
            assert a.nbytes == nbytes, (f, a.nbytes, nbytes)
            if f.startswith("rms"):
                assert a.dtype == np.float32, f

    def nbytes(self) -> int:
        return sum(getattr(self, f).nbytes for f in self.FIELDS)

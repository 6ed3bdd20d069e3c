"""Tensor-parallel sharding of the weight_module layout -- host-side logic.

The C library shards on upload (csrc/engine.cu); this module states the same split in numpy so
that it can be tested on CPU (tests/test_tp_cpu.py, gloo) and used for bookkeeping:

    attention  heads h in [rank*H/tp, (rank+1)*H/tp): their Wq rows, the Wk / Wv rows of their KV
               heads (KVH % tp == 0: a rank's heads map onto its own KV heads; tp % KVH == 0 with
               fewer KV heads than ranks: all of a rank's heads map onto ONE KV head, h / kv_mul, which
               is replicated on tp / KVH ranks), and the matching INPUT COLUMNS of Wo
               ->  partial Wo output, summed over ranks
    FFN        rows [rank*hid/tp, ...) of W1 and W3 and the matching input columns of W2
               ->  partial W2 output, summed over ranks
    classifier vocabulary rows [rank*V/tp, ...), logits all-gathered
    rmsnorm, residual stream, embedding table: replicated
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .fixtures import decode_matrix
from .layout import Config, Weights, F16, Q4_0, row_bytes


@dataclass(frozen=True)
class Shard:
    rank: int
    size: int
    heads: range
    kv_heads: range
    q_rows: range       # rows of wqkv (Wq part)
    k_rows: range       # rows of wqkv (Wk part)
    v_rows: range       # rows of wqkv (Wv part)
    att_cols: range     # input columns of Wo
    ffn_rows: range     # rows of W1 (gate); W3 rows are hidden_dim + the same
    vocab_rows: range


def check(cfg: Config, size: int) -> None:
    if size not in (1, 2, 4, 8):
        raise ValueError(f"tp size {size} not in 1, 2, 4, 8")
    if cfg.n_heads % size or (cfg.n_kv_heads % size and size % cfg.n_kv_heads):
        raise ValueError("n_heads must be a multiple of the tp size, n_kv_heads a multiple or a divisor")
    colmul = 32 if cfg.wtype == Q4_0 else (8 if cfg.wtype == F16 else 4)
    if cfg.hidden_dim % (size * colmul) or (cfg.emb_dim // size) % colmul or cfg.vocab_size % size:
        raise ValueError("hidden_dim / emb_dim / vocab_size do not split that many ways for this wtype")


def shard(cfg: Config, rank: int, size: int) -> Shard:
    check(cfg, size)
    hs, e, kv = cfg.head_size, cfg.emb_dim, cfg.kv_head_size
    hl, kvhl = cfg.n_heads // size, max(1, cfg.n_kv_heads // size)
    att, kvl, hid, vl = hl * hs, kvhl * hs, cfg.hidden_dim // size, cfg.vocab_size // size
    kvh0 = rank * cfg.n_kv_heads // size  # first KV head of this rank (shared with its neighbours when replicated)
    k0 = kvh0 * hs
    return Shard(rank, size, range(rank * hl, (rank + 1) * hl), range(kvh0, kvh0 + kvhl),
                 range(rank * att, (rank + 1) * att), range(e + k0, e + k0 + kvl),
                 range(e + kv + k0, e + kv + k0 + kvl), range(rank * att, (rank + 1) * att),
                 range(rank * hid, (rank + 1) * hid), range(rank * vl, (rank + 1) * vl))


def active_bytes_per_rank(cfg: Config, size: int) -> int:
    """Algorithmic weight bytes one token streams on ONE rank (the per-GPU roofline numerator)."""
    e, h, L, V = cfg.emb_dim, cfg.hidden_dim, cfg.n_layers, cfg.vocab_size
    att, kvl, hid, vl = e // size, max(1, cfg.n_kv_heads // size) * cfg.head_size, h // size, V // size
    wt = cfg.wtype
    per_layer = (att + 2 * kvl + 2 * hid) * row_bytes(wt, e) + e * row_bytes(wt, att) + e * row_bytes(wt, hid) + 2 * e * 4
    return L * per_layer + vl * row_bytes(wt, e) + e * 4 + row_bytes(wt, e)


def shard_f32(w: Weights, rank: int, size: int) -> dict[str, np.ndarray]:
    """This rank's slices as exact f32 arrays (tests only; the library slices the stored bytes)."""
    c = w.cfg
    s = shard(c, rank, size)
    e, h, L = c.emb_dim, c.hidden_dim, c.n_layers
    d = lambda a, n: decode_matrix(a, c.wtype, n)
    wqkv = d(w.wqkv, e).reshape(L, c.n_qkv, e)
    wo = d(w.wo, e).reshape(L, e, e)
    w13 = d(w.w13, e).reshape(L, 2 * h, e)
    w2 = d(w.w2, h).reshape(L, e, h)
    rows = lambda r: slice(r.start, r.stop)
    return dict(
        wq=wqkv[:, rows(s.q_rows)], wk=wqkv[:, rows(s.k_rows)], wv=wqkv[:, rows(s.v_rows)],
        wo=wo[:, :, rows(s.att_cols)],
        w1=w13[:, rows(s.ffn_rows)], w3=w13[:, h + s.ffn_rows.start:h + s.ffn_rows.stop],
        w2=w2[:, :, rows(s.ffn_rows)],
        wcls=d(w.wcls, e)[rows(s.vocab_rows)],
        emb=d(w.token_embedding_table, e),
    )

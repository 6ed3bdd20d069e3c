"""N > 1 host logic on CPU: the row-parallel split (llm.f90_b200.tp) computed by two gloo ranks
reproduces the single-rank oracle, i.e. heads / FFN rows / vocabulary rows are cut where the
reference's fused wqkv / w13 layout says they are."""
import os
import socket

import numpy as np
import pytest

from conftest import rel_err
from llm.f90_b200 import fixtures as fx, tp
from llm.f90_b200.layout import Config, TINY, SMALL, F32, F16, Q4_0, TINYLLAMA, LLAMA2_7B
from oracle.oracle_np import OracleNP


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_ranges_cover_everything_once():
    # (TINYLLAMA, 8): 4 KV heads on 8 ranks -- every KV head lives on two ranks
    for shape, size in ((TINYLLAMA, 4), (LLAMA2_7B, 8), (TINY, 2), (TINYLLAMA, 8)):
        cfg = Config(**shape)
        sh = [tp.shard(cfg, r, size) for r in range(size)]
        copies = max(1, size // cfg.n_kv_heads)
        for field, total, mult in (("heads", cfg.n_heads, 1), ("kv_heads", cfg.n_kv_heads, copies),
                                   ("att_cols", cfg.emb_dim, 1), ("ffn_rows", cfg.hidden_dim, 1),
                                   ("vocab_rows", cfg.vocab_size, 1)):
            got = sorted(i for s in sh for i in getattr(s, field))
            assert got == sorted(list(range(total)) * mult), field
        for s in sh:  # Wk / Wv rows are the rows of exactly those KV heads
            hs, e, kv = cfg.head_size, cfg.emb_dim, cfg.kv_head_size
            assert list(s.k_rows) == [e + g * hs + i for g in s.kv_heads for i in range(hs)]
            assert list(s.v_rows) == [e + kv + g * hs + i for g in s.kv_heads for i in range(hs)]
        # a rank's query heads read only its own KV heads (quirk Q3: kv head = h // kv_mul)
        kv_mul = cfg.n_heads // cfg.n_kv_heads
        for s in sh:
            assert {h // kv_mul for h in s.heads} == set(s.kv_heads)


def test_shard_rejects_impossible_splits():
    with pytest.raises(ValueError):  # 3 KV heads: neither a multiple nor a divisor of 2 ranks
        tp.shard(Config(emb_dim=768, hidden_dim=2048, n_layers=2, n_heads=24, n_kv_heads=3, vocab_size=512, seq_len=64), 0, 2)
    with pytest.raises(ValueError):
        tp.shard(Config(**TINY), 0, 3)


def test_active_bytes_per_rank_matches_baseline_md():
    assert tp.active_bytes_per_rank(Config(**LLAMA2_7B, wtype=F16), 8) == 1_652_842_496
    c = Config(**TINYLLAMA, wtype=F32)
    assert tp.active_bytes_per_rank(c, 1) == 4_138_057_728


@pytest.mark.parametrize("wt,kvh", [(F32, None), (Q4_0, None), (F32, 1)], ids=["f32", "q4_0", "f32-replicated-kv"])
def test_two_rank_gloo_forward_matches_oracle(tmp_path, wt, kvh):
    import torch.multiprocessing as mp
    from tp_worker import cpu_tp_forward
    shape = dict(SMALL, n_layers=2)
    if kvh:  # fewer KV heads than ranks: both ranks hold a copy of the one KV head
        shape["n_kv_heads"] = kvh
    cfg = Config(**shape, wtype=wt)
    tokens = [2, 17, 400, 33, 9]
    out = str(tmp_path / "tp_logits.npy")
    mp.spawn(cpu_tp_forward, args=(2, free_port(), shape, wt, 5, tokens, out), nprocs=2, join=True)
    got = np.load(out)
    ref = OracleNP(fx.synth_weights(cfg, 5))
    for i, t in enumerate(tokens):
        assert rel_err(got[i], ref.transformer(t, i + 1)) < 1e-9

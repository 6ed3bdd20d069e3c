"""The fused kernel's host-side plan (llmf90_b200_plan), checked without a GPU.

init decides the grid, the shared-memory ring and, for every CTA, the list of bulk copies one token
takes.  A wrong list does not fail loudly on the device -- a misaligned or oversized bulk copy faults,
a row streamed twice or never gives wrong logits only for some shapes -- so the properties are pinned
here for the full-size models of BASELINE.json and every tensor-parallel split:

  * every bulk copy is 16-byte aligned (cp.async.bulk) and non-empty, and the copies of one ring stage
    (the row segments of a tile's chunk) fit one ring slot together;
  * the stages of all CTAs cover each of the five streamed matrices exactly once (llama2.f90:529-531,
    :603-605, :610-612, :618-620, :634-636 read every row once per token);
  * the per-layer stride stays inside the matrix' allocation for the last layer;
  * the ring fits the B200's opt-in shared memory.
"""
import numpy as np
import pytest

from llm.f90_b200 import capi
from llm.f90_b200.layout import F16 as WT_F16, F32 as WT_F32, Q4_0 as WT_Q4_0
from llm.f90_b200.layout import Config, LLAMA2_7B, SMALL, TINYLLAMA

# "odd": a vocabulary that is not a multiple of the 16-row q4_0 tile (the last tile is padded) and FFN
# rows that do not divide by the grid
ODD = {**SMALL, "vocab_size": 1003, "hidden_dim": 1376}
MODELS = {"small": SMALL, "odd": ODD, "tinyllama": TINYLLAMA, "llama2-7b": LLAMA2_7B}
CASES = [(m, wt, tp) for m in MODELS for wt in (WT_F32, WT_F16, WT_Q4_0) for tp in (1, 2, 4, 8)]


def _splits(dims, wt, tp):
    colmul = {WT_F32: 4, WT_F16: 8, WT_Q4_0: 32}[wt]
    kv_ok = dims["n_kv_heads"] % tp == 0 or tp % dims["n_kv_heads"] == 0  # split, or replicated on tp / KVH ranks
    return (dims["n_heads"] % tp == 0 and kv_ok and dims["hidden_dim"] % (tp * colmul) == 0
            and (dims["emb_dim"] // tp) % colmul == 0 and dims["vocab_size"] % tp == 0)


def _region(src):
    return (np.asarray(src, np.uint64) >> np.uint64(40)).astype(np.int64) - 1


@pytest.mark.parametrize("model,wt,tp", CASES, ids=[f"{m}-{wt}-tp{tp}" for m, wt, tp in CASES])
def test_plan_covers_every_matrix_once(model, wt, tp):
    dims = MODELS[model]
    if not _splits(dims, wt, tp):
        pytest.skip("this model does not split that many ways")
    cfg = Config(**dims, wtype=wt)
    for rank in sorted({0, tp - 1}):
        info, st = capi.plan(cfg, tp_rank=rank, tp_size=tp)
        assert info["grid"] <= capi.B200_SMS and info["threads"] == 416
        assert info["n_slots"] >= 3 and info["slot_bytes"] % 128 == 0
        assert info["smem_bytes"] <= capi.B200_SMEM_OPTIN
        assert info["n_layers"] == cfg.n_layers
        used = st[st["bytes"] > 0]
        # ---- every bulk copy: aligned, fits a slot
        assert (used["src"] % 16 == 0).all() and (used["bytes"] % 16 == 0).all()
        assert (used["bytes"] <= info["slot_bytes"]).all()
        assert max(info["tile_chunks"]) <= 20 and all(r in (4, 8, 16) for r in info["tile_rows"])
        for cta in (0, info["grid"] // 2, info["grid"] - 1):
            row = st[cta][st[cta]["bytes"] > 0]
            per_stage = np.bincount(row["stage"], weights=row["bytes"])
            assert per_stage.max() <= info["slot_bytes"] and (per_stage > 0).all()
            assert (np.diff(row["stage"].astype(np.int64)) >= 0).all()
        reg = _region(used["src"])
        assert reg.min() >= 0 and reg.max() <= 8
        # ---- the five matrices: the union of all CTAs' stages is the matrix, nothing twice
        for k in range(5):
            m = used[reg == k]
            off = (m["src"] - np.uint64(capi.plan_vbase(k))).astype(np.int64)
            order = np.argsort(off)
            off, nb = off[order], m["bytes"][order].astype(np.int64)
            assert off[0] == 0, f"matrix {k}: first stage starts at {off[0]}"
            assert (off[1:] == off[:-1] + nb[:-1]).all(), f"matrix {k}: gap or overlap between stages"
            assert off[-1] + nb[-1] == info["matrix_bytes"][k], f"matrix {k}: stages end before / after the matrix"
            stride = m["layer_stride16"].astype(np.int64) * 16
            if k < 4:
                assert (stride == info["matrix_bytes"][k]).all()  # layer l = layer 0 + l * one layer's bytes
            else:
                assert (stride == 0).all()                         # the classifier is not per layer
        # ---- per CTA: one embedding row, the three norm vectors, phases flagged in order
        for cta in range(info["grid"]):
            row = st[cta][st[cta]["bytes"] > 0]
            r = _region(row["src"])
            assert r[0] == 5 and row["bytes"][0] == info["emb_row_bytes"]
            for k in (6, 7, 8):
                v = row[r == k]
                assert len(v) == 1 and v["bytes"][0] == info["vector_bytes"] and v["phase_start"][0] == 1
                assert v["layer_stride16"][0] * 16 == (info["vector_bytes"] if k < 8 else 0)
            # the order the kernel consumes: emb, rms_att, QKV, Wo, rms_ffn, W13, W2, rms_final, classifier
            rank_of = {5: 0, 6: 1, 0: 2, 1: 3, 7: 4, 2: 5, 3: 6, 8: 7, 4: 8}
            seq = np.array([rank_of[int(x)] for x in r])
            assert (np.diff(seq) >= 0).all()
            assert (r == 2).any(), "every CTA must own W13 rows (LL hand-over argument, stream.cu)"


def test_plan_rejects_what_init_rejects():
    bad = Config(**{**TINYLLAMA, "n_heads": 24}, wtype=WT_F32)  # head size 85.33
    with pytest.raises(capi.EngineError):
        capi.plan(bad)
    with pytest.raises(capi.EngineError):
        capi.plan(Config(**TINYLLAMA, wtype=WT_F32), tp_rank=0, tp_size=3)
    info, _ = capi.plan(Config(**TINYLLAMA, wtype=WT_F32), tp_rank=5, tp_size=8)  # 4 KV heads, each on 2 ranks
    assert info["rows"][0] == (2048 // 8) + 2 * 64  # 4 query heads + one (replicated) KV head: K and V rows


def test_plan_balance_is_within_one_unit():
    """Rows go to CTAs in units (row pairs; 16 rows for tiled q4_0): no CTA has more than one unit more
    than another in any phase."""
    for wt in (WT_F32, WT_Q4_0):
        info, st = capi.plan(Config(**LLAMA2_7B, wtype=wt))
        reg = np.where(st["bytes"] > 0, _region(st["src"]), -1)
        for k in range(5):
            per_cta = np.where(reg == k, st["bytes"], 0).sum(axis=1).astype(np.int64)
            rows = info["rows"][k]
            unit = 16 if wt == WT_Q4_0 else 2
            per_unit = info["matrix_bytes"][k] / (((rows + unit - 1) // unit))
            assert per_cta.max() - per_cta.min() <= per_unit + 1e-6


# ------------------------------------------------------------------ the batched prompt pass (prefill.cu)
@pytest.mark.parametrize("wt", [WT_F32, WT_F16, WT_Q4_0], ids=["f32", "f16", "q4_0"])
@pytest.mark.parametrize("shape", ["small", "odd", "tinyllama", "llama2-7b"])
@pytest.mark.parametrize("n_pos", [1, 16, 37, 128])
def test_prefill_plan_fits_a_b200_sm_and_covers_every_chunk(shape, wt, n_pos):
    """llmf90_b200_prefill_plan (no device): every GEMM's K splits cover the contraction chunks exactly once and none is
    empty, the accumulator fits tensor memory, the pipeline the opt-in shared memory, the operand-order copy has the
    bytes of its f16 planes, and the grid reaches the SMs wherever the matrix has the work for it."""
    from llm.f90_b200.layout import TINYLLAMA, LLAMA2_7B, SMALL
    dims = {"small": SMALL, "tinyllama": TINYLLAMA, "llama2-7b": LLAMA2_7B,
            "odd": dict(emb_dim=160, hidden_dim=416, n_layers=2, n_heads=5, n_kv_heads=5, vocab_size=300, seq_len=64)}[shape]
    cfg = Config(**dims, wtype=wt)
    gemms = capi.prefill_plan(cfg, n_pos)
    e, h, kv = cfg.emb_dim, cfg.hidden_dim, cfg.kv_head_size
    assert [(g["rows"], g["cols"]) for g in gemms] == [(e + 2 * kv, e), (e, e), (2 * h, e), (e, h)]
    for g in gemms:
        assert g["planes"] == (1 if wt == WT_F16 else 2)
        assert g["m_tiles"] * 128 >= g["rows"] > (g["m_tiles"] - 1) * 128
        assert g["k_chunks"] * 64 >= g["cols"] > (g["k_chunks"] - 1) * 64
        cps, ns = g["chunks_per_split"], g["n_splits"]
        assert 1 <= ns <= 16 and (ns - 1) * cps < g["k_chunks"] <= ns * cps        # exact cover, last split non-empty
        assert g["ppad"] % 16 == 0 and n_pos <= g["ppad"] < n_pos + 16
        assert g["tmem_cols"] in (32, 64, 128) and g["tmem_cols"] >= g["ppad"]       # of 512 columns per SM
        assert g["stage_bytes"] == g["planes"] * 16384 + g["ppad"] * 256
        assert 2 <= g["stages"] <= 8 and g["smem_bytes"] <= capi.B200_SMEM_OPTIN
        assert g["weight_bytes"] == g["m_tiles"] * g["k_chunks"] * g["planes"] * 16384
        assert g["partial_bytes"] == ns * g["ppad"] * g["rows"] * 4
        ctas = g["m_tiles"] * ns
        # small matrices are split along K until they cover at least half of the 148 SMs (or cannot be split further),
        # never beyond two waves
        assert ctas >= 0.5 * 148 or cps == -(-g["k_chunks"] // min(16, g["k_chunks"]))
        assert ctas <= max(2 * 148, g["m_tiles"])


def test_prefill_plan_cost_of_the_flag():
    """The second copy of the layer matrices: TinyLlama f32 as hi + lo f16 planes = its f32 layer bytes, f16 one plane."""
    from llm.f90_b200.layout import TINYLLAMA
    for wt, planes in ((WT_F32, 2), (WT_F16, 1), (WT_Q4_0, 2)):
        cfg = Config(**TINYLLAMA, wtype=wt)
        total = cfg.n_layers * sum(g["weight_bytes"] for g in capi.prefill_plan(cfg, 16))
        e, h, kv = cfg.emb_dim, cfg.hidden_dim, cfg.kv_head_size
        elems = cfg.n_layers * ((e + 2 * kv) * e + e * e + 2 * h * e + e * h)
        assert total == elems * 2 * planes  # TinyLlama's dimensions are multiples of 128 / 64: no padding
    with pytest.raises(capi.EngineError):
        capi.prefill_plan(Config(**TINYLLAMA), 129)

"""Synthetic models and GGUF files.

There are no real checkpoints and no network, so every test and benchmark runs on
seeded synthetic weights with the real architectures' shapes.  The GGUF files written
here only use what the reference loader accepts (read_ggml.f90:663-685: KV value types
4,5,6,8,9; tensor types 0 and 1 -- plus 2 for the q4_0 extension), so an f32 file from
this module is loadable by the unmodified reference.
"""
from __future__ import annotations

import struct
from typing import Iterable

import numpy as np

from .layout import (Config, Weights, F32, F16, Q4_0, Q6_K, GGML_TYPE, QK4_0, QK_K, Q6_K_BLOCK_BYTES, row_bytes)

GGUF_MAGIC = 1179993927  # read_ggml.f90:122
GGUF_VERSION = 3
ALIGNMENT = 32           # read_ggml.f90:104

KV_U32, KV_I32, KV_F32, KV_STR, KV_ARR = 4, 5, 6, 8, 9


# --------------------------------------------------------------------------- quantisation
def quantize_q4_0(w: np.ndarray) -> np.ndarray:
    """f32 [rows, n] -> uint8 [rows, n/32*18] ggml q4_0 blocks (public ggml algorithm:
    d = (signed value of largest magnitude) / -8, q = min(15, trunc(x/d + 8.5)))."""
    rows, n = w.shape
    assert n % QK4_0 == 0
    blk = np.ascontiguousarray(w, dtype=np.float32).reshape(rows, n // QK4_0, QK4_0)
    idx = np.argmax(np.abs(blk), axis=-1)
    mx = np.take_along_axis(blk, idx[..., None], axis=-1)[..., 0]
    d = mx / -8.0
    with np.errstate(divide="ignore"):
        inv = np.where(d != 0, 1.0 / d, 0.0).astype(np.float32)
    q = np.minimum(15, (blk * inv[..., None] + 8.5).astype(np.int32)).astype(np.uint8)
    packed = (q[..., :16] | (q[..., 16:] << 4)).astype(np.uint8)
    out = np.empty((rows, n // QK4_0, 18), dtype=np.uint8)
    out[..., 0:2] = d.astype(np.float16).view(np.uint8).reshape(rows, n // QK4_0, 2)
    out[..., 2:] = packed
    return out.reshape(rows, n // QK4_0 * 18)


def dequantize_q4_0(b: np.ndarray, n: int) -> np.ndarray:
    """uint8 [rows, n/32*18] -> f32 [rows, n], exact (value = d*(q-8))."""
    rows = b.shape[0]
    blk = b.reshape(rows, n // QK4_0, 18)
    d = blk[..., 0:2].copy().view(np.float16).astype(np.float32)  # [rows, nb, 1]
    qs = blk[..., 2:]
    lo = (qs & 0x0F).astype(np.int32) - 8
    hi = (qs >> 4).astype(np.int32) - 8
    vals = np.concatenate([lo, hi], axis=-1).astype(np.float32) * d
    return vals.reshape(rows, n)


def quantize_q6_k(w: np.ndarray) -> np.ndarray:
    """f32 [rows, n] -> uint8 [rows, n/256*210] ggml Q6_K blocks (ql[128] | qh[64] | int8 scales[16] | f16 d).
    A plain round-to-nearest quantiser (ggml searches the sub-block scales; any valid encoding serves a
    fixture): sub-block scale = max|x| / 31, d = max|scale| / 127, q = round(x / (d * sc)) in [-32, 31]."""
    rows, n = w.shape
    assert n % QK_K == 0
    nb = n // QK_K
    x = np.ascontiguousarray(w, np.float32).reshape(rows, nb, 16, 16)
    s = np.abs(x).max(-1) / 31.0                              # [rows, nb, 16]
    d = (s.max(-1) / 127.0).astype(np.float16)                # [rows, nb]
    df = d.astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        sc = np.where(df[..., None] > 0, np.rint(s / df[..., None]), 0).clip(-128, 127).astype(np.int8)
        eff = df[..., None] * sc.astype(np.float32)
        q = np.where(eff[..., None] != 0, np.rint(x / eff[..., None]), 0).clip(-32, 31).astype(np.int32) + 32
    q = q.reshape(rows, nb, 2, 4, 32)                         # [half][group of 32: +0, +32, +64, +96][l]
    ql = np.empty((rows, nb, 2, 64), np.uint8)
    ql[..., :32] = (q[..., 0, :] & 0xF) | ((q[..., 2, :] & 0xF) << 4)
    ql[..., 32:] = (q[..., 1, :] & 0xF) | ((q[..., 3, :] & 0xF) << 4)
    qh = ((q[..., 0, :] >> 4) | ((q[..., 1, :] >> 4) << 2) | ((q[..., 2, :] >> 4) << 4) | ((q[..., 3, :] >> 4) << 6)).astype(np.uint8)
    out = np.empty((rows, nb, Q6_K_BLOCK_BYTES), np.uint8)
    out[..., 0:128] = ql.reshape(rows, nb, 128)
    out[..., 128:192] = qh.reshape(rows, nb, 64)
    out[..., 192:208] = sc.view(np.uint8)
    out[..., 208:210] = d.view(np.uint8).reshape(rows, nb, 2)
    return out.reshape(rows, nb * Q6_K_BLOCK_BYTES)


def dequantize_q6_k(b: np.ndarray, n: int) -> np.ndarray:
    """uint8 [rows, n/256*210] -> f32 [rows, n], exact (d * scale * (q - 32): at most 24 significant bits)."""
    rows = b.shape[0]
    nb = n // QK_K
    blk = b.reshape(rows, nb, Q6_K_BLOCK_BYTES)
    ql = blk[..., 0:128].reshape(rows, nb, 2, 64).astype(np.int32)
    qh = blk[..., 128:192].reshape(rows, nb, 2, 32).astype(np.int32)
    sc = blk[..., 192:208].copy().view(np.int8).reshape(rows, nb, 2, 8).astype(np.float32)
    d = blk[..., 208:210].copy().view(np.float16).astype(np.float32).reshape(rows, nb, 1, 1, 1)
    q = np.empty((rows, nb, 2, 4, 32), np.int32)
    q[..., 0, :] = (ql[..., :32] & 0xF) | (((qh >> 0) & 3) << 4)
    q[..., 1, :] = (ql[..., 32:] & 0xF) | (((qh >> 2) & 3) << 4)
    q[..., 2, :] = (ql[..., :32] >> 4) | (((qh >> 4) & 3) << 4)
    q[..., 3, :] = (ql[..., 32:] >> 4) | (((qh >> 6) & 3) << 4)
    # element l of group g uses scale index l // 16 + 2 g of its half
    scale = sc.reshape(rows, nb, 2, 4, 2)[..., None].repeat(16, -1).reshape(rows, nb, 2, 4, 32)
    return ((d * scale) * (q - 32).astype(np.float32)).reshape(rows, n).astype(np.float32)


def encode_matrix(w: np.ndarray, wtype: int) -> np.ndarray:
    """f32 [rows, n] -> storage array for ``wtype`` (f32 array, f16 array, or q4_0 bytes)."""
    if wtype == F32:
        return np.ascontiguousarray(w, dtype=np.float32)
    if wtype == F16:
        return np.ascontiguousarray(w.astype(np.float16))
    if wtype == Q6_K:
        return quantize_q6_k(w)
    return quantize_q4_0(w)


def decode_matrix(a: np.ndarray, wtype: int, n: int) -> np.ndarray:
    """Storage array -> exact f32 values [rows, n]."""
    if wtype == F32:
        return a.reshape(-1, n).astype(np.float32)
    if wtype == F16:
        return a.reshape(-1, n).astype(np.float32)
    if wtype == Q6_K:
        return dequantize_q6_k(a.reshape(-1, row_bytes(Q6_K, n)), n)
    return dequantize_q4_0(a.reshape(-1, row_bytes(Q4_0, n)), n)


# --------------------------------------------------------------------------- weights
def _normal(rng: np.random.Generator, shape, std: float) -> np.ndarray:
    a = rng.standard_normal(shape, dtype=np.float32)
    a *= np.float32(std)
    return a


def synth_tensors(cfg: Config, seed: int = 0) -> dict[str, np.ndarray]:
    """Per-tensor f32 arrays keyed by their GGUF names (read_ggml.f90:238-410).
    Matrices ~ N(0, 1/fan_in), norm weights = 1 + N(0, 0.02^2), embeddings ~ N(0, 1)."""
    cfg.validate()
    rng = np.random.default_rng(seed)
    e, h, V, kv = cfg.emb_dim, cfg.hidden_dim, cfg.vocab_size, cfg.kv_head_size
    t: dict[str, np.ndarray] = {}
    t["token_embd.weight"] = _normal(rng, (V, e), 1.0)
    for l in range(cfg.n_layers):
        p = f"blk.{l}."
        t[p + "attn_norm.weight"] = 1.0 + _normal(rng, (e,), 0.02)
        t[p + "attn_q.weight"] = _normal(rng, (e, e), e ** -0.5)
        t[p + "attn_k.weight"] = _normal(rng, (kv, e), e ** -0.5)
        t[p + "attn_v.weight"] = _normal(rng, (kv, e), e ** -0.5)
        t[p + "attn_output.weight"] = _normal(rng, (e, e), e ** -0.5)
        t[p + "ffn_norm.weight"] = 1.0 + _normal(rng, (e,), 0.02)
        t[p + "ffn_gate.weight"] = _normal(rng, (h, e), e ** -0.5)
        t[p + "ffn_down.weight"] = _normal(rng, (e, h), h ** -0.5)
        t[p + "ffn_up.weight"] = _normal(rng, (h, e), e ** -0.5)
    t["output_norm.weight"] = 1.0 + _normal(rng, (e,), 0.02)
    t["output.weight"] = _normal(rng, (V, e), e ** -0.5)
    return t


def fuse_tensors(cfg: Config, t: dict[str, np.ndarray], cls_wtype: int | None = None) -> Weights:
    """GGUF-named f32 tensors -> ``Weights`` in the fused weight_module layout, stored as
    ``cfg.wtype`` (the same mapping load_ggml performs, read_ggml.f90:238-410); ``cls_wtype`` = Q6_K stores
    output.weight the way llama.cpp's q4_0 files do."""
    L, wt = cfg.n_layers, cfg.wtype
    enc = lambda a: encode_matrix(a, wt)

    def per_layer(names: Iterable[str]) -> np.ndarray:
        return np.stack([np.concatenate([enc(t[f"blk.{l}.{n}.weight"]) for n in names], axis=0)
                         for l in range(L)])

    return Weights(
        cfg,
        cls_wtype=cls_wtype,
        token_embedding_table=enc(t["token_embd.weight"]),
        rms_att_weight=np.stack([t[f"blk.{l}.attn_norm.weight"] for l in range(L)]),
        wqkv=per_layer(["attn_q", "attn_k", "attn_v"]),
        wo=per_layer(["attn_output"]),
        rms_ffn_weight=np.stack([t[f"blk.{l}.ffn_norm.weight"] for l in range(L)]),
        w13=per_layer(["ffn_gate", "ffn_up"]),
        w2=per_layer(["ffn_down"]),
        rms_final_weight=t["output_norm.weight"],
        wcls=encode_matrix(t["output.weight"], wt if cls_wtype is None else cls_wtype),
    )


def write_ak(path: str, cfg: Config, t: dict[str, np.ndarray]) -> None:
    """The reference's legacy ``--ak`` packed f32 model file (llama2.f90:158-294): 7 x int32 header,
    then the tensors in the llama2.c order without RoPE tables (all Wq, all Wk, all Wv, all Wo,
    ..., W1, W2, W3 grouped by kind over the layers)."""
    L = cfg.n_layers
    with open(path, "wb") as f:
        f.write(np.array([cfg.emb_dim, cfg.hidden_dim, L, cfg.n_heads, cfg.n_kv_heads, cfg.vocab_size, cfg.seq_len],
                         np.int32).tobytes())
        put = lambda a: f.write(np.ascontiguousarray(a, np.float32).tobytes())
        put(t["token_embd.weight"])
        for l in range(L):
            put(t[f"blk.{l}.attn_norm.weight"])
        for n in ("attn_q", "attn_k", "attn_v", "attn_output"):
            for l in range(L):
                put(t[f"blk.{l}.{n}.weight"])
        for l in range(L):
            put(t[f"blk.{l}.ffn_norm.weight"])
        for n in ("ffn_gate", "ffn_down", "ffn_up"):
            for l in range(L):
                put(t[f"blk.{l}.{n}.weight"])
        put(t["output_norm.weight"])
        put(t["output.weight"])


def write_tokenizer_bin(path: str, tokens: list[bytes], scores) -> None:
    """The reference's ``-s tokenizer.bin`` (llama2.f90:321-356): int32 max_len, then per token
    f32 score, int32 length, bytes."""
    import struct
    with open(path, "wb") as f:
        f.write(struct.pack("<i", max(len(s) for s in tokens)))
        for s, sc in zip(tokens, scores):
            f.write(struct.pack("<fi", float(sc), len(s)) + s)


def synth_weights(cfg: Config, seed: int = 0) -> Weights:
    return fuse_tensors(cfg, synth_tensors(cfg, seed))


def synth_weights_fast(cfg: Config, seed: int = 0) -> Weights:
    """Full-size synthetic weights directly in the fused layout, generated tensor by tensor
    with a cheap generator (used by bench.py / full-size tests where building per-tensor
    f32 copies first would need several times the model size in host RAM)."""
    cfg.validate()
    rng = np.random.default_rng(seed)
    e, h, L, V, wt = cfg.emb_dim, cfg.hidden_dim, cfg.n_layers, cfg.vocab_size, cfg.wtype

    def mat(rows: int, n: int, std: float) -> np.ndarray:
        out = np.empty((rows, row_bytes(wt, n)), dtype=np.uint8)
        step = max(1, (1 << 24) // n)
        for r0 in range(0, rows, step):
            r1 = min(rows, r0 + step)
            blk = _normal(rng, (r1 - r0, n), std)
            out[r0:r1] = encode_matrix(blk, wt).reshape(r1 - r0, -1).view(np.uint8)
        if wt == F32:
            return out.view(np.float32)
        if wt == F16:
            return out.view(np.float16)
        return out

    def stack(rows: int, n: int, std: float) -> np.ndarray:
        m = mat(L * rows, n, std)
        return m.reshape(L, rows, -1)

    return Weights(
        cfg,
        token_embedding_table=mat(V, e, 1.0),
        rms_att_weight=1.0 + _normal(rng, (L, e), 0.02),
        wqkv=stack(cfg.n_qkv, e, e ** -0.5),
        wo=stack(e, e, e ** -0.5),
        rms_ffn_weight=1.0 + _normal(rng, (L, e), 0.02),
        w13=stack(2 * h, e, e ** -0.5),
        w2=stack(e, h, h ** -0.5),
        rms_final_weight=1.0 + _normal(rng, (e,), 0.02),
        wcls=mat(V, e, e ** -0.5),
    )


def synth_weights_tiled(cfg: Config, seed: int = 0, pool_rows: int = 2048) -> Weights:
    """Full-size synthetic weights for benchmarking, built at memcpy speed: one seeded pool of
    `pool_rows` encoded rows per row length, tiled into every tensor at a different row offset.
    The values are as random as synth_weights' as far as the memory system is concerned (nothing
    is cached across tensors: every byte lives at its own address) but a 13 GB model is ready in
    seconds instead of minutes.  Deterministic: every tensor-parallel rank builds the same model."""
    cfg.validate()
    rng = np.random.default_rng(seed)
    e, h, L, V, wt = cfg.emb_dim, cfg.hidden_dim, cfg.n_layers, cfg.vocab_size, cfg.wtype
    pools: dict[int, np.ndarray] = {}
    cursor = [0]

    def pool(n: int) -> np.ndarray:
        if n not in pools:
            blk = _normal(rng, (pool_rows, n), n ** -0.5)
            pools[n] = encode_matrix(blk, wt).reshape(pool_rows, -1).view(np.uint8)
        return pools[n]

    def mat(rows: int, n: int, scale_rows: bool = False) -> np.ndarray:
        p = pool(n)
        out = np.empty((rows, p.shape[1]), dtype=np.uint8)
        r = 0
        while r < rows:
            start = cursor[0] % pool_rows
            take = min(rows - r, pool_rows - start)
            out[r:r + take] = p[start:start + take]
            r += take
            cursor[0] += take + 7
        if wt == F32:
            return out.view(np.float32)
        if wt == F16:
            return out.view(np.float16)
        return out

    def stack(rows: int, n: int) -> np.ndarray:
        return mat(L * rows, n).reshape(L, rows, -1)

    return Weights(
        cfg,
        token_embedding_table=mat(V, e),
        rms_att_weight=1.0 + _normal(rng, (L, e), 0.02),
        wqkv=stack(cfg.n_qkv, e),
        wo=stack(e, e),
        rms_ffn_weight=1.0 + _normal(rng, (L, e), 0.02),
        w13=stack(2 * h, e),
        w2=stack(e, h),
        rms_final_weight=1.0 + _normal(rng, (e,), 0.02),
        wcls=mat(V, e),
    )


# --------------------------------------------------------------------------- vocabulary
def synth_vocab(vocab_size: int) -> tuple[list[bytes], np.ndarray]:
    """A llama-style vocabulary: <unk>, <s>, </s>, the 95 printable ASCII characters with
    the space spelled as U+2581 (the loader rewrites a leading U+2581 to ' ',
    read_ggml.f90:483-503), then deterministic multi-character merges, some with a leading
    U+2581.  scores = -index, so earlier merges win (llama2.f90:691-703)."""
    assert vocab_size >= 3 + 95
    sp = "▁".encode("utf-8")
    toks: list[bytes] = [b"<unk>", b"<s>", b"</s>", sp]
    toks += [bytes([c]) for c in range(33, 127)]
    seen = set(toks)
    letters = "etaoinshrdlucmfwypvbgkqjxz"

    def gen():
        n = 2
        while True:
            idx = [0] * n
            while True:
                w = "".join(letters[i] for i in idx).encode()
                yield w
                yield sp + w
                k = n - 1
                while k >= 0:
                    idx[k] += 1
                    if idx[k] < len(letters):
                        break
                    idx[k] = 0
                    k -= 1
                if k < 0:
                    break
            n += 1

    g = gen()
    while len(toks) < vocab_size:
        w = next(g)
        if w not in seen and len(w) <= 48:
            seen.add(w)
            toks.append(w)
    scores = -np.arange(vocab_size, dtype=np.float32)
    return toks, scores


# --------------------------------------------------------------------------- GGUF writer
def _s(b: bytes) -> bytes:
    return struct.pack("<Q", len(b)) + b


def _kv(key: str, vtype: int, payload: bytes) -> bytes:
    return _s(key.encode()) + struct.pack("<I", vtype) + payload


def write_gguf(path: str, cfg: Config, tensors: dict[str, np.ndarray],
               vocab: list[bytes] | None = None, scores: np.ndarray | None = None,
               alignment: int = ALIGNMENT, name: str = "synthetic", cls_wtype: int | None = None) -> None:
    """Write a GGUF v3 file.  ``tensors`` are f32 arrays by GGUF name; 2-D ones are stored as
    ``cfg.wtype``, 1-D ones as f32.  Framing per read_ggml.f90:112-196,600-718."""
    if vocab is None:
        vocab, scores = synth_vocab(cfg.vocab_size)
    assert len(vocab) == cfg.vocab_size
    kvs = [
        _kv("general.architecture", KV_STR, _s(b"llama")),
        _kv("general.name", KV_STR, _s(name.encode())),
        _kv("llama.context_length", KV_U32, struct.pack("<I", cfg.seq_len)),
        _kv("llama.embedding_length", KV_U32, struct.pack("<I", cfg.emb_dim)),
        _kv("llama.block_count", KV_U32, struct.pack("<I", cfg.n_layers)),
        _kv("llama.feed_forward_length", KV_U32, struct.pack("<I", cfg.hidden_dim)),
        _kv("llama.rope.dimension_count", KV_U32, struct.pack("<I", cfg.head_size)),
        _kv("llama.attention.head_count", KV_U32, struct.pack("<I", cfg.n_heads)),
        _kv("llama.attention.head_count_kv", KV_U32, struct.pack("<I", cfg.n_kv_heads)),
        _kv("llama.attention.layer_norm_rms_epsilon", KV_F32, struct.pack("<f", 1e-5)),
        _kv("tokenizer.ggml.model", KV_STR, _s(b"llama")),
        _kv("tokenizer.ggml.tokens", KV_ARR,
            struct.pack("<IQ", KV_STR, len(vocab)) + b"".join(_s(t) for t in vocab)),
        _kv("tokenizer.ggml.scores", KV_ARR,
            struct.pack("<IQ", KV_F32, len(vocab)) + np.asarray(scores, "<f4").tobytes()),
        _kv("tokenizer.ggml.token_type", KV_ARR,
            struct.pack("<IQ", KV_I32, len(vocab)) + np.ones(len(vocab), "<i4").tobytes()),
        _kv("tokenizer.ggml.bos_token_id", KV_U32, struct.pack("<I", 1)),
        _kv("tokenizer.ggml.eos_token_id", KV_U32, struct.pack("<I", 2)),
    ]
    if alignment != ALIGNMENT:
        kvs.insert(2, _kv("general.alignment", KV_U32, struct.pack("<I", alignment)))

    infos, blobs, off = [], [], 0
    for tname, a in tensors.items():
        if a.ndim == 2:
            wt = cls_wtype if (tname == "output.weight" and cls_wtype is not None) else cfg.wtype
            ttype, data = GGML_TYPE[wt], encode_matrix(a, wt)
            dims = (a.shape[1], a.shape[0])  # innermost (contraction) dimension first
        else:
            ttype, data = GGML_TYPE[F32], np.ascontiguousarray(a, dtype=np.float32)
            dims = (a.shape[0],)
        raw = data.tobytes()
        infos.append(_s(tname.encode()) + struct.pack("<I", len(dims))
                     + b"".join(struct.pack("<Q", d) for d in dims)
                     + struct.pack("<IQ", ttype, off))
        pad = (-len(raw)) % alignment
        blobs.append(raw + b"\0" * pad)
        off += len(raw) + pad

    head = struct.pack("<IIQQ", GGUF_MAGIC, GGUF_VERSION, len(infos), len(kvs))
    meta = head + b"".join(kvs) + b"".join(infos)
    meta += b"\0" * ((-len(meta)) % alignment)
    with open(path, "wb") as f:
        f.write(meta)
        for b in blobs:
            f.write(b)


def write_synth_gguf(path: str, cfg: Config, seed: int = 0, **kw) -> Weights:
    """Write a synthetic model to ``path`` and return the fused ``Weights`` it must load to."""
    t = synth_tensors(cfg, seed)
    write_gguf(path, cfg, t, **kw)
    return fuse_tensors(cfg, t, cls_wtype=kw.get("cls_wtype"))

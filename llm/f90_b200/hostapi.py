"""ctypes binding of libllmf90_host.so -- the C++ mirror of the reference's host program
(GGUF loader, tokenizer, sampler; include/llmf90_host.h).  Used by the tests, the benchmark
(prompt tokenisation) and anyone who wants the reference's `load_ggml` result as numpy arrays."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .layout import Config, Weights, F32, F16

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libllmf90_host.so")
LLM_BIN = os.path.join(HERE, "bin", "llm")


class HostError(RuntimeError):
    pass


class CHostConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("emb_dim", "hidden_dim", "n_layers", "n_heads", "n_kv_heads",
                                         "vocab_size", "seq_len", "wtype", "cls_wtype")]


_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HostError(f"{LIB_PATH} is missing: run `python -m llm.f90_b200.build`")
        L = C.CDLL(LIB_PATH)
        L.llmf90_host_load.restype = C.c_void_p
        L.llmf90_host_load.argtypes = [C.c_char_p, C.c_int32]
        L.llmf90_host_load_ak.restype = C.c_void_p
        L.llmf90_host_load_ak.argtypes = [C.c_char_p, C.c_int32]
        L.llmf90_host_free.argtypes = [C.c_void_p]
        L.llmf90_host_last_error.restype = C.c_char_p
        L.llmf90_host_get_config.argtypes = [C.c_void_p, C.POINTER(CHostConfig)]
        L.llmf90_host_data_offset.restype = C.c_uint64
        L.llmf90_host_data_offset.argtypes = [C.c_void_p]
        L.llmf90_host_tensor.restype = C.c_void_p
        L.llmf90_host_tensor.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_uint64)]
        L.llmf90_host_vocab.restype = C.c_int32
        L.llmf90_host_vocab.argtypes = [C.c_void_p, C.c_int32, C.c_char_p, C.c_int32, C.POINTER(C.c_float)]
        L.llmf90_host_load_tokenizer.argtypes = [C.c_void_p, C.c_char_p]
        L.llmf90_host_encode.restype = C.c_int32
        L.llmf90_host_encode.argtypes = [C.c_void_p, C.c_char_p, C.c_int32, C.POINTER(C.c_int32), C.c_int32]
        L.llmf90_host_argmax.restype = C.c_int32
        L.llmf90_host_argmax.argtypes = [C.POINTER(C.c_float), C.c_int32]
        L.llmf90_host_sample.restype = C.c_int32
        L.llmf90_host_sample.argtypes = [C.POINTER(C.c_float), C.c_int32, C.c_float, C.c_float]
        _lib = L
    return _lib


class HostModel:
    """What `load_ggml` (read_ggml.f90:53) returns: weights in the fused layout, vocabulary, scores."""

    def __init__(self, path: str, verbose: bool = False, quiet: bool = True, ak: bool = False):
        self.L = load()
        if ak:  # the legacy packed f32 file (llama2.f90:158-294)
            self.h = self.L.llmf90_host_load_ak(path.encode(), 1 if verbose else 0)
        else:
            self.h = self.L.llmf90_host_load(path.encode(), 1 if verbose else (-1 if quiet else 0))
        if not self.h:
            raise HostError(self.L.llmf90_host_last_error().decode(errors="replace"))
        cc = CHostConfig()
        self.L.llmf90_host_get_config(self.h, C.byref(cc))
        self.cfg = Config(**{n: getattr(cc, n) for n, _ in CHostConfig._fields_ if n != "cls_wtype"})
        self.cls_wtype = int(cc.cls_wtype)
        self.data_offset = int(self.L.llmf90_host_data_offset(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.L.llmf90_host_free(self.h)
            self.h = None

    __del__ = close

    def _tensor(self, which: int, dtype) -> np.ndarray:
        n = C.c_uint64(0)
        p = self.L.llmf90_host_tensor(self.h, which, C.byref(n))
        buf = (C.c_uint8 * n.value).from_address(p)
        return np.frombuffer(buf, dtype=np.uint8).view(dtype)  # a view: valid while the model is open

    def weights(self) -> Weights:
        wt = self.cfg.wtype
        dt = np.float32 if wt == F32 else (np.float16 if wt == F16 else np.uint8)
        names = Weights.FIELDS
        arrs = {}
        for i, f in enumerate(names):
            arrs[f] = self._tensor(i, np.float32 if f.startswith("rms") else dt).copy()
        if self.cls_wtype != wt:  # Q6_K classifier: raw 210-byte blocks
            arrs["wcls"] = self._tensor(8, np.uint8).copy()
        return Weights(self.cfg, cls_wtype=self.cls_wtype, **arrs)

    def vocab(self) -> tuple[list[bytes], np.ndarray]:
        toks, scores = [], np.empty(self.cfg.vocab_size, np.float32)
        buf = C.create_string_buffer(256)
        sc = C.c_float(0)
        for i in range(self.cfg.vocab_size):
            n = self.L.llmf90_host_vocab(self.h, i, buf, 256, C.byref(sc))
            toks.append(buf.raw[:n])
            scores[i] = sc.value
        return toks, scores

    def load_tokenizer(self, path: str) -> None:
        if self.L.llmf90_host_load_tokenizer(self.h, path.encode()):
            raise HostError(self.L.llmf90_host_last_error().decode(errors="replace"))

    def encode(self, text: bytes | str) -> list[int]:
        """bpe_encode (llama2.f90:658-724): 1-based token ids."""
        if isinstance(text, str):
            text = text.encode("utf-8")
        out = (C.c_int32 * max(1, len(text)))()
        n = self.L.llmf90_host_encode(self.h, text, len(text), out, max(1, len(text)))
        if n < 0:
            raise HostError(self.L.llmf90_host_last_error().decode(errors="replace"))
        return list(out[:n])


def argmax(logits: np.ndarray) -> int:
    return int(load().llmf90_host_argmax(logits.ctypes.data_as(C.POINTER(C.c_float)), len(logits)))


def sample(logits: np.ndarray, temperature: float, r: float) -> int:
    return int(load().llmf90_host_sample(logits.ctypes.data_as(C.POINTER(C.c_float)), len(logits), temperature, r))


def encode_with_synth_vocab(text: str, vocab_size: int) -> list[int]:
    """Tokenise with the synthetic vocabulary of fixtures.synth_vocab (bench.py's prompt): writes a
    one-layer GGUF carrying that vocabulary, loads it with the host loader, runs bpe_encode."""
    import tempfile

    from . import fixtures as fx
    cfg = Config(emb_dim=32, hidden_dim=32, n_layers=1, n_heads=1, n_kv_heads=1, vocab_size=vocab_size, seq_len=8)
    rng = np.random.default_rng(0)
    t = {"token_embd.weight": np.zeros((vocab_size, 32), np.float32), "output.weight": np.zeros((vocab_size, 32), np.float32),
         "output_norm.weight": np.ones(32, np.float32)}
    for n, shp in (("attn_norm", (32,)), ("attn_q", (32, 32)), ("attn_k", (32, 32)), ("attn_v", (32, 32)),
                   ("attn_output", (32, 32)), ("ffn_norm", (32,)), ("ffn_gate", (32, 32)), ("ffn_down", (32, 32)),
                   ("ffn_up", (32, 32))):
        t[f"blk.0.{n}.weight"] = rng.standard_normal(shp).astype(np.float32)
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "vocab.gguf")
        fx.write_gguf(p, cfg, t)
        m = HostModel(p)
        try:
            return m.encode(text.replace(" ", " "))
        finally:
            m.close()

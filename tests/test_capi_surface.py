"""CPU-side checks of the drop-in boundary: the shared library builds, loads and exports every
symbol include/llmf90_b200.h declares; without a GPU the compute entry points fail loudly
(no CPU fallback); the product package never touches the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "llmf90_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(llmf90_b200_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for s in ("llmf90_b200_init", "llmf90_b200_transformer", "llmf90_b200_times", "llmf90_b200_free",
              "llmf90_b200_matvec", "llmf90_b200_rmsnorm", "llmf90_b200_softmax", "llmf90_b200_rope",
              "llmf90_b200_last_error"):
        assert s in syms


def test_library_exports_every_declared_symbol(built):
    from llm.f90_b200 import capi
    lib = capi.load()
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    assert set(capi.EXPORTS) == set(declared_symbols())


def test_config_struct_matches_header():
    from llm.f90_b200 import capi
    src = open(HEADER).read()
    body = re.search(r"typedef struct llmf90_b200_config \{(.*?)\}", src, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"u?int32_t\s+(\w+);", body)
    assert fields == [n for n, _ in capi.CConfig._fields_]
    assert C.sizeof(capi.CConfig) == 4 * len(fields)


def _cuda():
    import torch
    return torch.cuda.is_available()


@pytest.mark.skipif(_cuda(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(built):
    from llm.f90_b200 import capi, fixtures as fx
    from llm.f90_b200.layout import Config, TINY
    x = np.ones(8, np.float32)
    with pytest.raises(capi.EngineError, match="no CPU fallback|no CUDA"):
        capi.rmsnorm(x, x)
    with pytest.raises(capi.EngineError):
        capi.matvec(np.ones((2, 8), np.float32), 0, 2, 8, x)
    w = fx.synth_weights(Config(**TINY), 0)
    with pytest.raises(capi.EngineError):
        capi.Engine(w)
    lg = np.zeros(512, np.float32)
    assert capi.load().llmf90_b200_transformer(2, 1, capi._fp(lg)) != 0


def test_argument_validation_needs_no_gpu(built):
    from llm.f90_b200 import capi
    L = capi.load()
    assert L.llmf90_b200_init(None, *([None] * 9)) != 0
    assert b"null" in L.llmf90_b200_last_error()
    x = np.ones(8, np.float32)
    assert L.llmf90_b200_softmax(capi._fp(x), 8, 9, capi._fp(x)) != 0  # s > n
    assert L.llmf90_b200_matvec(None, 0, 1, 4, capi._fp(x), capi._fp(x)) != 0


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "llm")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", ".f90")):
                txt = open(os.path.join(dp, fn), errors="ignore").read()
                assert "oracle" not in txt.replace("test oracle", ""), f"{fn} mentions the oracle"

"""llm.f90_b200 -- B200-native decode engine behind llm.f90's ``transformer(token,pos,s,w)``.

Only what the hot path needs lives here: ``csrc/`` (sm_100a CUDA kernels + the C-ABI
library ``libllmf90_b200.so`` + the C++ host mirror of the reference program) and thin
ctypes plumbing used by the tests and the benchmark.  See DESIGN.md.
"""
from .layout import Config, Weights, F32, F16, Q4_0, WTYPE_NAMES, WTYPE_BY_NAME  # noqa: F401

"""Build the native artefacts in-tree (they travel to the GPU box with the repo snapshot).

    libllmf90_b200.so   CUDA kernels for sm_100a + the C ABI (include/llmf90_b200.h)
    libllmf90_host.so   C++ mirror of the reference's host program (loader, tokenizer, sampler)
    llm                 the `./llm -m <gguf>` command-line program

nvcc cross-compiles sm_100a without a GPU, so this runs in the CPU container too.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(os.path.dirname(HERE))

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]
CUDA_SRCS = ["ops.cu", "stream.cu", "prefill.cu", "engine.cu"]
CUDA_HDRS = ["common.cuh", "kernels.cuh", os.path.join(ROOT, "include", "llmf90_b200.h")]
CUDA_LIB = os.path.join(HERE, "libllmf90_b200.so")
HOST_LIB = os.path.join(HERE, "libllmf90_host.so")
HOST_SRCS = ["host/gguf_loader.cpp", "host/ak_loader.cpp", "host/tokenizer.cpp", "host/host_api.cpp"]
HOST_HDRS = ["host/host.hpp", os.path.join(ROOT, "include", "llmf90_host.h")]
LLM_BIN = os.path.join(HERE, "bin", "llm")


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def _abs(names):
    return [n if os.path.isabs(n) else os.path.join(CSRC, n) for n in names]


def _cxx() -> str:
    for c in ("/usr/bin/g++", shutil.which("g++") or ""):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("no g++ found")


# device functions of the decode kernel whose register spills are bounded at build time: (substring of the mangled
# name, bytes of spill loads allowed in the production / in the instrumented instantiation)
HOT_FUNCTIONS = (("producer_loop", 0, 0), ("consumer_main", 256, 512), ("gather_tp_n", 16, 16))


def _check_hot_functions(ptxas_log: str) -> None:
    """The phase loop of the decode kernel (consumer_main: prologue, tiles and epilogue inlined) lives within the
    128 registers that 13 warps per CTA leave a thread; ptxas' interprocedural allocation parks some words per
    thread in local memory, and where they land moves with unrelated edits.  A few dozen bytes on the per-phase
    path are the measured state; a spilled weight register in a mat-vec loop costs a factor of
    two.  Fail the build when the spill volume leaves the known range instead of finding out on the GPU
    (tools/sass_local_lines.py shows where they are)."""
    lines = ptxas_log.splitlines()
    bad = []
    for i, ln in enumerate(lines):
        if "Function properties for" not in ln or i + 1 >= len(lines) or "spill loads" not in lines[i + 1]:
            continue
        for name, lim, lim_prof in HOT_FUNCTIONS:
            if name in ln:
                loads = int(lines[i + 1].split("bytes spill stores,")[1].split("bytes spill loads")[0])
                instrumented = "Lb1E" in ln  # the profiling instantiation (template argument PROF = true)
                if loads > (lim_prof if instrumented else lim):
                    bad.append(ln.split("Function properties for ")[1][:80] + ":" + lines[i + 1].strip())
    if bad:
        raise RuntimeError("hot device functions spill more than the known state:\n  " + "\n  ".join(bad))


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    deps = _abs(CUDA_SRCS) + _abs(CUDA_HDRS)
    if force or _newer(CUDA_LIB, deps):
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        extra = ["-DLLMF90_WATCHDOG"] if os.environ.get("LLMF90_BUILD_WATCHDOG") else []  # debug builds only
        cmd = [nvcc] + NVCC_FLAGS + extra + ["-Xptxas", "-v", "-o", CUDA_LIB] + _abs(CUDA_SRCS)
        r = subprocess.run(cmd, cwd=CSRC, stderr=subprocess.PIPE, text=True)
        if verbose or r.returncode:
            sys.stderr.write(r.stderr)
        if r.returncode:
            raise subprocess.CalledProcessError(r.returncode, cmd)
        _check_hot_functions(r.stderr)
    return CUDA_LIB


def build_host(force: bool = False) -> str:
    srcs = _abs(HOST_SRCS)
    if not all(os.path.exists(s) for s in srcs):
        return ""
    if force or _newer(HOST_LIB, srcs + _abs(HOST_HDRS)):
        cmd = [_cxx(), "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-o", HOST_LIB] + srcs
        subprocess.run(cmd, check=True, cwd=CSRC)
    main = os.path.join(CSRC, "host", "llm_main.cpp")
    if os.path.exists(main) and (force or _newer(LLM_BIN, [main, HOST_LIB, CUDA_LIB])):
        os.makedirs(os.path.dirname(LLM_BIN), exist_ok=True)
        cmd = [_cxx(), "-O2", "-std=c++17", "-Wall", "-o", LLM_BIN, main, "-L" + HERE, "-lllmf90_host",
               "-lllmf90_b200", "-Wl,-rpath,$ORIGIN/..", "-Wl,--allow-shlib-undefined"]
        subprocess.run(cmd, check=True, cwd=CSRC)
    return HOST_LIB


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_cuda(force, verbose)
    build_host(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built:", CUDA_LIB, HOST_LIB if os.path.exists(HOST_LIB) else "(host lib pending)")

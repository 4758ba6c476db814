"""In-tree build of libjsd_b200.so (sm_100a only; nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB_PATH = os.path.join(CSRC, "libjsd_b200.so")
SOURCES = ["jsd_capi.cu"]
HEADERS = ["ptx.cuh", "jsd_dense.cuh", "jsd_fused.cuh", "jsd_heads.cuh", "jsd_rowwise.cuh", "jsd_score.cuh",
           os.path.join("..", "..", "include", "jsd_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libjsd_b200.so")


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > built for d in deps if os.path.exists(d))


TRACE_LIB_PATH = os.path.join(CSRC, "libjsd_b200_trace.so")


def build_library(force: bool = False, verbose: bool = False, trace: bool = False) -> str:
    """Compile csrc/*.cu into csrc/libjsd_b200.so; returns the library path.  trace=True builds the instrumented
    variant libjsd_b200_trace.so (-DJSD_TRACE=1, see clip_lite_b200/trace.py; select it with JSD_LIB=...)."""
    if trace:
        cmd = [_nvcc(), *NVCC_FLAGS, "-DJSD_TRACE=1", "-o", TRACE_LIB_PATH] + SOURCES
        res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
        return TRACE_LIB_PATH
    if not force and not is_stale():
        return LIB_PATH
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", LIB_PATH] + SOURCES
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))

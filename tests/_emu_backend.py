"""TEST-ONLY: route the row-wise entry points of the C ABI to the CPU emulation of the same kernel source
(tests/emu/emu_kernels.cpp exports them under their product names, compiled against include/jsd_b200.h), so that
clip_lite_b200/kernels.py, ops.py and loss.py run UNMODIFIED on CPU tensors: the ctypes marshalling of the long
argument lists, the autograd wiring and the module option are exercised without a GPU.  The tensor-core entry points
are replaced by the oracle-backed stand-ins of tests/_standin_kernels.py.  The product never imports this file."""
import ctypes
import os
import subprocess

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
EMU_DIR = os.path.join(HERE, "emu")
CSRC = os.path.join(ROOT, "clip_lite_b200", "csrc")
LIB = os.path.join(EMU_DIR, "_build", "libjsd_emu.so")
CUDA_INC = os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/include"
EMULATED = ("jsd_ln_normalize_pair", "jsd_ln_normalize_bwd_pair", "jsd_index_fwd_bwd")


def available() -> bool:
    import shutil
    return shutil.which("g++") is not None and os.path.exists(CUDA_INC + "/cuda_runtime.h")


def build() -> str:
    srcs = [os.path.join(EMU_DIR, f) for f in ("emu_kernels.cpp", "cuda_emu.h", "ptx_emu.cuh")] + \
           [os.path.join(CSRC, f) for f in ("jsd_heads.cuh", "jsd_rowwise.cuh")] + \
           [os.path.join(ROOT, "include", "jsd_b200.h"), __file__]
    if not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        tmp = LIB + f".{os.getpid()}.tmp"
        cmd = ["g++", "-std=c++20", "-O1", "-pthread", "-fPIC", "-shared", "-DJSD_HOST_EMU",
               "-fsanitize=alignment", "-fno-sanitize-recover=alignment",   # a misaligned vector access aborts
               "-I" + CUDA_INC, "-I" + EMU_DIR, "-I" + CSRC, srcs[0], "-o", tmp]
        res = subprocess.run(cmd, capture_output=True, text=True)
        assert res.returncode == 0, "g++ failed:\n" + res.stderr[-4000:]
        os.replace(tmp, LIB)
    return LIB


def load() -> ctypes.CDLL:
    return ctypes.CDLL(build())


def install(monkeypatch, bwd_blocks: int = 3):
    """Patch clip_lite_b200 so that CPU tensors flow through kernels.py into the emulated kernels."""
    from clip_lite_b200 import _lib, kernels as K, ops
    from tests import _standin_kernels as S
    emu = load()
    emu.emu_set_bwd_blocks(bwd_blocks)
    for name in EMULATED:
        fn = getattr(emu, name)
        fn.restype, fn.argtypes = _lib.SIGNATURES[name]          # the binding's own declaration of the entry point
    emu.jsd_ln_workspace_bytes.restype, emu.jsd_ln_workspace_bytes.argtypes = _lib.SIGNATURES["jsd_ln_workspace_bytes"]
    emu.jsd_index_workspace_bytes.restype, emu.jsd_index_workspace_bytes.argtypes = \
        _lib.SIGNATURES["jsd_index_workspace_bytes"]
    calls = []

    def call(name, *args):
        assert name in EMULATED, f"{name} is not emulated on the CPU"
        calls.append(name)
        rc = getattr(emu, name)(*args)
        assert rc == 0, f"{name} -> {rc}"

    monkeypatch.setattr(_lib, "call", call)
    monkeypatch.setattr(_lib, "load", lambda: emu)

    def req(t, name, dtype=None, ndim=None):
        assert isinstance(t, torch.Tensor), name
        if dtype is not None and t.dtype != dtype:
            raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
        if ndim is not None and t.dim() != ndim:
            raise ValueError(f"{name} must be {ndim}-dimensional, got shape {tuple(t.shape)}")
        if not t.is_contiguous():
            raise ValueError(f"{name} must be contiguous")
        return t

    def scalar(t, name):
        return t if (t.dtype == torch.float32 and t.numel() == 1) else t.detach().float().reshape(1).contiguous()

    monkeypatch.setattr(K, "_req", req)
    monkeypatch.setattr(K, "_scalar", scalar)
    monkeypatch.setattr(K, "_stream", lambda: None)
    # tensor-core entry points: oracle-backed stand-ins with the kernels' contracts
    monkeypatch.setattr(K, "dense_fwd", S.dense_fwd)
    monkeypatch.setattr(K, "dense_bwd_du", lambda gmat, v, t, gamma=None, stream_k=False: S.dense_bwd_du(gmat, v, t, gamma))
    monkeypatch.setattr(K, "dense_bwd_dv", lambda gmat, u, n, t, gamma=None, stream_k=False: S.dense_bwd_dv(gmat, u, n, t, gamma))
    monkeypatch.setattr(K, "fused_supported", lambda b, d: False)

    def dense_forward(f, g, t, want_grad=True):                  # the default (unfused) module route, for comparison
        u, v, inv_f, inv_g = S.normalize_cast_pair(f, g)
        out4, loss, gmat, gdiag = S.dense_fwd(u, v, t, 0, want_grad)
        return out4, loss, (u, v, inv_f, inv_g, gmat, gdiag)

    def dense_backward(f, g, t, gamma, saved):
        u, v, inv_f, inv_g, gmat, gdiag = saved
        b = f.shape[0]
        du, dv = S.dense_bwd_du(gmat, v, t, gamma), S.dense_bwd_dv(gmat, u, b, t, gamma)
        df, dt = S.normalize_bwd(f, inv_f, du, v, 0, gdiag, t, gamma, b, want_dt=True)
        return df, S.normalize_bwd(g, inv_g, dv, u, 0, gdiag, t, gamma, b), dt

    monkeypatch.setattr(K, "dense_forward", dense_forward)
    monkeypatch.setattr(K, "dense_backward", dense_backward)

    def loose_pair(xf, xg):
        dt = torch.promote_types(xf.dtype, xg.dtype)
        return xf.to(dt).contiguous(), xg.to(dt).contiguous()

    monkeypatch.setattr(ops, "_pair_inputs", loose_pair)
    monkeypatch.setattr(ops, "_common", loose_pair)
    return calls

"""ctypes binding of libjsd_b200.so (the C ABI declared in include/jsd_b200.h).

There is deliberately no fallback: if the library has not been built, or a
call fails, this module raises -- the product path never silently drops to
PyTorch or to the CPU oracle.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

from .build import LIB_PATH as _DEFAULT_LIB_PATH

# development aid: A/B-test an alternative build of the library
LIB_PATH = os.environ.get("JSD_LIB", _DEFAULT_LIB_PATH)

# name -> (restype, argtypes); mirrors include/jsd_b200.h one to one
SIGNATURES = {
    "jsd_abi_version": (c_int, []),
    "jsd_last_error": (c_char_p, []),
    "jsd_sm_count": (c_int, []),
    "jsd_index_workspace_bytes": (c_size_t, [c_int64]),
    "jsd_index_fwd_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p,
                                  c_void_p]),
    "jsd_normalize_cast": (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
    "jsd_dense_workspace_bytes": (c_size_t, []),
    "jsd_dense_fwd": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_int64,
                              c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "jsd_dense_forward": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p]),
    "jsd_dense_backward": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "jsd_dense_fused_supported": (c_int, [c_int64, c_int64]),
    "jsd_dense_fused_splits": (c_int, [c_int64, c_int64]),
    "jsd_dense_fused_fwd_bwd": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_void_p, c_void_p]),
    "jsd_dense_fused_forward": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_void_p]),
    "jsd_dense_fused_backward": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "jsd_normalize_cast_pair": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_void_p]),
    "jsd_dense_backward_image_side": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_int64, c_void_p,
                                              c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                              c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "jsd_streamk_workspace_bytes": (c_size_t, []),
    "jsd_streamk_flag_bytes": (c_size_t, []),
    "jsd_dense_bwd_du": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_void_p]),
    "jsd_dense_bwd_dv": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_void_p]),
    "jsd_normalize_bwd": (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_int64,
                                  c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p]),
    "jsd_ln_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "jsd_ln_normalize_pair": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p, c_float,
                                      c_void_p, c_void_p, c_float, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_void_p]),
    "jsd_ln_normalize_bwd_pair": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                                          c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_float,
                                          c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64,
                                          c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_void_p, c_void_p, c_void_p]),
    "jsd_peer_flag_bytes": (c_size_t, []),
    "jsd_peer_alloc": (c_int, [c_size_t, c_void_p]),
    "jsd_peer_free": (c_int, [c_void_p]),
    "jsd_peer_export": (c_int, [c_void_p, c_void_p]),
    "jsd_peer_open": (c_int, [c_void_p, c_void_p]),
    "jsd_peer_close": (c_int, [c_void_p]),
    "jsd_peer_normalize_push": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                        c_void_p]),
    "jsd_peer_dense_fwd": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p]),
    "jsd_peer_dense_bwd_dv": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                      c_void_p]),
    "jsd_peer_normalize_bwd_text": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                            c_void_p, c_int, c_void_p, c_void_p]),
    "jsd_peer_dense_backward": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "jsd_peer_set_timeout": (c_int, [ctypes.c_double]),
    "jsd_peer_wait_error": (c_int, [c_void_p, c_void_p, c_void_p]),
    "jsd_split_bf16x3": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int, c_int, c_void_p, c_void_p]),
    "jsd_score_ranks": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "jsd_score_argmax": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p]),
    "jsd_trace_enable": (c_int, [c_void_p, c_int, c_void_p]),
    "jsd_gemm_bf16": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_int64, c_int64, c_int64,
                              c_void_p, c_void_p, c_void_p]),
}

MAX_PEERS = 8
PEER_HANDLE_BYTES = 64


class PeerCtx(ctypes.Structure):
    """struct jsd_peer_ctx of include/jsd_b200.h."""
    _fields_ = [("rank", c_int32), ("world", c_int32), ("rows", c_int64), ("dim", c_int64),
                ("v_all", (c_void_p * MAX_PEERS) * 2), ("stage", c_void_p * MAX_PEERS),
                ("flags", c_void_p * MAX_PEERS)]


ABI_VERSION = 12
_lib = None


class JSDLibraryError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load (once) and type the shared library.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise JSDLibraryError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no PyTorch/CPU fallback for the JSD kernels)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here == header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.jsd_abi_version() != ABI_VERSION:
        raise JSDLibraryError(f"libjsd_b200.so ABI {lib.jsd_abi_version()} != expected {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


ERROR_HOOKS = []     # callables run (best effort) before a failing call raises, e.g. kernels._rearm_workspaces


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().jsd_last_error().decode("utf-8", "replace")
        for hook in ERROR_HOOKS:
            hook()
        raise JSDLibraryError(f"{what} failed: {msg}")


# JSD_NVTX=1: every library call is wrapped in an NVTX range named after the entry point, so that the ranges an
# `ncu --nvtx --nvtx-include "jsd_dense_forward/"` capture (or any NVTX-aware profiler) selects are the library's own
# call boundaries (the reference has no NVTX either).  Off by default: two extra host calls per FFI
# crossing are measurable on the latency-bound configs[1] steps when they run eagerly.
NVTX = os.environ.get("JSD_NVTX", "0") == "1"


def call(name: str, *args) -> None:
    if NVTX:
        import torch
        torch.cuda.nvtx.range_push(name)
        try:
            rc = getattr(_lib or load(), name)(*args)
        finally:
            torch.cuda.nvtx.range_pop()
    else:
        rc = getattr(_lib or load(), name)(*args)
    if rc != 0:
        check(rc, name)

"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Imports the *unmodified* reference ``loss.py`` from /root/reference when that tree
is present (the build container only -- it does not exist on the GPU box, so
nothing in the gpu tests, smoke() or bench.py may depend on this module).

Used by tests/golden/make_golden.py to generate the committed golden vectors
and by the CPU tests that compare against the live reference when available.
"""
from __future__ import annotations

import contextlib
import importlib.util
import os
import sys

import torch

REFERENCE_ROOT = os.environ.get("CLIPLITE_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "loss.py"))


_cached = None


def load_reference_loss():
    """Return the reference's ``loss`` module (loss.py, imported as-is)."""
    global _cached
    if _cached is not None:
        return _cached
    if not reference_available():
        raise FileNotFoundError(f"reference loss.py not found under {REFERENCE_ROOT}")
    had_utils = sys.modules.get("utils")
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        sys.modules.pop("utils", None)          # loss.py:9 does `from utils import *`
        spec = importlib.util.spec_from_file_location(
            "cliplite_reference_loss", os.path.join(REFERENCE_ROOT, "loss.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove(REFERENCE_ROOT)
        sys.modules.pop("utils", None)
        if had_utils is not None:
            sys.modules["utils"] = had_utils
    _cached = mod
    return mod


@contextlib.contextmanager
def cuda_calls_neutralised():
    """loss.py:186,257,280 hard-code ``.cuda()``.  On a host without a GPU make
    ``Tensor.cuda`` the identity for the duration of the call; the reference
    files themselves are never edited."""
    if torch.cuda.is_available():
        yield
        return
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = orig


def reference_estimator_module(**kw):
    """The reference JSDInfoMaxLoss with both projection heads swapped for
    nn.Identity so that its own estimator code (loss.py:94-105,204-254) runs on
    arbitrary (B, D) embeddings."""
    ref = load_reference_loss()
    kw.setdefault("image_dim", 8)
    kw.setdefault("text_dim", 8)
    kw.setdefault("type", "dot")
    kw.setdefault("image_prior", False)
    kw.setdefault("text_prior", False)
    m = ref.JSDInfoMaxLoss(**kw)
    m.global_d.img_block = torch.nn.Identity()
    m.global_d.text_block = torch.nn.Identity()
    return m

"""Tensor-level wrappers over the C ABI: one Python function per entry point of
include/jsd_b200.h.  Inputs are CUDA tensors; pointers, sizes and the current
CUDA stream are handed to libjsd_b200.so.  Nothing here computes anything in
PyTorch -- allocation and plumbing only."""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib

_DTYPE_CODE = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}


def _code(t: torch.Tensor) -> int:
    try:
        return _DTYPE_CODE[t.dtype]
    except KeyError:
        raise TypeError(f"unsupported dtype {t.dtype}; expected float32, bfloat16 or float16") from None


def _req(t: torch.Tensor, name: str, dtype=None, ndim: Optional[int] = None) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the JSD kernels have no CPU path")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if ndim is not None and t.dim() != ndim:
        raise ValueError(f"{name} must be {ndim}-dimensional, got shape {tuple(t.shape)}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    return t


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _scalar(t: torch.Tensor, name: str) -> torch.Tensor:
    """Device fp32 scalar (0-dim or 1-element) passed by pointer."""
    if not t.is_cuda:
        raise RuntimeError(f"{name} must live on the GPU")
    if t.dtype == torch.float32 and t.numel() == 1:
        return t                       # data_ptr() is all the kernels need; no torch op on the fast path
    return t.detach().float().reshape(1).contiguous()


class _on_device:
    """`with _on_device(dev)` only when dev is not already current (the common case costs nothing)."""

    def __init__(self, device):
        self.ctx = None
        if device.index is not None and device.index != torch.cuda.current_device():
            self.ctx = torch.cuda.device(device)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *a):
        if self.ctx is not None:
            self.ctx.__exit__(*a)


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def sm_count() -> int:
    return _lib.load().jsd_sm_count()


# ------------------------------------------------------------------ index mode
def index_fwd_bwd(f: torch.Tensor, g: torch.Tensor, t: torch.Tensor, neg_index: Optional[torch.Tensor] = None,
                  inv_ptr: Optional[torch.Tensor] = None, inv_idx: Optional[torch.Tensor] = None,
                  want_grad: bool = True, grad_scale: float = 1.0, gamma: Optional[torch.Tensor] = None):
    """Reference-semantics estimator, fused fwd+bwd.  Returns (out4, loss, dF, dG, grad_scale):
    out4 = [pos, neg, pos+neg, dL/dt] (fp32, device), loss = 0-dim copy of out4[2], dF/dG = grad_scale * gamma
    times the gradients (None with want_grad=False).  ``gamma`` is the upstream gradient as a device scalar
    (default 1); it is applied in fp32 inside the kernel, so reduced-precision gradients are rounded once."""
    _req(f, "F", ndim=2)
    _req(g, "G", dtype=f.dtype, ndim=2)
    if f.shape != g.shape:
        raise ValueError(f"F {tuple(f.shape)} and G {tuple(g.shape)} must have the same shape")
    b, d = f.shape
    if b == 0 or d == 0:
        raise ValueError("empty batch")
    for name, ix, n in (("neg_index", neg_index, b), ("inv_ptr", inv_ptr, b + 1), ("inv_idx", inv_idx, b)):
        if ix is not None:
            _req(ix, name, dtype=torch.int32, ndim=1)
            if ix.numel() != n:
                raise ValueError(f"{name} must have {n} entries")
    tt = _scalar(t, "temperature")
    gs = float(grad_scale)
    gg = None if gamma is None else _scalar(gamma, "gamma")
    lib = _lib.load()
    ws = torch.empty(lib.jsd_index_workspace_bytes(b) // 4, dtype=torch.float32, device=f.device)
    out4 = torch.empty(4, dtype=torch.float32, device=f.device)
    loss = torch.empty((), dtype=torch.float32, device=f.device)
    df = torch.empty_like(f) if want_grad else None
    dg = torch.empty_like(g) if want_grad else None
    with _on_device(f.device):
        _lib.call("jsd_index_fwd_bwd", _ptr(f), _ptr(g), _code(f), b, d, _ptr(neg_index), _ptr(inv_ptr),
                  _ptr(inv_idx), _ptr(tt), _ptr(ws), _ptr(out4), _ptr(loss), _ptr(df), _ptr(dg), gs, _ptr(gg),
                  _stream())
    return out4, loss, df, dg, gs


# ------------------------------------------------------------------ dense mode
# Scratch buffers of the dense path, one of each per DEVICE, allocated zeroed on first use and kept: the kernels
# leave their flag / ticket words zero after every launch.  They are deliberately not per stream -- a CUDA-graph
# capture must not allocate (and memset) 42 MB inside the graph, it re-uses the buffers of the eager warm-up --
# so at most one dense step may be in flight per device at a time.  That rule is ENFORCED, not just documented:
# when a different stream than the last user asks for the buffers, it is first made to wait for that stream
# (two loss modules on two streams, or per-device threads, are serialised instead of sharing tickets and partial
# sums), and a failing library call re-zeroes the ticket words (an aborted launch may have left them armed).
_SK_WORKSPACES = {}
_FWD_WORKSPACES = {}
_LAST_STREAM = {}


def _device_workspace(cache: dict, nbytes: int, device) -> torch.Tensor:
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    ws = cache.get(key)
    capturing = torch.cuda.is_current_stream_capturing()
    if ws is None:
        if capturing:
            raise RuntimeError("run one eager step of the dense path before capturing it into a CUDA graph "
                               "(its scratch buffers are allocated on first use)")
        ws = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        cache[key] = ws
    if not capturing:                  # inside a capture the graph's own edges order the kernels
        cur = torch.cuda.current_stream(ws.device)
        last = _LAST_STREAM.get(key)
        if last is not None and last != cur:
            cur.wait_stream(last)
        _LAST_STREAM[key] = cur
    return ws


def _rearm_workspaces() -> None:
    """After a failed library call: zero the "last block" tickets and hand-off flags of every cached scratch buffer
    (an aborted launch sequence may have left them armed, which would corrupt every later step)."""
    try:
        flag_bytes = _lib.load().jsd_streamk_flag_bytes()
        for ws in _FWD_WORKSPACES.values():
            ws[:16].zero_()
        for ws in _SK_WORKSPACES.values():
            ws[:flag_bytes].zero_()
    except Exception:
        pass                           # the context may be gone; the original error is what matters


_lib.ERROR_HOOKS.append(_rearm_workspaces)


def streamk_workspace(device) -> torch.Tensor:
    """Scratch of the backward contractions (stream-K partial tiles + hand-off flags, 33 MB)."""
    return _device_workspace(_SK_WORKSPACES, _lib.load().jsd_streamk_workspace_bytes(), device)


def dense_workspace(device) -> torch.Tensor:
    """Scratch of the forward / Jacobian kernels: "last block" tickets followed by the per-warp loss partials."""
    return _device_workspace(_FWD_WORKSPACES, _lib.load().jsd_dense_workspace_bytes(), device)


def normalize_cast(x: torch.Tensor):
    """(Xn bf16 [rows, D], inv_norm fp32 [rows])."""
    _req(x, "X", ndim=2)
    rows, d = x.shape
    xn = torch.empty(rows, d, dtype=torch.bfloat16, device=x.device)
    inv = torch.empty(rows, dtype=torch.float32, device=x.device)
    with _on_device(x.device):
        _lib.call("jsd_normalize_cast", _ptr(x), _code(x), rows, d, _ptr(xn), _ptr(inv), _stream())
    return xn, inv


def dense_fwd(u: torch.Tensor, v: torch.Tensor, t: torch.Tensor, row_offset: int = 0, want_grad: bool = True):
    """Row-slab dense forward.  Returns (out4, loss, Gmat or None, gdiag)."""
    _req(u, "U", dtype=torch.bfloat16, ndim=2)
    _req(v, "V", dtype=torch.bfloat16, ndim=2)
    m, d = u.shape
    n, d2 = v.shape
    if d != d2:
        raise ValueError(f"U and V disagree on D: {d} vs {d2}")
    tt = _scalar(t, "temperature")
    out4 = torch.empty(4, dtype=torch.float32, device=u.device)
    loss = torch.empty((), dtype=torch.float32, device=u.device)
    gdiag = torch.empty(m, dtype=torch.float32, device=u.device)
    gmat = None
    ldg = 0
    if want_grad:
        ldg = round_up(n, 64)
        gmat = torch.empty(m, ldg, dtype=torch.bfloat16, device=u.device)
    with _on_device(u.device):
        ws = dense_workspace(u.device)
        _lib.call("jsd_dense_fwd", _ptr(u), _ptr(v), m, n, d, row_offset, _ptr(tt), _ptr(gmat), ldg, _ptr(gdiag),
                  _ptr(ws), _ptr(out4), _ptr(loss), _stream())
    return out4, loss, gmat, gdiag


def dense_forward(f: torch.Tensor, g: torch.Tensor, t: torch.Tensor, want_grad: bool = True):
    """Whole single-GPU forward in ONE library call (normalise+cast x2, tensor-core forward, loss).
    f, g: [B, D] contiguous CUDA tensors of the same dtype.  Returns
    (out4, loss, saved) with saved = (u, v, inv_f, inv_g, gmat, gdiag) for dense_backward."""
    b, d = f.shape
    dev = f.device
    tt = _scalar(t, "temperature")
    u = torch.empty(b, d, dtype=torch.bfloat16, device=dev)
    v = torch.empty(b, d, dtype=torch.bfloat16, device=dev)
    # one fp32 allocation for the small vectors: inv_f | inv_g | gdiag | out4 | loss
    small = torch.empty(3 * b + 8, dtype=torch.float32, device=dev)
    inv_f, inv_g, gdiag = small[:b], small[b:2 * b], small[2 * b:3 * b]
    out4, loss = small[3 * b:3 * b + 4], small[3 * b + 4]
    gmat, ldg = None, 0
    if want_grad:
        ldg = round_up(b, 64)
        gmat = torch.empty(b, ldg, dtype=torch.bfloat16, device=dev)
    with _on_device(dev):
        ws = dense_workspace(dev)
        _lib.call("jsd_dense_forward", f.data_ptr(), g.data_ptr(), _code(f), b, d, tt.data_ptr(), u.data_ptr(),
                  v.data_ptr(), inv_f.data_ptr(), inv_g.data_ptr(), _ptr(gmat), ldg, gdiag.data_ptr(),
                  ws.data_ptr(), out4.data_ptr(), loss.data_ptr(), _stream())
    return out4, loss, (u, v, inv_f, inv_g, gmat, gdiag)


FUSED_AUTO_MAX_BATCH = 4096


def fused_supported(b: int, d: int) -> bool:
    """The single-pass fused kernel (score tile and gradient accumulators resident in TMEM, no B x B buffer in HBM)
    takes D in {64, 128, 192, 256}.  It is bound by the softplus / sigmoid epilogue (two score tiles are recomputed
    for the text side), so by default it is used where it measures faster than the staged path -- B <= 4096 (1.1x there, 1.4-1.5x at
    BASELINE configs[1]'s B = 1024 (profiles/fused_ab_*.log); JSD_FUSED=1 forces it wherever it is
    supported (no O(B^2) memory: B = 65536, D = 128 needs 67 MB of accumulators instead of an 8.6 GB sigma matrix),
    JSD_FUSED=0 disables it."""
    import os
    mode = os.environ.get("JSD_FUSED", "auto")
    if mode == "0" or not _lib.load().jsd_dense_fused_supported(b, d):
        return False
    return mode == "1" or b <= FUSED_AUTO_MAX_BATCH


def dense_fused_forward(f: torch.Tensor, g: torch.Tensor, t: torch.Tensor):
    """Whole forward AND the tensor-core part of the backward in ONE library call (D <= 256).  Returns
    (out4, loss, saved) with saved = (u, v, inv_f, inv_g, acc, gdiag) for dense_fused_backward."""
    b, d = f.shape
    dev = f.device
    tt = _scalar(t, "temperature")
    uv = torch.empty(2, b, d, dtype=torch.bfloat16, device=dev)
    with _on_device(dev):
        ns = _lib.load().jsd_dense_fused_splits(b, d)
    acc = torch.empty(2, ns, b, d, dtype=torch.float32, device=dev)
    small = torch.empty(3 * b + 8, dtype=torch.float32, device=dev)
    inv_f, inv_g, gdiag = small[:b], small[b:2 * b], small[2 * b:3 * b]
    out4, loss = small[3 * b:3 * b + 4], small[3 * b + 4]
    with _on_device(dev):
        ws = dense_workspace(dev)
        _lib.call("jsd_dense_fused_forward", f.data_ptr(), g.data_ptr(), _code(f), b, d, tt.data_ptr(),
                  uv[0].data_ptr(), uv[1].data_ptr(), inv_f.data_ptr(), inv_g.data_ptr(), acc[0].data_ptr(),
                  acc[1].data_ptr(), gdiag.data_ptr(), ws.data_ptr(), out4.data_ptr(), loss.data_ptr(), _stream())
    return out4, loss, (uv[0], uv[1], inv_f, inv_g, acc, gdiag)


def dense_fused_backward(f: torch.Tensor, g: torch.Tensor, t: torch.Tensor, gamma: torch.Tensor, saved):
    """Both normalisation Jacobians (+ dL/dt) on the fused kernel's accumulators: one launch.  Returns (dF, dG, dt)."""
    u, v, inv_f, inv_g, acc, gdiag = saved
    b, d = f.shape
    dev = f.device
    tt = _scalar(t, "temperature")
    gg = _scalar(gamma, "gamma")
    df = torch.empty_like(f)
    dg = torch.empty_like(g)
    small = torch.empty(b + 1, dtype=torch.float32, device=dev)       # row dots | dt
    with _on_device(dev):
        ws = dense_workspace(dev)
        _lib.call("jsd_dense_fused_backward", f.data_ptr(), g.data_ptr(), _code(f), b, d, u.data_ptr(), v.data_ptr(),
                  inv_f.data_ptr(), inv_g.data_ptr(), gdiag.data_ptr(), tt.data_ptr(), gg.data_ptr(),
                  acc[0].data_ptr(), acc[1].data_ptr(), small.data_ptr(), ws.data_ptr(), df.data_ptr(),
                  dg.data_ptr(), small[b:].data_ptr(), _stream())
    return df, dg, small[b]


def dense_backward(f: torch.Tensor, g: torch.Tensor, t: torch.Tensor, gamma: torch.Tensor, saved):
    """Whole single-GPU backward in ONE library call.  Returns (dF, dG, dt)."""
    u, v, inv_f, inv_g, gmat, gdiag = saved
    b, d = f.shape
    dev = f.device
    tt = _scalar(t, "temperature")
    gg = _scalar(gamma, "gamma")
    acc = torch.empty(2, b, d, dtype=torch.float32, device=dev)
    df = torch.empty_like(f)
    dg = torch.empty_like(g)
    small = torch.empty(b + 1, dtype=torch.float32, device=dev)       # row dots | dt
    with _on_device(dev):
        ws = dense_workspace(dev)
        _lib.call("jsd_dense_backward", f.data_ptr(), g.data_ptr(), _code(f), b, d, u.data_ptr(), v.data_ptr(),
                  inv_f.data_ptr(), inv_g.data_ptr(), gmat.data_ptr(), gmat.shape[1], gdiag.data_ptr(),
                  tt.data_ptr(), gg.data_ptr(), acc[0].data_ptr(), acc[1].data_ptr(), small.data_ptr(),
                  ws.data_ptr(), streamk_workspace(dev).data_ptr(), df.data_ptr(), dg.data_ptr(),
                  small[b:].data_ptr(), _stream())
    return df, dg, small[b]


def _dense_bwd(name: str, gmat: torch.Tensor, x: torch.Tensor, m: int, n: int, rows_out: int, t, gamma,
               stream_k: bool = False):
    _req(gmat, "Gmat", dtype=torch.bfloat16, ndim=2)
    _req(x, "operand", dtype=torch.bfloat16, ndim=2)
    if gmat.shape[0] != m or gmat.shape[1] < n:
        raise ValueError(f"Gmat {tuple(gmat.shape)} does not cover the [{m}, {n}] slab")
    d = x.shape[1]
    tt = _scalar(t, "temperature")
    gg = None if gamma is None else _scalar(gamma, "gamma")
    out = torch.empty(rows_out, d, dtype=torch.float32, device=gmat.device)
    with _on_device(gmat.device):
        # stream-K is opt-in: on B200 the ragged second wave costs less than the partial-tile exchange
        # (the chip is power-limited, idle SMs let the busy ones clock higher) -- see DESIGN.md
        ws = streamk_workspace(gmat.device) if stream_k else None
        _lib.call(name, _ptr(gmat), gmat.shape[1], _ptr(x), m, n, d, _ptr(tt), _ptr(gg), _ptr(ws), _ptr(out),
                  _stream())
    return out


def dense_bwd_du(gmat: torch.Tensor, v: torch.Tensor, t: torch.Tensor,
                 gamma: Optional[torch.Tensor] = None, stream_k: bool = False) -> torch.Tensor:
    """dUacc [M, D] fp32 = gamma tau / (M (N-1)) Gmat . V  (v = V [N, D] bf16, read in place)."""
    return _dense_bwd("jsd_dense_bwd_du", gmat, v, gmat.shape[0], v.shape[0], gmat.shape[0], t, gamma, stream_k)


def dense_bwd_dv(gmat: torch.Tensor, u: torch.Tensor, n: int, t: torch.Tensor,
                 gamma: Optional[torch.Tensor] = None, stream_k: bool = False) -> torch.Tensor:
    """dVacc [N, D] fp32 = gamma tau / (M (N-1)) Gmat^T . U  (u = U [M, D] bf16, read in place)."""
    return _dense_bwd("jsd_dense_bwd_dv", gmat, u, u.shape[0], n, n, t, gamma, stream_k)


def normalize_bwd(x: torch.Tensor, inv_norm: torch.Tensor, acc: torch.Tensor, partner: torch.Tensor,
                  partner_offset: int, gdiag: Optional[torch.Tensor], t: torch.Tensor,
                  gamma: Optional[torch.Tensor], m_rows: int, want_dt: bool = False):
    """Positive-pair term + Jacobian of F.normalize; returns dX in x's dtype, or (dX, dt) with
    want_dt: dt = sum_rows <u_row, d_row> = gamma * dL/dt when x are the image rows."""
    _req(x, "X", ndim=2)
    rows, d = x.shape
    _req(inv_norm, "inv_norm", dtype=torch.float32, ndim=1)
    _req(acc, "acc", dtype=torch.float32, ndim=2)
    _req(partner, "partner", dtype=torch.bfloat16, ndim=2)
    if acc.shape != x.shape or partner.shape[1] != d or partner.shape[0] < rows + partner_offset:
        raise ValueError("normalize_bwd: shape mismatch")
    if gdiag is not None:
        _req(gdiag, "gdiag", dtype=torch.float32, ndim=1)
    tt = _scalar(t, "temperature")
    gg = None if gamma is None else _scalar(gamma, "gamma")
    dx = torch.empty_like(x)
    rowdot = torch.empty(rows + 1, dtype=torch.float32, device=x.device) if want_dt else None
    with _on_device(x.device):
        ws = dense_workspace(x.device) if want_dt else None
        _lib.call("jsd_normalize_bwd", _ptr(x), _code(x), rows, d, _ptr(inv_norm), _ptr(acc), _ptr(partner),
                  partner_offset, _ptr(gdiag), _ptr(tt), _ptr(gg), m_rows, _ptr(dx), _ptr(rowdot), _ptr(ws),
                  rowdot[rows:].data_ptr() if want_dt else None, _stream())
    return (dx, rowdot[rows]) if want_dt else dx


def normalize_cast_pair(f: torch.Tensor, g: torch.Tensor):
    """(U, V, inv_f, inv_g) for two [rows, D] tensors of the same dtype in one library call."""
    rows, d = f.shape
    dev = f.device
    uv = torch.empty(2, rows, d, dtype=torch.bfloat16, device=dev)
    inv = torch.empty(2, rows, dtype=torch.float32, device=dev)
    with _on_device(dev):
        _lib.call("jsd_normalize_cast_pair", f.data_ptr(), g.data_ptr(), _code(f), rows, d, uv[0].data_ptr(),
                  uv[1].data_ptr(), inv[0].data_ptr(), inv[1].data_ptr(), _stream())
    return uv[0], uv[1], inv[0], inv[1]


def dense_fused_fwd_bwd(u: torch.Tensor, v: torch.Tensor, t: torch.Tensor):
    """The single-pass fused kernel on ready bf16 unit rows (D in {64, 128, 192, 256}).  Returns
    (out4, loss, gdiag, acc) with acc fp32 [2, S, B, D]: the UNSCALED accumulators of the image / text side in S
    column-split slices (jsd_dense_fused_splits)."""
    _req(u, "U", dtype=torch.bfloat16, ndim=2)
    _req(v, "V", dtype=torch.bfloat16, ndim=2)
    if u.shape != v.shape:
        raise ValueError("the fused kernel needs M == N == B")
    b, d = u.shape
    dev = u.device
    tt = _scalar(t, "temperature")
    lib = _lib.load()
    if not lib.jsd_dense_fused_supported(b, d):
        raise ValueError(f"the fused single-pass kernel does not take B={b}, D={d}")
    with _on_device(dev):
        ns = lib.jsd_dense_fused_splits(b, d)
    acc = torch.empty(2, ns, b, d, dtype=torch.float32, device=dev)
    small = torch.empty(b + 8, dtype=torch.float32, device=dev)
    gdiag, out4, loss = small[:b], small[b:b + 4], small[b + 4]
    with _on_device(dev):
        ws = dense_workspace(dev)
        _lib.call("jsd_dense_fused_fwd_bwd", u.data_ptr(), v.data_ptr(), b, d, tt.data_ptr(), acc[0].data_ptr(),
                  acc[1].data_ptr(), gdiag.data_ptr(), ws.data_ptr(), out4.data_ptr(), loss.data_ptr(), _stream())
    return out4, loss, gdiag, acc


# ------------------------------------------------------------------ projection-head tail (LayerNorm + normalise)
def _ln_param(p: Optional[torch.Tensor], d: int, name: str, device=None) -> Optional[torch.Tensor]:
    if p is None:
        return None
    _req(p, name, ndim=1)
    if p.numel() != d:
        raise ValueError(f"{name} must have {d} entries, got {p.numel()}")
    if device is not None and p.device != device:
        raise ValueError(f"{name} lives on {p.device}, the rows on {device}")
    return p if p.dtype == torch.float32 else p.float()


def ln_normalize_pair(x0: torch.Tensor, x1: Optional[torch.Tensor], ln0, ln1=None, out_bf16: bool = False):
    """u = LN(x) / max(||LN(x)||, 1e-12) for one or two [rows, D] row sets of the same shape and dtype in one launch
    (the image and the text head, reference loss.py:36-38 + :94-95).  ln = (weight or None, bias or None, eps).
    Returns (out0, out1, stats0, stats1): unit rows in fp32 (bf16 with out_bf16) and stats [3, rows] fp32 =
    mean | rstd | 1 / max(||LN(x)||, 1e-12); the second pair is None when x1 is None."""
    _req(x0, "X0", ndim=2)
    rows, d = x0.shape
    if rows == 0 or d == 0:
        raise ValueError("empty batch")
    two = x1 is not None
    if two:
        _req(x1, "X1", dtype=x0.dtype, ndim=2)
        if x1.shape != x0.shape:
            raise ValueError(f"X0 {tuple(x0.shape)} and X1 {tuple(x1.shape)} must have the same shape")
    dev = x0.device
    w0, b0, e0 = (_ln_param(ln0[0], d, "LayerNorm weight", dev), _ln_param(ln0[1], d, "LayerNorm bias", dev),
                  float(ln0[2]))
    w1 = b1 = None
    e1 = 0.0
    if two:
        w1, b1, e1 = (_ln_param(ln1[0], d, "LayerNorm weight", dev), _ln_param(ln1[1], d, "LayerNorm bias", dev),
                      float(ln1[2]))
    odt = torch.bfloat16 if out_bf16 else torch.float32
    out = torch.empty(2 if two else 1, rows, d, dtype=odt, device=dev)
    stats = torch.empty(2 if two else 1, 3, rows, dtype=torch.float32, device=dev)
    with _on_device(dev):
        _lib.call("jsd_ln_normalize_pair", _ptr(x0), _ptr(x1), _code(x0), rows, d, _ptr(w0), _ptr(b0), e0, _ptr(w1),
                  _ptr(b1), e1, int(out_bf16), out[0].data_ptr(), out[1].data_ptr() if two else None,
                  stats[0].data_ptr(), stats[1].data_ptr() if two else None, _stream())
    return (out[0], out[1], stats[0], stats[1]) if two else (out[0], None, stats[0], None)


def ln_normalize_bwd_pair(x0: torch.Tensor, x1: Optional[torch.Tensor], ln0, ln1, stats0: torch.Tensor,
                          stats1: Optional[torch.Tensor], acc0: torch.Tensor, acc1: Optional[torch.Tensor],
                          acc_scale: float = 0.0, partner0: Optional[torch.Tensor] = None,
                          partner1: Optional[torch.Tensor] = None, gdiag: Optional[torch.Tensor] = None,
                          t: Optional[torch.Tensor] = None, gamma: Optional[torch.Tensor] = None,
                          m_rows: Optional[int] = None, want_dt: bool = False):
    """Positive-pair term + Jacobian of F.normalize + LayerNorm backward in one pass (and one small reduction launch).
    acc: fp32 [rows, D] or [S, rows, D] (S slices summed in order); see include/jsd_b200.h for the flavours.
    Returns (dx0, dx1, dw0, db0, dw1, db1, dt): dx in x's dtype, dw / db fp32 [D] (None where the LayerNorm has no
    such parameter), dt = sum_rows <u, dU> of row set 0 (None unless want_dt)."""
    _req(x0, "X0", ndim=2)
    rows, d = x0.shape
    two = x1 is not None
    dev = x0.device

    def _acc(a, name):
        _req(a, name, dtype=torch.float32)
        if a.dim() == 2:
            a = a.unsqueeze(0)
        if a.dim() != 3 or a.shape[1:] != (rows, d):
            raise ValueError(f"{name} must be [rows, D] or [S, rows, D] for rows={rows}, D={d}; got {tuple(a.shape)}")
        return a

    acc0 = _acc(acc0, "acc0")
    ns = acc0.shape[0]
    if two:
        _req(x1, "X1", dtype=x0.dtype, ndim=2)
        acc1 = _acc(acc1, "acc1")
        if x1.shape != x0.shape or acc1.shape != acc0.shape:
            raise ValueError("the two row sets must have the same shape and the same number of accumulator slices")
    for st in (stats0,) + ((stats1,) if two else ()):
        _req(st, "stats", dtype=torch.float32, ndim=2)
        if st.shape != (3, rows):
            raise ValueError(f"stats must be [3, {rows}], got {tuple(st.shape)}")
    if gdiag is not None:
        _req(gdiag, "gdiag", dtype=torch.float32, ndim=1)
        for pr in (partner0,) + ((partner1,) if two else ()):
            _req(pr, "partner", dtype=torch.bfloat16, ndim=2)
            if pr.shape[1] != d or pr.shape[0] < rows:
                raise ValueError("partner rows do not cover the row set")
    w0, b0 = _ln_param(ln0[0], d, "LayerNorm weight", dev), _ln_param(ln0[1], d, "LayerNorm bias", dev)
    w1 = b1 = None
    if two:
        w1, b1 = _ln_param(ln1[0], d, "LayerNorm weight", dev), _ln_param(ln1[1], d, "LayerNorm bias", dev)
    tt = None if t is None else _scalar(t, "temperature")
    gg = None if gamma is None else _scalar(gamma, "gamma")
    dx = [torch.empty_like(x0), torch.empty_like(x1) if two else None]
    dwb = torch.empty(4, d, dtype=torch.float32, device=dev)       # dw0 | db0 | dw1 | db1
    small = torch.empty(rows + 1, dtype=torch.float32, device=dev) if want_dt else None   # row dots | dt
    lib = _lib.load()
    with _on_device(dev):
        ws = torch.empty(lib.jsd_ln_workspace_bytes(rows, d) // 4, dtype=torch.float32, device=dev)
        _lib.call("jsd_ln_normalize_bwd_pair", _ptr(x0), _ptr(x1), _code(x0), rows, d, _ptr(w0), _ptr(b0), _ptr(w1),
                  _ptr(b1), _ptr(stats0), _ptr(stats1) if two else None, acc0.data_ptr(),
                  acc1.data_ptr() if two else None, ns, acc0.stride(0) if ns > 1 else 0, float(acc_scale),
                  _ptr(partner0) if gdiag is not None else None, 0,
                  _ptr(partner1) if (gdiag is not None and two) else None, 0, _ptr(gdiag), _ptr(tt), _ptr(gg),
                  int(m_rows if m_rows is not None else rows), ws.data_ptr(), dx[0].data_ptr(),
                  dx[1].data_ptr() if two else None, dwb[0].data_ptr() if w0 is not None else None,
                  dwb[1].data_ptr() if b0 is not None else None, dwb[2].data_ptr() if w1 is not None else None,
                  dwb[3].data_ptr() if b1 is not None else None, small.data_ptr() if want_dt else None,
                  small[rows:].data_ptr() if want_dt else None, _stream())
    return (dx[0], dx[1], dwb[0] if w0 is not None else None, dwb[1] if b0 is not None else None,
            dwb[2] if w1 is not None else None, dwb[3] if b1 is not None else None, small[rows] if want_dt else None)


def dense_backward_image_side(f, v_all, inv_f, gmat, gdiag, t, gamma, row_offset: int):
    """dU GEMM + positive-pair term + normalisation Jacobian + dL/dt for a row slab, one library call.
    Returns (dF, dt) with dt = gamma * dL_r/dt."""
    m, d = f.shape
    n = v_all.shape[0]
    dev = f.device
    tt = _scalar(t, "temperature")
    gg = _scalar(gamma, "gamma")
    acc = torch.empty(m, d, dtype=torch.float32, device=dev)
    small = torch.empty(m + 1, dtype=torch.float32, device=dev)
    df = torch.empty_like(f)
    with _on_device(dev):
        ws = dense_workspace(dev)
        _lib.call("jsd_dense_backward_image_side", f.data_ptr(), _code(f), m, n, d, row_offset, v_all.data_ptr(),
                  inv_f.data_ptr(), gmat.data_ptr(), gmat.shape[1], gdiag.data_ptr(), tt.data_ptr(), gg.data_ptr(),
                  acc.data_ptr(), small.data_ptr(), ws.data_ptr(), streamk_workspace(dev).data_ptr(), df.data_ptr(),
                  small[m:].data_ptr(), _stream())
    return df, small[m]


def gemm_bf16(a: torch.Tensor, b: torch.Tensor, a_mn_major: bool = False, b_mn_major: bool = False,
              stream_k: bool = True) -> torch.Tensor:
    """C [M, N] fp32 = A . B^T on the tcgen05 kernel.  a is [M, K] (K-major) or, with a_mn_major,
    A^T stored [K, M]; b is [N, K] or, with b_mn_major, B^T stored [K, N]."""
    _req(a, "A", dtype=torch.bfloat16, ndim=2)
    _req(b, "B", dtype=torch.bfloat16, ndim=2)
    k, m = (a.shape if a_mn_major else a.shape[::-1])
    k2, n = (b.shape if b_mn_major else b.shape[::-1])
    if k != k2:
        raise ValueError("gemm_bf16: K mismatch")
    out = torch.empty(m, n, dtype=torch.float32, device=a.device)
    with _on_device(a.device):
        ws = streamk_workspace(a.device) if stream_k else None
        _lib.call("jsd_gemm_bf16", _ptr(a), a.shape[1], int(a_mn_major), _ptr(b), b.shape[1], int(b_mn_major),
                  m, n, k, _ptr(ws), _ptr(out), _stream())
    return out

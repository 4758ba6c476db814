"""Autograd entry points of the JSD estimator.  Each Function only moves tensors
in and out of the C ABI (clip_lite_b200.kernels); the arithmetic of
loss.py:94-105,204-254 and of its backward runs in libjsd_b200.so.

    jsd_index_loss(f, g, t, neg_index=None)   reference semantics: one indexed negative per row
    jsd_dense_loss(f, g, t)                   all off-diagonal pairs as negatives (single GPU)
    ln_normalize_pair(xf, xg, ln_f, ln_g)     tail of the projection heads: LayerNorm + F.normalize in one pass
    jsd_dense_loss_ln(xf, xg, ln_f, ln_g, t)  jsd_dense_loss on the heads' pre-LayerNorm outputs, tail fused in

f, g are the projected features ([B, D]; fp32, bf16 or fp16), t the 0-dim
`temperature` parameter.  Both return (loss, stats) where loss is the 0-dim
CROSS_MODAL_LOSS = Em - Ej of loss.py:254 and stats = [pos, neg, loss, dL/dt]
(detached, for logging without a host sync; the dense mode reports 0 for dL/dt --
its temperature gradient falls out of the backward pass).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import kernels as K


class NegativeIndex:
    """Device-side negative index of the index mode together with its CSR inverse.

    The reference builds its negatives by concatenating / rolling tensors
    (loss.py:214-216 normal mode, loss.py:225-252 cluster mode); here the same
    pairing is an index vector so that the projection heads run once.
    """

    def __init__(self, neg_index: torch.Tensor):
        if neg_index.dim() != 1:
            raise ValueError("neg_index must be 1-D")
        n = neg_index.numel()
        idx = neg_index.detach().to(torch.int64).cpu()
        if n == 0 or int(idx.min()) < 0 or int(idx.max()) >= n:
            raise ValueError("neg_index entries must lie in [0, B)")
        order = torch.argsort(idx, stable=True)
        ptr = torch.zeros(n + 1, dtype=torch.int64)
        ptr[1:] = torch.cumsum(torch.bincount(idx, minlength=n), 0)
        self.n = n
        self._host = (idx.to(torch.int32), ptr.to(torch.int32), order.to(torch.int32))
        self._dev = {}

    def on(self, device) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        key = str(device)
        if key not in self._dev:
            self._dev[key] = tuple(x.to(device) for x in self._host)
        return self._dev[key]

    @staticmethod
    def cluster(half: int) -> "NegativeIndex":
        """Row i < half -> half + i (its hard negative), row half + i -> (i + 1) mod half."""
        i = torch.arange(half)
        return NegativeIndex(torch.cat((half + i, (i + 1) % half)))


def _common(f: torch.Tensor, g: torch.Tensor):
    if f.dim() != 2 or g.dim() != 2:
        raise ValueError(f"features must be [B, D]; got {tuple(f.shape)} and {tuple(g.shape)}")
    if f.shape != g.shape:
        raise ValueError(f"image features {tuple(f.shape)} and text features {tuple(g.shape)} must match after projection")
    if not (f.is_cuda and g.is_cuda):
        raise RuntimeError("the JSD estimator runs on CUDA only (no CPU fallback)")
    dt = torch.promote_types(f.dtype, g.dtype)
    if dt not in (torch.float32, torch.bfloat16, torch.float16):
        dt = torch.float32
    return f.to(dt).contiguous(), g.to(dt).contiguous()


class _JSDIndexFn(torch.autograd.Function):
    """Two calls of the fused kernel per training step: the forward reads F, G and returns the loss (no write-back
    pass); the backward reads them again with the upstream gradient as a device scalar and writes dF, dG in the
    feature dtype -- 6 B D elements of HBM traffic, where "gradients in the forward + one scaling pass in the
    backward" moves 8, and the upstream gradient (GradScaler's factor included) is applied in fp32 before the
    single rounding, so fp16 gradients never pass through the subnormal range."""

    @staticmethod
    def forward(ctx, f, g, t, neg: Optional[NegativeIndex]):
        need_grad = any(ctx.needs_input_grad)
        with torch.autocast("cuda", enabled=False):
            fc, gc = _common(f, g)
            if neg is not None and neg.n != fc.shape[0]:
                raise ValueError(f"neg_index has {neg.n} entries for a batch of {fc.shape[0]}")
            ix = neg.on(fc.device) if neg is not None else (None, None, None)
            out4, loss, _, _, _ = K.index_fwd_bwd(fc, gc, t, *ix, want_grad=False)
        if need_grad:
            ctx.save_for_backward(fc, gc, t, out4)
            ctx.ix = ix
        ctx.dtypes = (f.dtype, g.dtype, t.dtype)
        ctx.mark_non_differentiable(out4)
        return loss, out4

    @staticmethod
    def backward(ctx, grad_loss, _grad_stats):
        fc, gc, t, out4 = ctx.saved_tensors
        with torch.autocast("cuda", enabled=False):
            go = grad_loss.float()
            _, _, df, dg, _ = K.index_fwd_bwd(fc, gc, t, *ctx.ix, want_grad=True, gamma=go)
        fd, gd, td = ctx.dtypes
        return df.to(fd), dg.to(gd), (go * out4[3]).to(td), None


class _JSDDenseFn(torch.autograd.Function):
    """D in {64, 128, 192, 256}: the single-pass fused kernel computes the loss AND both gradient contractions in
    the forward call (score tile, sigma(S) and the accumulators never leave the SM; nothing B x B is stored); the
    backward only applies the upstream gradient and the normalisation Jacobians.  Other D: staged path (forward
    writes the bf16 sigma matrix, backward runs the two contractions)."""

    @staticmethod
    def forward(ctx, f, g, t):
        need_grad = any(ctx.needs_input_grad)
        with torch.autocast("cuda", enabled=False):
            fc, gc = _common(f, g)
            ctx.fused = need_grad and K.fused_supported(fc.shape[0], fc.shape[1])
            if ctx.fused:
                out4, loss, saved = K.dense_fused_forward(fc, gc, t)
            else:
                out4, loss, saved = K.dense_forward(fc, gc, t, want_grad=need_grad)
        if need_grad:
            ctx.save_for_backward(fc, gc, t, *saved)
        ctx.dtypes = (f.dtype, g.dtype, t.dtype)
        ctx.mark_non_differentiable(out4)
        return loss, out4

    @staticmethod
    def backward(ctx, grad_loss, _grad_stats):
        fc, gc, t, *saved = ctx.saved_tensors
        with torch.autocast("cuda", enabled=False):
            fn = K.dense_fused_backward if ctx.fused else K.dense_backward
            df, dg, dt = fn(fc, gc, t, grad_loss, saved)
        fd, gd, td = ctx.dtypes
        return df.to(fd), dg.to(gd), dt.to(td)


def _ln_args(ln):
    """(weight, bias, eps) of an nn.LayerNorm (or of a plain tuple)."""
    if isinstance(ln, torch.nn.LayerNorm):
        return ln.weight, ln.bias, float(ln.eps)
    w, b, eps = ln
    return w, b, float(eps)


def _pair_inputs(xf: torch.Tensor, xg: torch.Tensor):
    if xf.dim() != 2 or xg.dim() != 2 or xf.shape != xg.shape:
        raise ValueError(f"head outputs must be two [B, D] tensors of the same shape; got {tuple(xf.shape)} and "
                         f"{tuple(xg.shape)}")
    if not (xf.is_cuda and xg.is_cuda):
        raise RuntimeError("the projection-head tail runs on CUDA only (no CPU fallback)")
    dt = torch.promote_types(xf.dtype, xg.dtype)
    if dt not in (torch.float32, torch.bfloat16, torch.float16):
        dt = torch.float32
    return xf.to(dt).contiguous(), xg.to(dt).contiguous()


def _param_grad(g: Optional[torch.Tensor], p: Optional[torch.Tensor]):
    return None if (g is None or p is None) else g.to(p.dtype)


class _LnNormalizePairFn(torch.autograd.Function):
    """(xf, xg) -> fp32 unit rows of LN(xf), LN(xg): one launch forward; backward = Jacobian of F.normalize +
    LayerNorm backward + the weight / bias sums in one pass over (x, upstream gradient) and one small reduction
    launch.  The [B, D] LayerNorm output is never stored: 12 bytes per row (mean, rstd, 1/||.||) are kept."""

    @staticmethod
    def forward(ctx, xf, xg, wf, bf, wg, bg, eps_f, eps_g):
        with torch.autocast("cuda", enabled=False):
            xfc, xgc = _pair_inputs(xf, xg)
            uf, ug, st_f, st_g = K.ln_normalize_pair(xfc, xgc, (wf, bf, eps_f), (wg, bg, eps_g), out_bf16=False)
        ctx.save_for_backward(xfc, xgc, wf, bf, wg, bg, st_f, st_g)
        ctx.dtypes = (xf.dtype, xg.dtype)
        return uf, ug

    @staticmethod
    def backward(ctx, guf, gug):
        xfc, xgc, wf, bf, wg, bg, st_f, st_g = ctx.saved_tensors
        with torch.autocast("cuda", enabled=False):
            guf = torch.zeros_like(xfc, dtype=torch.float32) if guf is None else guf.float().contiguous()
            gug = torch.zeros_like(xgc, dtype=torch.float32) if gug is None else gug.float().contiguous()
            dxf, dxg, dwf, dbf, dwg, dbg, _ = K.ln_normalize_bwd_pair(xfc, xgc, (wf, bf, 0.0), (wg, bg, 0.0), st_f,
                                                                      st_g, guf, gug)
        fd, gd = ctx.dtypes
        return (dxf.to(fd), dxg.to(gd), _param_grad(dwf, wf), _param_grad(dbf, bf), _param_grad(dwg, wg),
                _param_grad(dbg, bg), None, None)


def ln_normalize_pair(xf: torch.Tensor, xg: torch.Tensor, ln_f, ln_g):
    """Tail of the two projection heads (reference loss.py:36-38 then :94-95) in one launch: fp32 unit rows
    LN(x) / max(||LN(x)||, 1e-12) of the image and the text head.  ln_* is the head's nn.LayerNorm (or a
    (weight, bias, eps) tuple).  Feed the result to jsd_index_loss / any gathered estimator: their own
    normalisation of a unit row is the identity, and its Jacobian is the projection this backward applies anyway."""
    wf, bf, ef = _ln_args(ln_f)
    wg, bg, eg = _ln_args(ln_g)
    return _LnNormalizePairFn.apply(xf, xg, wf, bf, wg, bg, ef, eg)


class _JSDDenseLnFn(torch.autograd.Function):
    """jsd_dense_loss with the heads' tail fused into its row passes (single GPU): the forward's normalise pass
    applies LayerNorm first (bf16 unit rows straight from the pre-LayerNorm head output), the backward's Jacobian
    pass continues through LayerNorm and sums its weight / bias gradients -- no LayerNorm output, no separate
    LayerNorm kernels in either direction."""

    @staticmethod
    def forward(ctx, xf, xg, wf, bf, wg, bg, t, eps_f, eps_g):
        need_grad = any(ctx.needs_input_grad)
        with torch.autocast("cuda", enabled=False):
            xfc, xgc = _pair_inputs(xf, xg)
            b, d = xfc.shape
            u, v, st_f, st_g = K.ln_normalize_pair(xfc, xgc, (wf, bf, eps_f), (wg, bg, eps_g), out_bf16=True)
            ctx.fused = need_grad and K.fused_supported(b, d)
            if ctx.fused:
                out4, loss, gdiag, acc = K.dense_fused_fwd_bwd(u, v, t)
                keep = (acc,)
            else:
                out4, loss, gmat, gdiag = K.dense_fwd(u, v, t, 0, want_grad=need_grad)
                keep = (gmat,) if need_grad else ()
        if need_grad:
            ctx.save_for_backward(xfc, xgc, wf, bf, wg, bg, t, u, v, st_f, st_g, gdiag, *keep)
        ctx.dtypes = (xf.dtype, xg.dtype, t.dtype)
        ctx.mark_non_differentiable(out4)
        return loss, out4

    @staticmethod
    def backward(ctx, grad_loss, _grad_stats):
        xfc, xgc, wf, bf, wg, bg, t, u, v, st_f, st_g, gdiag, kept = ctx.saved_tensors
        b = xfc.shape[0]
        with torch.autocast("cuda", enabled=False):
            gamma = grad_loss.float()
            if ctx.fused:
                acc_u, acc_v, scale = kept[0], kept[1], 1.0 / (b * (b - 1.0))
            else:
                acc_u = K.dense_bwd_du(kept, v, t, gamma)
                acc_v = K.dense_bwd_dv(kept, u, b, t, gamma)
                scale = 0.0
            dxf, dxg, dwf, dbf, dwg, dbg, dt = K.ln_normalize_bwd_pair(
                xfc, xgc, (wf, bf, 0.0), (wg, bg, 0.0), st_f, st_g, acc_u, acc_v, acc_scale=scale, partner0=v,
                partner1=u, gdiag=gdiag, t=t, gamma=gamma, m_rows=b, want_dt=True)
        fd, gd, td = ctx.dtypes
        return (dxf.to(fd), dxg.to(gd), _param_grad(dwf, wf), _param_grad(dbf, bf), _param_grad(dwg, wg),
                _param_grad(dbg, bg), dt.to(td), None, None)


def jsd_dense_loss_ln(xf: torch.Tensor, xg: torch.Tensor, ln_f, ln_g, t: torch.Tensor):
    """All-pairs estimator on the projection heads' PRE-LayerNorm outputs (loss.py:36 `f` before
    feature_block_ln): LayerNorm, F.normalize, the bf16 cast and 1/||.|| happen in the estimator's own row pass,
    and so does their backward.  Same value and gradients as jsd_dense_loss(ln_f(xf), ln_g(xg), t), plus the
    LayerNorm parameter gradients."""
    if xf.shape[0] < 2:
        raise ValueError("the dense estimator needs at least two rows (one negative per row)")
    wf, bf, ef = _ln_args(ln_f)
    wg, bg, eg = _ln_args(ln_g)
    return _JSDDenseLnFn.apply(xf, xg, wf, bf, wg, bg, t, ef, eg)


def jsd_index_loss(f: torch.Tensor, g: torch.Tensor, t: torch.Tensor, neg_index: Optional[NegativeIndex] = None):
    """Reference estimator (loss.py:204-254): negatives = text rows rolled by one,
    or the rows named by ``neg_index``."""
    return _JSDIndexFn.apply(f, g, t, neg_index)


def jsd_dense_loss(f: torch.Tensor, g: torch.Tensor, t: torch.Tensor):
    """All-pairs estimator: mean_i sp(-S_ii) + mean_{i != j} sp(S_ij) on tensor cores."""
    if f.shape[0] < 2:
        raise ValueError("the dense estimator needs at least two rows (one negative per row)")
    return _JSDDenseFn.apply(f, g, t)

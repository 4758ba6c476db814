"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

CPU / fp64 restatement of the TAIL of CLIP-Lite's projection heads, the step in front of the estimator
(SURVEY 8-f #1):

* ``layer_norm_rows``     <- ``self.feature_block_ln(f)`` at the end of ``MILinearBlock.forward``, loss.py:36-38
                             (nn.LayerNorm over the last dimension: biased variance, eps inside the square root)
* ``ln_unit``             <- that LayerNorm followed by ``F.normalize(t, p=2, dim=-1)`` of
                             ``GlobalDiscriminatorDot.forward``, loss.py:94-95 (eps 1e-12 on the norm)
* ``ln_unit_grads``       <- closed form of the autograd backward of both (what PyTorch derives for train.py:218):
                             Jacobian of the normalisation, LayerNorm's input gradient and its weight / bias sums

Pinned in tests/test_oracle.py against fp64 autograd of the very PyTorch ops the reference calls, and against the
live reference module (its MILinearBlock's LayerNorm + F.normalize) when /root/reference is present.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

NORM_EPS = 1e-12          # F.normalize default eps (loss.py:94)


def layer_norm_rows(x: torch.Tensor, w: Optional[torch.Tensor], b: Optional[torch.Tensor], eps: float):
    """(y, xhat, rstd): y = xhat * w + b with xhat = (x - mean) * rstd, rstd = 1 / sqrt(var_biased + eps)."""
    mean = x.mean(dim=-1, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=-1, keepdim=True)
    rstd = 1.0 / torch.sqrt(var + eps)
    xhat = (x - mean) * rstd
    y = xhat if w is None else xhat * w
    if b is not None:
        y = y + b
    return y, xhat, rstd


def ln_unit(x: torch.Tensor, w: Optional[torch.Tensor], b: Optional[torch.Tensor], eps: float):
    """(u, stats): u = LN(x) / max(||LN(x)||, 1e-12); stats = (mean, rstd, 1 / max(||LN(x)||, 1e-12)) per row."""
    y, _, rstd = layer_norm_rows(x, w, b, eps)
    n = torch.sqrt((y * y).sum(dim=-1, keepdim=True)).clamp_min(NORM_EPS)
    return y / n, (x.mean(dim=-1), rstd.squeeze(-1), (1.0 / n).squeeze(-1))


def ln_unit_grads(x: torch.Tensor, w: Optional[torch.Tensor], b: Optional[torch.Tensor], eps: float,
                  du: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """(dx, dw, db, rowdot) for an upstream gradient ``du`` with respect to the unit rows:
         rowdot = <u, du>                                   (per row; its sum is gamma dL/dt in the dense estimator)
         dy     = (du - u rowdot) / max(||y||, eps)          Jacobian of F.normalize
         dw     = sum_rows dy xhat,  db = sum_rows dy        LayerNorm parameters
         dx     = rstd (dy w - mean_d(dy w) - xhat mean_d(dy w xhat))
    (dw / db are returned for w = 1 / b = 0 as well when the LayerNorm has no affine parameters.)"""
    y, xhat, rstd = layer_norm_rows(x, w, b, eps)
    n = torch.sqrt((y * y).sum(dim=-1, keepdim=True)).clamp_min(NORM_EPS)
    u = y / n
    rowdot = (u * du).sum(dim=-1, keepdim=True)
    dy = (du - u * rowdot) / n
    g = dy if w is None else dy * w
    dx = rstd * (g - g.mean(dim=-1, keepdim=True) - xhat * (g * xhat).mean(dim=-1, keepdim=True))
    return dx, (dy * xhat).sum(dim=0), dy.sum(dim=0), rowdot.squeeze(-1)

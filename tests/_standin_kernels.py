"""TEST-ONLY stand-in for clip_lite_b200.kernels backed by the oracle, so that the
host-side sharding / collective logic of clip_lite_b200.parallel can be exercised on
CPU with the gloo backend.  It mirrors the kernels' contracts (bf16 unit operands,
padded pitches, fp32 accumulators); the product never imports this file."""
import torch

from oracle import jsd_oracle as orc


def _round_up(x, m):
    return (x + m - 1) // m * m


def normalize_cast(x):
    u, n = orc.l2_normalize(x.float())
    return u.bfloat16(), (1.0 / n.squeeze(-1)).float()


def dense_fwd(u, v, t, row_offset=0, want_grad=True):
    d = orc.dense_from_unit(u.double(), v.double(), float(t), row_offset)
    out4 = torch.stack((d["pos"], d["neg"], d["loss"], torch.zeros_like(d["loss"]))).float()
    gmat = None
    if want_grad:
        gmat = torch.zeros(u.shape[0], _round_up(v.shape[0], 64), dtype=torch.bfloat16)
        gmat[:, :v.shape[0]] = d["gmat"].bfloat16()
    return out4, out4[2].clone(), gmat, d["gdiag"].float()


def _scale(m, n, t, gamma):
    g = 1.0 if gamma is None else float(gamma)
    return g * float(torch.as_tensor(float(t)).exp()) / (m * (n - 1))


def dense_bwd_du(gmat, v, t, gamma=None):
    m, n = gmat.shape[0], v.shape[0]
    return (_scale(m, n, t, gamma) * (gmat[:, :n].double() @ v.double())).float()


def dense_bwd_dv(gmat, u, n, t, gamma=None):
    m = gmat.shape[0]
    return (_scale(m, n, t, gamma) * (gmat[:, :n].double().t() @ u.double())).float()


def normalize_bwd(x, inv_norm, acc, partner, partner_offset, gdiag, t, gamma, m_rows, want_dt=False):
    g = 1.0 if gamma is None else float(gamma)
    rows = x.shape[0]
    d = acc.double()
    if gdiag is not None:
        c = g * float(torch.as_tensor(float(t)).exp()) / m_rows
        d = d + c * gdiag.double()[:, None] * partner[partner_offset:partner_offset + rows].double()
    inv = inv_norm.double()[:, None]
    u = x.double() * inv
    dx = ((d - u * (u * d).sum(-1, keepdim=True)) * inv).to(x.dtype)
    return (dx, (u * d).sum().float()) if want_dt else dx


def normalize_cast_pair(f, g):
    u, inv_f = normalize_cast(f)
    v, inv_g = normalize_cast(g)
    return u, v, inv_f, inv_g


def dense_backward_image_side(f, v_all, inv_f, gmat, gdiag, t, gamma, row_offset):
    du = dense_bwd_du(gmat, v_all, t, gamma)
    return normalize_bwd(f, inv_f, du, v_all, row_offset, gdiag, t, gamma, f.shape[0], want_dt=True)

"""CPU tests of the oracle itself: pinned against the golden vectors generated from
the unmodified reference loss.py (tests/golden/make_golden.py), against the live
reference when /root/reference is present, and against its own identities."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import jsd_oracle as orc
from oracle import reference_loader as rl


def _cases(golden_dir, pattern):
    files = sorted(glob.glob(os.path.join(golden_dir, pattern)))
    assert files, "golden vectors missing"
    return [np.load(p, allow_pickle=True) for p in files]


def _oracle_on_case(z):
    """Re-assemble the reference forward of one estimator_* case from oracle pieces."""
    t = float(z["t"])
    f = torch.from_numpy(z["in_image_features"]).double()
    g = torch.from_numpy(z["in_text_features"]).double()
    grads = {}
    if str(z["mode"]) == "cluster":
        half = f.shape[0]
        fa = torch.cat((f, torch.from_numpy(z["in_neg_image_features"]).double()))
        ga = torch.cat((g, torch.from_numpy(z["in_neg_text_features"]).double()))
        neg = orc.cluster_index(half)
        cross = orc.jsd_index(fa, ga, t, neg)["loss"]
        df, dg, dt = orc.jsd_index_grads(fa, ga, t, neg, gamma=0.9)
        grads.update(image_features=df[:half], neg_image_features=df[half:],
                     text_features=dg[:half], neg_text_features=dg[half:])
        g_for_ssl = torch.cat((g[1:], g[:1]))          # loss.py:237-239 re-binds text_features
        unroll = True
    else:
        cross = orc.jsd_index(f, g, t)["loss"]
        df, dg, dt = orc.jsd_index_grads(f, g, t, gamma=0.9)
        grads.update(image_features=df, text_features=dg)
        g_for_ssl, unroll = g, False
    visual = textual = torch.zeros((), dtype=torch.float64)
    if bool(z["ssl"]):
        ts = float(z["t_ssl"])
        af = torch.from_numpy(z["in_aug_image_features"]).double()
        ag = torch.from_numpy(z["in_aug_text_features"]).double()
        visual = orc.jsd_index(f, af, ts)["loss"]
        d1, d2, _ = orc.jsd_index_grads(f, af, ts, gamma=0.9)
        grads["image_features"] = grads["image_features"] + d1
        grads["aug_image_features"] = d2
        textual = orc.jsd_index(g_for_ssl, ag, ts)["loss"]
        d1, d2, _ = orc.jsd_index_grads(g_for_ssl, ag, ts, gamma=0.9)
        if unroll:
            d1 = torch.cat((d1[-1:], d1[:-1]))
        grads["text_features"] = grads["text_features"] + d1
        grads["aug_text_features"] = d2
    return cross, visual, textual, grads, dt


def test_oracle_matches_reference_golden(golden_dir):
    for z in _cases(golden_dir, "estimator_*.npz"):
        cross, visual, textual, grads, dt = _oracle_on_case(z)
        assert abs(float(cross) - float(z["out_cross_modal_loss"])) < 1e-12
        assert abs(float(visual) - float(z["out_visual_loss"])) < 1e-12
        assert abs(float(textual) - float(z["out_textual_loss"])) < 1e-12
        total = 0.9 * (cross + visual + textual)
        assert abs(float(total) - float(z["out_total_loss"])) < 1e-12
        for k, v in grads.items():
            assert np.abs(v.numpy() - z["grad_" + k]).max() < 1e-13, k
        assert abs(float(dt) - float(z["grad_temperature"])) < 1e-12


def test_oracle_fp32_within_tolerance_of_golden(golden_dir):
    """The fp32 run of the oracle (the timed CPU baseline) agrees with the fp64 golden far inside 1e-3."""
    for z in _cases(golden_dir, "estimator_b*.npz"):
        f = torch.from_numpy(z["in_image_features"])
        g = torch.from_numpy(z["in_text_features"])
        out = orc.jsd_index(f, g, float(z["t"]))
        ref = float(z["out_cross_modal_loss"])
        assert abs(float(out["loss"]) - ref) < 1e-5 * abs(ref)


@pytest.mark.skipif(not rl.reference_available(), reason="reference tree not mounted")
@pytest.mark.parametrize("b,d,seed", [(2, 4, 0), (5, 16, 1), (32, 64, 2)])
def test_oracle_matches_live_reference(b, d, seed):
    f, g = orc.synth_embeddings(b, d, seed, correlated=True, dtype=torch.float64)
    m = rl.reference_estimator_module().double()
    fl, gl = f.clone().requires_grad_(True), g.clone().requires_grad_(True)
    with rl.cuda_calls_neutralised():
        out = m(fl, gl)
    out["total_loss"].backward()
    t = float(m.global_d.temperature)          # log(1/0.07) rounded to fp32 at construction (loss.py:82)
    ref = orc.jsd_index(f, g, t)
    assert abs(float(out["cross_modal_loss"]) - float(ref["loss"])) < 1e-12
    df, dg, dt = orc.jsd_index_grads(f, g, t, gamma=0.9)
    assert (fl.grad - df).abs().max() < 1e-13 and (gl.grad - dg).abs().max() < 1e-13
    assert abs(float(m.global_d.temperature.grad) - float(dt)) < 1e-12


@pytest.mark.parametrize("kind", ["index", "dense"])
def test_closed_form_grads_match_autograd(kind):
    f, g = orc.synth_embeddings(12, 16, 0, True, torch.float64)
    kw = {"row_offset": 0} if kind == "dense" else {}
    _, df, dg, dt = orc.autograd_step(kind, f, g, 2.1, gamma=0.7, **kw)
    fn = orc.jsd_index_grads if kind == "index" else orc.jsd_dense_grads
    df2, dg2, dt2 = fn(f, g, 2.1, gamma=0.7)
    assert (df - df2).abs().max() < 1e-14 and (dg - dg2).abs().max() < 1e-14 and abs(dt - dt2) < 1e-13


def test_dense_is_mean_over_shifts_of_reference_estimator():
    b = 9
    f, g = orc.synth_embeddings(b, 8, 3, True, torch.float64)
    negs = [orc.jsd_index(f, g, 2.0, (torch.arange(b) + k) % b)["neg"] for k in range(1, b)]
    dense = orc.jsd_dense(f, g, 2.0)
    assert abs(float(torch.stack(negs).mean()) - float(dense["neg"])) < 1e-13
    assert abs(float(orc.jsd_index(f, g, 2.0)["pos"]) - float(dense["pos"])) < 1e-14


def test_row_slabs_compose_to_the_global_loss_and_grads():
    b, world = 12, 3
    m = b // world
    f, g = orc.synth_embeddings(b, 8, 5, True, torch.float64)
    full = orc.jsd_dense(f, g, 2.3)["loss"]
    df, dg, dt = orc.jsd_dense_grads(f, g, 2.3)
    slab_loss, sdf, sdg, sdt = 0.0, [], 0.0, 0.0
    for r in range(world):
        sl = slice(r * m, (r + 1) * m)
        slab_loss = slab_loss + orc.jsd_dense(f[sl], g, 2.3, row_offset=r * m)["loss"]
        a, bb, c = orc.jsd_dense_grads(f[sl], g, 2.3, row_offset=r * m)
        sdf.append(a)
        sdg = sdg + bb
        sdt = sdt + c
    assert abs(float(slab_loss / world) - float(full)) < 1e-13
    assert (torch.cat(sdf) / world - df).abs().max() < 1e-14
    assert (sdg / world - dg).abs().max() < 1e-14 and abs(sdt / world - dt) < 1e-13


def test_unit_level_restatement_is_consistent():
    f, g = orc.synth_embeddings(10, 8, 1, True, torch.float64)
    u, _ = orc.l2_normalize(f)
    v, _ = orc.l2_normalize(g)
    d = orc.dense_from_unit(u, v, 2.0, gamma=0.5)
    assert abs(float(d["loss"]) - float(orc.jsd_dense(f, g, 2.0)["loss"])) < 1e-13
    uu = u.clone().requires_grad_(True)
    vv = v.clone().requires_grad_(True)
    tt = torch.tensor(2.0, dtype=torch.float64, requires_grad=True)
    s = tt.exp() * (uu @ vv.t())
    eye = torch.eye(10, dtype=torch.bool)
    loss = orc.softplus(-s[eye]).mean() + orc.softplus(s[~eye]).sum() / 90
    (0.5 * loss).backward()
    assert (uu.grad - d["du"]).abs().max() < 1e-14 and (vv.grad - d["dv"]).abs().max() < 1e-14
    assert abs(float(tt.grad) - 0.5 * float(d["dt"])) < 1e-13


def test_softplus_threshold_and_edge_cases():
    x = torch.tensor([-50.0, 0.0, 19.9, 20.0, 20.1, 80.0], dtype=torch.float64)
    assert torch.allclose(orc.softplus(x), torch.nn.functional.softplus(x), atol=0, rtol=1e-15)
    z, n = orc.l2_normalize(torch.zeros(2, 4))
    assert torch.equal(z, torch.zeros(2, 4)) and float(n[0]) == pytest.approx(1e-12)
    assert orc.shift1_index(1).tolist() == [0]
    assert orc.cluster_index(3).tolist() == [3, 4, 5, 1, 2, 0]


# ------------------------------------------------------------------ projection-head tail
@pytest.mark.parametrize("affine", [True, False])
@pytest.mark.parametrize("rows,d", [(7, 16), (33, 100), (4, 2048)])
def test_heads_oracle_closed_forms_equal_autograd_of_the_reference_ops(rows, d, affine):
    """oracle.heads_oracle against fp64 autograd of nn.LayerNorm (loss.py:36-38) + F.normalize (loss.py:94-95)."""
    from oracle import heads_oracle as ho
    g = torch.Generator().manual_seed(rows * d)
    x = (torch.randn(rows, d, generator=g, dtype=torch.float64) * 1.7 + 0.4).requires_grad_(True)
    ln = torch.nn.LayerNorm(d, eps=1e-5, elementwise_affine=affine).double()
    if affine:
        with torch.no_grad():
            ln.weight.copy_(1.0 + 0.3 * torch.randn(d, generator=g, dtype=torch.float64))
            ln.bias.copy_(0.2 * torch.randn(d, generator=g, dtype=torch.float64))
    du = torch.randn(rows, d, generator=g, dtype=torch.float64)
    ref_u = torch.nn.functional.normalize(ln(x), p=2, dim=-1)
    (ref_u * du).sum().backward()
    w, b = (ln.weight.detach(), ln.bias.detach()) if affine else (None, None)
    u, (mean, rstd, inv) = ho.ln_unit(x.detach(), w, b, ln.eps)
    dx, dw, db, rowdot = ho.ln_unit_grads(x.detach(), w, b, ln.eps, du)
    assert (u - ref_u).abs().max() < 1e-14
    assert (dx - x.grad).abs().max() < 1e-12 * x.grad.abs().max().clamp_min(1.0)
    if affine:
        assert (dw - ln.weight.grad).abs().max() < 1e-12 * ln.weight.grad.abs().max().clamp_min(1.0)
        assert (db - ln.bias.grad).abs().max() < 1e-12 * ln.bias.grad.abs().max().clamp_min(1.0)
    assert (rowdot - (ref_u.detach() * du).sum(-1)).abs().max() < 1e-13
    assert (mean - x.detach().mean(-1)).abs().max() < 1e-14 and (inv - 1.0 / ln(x).detach().norm(dim=-1)).abs().max() < 1e-12
    assert (rstd - 1.0 / torch.sqrt(x.detach().var(-1, unbiased=False) + ln.eps)).abs().max() < 1e-13


@pytest.mark.skipif(not rl.reference_available(), reason="needs /root/reference (build container)")
def test_heads_oracle_matches_the_live_reference_head():
    """The reference's own MILinearBlock (loss.py:12-40) followed by its F.normalize (loss.py:94-95): the oracle's
    ln_unit applied to the head's pre-LayerNorm sum reproduces the unit rows the reference's dot critic scores."""
    from oracle import heads_oracle as ho
    ref = rl.load_reference_loss()
    torch.manual_seed(0)
    block = ref.MILinearBlock(24, units=64).double().eval()
    with torch.no_grad():
        block.feature_block_ln.weight.uniform_(0.5, 1.5)
        block.feature_block_ln.bias.uniform_(-0.3, 0.3)
    x = torch.randn(9, 24, dtype=torch.float64)
    with torch.no_grad():
        want = torch.nn.functional.normalize(block(x), p=2, dim=-1)          # loss.py:91-95 on one head
        pre = block.feature_nonlinear(x) + block.feature_shortcut(x)         # loss.py:36
    ln = block.feature_block_ln
    got, _ = ho.ln_unit(pre, ln.weight.detach(), ln.bias.detach(), ln.eps)
    assert (got - want).abs().max() < 1e-13

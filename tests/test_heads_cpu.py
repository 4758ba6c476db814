"""Host side of the fused projection-head tail (fused_heads=True; reference loss.py:36-38 + :94-95) on CPU:
clip_lite_b200/kernels.py, ops.py and loss.py run unmodified on CPU tensors, the row-wise entry points land in the
CPU emulation of the same kernel source (tests/_emu_backend.py), the tensor-core entry points in the oracle-backed
stand-ins.  Checks the ctypes marshalling of the long argument lists, the autograd wiring (which gradient goes to
which input, LayerNorm parameters included) and the module option -- against fp64 PyTorch autograd of
LayerNorm -> estimator and against the golden vectors generated from the unmodified reference module.
The GPU tier (tests/test_zz_gpu_heads.py) repeats the same comparisons on the real library."""
import os

import numpy as np
import pytest
import torch

from oracle import jsd_oracle as orc
from tests import _emu_backend

pytestmark = pytest.mark.skipif(not _emu_backend.available(), reason="needs g++ and the CUDA headers")


def rel(a, b):
    a, b = torch.as_tensor(a).detach().double(), torch.as_tensor(b).detach().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def make_heads(b, d, seed, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    xf = (torch.randn(b, d, generator=g) * 1.5 + 0.2).to(dtype)
    xg = (0.5 * xf.float() + torch.randn(b, d, generator=g)).to(dtype)
    lns = []
    for _ in range(2):
        ln = torch.nn.LayerNorm(d)
        with torch.no_grad():
            ln.weight.copy_(1.0 + 0.3 * torch.randn(d, generator=g))
            ln.bias.copy_(0.2 * torch.randn(d, generator=g))
        lns.append(ln)
    return xf, xg, lns[0], lns[1]


def reference(xf, xg, ln_f, ln_g, t, estimator, gamma=1.0, **kw):
    """fp64 autograd of LayerNorm -> estimator (the oracle normalises internally, as loss.py:94-95)."""
    leaves = [xf.double().clone().requires_grad_(True), xg.double().clone().requires_grad_(True)]
    params = [p.detach().double().clone().requires_grad_(True) for p in (ln_f.weight, ln_f.bias, ln_g.weight, ln_g.bias)]
    tt = torch.tensor(float(t), dtype=torch.float64, requires_grad=True)
    d = xf.shape[1]
    f = torch.nn.functional.layer_norm(leaves[0], (d,), params[0], params[1], ln_f.eps)
    g = torch.nn.functional.layer_norm(leaves[1], (d,), params[2], params[3], ln_g.eps)
    loss = estimator(f, g, tt, **kw)["loss"]
    (gamma * loss).backward()
    return loss.detach(), [x.grad for x in leaves], [p.grad for p in params], tt.grad


@pytest.mark.parametrize("b,d,blocks", [(8, 2048, 3), (5, 100, 2), (12, 256, 12)])
def test_ln_normalize_pair_autograd(monkeypatch, b, d, blocks):
    from clip_lite_b200 import ops
    calls = _emu_backend.install(monkeypatch, bwd_blocks=blocks)
    xf, xg, ln_f, ln_g = make_heads(b, d, seed=b + d)
    xf.requires_grad_(True)
    xg.requires_grad_(True)
    uf, ug = ops.ln_normalize_pair(xf, xg, ln_f, ln_g)
    gen = torch.Generator().manual_seed(1)
    cf, cg = torch.randn(b, d, generator=gen), torch.randn(b, d, generator=gen)
    ((uf * cf).sum() + (ug * cg).sum()).backward()
    assert calls == ["jsd_ln_normalize_pair", "jsd_ln_normalize_bwd_pair"]
    for x, ln, c, u in ((xf, ln_f, cf, uf), (xg, ln_g, cg, ug)):
        xr = x.detach().double().requires_grad_(True)
        w, bb = ln.weight.detach().double().requires_grad_(True), ln.bias.detach().double().requires_grad_(True)
        y = torch.nn.functional.layer_norm(xr, (d,), w, bb, ln.eps)
        ur = y / y.norm(dim=-1, keepdim=True)
        (ur * c.double()).sum().backward()
        assert u.dtype == torch.float32 and rel(u, ur) < 1e-5
        assert rel(x.grad, xr.grad) < 5e-5
        assert rel(ln.weight.grad, w.grad) < 5e-5 and rel(ln.bias.grad, bb.grad) < 5e-5


def test_ln_normalize_pair_one_output_unused(monkeypatch):
    """Only one of the two unit-row outputs reaches the loss: the other side's gradients are exact zeros."""
    from clip_lite_b200 import ops
    _emu_backend.install(monkeypatch)
    xf, xg, ln_f, ln_g = make_heads(6, 64, seed=2)
    xf.requires_grad_(True)
    xg.requires_grad_(True)
    uf, _ = ops.ln_normalize_pair(xf, xg, ln_f, ln_g)
    uf[:, 0].sum().backward()
    assert float(xg.grad.abs().max()) == 0.0 and float(ln_g.weight.grad.abs().max()) == 0.0
    assert float(xf.grad.abs().max()) > 0.0


@pytest.mark.parametrize("mode", ["shift1", "cluster"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_index_mode_with_fused_tail(monkeypatch, mode, dtype):
    """ln_normalize_pair -> jsd_index_loss (both through the emulated kernels) == LayerNorm -> reference estimator."""
    from clip_lite_b200 import ops
    _emu_backend.install(monkeypatch)
    b, d = 12, 256
    xf, xg, ln_f, ln_g = make_heads(b, d, seed=4, dtype=dtype)
    xf.requires_grad_(True)
    xg.requires_grad_(True)
    t = torch.tensor(orc.T_INIT, requires_grad=True)
    neg = ops.NegativeIndex.cluster(b // 2) if mode == "cluster" else None
    f, g = ops.ln_normalize_pair(xf, xg, ln_f, ln_g)
    loss, _ = ops.jsd_index_loss(f, g, t, neg)
    (0.7 * loss).backward()
    kw = {"neg_index": orc.cluster_index(b // 2)} if mode == "cluster" else {}
    rl, rx, rp, rt = reference(xf.detach(), xg.detach(), ln_f, ln_g, orc.T_INIT, orc.jsd_index, gamma=0.7, **kw)
    tol = 1e-4 if dtype == torch.float32 else 1e-2          # bf16 inputs: the gradient is rounded to bf16
    assert rel(loss, rl) < 1e-5
    assert xf.grad.dtype == dtype
    assert rel(xf.grad, rx[0]) < tol and rel(xg.grad, rx[1]) < tol
    for p, r in zip((ln_f.weight, ln_f.bias, ln_g.weight, ln_g.bias), rp):
        assert rel(p.grad, r) < 1e-4
    assert rel(t.grad, rt) < (1e-4 if dtype == torch.float32 else 1e-3)   # a sum with cancellation (|dt| ~ 1e-4)


@pytest.mark.parametrize("fused_slices", [0, 3])
def test_dense_loss_with_fused_tail(monkeypatch, fused_slices):
    """jsd_dense_loss_ln == LayerNorm -> dense estimator: staged flavour (scaled accumulators of the two
    contractions) and fused flavour (unscaled accumulator slices, gamma * tau / (B (B - 1)) applied in the tail)."""
    from clip_lite_b200 import kernels as K, ops
    calls = _emu_backend.install(monkeypatch)
    b, d = 16, 128
    if fused_slices:
        def fused_fwd_bwd(u, v, t):
            r = orc.dense_from_unit(u.double(), v.double(), float(t))
            out4 = torch.stack((r["pos"], r["neg"], r["loss"], torch.zeros_like(r["loss"]))).float()
            acc = torch.stack((r["gmat"] @ v.double(), r["gmat"].t() @ u.double())).float()     # unscaled sums
            parts = torch.rand(2, fused_slices, b, d)
            parts = parts / parts.sum(1, keepdim=True)                                         # slices that sum to it
            return out4, out4[2].clone(), r["gdiag"].float(), (acc[:, None] * parts).contiguous()
        monkeypatch.setattr(K, "fused_supported", lambda bb, dd: True)
        monkeypatch.setattr(K, "dense_fused_fwd_bwd", fused_fwd_bwd)
    xf, xg, ln_f, ln_g = make_heads(b, d, seed=9)
    xf.requires_grad_(True)
    xg.requires_grad_(True)
    t = torch.tensor(1.3, requires_grad=True)
    loss, stats = ops.jsd_dense_loss_ln(xf, xg, ln_f, ln_g, t)
    (0.9 * loss).backward()
    assert calls == ["jsd_ln_normalize_pair", "jsd_ln_normalize_bwd_pair"]
    rl, rx, rp, rt = reference(xf.detach(), xg.detach(), ln_f, ln_g, 1.3, orc.jsd_dense, gamma=0.9)
    assert rel(loss, rl) < 1e-3                                   # BASELINE tolerance: bf16 unit rows
    assert rel(xf.grad, rx[0]) < 1e-2 and rel(xg.grad, rx[1]) < 1e-2
    for p, r in zip((ln_f.weight, ln_f.bias, ln_g.weight, ln_g.bias), rp):
        assert rel(p.grad, r) < 1e-2
    assert rel(t.grad, rt) < 1e-2
    assert stats.shape == (4,)


def test_dense_loss_with_fused_tail_no_grad(monkeypatch):
    from clip_lite_b200 import ops
    calls = _emu_backend.install(monkeypatch)
    xf, xg, ln_f, ln_g = make_heads(8, 64, seed=3)
    with torch.no_grad():
        loss, _ = ops.jsd_dense_loss_ln(xf, xg, ln_f, ln_g, torch.tensor(1.0))
    ref = orc.jsd_dense(ln_f(xf).double(), ln_g(xg).double(), 1.0)["loss"]
    assert calls == ["jsd_ln_normalize_pair"] and rel(loss, ref) < 1e-3


# ------------------------------------------------------------------ the drop-in module with fused_heads=True
def _seeded_state_dict(shapes, seed):
    from make_golden import seeded_state_dict
    return seeded_state_dict(shapes, seed)


@pytest.mark.parametrize("neg_mode", ["shift1", "dense"])
@pytest.mark.parametrize("case", ["module_b8_train", "module_b8_eval", "module_cluster_b6_train"])
def test_module_fused_heads_matches_reference_golden(monkeypatch, golden_dir, case, neg_mode):
    """fused_heads=True against the golden vectors written from the UNMODIFIED reference module (loss values, input
    gradients, every parameter gradient -- LayerNorm weight / bias included -- and the BatchNorm buffers).  With
    neg_mode="dense" there is no reference number: the fused-tail module must equal the default (unfused) module."""
    from clip_lite_b200 import loss as L
    _emu_backend.install(monkeypatch)
    monkeypatch.setattr(L.JSDInfoMaxLoss, "_require_cuda", staticmethod(lambda t: None))
    z = np.load(os.path.join(golden_dir, case + ".npz"), allow_pickle=True)

    def run(fused):
        m = L.JSDInfoMaxLoss(image_dim=int(z["image_dim"]), text_dim=int(z["text_dim"]), type="dot",
                             image_prior=False, text_prior=False, neg_mode=neg_mode, fused_heads=fused)
        shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        m.load_state_dict(_seeded_state_dict(shapes, int(z["seed"])))
        m.train(bool(z["train"]))
        names = [k[3:] for k in z.files if k.startswith("in_")]
        leaves = {k: torch.from_numpy(z["in_" + k]).requires_grad_(True) for k in names}
        out = m(**leaves)
        out["total_loss"].backward()
        return m, names, leaves, out

    m, names, leaves, out = run(True)
    if neg_mode == "shift1":
        for k in out:
            assert abs(float(out[k]) - float(z["out_" + k])) <= 1e-3 * max(abs(float(z["out_" + k])), 1e-30), k
        for k in names:
            assert rel(leaves[k].grad, z["grad_" + k]) < 1e-2, k
        for k, p in m.named_parameters():
            g = p.grad if p.grad is not None else torch.zeros_like(p)
            w = torch.from_numpy(np.random.RandomState(7).standard_normal(tuple(g.shape) or (1,))).reshape(g.shape)
            scale = max(float(z["pgrad_abs/" + k]), 1e-30)
            assert abs(float((g.double() * w).sum()) - float(z["pgrad_proj/" + k])) < 1e-2 * scale, k
            assert abs(float(g.sum()) - float(z["pgrad_sum/" + k])) < 1e-2 * scale, k
        sd = m.state_dict()
        for k in z.files:
            if k.startswith("buf/"):
                assert np.allclose(sd[k[4:]].numpy(), z[k], rtol=1e-4, atol=1e-5), k
    else:
        m0, _, leaves0, out0 = run(False)
        for k in out:
            assert abs(float(out[k]) - float(out0[k])) <= 1e-3 * max(abs(float(out0[k])), 1e-30), k
        for k in names:
            assert rel(leaves[k].grad, leaves0[k].grad) < 1e-2, k
        for (k, p), (_, p0) in zip(m.named_parameters(), m0.named_parameters()):
            if p0.grad is None:
                assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            else:
                assert rel(p.grad, p0.grad) < 1e-2, k


def test_fused_heads_skips_blocks_without_layernorm(monkeypatch):
    """bln=False heads and replaced (Identity) heads have no LayerNorm to fuse: the default route is taken."""
    from clip_lite_b200 import loss as L
    calls = _emu_backend.install(monkeypatch)
    monkeypatch.setattr(L.JSDInfoMaxLoss, "_require_cuda", staticmethod(lambda t: None))
    m = L.JSDInfoMaxLoss(image_dim=16, text_dim=16, image_prior=False, fused_heads=True)
    m.global_d.img_block = torch.nn.Identity()
    m.global_d.text_block = torch.nn.Identity()
    out = m(torch.randn(6, 16), torch.randn(6, 16))
    assert torch.isfinite(out["total_loss"]) and calls == ["jsd_index_fwd_bwd"]


@pytest.mark.parametrize("fused", [False, True])
def test_heads_dtype_runs_the_heads_under_autocast(monkeypatch, fused):
    """heads_dtype=torch.bfloat16: the heads' GEMMs run under torch.autocast(bf16) whatever the caller's context is
    (the tail / estimator then read bf16 head outputs); the result stays within bf16 rounding of the fp32 module and
    the BatchNorm double-update replay is unaffected."""
    from clip_lite_b200 import loss as L
    _emu_backend.install(monkeypatch)
    monkeypatch.setattr(L.JSDInfoMaxLoss, "_require_cuda", staticmethod(lambda t: None))
    seen = []
    orig = L.MILinearBlock.pre_norm

    def spy(self, feat):
        out = orig(self, feat)
        seen.append(out.dtype)
        return out

    monkeypatch.setattr(L.MILinearBlock, "pre_norm", spy)
    img, txt = torch.randn(16, 24), torch.randn(16, 20)

    def run(dtype):
        torch.manual_seed(0)
        m = L.JSDInfoMaxLoss(image_dim=24, text_dim=20, image_prior=False, fused_heads=fused, heads_dtype=dtype)
        a, b = img.clone().requires_grad_(True), txt.clone().requires_grad_(True)
        out = m(a, b)
        out["total_loss"].backward()
        return m, out, a.grad

    m16, o16, g16 = run(torch.bfloat16)
    assert seen and all(d == torch.bfloat16 for d in seen)
    seen.clear()
    m32, o32, g32 = run(None)
    assert all(d == torch.float32 for d in seen)
    assert abs(float(o16["total_loss"]) - float(o32["total_loss"])) < 3e-2 * abs(float(o32["total_loss"]))
    assert g16.dtype == torch.float32 and rel(g16, g32) < 0.15
    bn16, bn32 = m16.global_d.img_block.feature_nonlinear[1], m32.global_d.img_block.feature_nonlinear[1]
    assert int(bn16.num_batches_tracked) == int(bn32.num_batches_tracked) == 2
    assert rel(bn16.running_mean, bn32.running_mean) < 3e-2
    assert all(p.grad is None or p.grad.dtype == p.dtype for p in m16.parameters())


def test_estimator_goldens_through_the_emulated_index_kernel(monkeypatch, golden_dir):
    """The reference's own estimator semantics (normal, cluster, SSL, hot temperature: tests/golden/estimator_*.npz,
    written from the unmodified loss.py) through the drop-in module with the index kernel's SOURCE running on the
    CPU: same check as tests/test_gpu_module.py::test_module_estimator_cases_match_reference_golden, without a GPU."""
    import glob
    from clip_lite_b200 import loss as L
    calls = _emu_backend.install(monkeypatch)
    monkeypatch.setattr(L.JSDInfoMaxLoss, "_require_cuda", staticmethod(lambda t: None))
    paths = sorted(glob.glob(os.path.join(golden_dir, "estimator_*.npz")))
    assert len(paths) >= 5
    for path in paths:
        z = np.load(path, allow_pickle=True)
        ssl = bool(z["ssl"])
        d = z["in_image_features"].shape[1]
        m = L.JSDInfoMaxLoss(image_dim=d, text_dim=d, type="dot", image_prior=False, text_prior=False,
                             visual_self_supervised=ssl, textual_self_supervised=ssl)
        critics = [m.global_d] + ([m.visual_d, m.textual_d] if ssl else [])
        for c in critics:
            c.img_block = torch.nn.Identity()
            c.text_block = torch.nn.Identity()
        m.global_d.temperature.data.fill_(float(z["t"]))
        if ssl:
            m.visual_d.temperature.data.fill_(float(z["t_ssl"]))
            m.textual_d.temperature.data.fill_(float(z["t_ssl"]))
        names = [k[3:] for k in z.files if k.startswith("in_")]
        leaves = {k: torch.from_numpy(z["in_" + k]).requires_grad_(True) for k in names}
        out = m(**leaves)
        out["total_loss"].backward()
        for k in out:
            ref = float(z["out_" + k])
            assert abs(float(out[k].detach()) - ref) <= 1e-5 * max(abs(ref), 1e-30), (path, k)
        for k in names:
            assert rel(leaves[k].grad, z["grad_" + k]) < 1e-4, (path, k)
        assert rel(m.global_d.temperature.grad, z["grad_temperature"]) < 1e-4, path
        if ssl:
            assert rel(m.visual_d.temperature.grad, z["grad_temperature_visual"]) < 1e-4
            assert rel(m.textual_d.temperature.grad, z["grad_temperature_textual"]) < 1e-4
    assert set(calls) == {"jsd_index_fwd_bwd"}

"""GPU parity of the fused projection-head tail (jsd_heads.cuh; reference loss.py:36-38 LayerNorm + loss.py:94-95
F.normalize, forward and backward): every entry point through the C ABI against fp64 PyTorch, the autograd entry
points against LayerNorm -> oracle, and the drop-in module with fused_heads=True against the golden vectors written
from the unmodified reference module (LayerNorm weight / bias gradients included).

These kernels were written while the round's GPU budget was spent; their logic is covered on the CPU by
tests/test_emu_kernels.py and tests/test_heads_cpu.py (same kernel source under a thread-per-OS-thread shim).  This
file is named to run LAST in the GPU tier so that a failure here cannot hide the result of any other GPU test."""
import os

import numpy as np
import pytest
import torch

from _weights import seeded_state_dict
from oracle import heads_oracle as ho
from oracle import jsd_oracle as orc

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]

LOSS_RTOL = 1e-3     # BASELINE.json
GRAD_RTOL = 1e-2     # BASELINE.json


def relerr(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.fixture(scope="module")
def K():
    from clip_lite_b200 import kernels
    return kernels


@pytest.fixture(scope="module")
def ops():
    from clip_lite_b200 import ops
    return ops


@pytest.fixture(scope="module")
def L():
    from clip_lite_b200 import loss
    return loss


def make_ln(d, seed, affine=True):
    if not affine:
        return None, None
    g = torch.Generator().manual_seed(seed)
    return ((1.0 + 0.3 * torch.randn(d, generator=g)).cuda(), (0.2 * torch.randn(d, generator=g)).cuda())


def ln_unit_reference(x, w, b, eps):
    """(unit rows, mean, rstd, 1/||LN(x)||) by the fp64 oracle (oracle/heads_oracle.py <- loss.py:36-38, :94-95)."""
    u, (mean, rstd, inv) = ho.ln_unit(x.double(), None if w is None else w.double(),
                                      None if b is None else b.double(), eps)
    return u, mean, rstd, inv


def ln_bwd_reference(x, w, b, eps, du):
    """(dx, dw, db, <u, dU>) for the upstream gradient dU by the fp64 oracle's closed forms (pinned against
    autograd of the reference's ops in tests/test_oracle.py)."""
    return ho.ln_unit_grads(x.double(), None if w is None else w.double(), None if b is None else b.double(), eps,
                            du.double())


# ------------------------------------------------------------------ kernels through the C ABI
@pytest.mark.parametrize("out_bf16", [False, True])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("rows,d", [(19, 102), (19, 100), (9, 64), (64, 256), (33, 2048), (40, 2176), (7, 4096),
                                    (1, 128), (8192, 2048)])
def test_ln_normalize_forward(K, rows, d, dtype, out_bf16):
    g = torch.Generator().manual_seed(rows + d)
    x0 = (torch.randn(rows, d, generator=g) * 2 + 0.5).to(dtype).cuda()
    x1 = (torch.randn(rows, d, generator=g) * 0.1 - 3).to(dtype).cuda()
    ln0 = make_ln(d, 1) + (1e-5,)
    ln1 = make_ln(d, 2, affine=(rows % 2 == 1)) + (1e-3,)
    out0, out1, st0, st1 = K.ln_normalize_pair(x0, x1, ln0, ln1, out_bf16=out_bf16)
    assert out0.dtype == (torch.bfloat16 if out_bf16 else torch.float32)
    for x, ln, out, st in ((x0, ln0, out0, st0), (x1, ln1, out1, st1)):
        u, mean, rstd, inv = ln_unit_reference(x, *ln)
        assert float((out.double() - u).abs().max()) < (2.0 ** -8 if out_bf16 else 3e-6)
        assert relerr(st[0], mean) < 1e-5 and relerr(st[1], rstd) < 1e-5 and relerr(st[2], inv) < 1e-5
    # one row set only
    o, none, s, none2 = K.ln_normalize_pair(x0, None, ln0, out_bf16=out_bf16)
    # (the alignment of the second output can select another variant for the pair: same values up to summation order)
    assert none is None and none2 is None and relerr(s, st0) < 1e-6
    assert float((o.double() - out0.double()).abs().max()) < (2.0 ** -8 if out_bf16 else 1e-6)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("rows,d", [(19, 102), (19, 100), (7, 64), (300, 1024), (1000, 2048), (37, 2176), (9, 4096),
                                    (6, 300), (8192, 2048)])
def test_ln_normalize_backward_gradient_in(K, rows, d, dtype):
    """Generic flavour: the gradient with respect to the unit rows comes in as one fp32 tensor."""
    g = torch.Generator().manual_seed(rows * 31 + d)
    x = (torch.randn(rows, d, generator=g) * 1.5 + 0.3).to(dtype).cuda()
    w, b = make_ln(d, 21)
    _, _, st, _ = K.ln_normalize_pair(x, None, (w, b, 1e-5))
    du = torch.randn(rows, d, generator=g).cuda()
    dx, _, dw, db, _, _, dt = K.ln_normalize_bwd_pair(x, None, (w, b, 0.0), None, st, None, du, None, want_dt=True)
    rdx, rdw, rdb, rdot = ln_bwd_reference(x, w, b, 1e-5, du.double())
    tol = {torch.float32: 2e-5, torch.bfloat16: 1.2e-2, torch.float16: 1.5e-3}[dtype]
    assert dx.dtype == dtype and relerr(dx, rdx) < tol
    assert relerr(dw, rdw) < 1e-4 and relerr(db, rdb) < 1e-4
    assert abs(float(dt) - float(rdot.sum())) < 1e-4 * float(rdot.abs().sum())
    # deterministic: a second call gives bit-identical results
    dx2, _, dw2, db2, _, _, dt2 = K.ln_normalize_bwd_pair(x, None, (w, b, 0.0), None, st, None, du, None, want_dt=True)
    assert torch.equal(dx, dx2) and torch.equal(dw, dw2) and torch.equal(db, db2) and torch.equal(dt, dt2)


@pytest.mark.parametrize("rows,d,n_slices,fused", [(64, 256, 1, False), (200, 256, 3, True), (128, 2048, 2, True),
                                                    (13, 102, 1, False), (13, 100, 4, True), (1024, 2048, 1, False)])
def test_ln_normalize_backward_dense_pair(K, rows, d, n_slices, fused):
    """Dense flavour, both heads in one launch: accumulator slices (+ gamma tau acc_scale for the fused kernel's
    unscaled sums) + positive-pair term + both Jacobians + LayerNorm backward; dt = sum of the image-side row dots."""
    g = torch.Generator().manual_seed(rows + 7 * d + n_slices)
    xs = [(torch.randn(rows, d, generator=g) * 2).cuda(), (torch.randn(rows, d, generator=g) - 1.0).cuda()]
    lns = [make_ln(d, 31), make_ln(d, 32)]
    u16, v16, st0, st1 = K.ln_normalize_pair(xs[0], xs[1], lns[0] + (1e-5,), lns[1] + (1e-5,), out_bf16=True)
    t, gamma = torch.tensor(1.1, device="cuda"), torch.tensor(0.6, device="cuda")
    gdiag = -torch.rand(rows, generator=g).cuda()
    accs = [(torch.randn(n_slices, rows, d, generator=g) * (1.0 if fused else 1e-3)).cuda() for _ in range(2)]
    acc_scale = 1.0 / (rows * (rows - 1)) if fused else 0.0
    dx0, dx1, dw0, db0, dw1, db1, dt = K.ln_normalize_bwd_pair(
        xs[0], xs[1], lns[0] + (0.0,), lns[1] + (0.0,), st0, st1, accs[0], accs[1], acc_scale=acc_scale, partner0=v16,
        partner1=u16, gdiag=gdiag, t=t, gamma=gamma, m_rows=rows, want_dt=True)
    c = float(gamma) * float(t.exp()) / rows
    scale = float(gamma) * float(t.exp()) * acc_scale if fused else 1.0
    got = ((dx0, dw0, db0), (dx1, dw1, db1))
    for j, (x, (w, b), partner) in enumerate(zip(xs, lns, (v16, u16))):
        du = accs[j].double().sum(0) * scale + c * gdiag.double()[:, None] * partner.double()
        rdx, rdw, rdb, rdot = ln_bwd_reference(x, w, b, 1e-5, du)
        assert relerr(got[j][0], rdx) < 5e-5, j
        assert relerr(got[j][1], rdw) < 1e-4 and relerr(got[j][2], rdb) < 1e-4, j
        if j == 0:
            assert abs(float(dt) - float(rdot.sum())) < 1e-4 * float(rdot.abs().sum())


def test_bad_arguments_fail_loudly(K):
    from clip_lite_b200._lib import JSDLibraryError
    x = torch.randn(4, 4104, device="cuda")
    _, _, st, _ = K.ln_normalize_pair(x, None, (None, None, 1e-5))          # forward: any D
    with pytest.raises(JSDLibraryError):
        K.ln_normalize_bwd_pair(x, None, (None, None, 0.0), None, st, None, torch.randn(4, 4104, device="cuda"), None)
    with pytest.raises(RuntimeError):
        K.ln_normalize_pair(torch.randn(4, 8), None, (None, None, 1e-5))    # CPU tensors: no fallback


# ------------------------------------------------------------------ autograd entry points
def make_heads(b, d, seed, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    xf = (torch.randn(b, d, generator=g) * 1.5 + 0.2).to(dtype).cuda()
    xg = (0.5 * xf.float().cpu() + torch.randn(b, d, generator=g)).to(dtype).cuda()
    lns = []
    for _ in range(2):
        ln = torch.nn.LayerNorm(d)
        with torch.no_grad():
            ln.weight.copy_(1.0 + 0.3 * torch.randn(d, generator=g))
            ln.bias.copy_(0.2 * torch.randn(d, generator=g))
        lns.append(ln.cuda())
    return xf, xg, lns[0], lns[1]


def reference(xf, xg, ln_f, ln_g, t, estimator, gamma=1.0, **kw):
    """fp64 autograd of LayerNorm -> estimator on the device (the oracle normalises internally, as loss.py:94-95)."""
    leaves = [xf.detach().double().clone().requires_grad_(True), xg.detach().double().clone().requires_grad_(True)]
    params = [p.detach().double().clone().requires_grad_(True) for p in (ln_f.weight, ln_f.bias, ln_g.weight, ln_g.bias)]
    tt = torch.tensor(float(t), dtype=torch.float64, device=xf.device, requires_grad=True)
    d = xf.shape[1]
    f = torch.nn.functional.layer_norm(leaves[0], (d,), params[0], params[1], ln_f.eps)
    g = torch.nn.functional.layer_norm(leaves[1], (d,), params[2], params[3], ln_g.eps)
    loss = estimator(f, g, tt, **kw)["loss"]
    (gamma * loss).backward()
    return loss.detach(), [x.grad for x in leaves], [p.grad for p in params], tt.grad


@pytest.mark.parametrize("mode", ["shift1", "cluster"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("b,d", [(12, 256), (96, 100), (512, 2048)])
def test_index_mode_with_fused_tail(ops, b, d, dtype, mode):
    xf, xg, ln_f, ln_g = make_heads(b, d, seed=4 + b, dtype=dtype)
    xf.requires_grad_(True)
    xg.requires_grad_(True)
    t = torch.tensor(orc.T_INIT, device="cuda", requires_grad=True)
    neg = ops.NegativeIndex.cluster(b // 2) if mode == "cluster" else None
    f, g = ops.ln_normalize_pair(xf, xg, ln_f, ln_g)
    loss, _ = ops.jsd_index_loss(f, g, t, neg)
    (0.7 * loss).backward()
    kw = {"neg_index": orc.cluster_index(b // 2, device="cuda")} if mode == "cluster" else {}
    rl, rx, rp, rt = reference(xf, xg, ln_f, ln_g, orc.T_INIT, orc.jsd_index, gamma=0.7, **kw)
    tol = 1e-4 if dtype == torch.float32 else GRAD_RTOL
    assert relerr(loss, rl) < 1e-5
    assert xf.grad.dtype == dtype
    assert relerr(xf.grad, rx[0]) < tol and relerr(xg.grad, rx[1]) < tol
    for p, r in zip((ln_f.weight, ln_f.bias, ln_g.weight, ln_g.bias), rp):
        assert relerr(p.grad, r) < 2e-4
    assert abs(float(t.grad) - float(rt)) < 1e-3 * max(abs(float(rt)), 1e-2)       # a sum with cancellation


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("b,d", [(256, 128), (1000, 256), (200, 72), (512, 1024), (1024, 2048)])
def test_dense_loss_with_fused_tail(ops, b, d, dtype):
    """jsd_dense_loss_ln (fused single-pass kernel at D <= 256, staged tensor-core path above) against fp64
    LayerNorm -> dense oracle, and against the unfused route jsd_dense_loss(LayerNorm(x))."""
    xf, xg, ln_f, ln_g = make_heads(b, d, seed=9 + d, dtype=dtype)
    xf.requires_grad_(True)
    xg.requires_grad_(True)
    t = torch.tensor(1.3, device="cuda", requires_grad=True)
    loss, stats = ops.jsd_dense_loss_ln(xf, xg, ln_f, ln_g, t)
    (0.9 * loss).backward()
    rl, rx, rp, rt = reference(xf, xg, ln_f, ln_g, 1.3, orc.jsd_dense, gamma=0.9)
    assert relerr(loss, rl) < LOSS_RTOL
    assert relerr(xf.grad, rx[0]) < GRAD_RTOL and relerr(xg.grad, rx[1]) < GRAD_RTOL
    for p, r in zip((ln_f.weight, ln_f.bias, ln_g.weight, ln_g.bias), rp):
        assert relerr(p.grad, r) < GRAD_RTOL
    assert relerr(t.grad, rt) < GRAD_RTOL
    assert stats.shape == (4,) and relerr(stats[2], rl) < LOSS_RTOL
    if dtype != torch.float32:
        return
    # the unfused route on the same inputs
    fused = [xf.grad.clone(), xg.grad.clone(), ln_f.weight.grad.clone(), ln_g.bias.grad.clone(), t.grad.clone()]
    for p in (xf, xg, ln_f.weight, ln_f.bias, ln_g.weight, ln_g.bias, t):
        p.grad = None
    loss2, _ = ops.jsd_dense_loss(ln_f(xf.float()), ln_g(xg.float()), t)
    (0.9 * loss2).backward()
    assert relerr(loss, loss2) < LOSS_RTOL
    for a, p in zip(fused, (xf, xg, ln_f.weight, ln_g.bias, t)):
        assert relerr(a, p.grad) < GRAD_RTOL


def test_dense_loss_with_fused_tail_no_grad(ops):
    xf, xg, ln_f, ln_g = make_heads(64, 128, seed=3)
    with torch.no_grad():
        loss, _ = ops.jsd_dense_loss_ln(xf, xg, ln_f, ln_g, torch.tensor(1.0, device="cuda"))
    ref = orc.jsd_dense(ln_f(xf).double(), ln_g(xg).double(), 1.0)["loss"]
    assert relerr(loss, ref) < LOSS_RTOL and not loss.requires_grad


# ------------------------------------------------------------------ the drop-in module with fused_heads=True
@pytest.mark.parametrize("case", ["module_b8_train", "module_b8_eval", "module_cluster_b6_train"])
def test_module_fused_heads_matches_reference_golden(L, golden_dir, case):
    """Same comparison as tests/test_gpu_module.py::test_module_matches_reference_golden, with the fused tail: loss
    values, input gradients, every parameter gradient (LayerNorm weight / bias included) and the BatchNorm buffers
    of the unmodified reference module."""
    z = np.load(os.path.join(golden_dir, case + ".npz"), allow_pickle=True)
    m = L.JSDInfoMaxLoss(image_dim=int(z["image_dim"]), text_dim=int(z["text_dim"]), type="dot",
                         image_prior=False, text_prior=False, fused_heads=True)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict(seeded_state_dict(shapes, int(z["seed"])))
    m.cuda().train(bool(z["train"]))
    names = [k[3:] for k in z.files if k.startswith("in_")]
    leaves = {k: torch.from_numpy(z["in_" + k]).cuda().requires_grad_(True) for k in names}
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        out = m(**leaves)
        out["total_loss"].backward()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    for k in out:
        assert abs(float(out[k]) - float(z["out_" + k])) <= LOSS_RTOL * max(abs(float(z["out_" + k])), 1e-30), k
    for k in names:
        assert relerr(leaves[k].grad, z["grad_" + k]) < GRAD_RTOL, k
    for k, p in m.named_parameters():
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        w = torch.from_numpy(np.random.RandomState(7).standard_normal(tuple(g.shape) or (1,))).reshape(g.shape)
        proj = float((g.double().cpu() * w).sum())
        scale = max(float(z["pgrad_abs/" + k]), 1e-30)
        assert abs(proj - float(z["pgrad_proj/" + k])) < GRAD_RTOL * scale, k
        assert abs(float(g.sum()) - float(z["pgrad_sum/" + k])) < GRAD_RTOL * scale, k
    sd = m.state_dict()
    for k in z.files:
        if k.startswith("buf/"):
            assert np.allclose(sd[k[4:]].cpu().numpy(), z[k], rtol=1e-4, atol=1e-5), k


@pytest.mark.parametrize("neg_mode", ["shift1", "dense"])
def test_module_fused_heads_equals_default_route(L, neg_mode):
    """Both estimator modes, priors and SSL critics on: fused_heads=True and the default module give the same
    losses and gradients (same weights, same inputs, same RNG for the prior noise)."""
    def run(fused):
        torch.manual_seed(0)
        m = L.JSDInfoMaxLoss(image_dim=64, text_dim=48, image_prior=True, text_prior=True,
                             visual_self_supervised=True, textual_self_supervised=True, neg_mode=neg_mode,
                             fused_heads=fused).cuda()
        g = torch.Generator(device="cuda").manual_seed(1)
        leaves = {k: torch.randn(32, dim, device="cuda", generator=g).requires_grad_(True)
                  for k, dim in (("image_features", 64), ("text_features", 48), ("aug_image_features", 64),
                                 ("aug_text_features", 48))}
        torch.manual_seed(2)
        out = m(**leaves)
        out["total_loss"].backward()
        return m, leaves, out

    m1, l1, o1 = run(True)
    m0, l0, o0 = run(False)
    for k in o0:
        assert abs(float(o1[k]) - float(o0[k])) <= LOSS_RTOL * max(abs(float(o0[k])), 1e-30), k
    for k in l0:
        assert relerr(l1[k].grad, l0[k].grad) < GRAD_RTOL, k
    for (k, p1), (_, p0) in zip(m1.named_parameters(), m0.named_parameters()):
        if p0.grad is None:
            assert p1.grad is None or float(p1.grad.abs().max()) == 0.0, k
        else:
            assert relerr(p1.grad, p0.grad) < GRAD_RTOL, k


def test_module_fused_heads_under_autocast_and_grad_scaler(L):
    """train.py:214-225 shape: the heads' GEMMs run in fp16 under autocast, the tail reads their fp16 output and
    hands fp32 unit rows to the estimator; gradients arrive in the parameters' dtype."""
    m = L.JSDInfoMaxLoss(image_dim=64, text_dim=48, image_prior=True, text_prior=True, fused_heads=True).cuda()
    img = torch.randn(32, 64, device="cuda", requires_grad=True)
    txt = torch.randn(32, 48, device="cuda", requires_grad=True)
    with torch.autocast("cuda", dtype=torch.float16):
        out = m(image_features=img, text_features=txt)
    (out["total_loss"] * 1024.0).backward()
    assert torch.isfinite(img.grad).all() and torch.isfinite(txt.grad).all()
    ln = m.global_d.img_block.feature_block_ln
    assert ln.weight.grad is not None and torch.isfinite(ln.weight.grad).all() and float(ln.weight.grad.abs().max()) > 0
    assert all(p.grad is None or p.grad.dtype == p.dtype for p in m.parameters())


@pytest.mark.parametrize("fused", [False, True])
def test_heads_dtype_bf16(L, fused):
    """heads_dtype=torch.bfloat16 (no outer autocast, no GradScaler): the heads' GEMMs run as bf16 library GEMMs,
    the tail / estimator read their bf16 output; losses stay within bf16 rounding of the fp32 module."""
    def run(dtype):
        torch.manual_seed(0)
        m = L.JSDInfoMaxLoss(image_dim=64, text_dim=48, image_prior=False, fused_heads=fused, heads_dtype=dtype).cuda()
        g = torch.Generator(device="cuda").manual_seed(1)
        img = torch.randn(64, 64, device="cuda", generator=g).requires_grad_(True)
        txt = torch.randn(64, 48, device="cuda", generator=g).requires_grad_(True)
        out = m(img, txt)
        out["total_loss"].backward()
        return m, out, img.grad

    m16, o16, g16 = run(torch.bfloat16)
    m32, o32, g32 = run(None)
    assert abs(float(o16["total_loss"]) - float(o32["total_loss"])) < 3e-2 * abs(float(o32["total_loss"]))
    assert g16.dtype == torch.float32 and torch.isfinite(g16).all() and relerr(g16, g32) < 0.2
    assert all(p.grad is None or (p.grad.dtype == p.dtype and torch.isfinite(p.grad).all()) for p in m16.parameters())


@pytest.mark.parametrize("mode,b,d", [("dense", 512, 128), ("dense", 256, 2048), ("index", 512, 2048)])
def test_fused_tail_step_cuda_graph_replay_matches_eager(ops, mode, b, d):
    """The fused-tail step is graph-capturable (no host sync, no allocation outside the graph's pool, the occupancy
    query is not a stream operation) and a replay is bit-identical to the eager step on the same inputs."""
    from clip_lite_b200.graph import GraphedStep
    xf, xg, ln_f, ln_g = make_heads(b, d, seed=7)
    t = torch.tensor(orc.T_INIT, device="cuda", requires_grad=True)

    def loss_fn(a, c, tt):
        if mode == "dense":
            return ops.jsd_dense_loss_ln(a, c, ln_f, ln_g, tt)
        f, g = ops.ln_normalize_pair(a, c, ln_f, ln_g)
        return ops.jsd_index_loss(f, g, tt)

    gs = GraphedStep(loss_fn, xf, xg, t)
    for seed in (7, 8):
        x2, y2, _, _ = make_heads(b, d, seed=seed)
        loss, df, dg, dt = gs(x2, y2)
        xl, yl = x2.clone().requires_grad_(True), y2.clone().requires_grad_(True)
        ref_loss = loss_fn(xl, yl, t)[0]
        rdf, rdg, rdt = torch.autograd.grad(ref_loss, (xl, yl, t))
        assert torch.equal(loss, ref_loss) and torch.equal(df, rdf) and torch.equal(dg, rdg) and torch.equal(dt, rdt)

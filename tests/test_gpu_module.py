"""GPU parity of the drop-in module and the autograd entry points: against the golden
vectors generated from the unmodified reference (full module with projection heads,
cluster mode, SSL call sites), against the oracle, and -- at BASELINE.json's full sizes
-- through size-independent properties."""
import glob
import os

import numpy as np
import pytest
import torch

from _weights import seeded_state_dict
from oracle import jsd_oracle as orc

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-3     # BASELINE.json
GRAD_RTOL = 1e-2     # BASELINE.json


def relerr(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.fixture(scope="module")
def L():
    from clip_lite_b200 import loss
    return loss


@pytest.fixture(scope="module")
def ops():
    from clip_lite_b200 import ops
    return ops


# ------------------------------------------------------------------ golden: full module
@pytest.mark.parametrize("case", ["module_b8_train", "module_b8_eval", "module_cluster_b6_train"])
def test_module_matches_reference_golden(L, golden_dir, case):
    z = np.load(os.path.join(golden_dir, case + ".npz"), allow_pickle=True)
    m = L.JSDInfoMaxLoss(image_dim=int(z["image_dim"]), text_dim=int(z["text_dim"]), type="dot",
                         image_prior=False, text_prior=False)
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
    assert set(out) == {"total_loss", "cross_modal_loss", "visual_loss", "textual_loss"}
    for k in out:
        assert out[k].dim() == 0 and out[k].is_cuda
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


# ------------------------------------------------------------------ golden: estimator through the module (incl. SSL)
def test_module_estimator_cases_match_reference_golden(L, golden_dir):
    for path in sorted(glob.glob(os.path.join(golden_dir, "estimator_*.npz"))):
        z = np.load(path, allow_pickle=True)
        ssl = bool(z["ssl"])
        d = z["in_image_features"].shape[1]
        m = L.JSDInfoMaxLoss(image_dim=d, text_dim=d, type="dot", image_prior=False, text_prior=False,
                             visual_self_supervised=ssl, textual_self_supervised=ssl)
        critics = [m.global_d] + ([m.visual_d, m.textual_d] if ssl else [])
        for c in critics:
            c.img_block = torch.nn.Identity()
            c.text_block = torch.nn.Identity()
        m.cuda()
        m.global_d.temperature.data.fill_(float(z["t"]))
        if ssl:
            m.visual_d.temperature.data.fill_(float(z["t_ssl"]))
            m.textual_d.temperature.data.fill_(float(z["t_ssl"]))
        names = [k[3:] for k in z.files if k.startswith("in_")]
        leaves = {k: torch.from_numpy(z["in_" + k]).cuda().requires_grad_(True) for k in names}
        out = m(**leaves)
        out["total_loss"].backward()
        for k in out:
            ref = float(z["out_" + k])
            assert abs(float(out[k]) - ref) <= LOSS_RTOL * max(abs(ref), 1e-30), (path, k)
        for k in names:
            assert relerr(leaves[k].grad, z["grad_" + k]) < 1e-4, (path, k)
        assert relerr(m.global_d.temperature.grad, z["grad_temperature"]) < 1e-4, path
        if ssl:
            assert relerr(m.visual_d.temperature.grad, z["grad_temperature_visual"]) < 1e-4
            assert relerr(m.textual_d.temperature.grad, z["grad_temperature_textual"]) < 1e-4


# ------------------------------------------------------------------ autograd entry points vs oracle
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("b,d", [(32, 64), (96, 72), (1024, 128), (1024, 1024)])
def test_dense_autograd_vs_oracle(ops, dtype, b, d):
    f, g = orc.synth_embeddings(b, d, seed=1, correlated=True)
    f, g = f.to(dtype), g.to(dtype)
    fl, gl = f.cuda().requires_grad_(True), g.cuda().requires_grad_(True)
    t = torch.tensor(orc.T_INIT, device="cuda", requires_grad=True)
    loss, stats = ops.jsd_dense_loss(fl, gl, t)
    # fp16 gradients of a mean over B^2 pairs underflow without loss scaling (train.py uses a GradScaler)
    gamma = 0.9 * (65536.0 if dtype == torch.float16 else 1.0)
    (gamma * loss).backward()
    ref = orc.jsd_dense(f.double(), g.double(), orc.T_INIT)
    df, dg, dt = orc.jsd_dense_grads(f.double(), g.double(), orc.T_INIT, gamma=gamma)
    assert relerr(loss, ref["loss"]) < LOSS_RTOL
    assert fl.grad.dtype == dtype and gl.grad.dtype == dtype
    assert relerr(fl.grad, df) < GRAD_RTOL and relerr(gl.grad, dg) < GRAD_RTOL
    assert relerr(t.grad, dt) < GRAD_RTOL
    assert relerr(stats[0] + stats[1], ref["loss"]) < LOSS_RTOL


def test_dense_equals_mean_over_shifts_of_the_reference_estimator(ops):
    """Ties the tensor-core path to the reference semantics: the dense negative term is the
    mean over k = 1..B-1 of the reference's roll-by-k negative term (index kernel, golden-pinned)."""
    b, d = 48, 64
    f, g = orc.synth_embeddings(b, d, seed=2, correlated=True)
    f, g = f.cuda(), g.cuda()
    t = torch.tensor(orc.T_INIT, device="cuda")
    _, dense = ops.jsd_dense_loss(f, g, t)
    negs = []
    for k in range(1, b):
        ni = ops.NegativeIndex((torch.arange(b) + k) % b)
        negs.append(ops.jsd_index_loss(f, g, t, ni)[1][1])
    assert relerr(dense[1], torch.stack(negs).mean()) < LOSS_RTOL
    assert relerr(dense[0], ops.jsd_index_loss(f, g, t)[1][0]) < LOSS_RTOL


def test_cluster_batches_with_dense_negatives(L):
    """neg_mode="dense" on a cluster-mode call (loss.py:225-252 inputs): the estimator runs over the concatenated
    2B' rows with EVERY other text row -- the B' hard negatives included -- as a negative of each image row
    ("dense + hard-negative columns").  Checked against oracle.jsd_dense on the concatenation."""
    half, d = 96, 64
    gen = torch.Generator().manual_seed(11)
    img, txt, nimg, ntxt = (torch.randn(half, d, generator=gen) for _ in range(4))
    m = L.JSDInfoMaxLoss(image_dim=d, text_dim=d, type="dot", image_prior=False, text_prior=False,
                         neg_mode="dense").cuda()
    m.global_d.img_block = torch.nn.Identity()
    m.global_d.text_block = torch.nn.Identity()
    leaves = [x.cuda().requires_grad_(True) for x in (img, txt, nimg, ntxt)]
    out = m(image_features=leaves[0], text_features=leaves[1], neg_image_features=leaves[2],
            neg_text_features=leaves[3])
    out["total_loss"].backward()
    f_all, g_all = torch.cat((img, nimg)).double(), torch.cat((txt, ntxt)).double()
    ref = orc.jsd_dense(f_all, g_all, orc.T_INIT)
    rdf, rdg, rdt = orc.jsd_dense_grads(f_all, g_all, orc.T_INIT, gamma=0.9)
    assert relerr(out["cross_modal_loss"], ref["loss"]) < LOSS_RTOL
    assert relerr(torch.cat((leaves[0].grad, leaves[2].grad)), rdf) < GRAD_RTOL
    assert relerr(torch.cat((leaves[1].grad, leaves[3].grad)), rdg) < GRAD_RTOL
    assert relerr(m.global_d.temperature.grad, rdt) < GRAD_RTOL


def test_no_grad_forward_and_module_eval(L):
    m = L.JSDInfoMaxLoss(image_dim=32, text_dim=24, image_prior=True, text_prior=True, neg_mode="dense").cuda().eval()
    with torch.no_grad():
        out = m(torch.randn(16, 32, device="cuda"), torch.randn(16, 24, device="cuda"))
    assert all(torch.isfinite(v) for v in out.values()) and not out["total_loss"].requires_grad


def test_module_under_autocast_and_grad_scaler(L):
    """train.py:214-225 shape: autocast forward, scaled backward; grads arrive in the parameters' dtype."""
    m = L.JSDInfoMaxLoss(image_dim=64, text_dim=48, image_prior=True, text_prior=True).cuda()
    img = torch.randn(32, 64, device="cuda", requires_grad=True)
    txt = torch.randn(32, 48, device="cuda", requires_grad=True)
    with torch.autocast("cuda", dtype=torch.float16):
        out = m(image_features=img, text_features=txt)
    (out["total_loss"] * 1024.0).backward()
    assert torch.isfinite(img.grad).all() and torch.isfinite(txt.grad).all()
    assert m.global_d.temperature.grad is not None and torch.isfinite(m.global_d.temperature.grad)
    assert all(p.grad is None or p.grad.dtype == p.dtype for p in m.parameters())


def test_gathered_loss_single_process_equals_dense(ops):
    from clip_lite_b200 import parallel
    f, g = orc.synth_embeddings(256, 128, seed=4, correlated=True)
    outs = []
    for fn in (ops.jsd_dense_loss, parallel.gathered_dense_loss):
        fl, gl = f.cuda().requires_grad_(True), g.cuda().requires_grad_(True)
        t = torch.tensor(orc.T_INIT, device="cuda", requires_grad=True)
        loss, _ = fn(fl, gl, t)
        loss.backward()
        outs.append((loss.detach(), fl.grad, gl.grad, t.grad))
    for a, b in zip(*outs):
        assert relerr(a, b) < 1e-6


# ------------------------------------------------------------------ BASELINE sizes: properties + oracle on the GPU
@pytest.mark.parametrize("b,d", [(8192, 1024), (4096, 512)])
def test_full_size_dense_properties(ops, b, d):
    f, g = orc.synth_embeddings(b, d, seed=0, correlated=True)
    f, g = f.cuda(), g.cuda()

    def run(ff, gg, gamma):
        fl, gl = ff.clone().requires_grad_(True), gg.clone().requires_grad_(True)
        t = torch.tensor(orc.T_INIT, device="cuda", requires_grad=True)
        loss, _ = ops.jsd_dense_loss(fl, gl, t)
        (gamma * loss).backward()
        return loss.detach(), fl.grad, gl.grad, t.grad

    loss, df, dg, dt = run(f, g, 1.0)
    # (1) oracle restatement evaluated on the GPU in fp64 (checker only)
    ref = orc.jsd_dense(f.double(), g.double(), orc.T_INIT)
    rdf, rdg, rdt = orc.jsd_dense_grads(f.double(), g.double(), orc.T_INIT)
    assert relerr(loss, ref["loss"]) < LOSS_RTOL
    assert relerr(df, rdf) < GRAD_RTOL and relerr(dg, rdg) < GRAD_RTOL and relerr(dt, rdt) < GRAD_RTOL
    # (2) the gradient is orthogonal to each input row (Jacobian of the L2 normalisation)
    assert float(((f * df).sum(-1).abs().max())) < 1e-3 * float(df.abs().max() * f.norm(dim=-1).max())
    assert float(((g * dg).sum(-1).abs().max())) < 1e-3 * float(dg.abs().max() * g.norm(dim=-1).max())
    # (3) linear in the upstream gradient
    _, df2, dg2, dt2 = run(f, g, 3.0)
    assert relerr(df2, 3.0 * df) < 1e-5 and relerr(dg2, 3.0 * dg) < 1e-5 and relerr(dt2, 3.0 * dt) < 1e-5
    # (4) invariant to the scale of each row, equivariant to a joint row permutation
    perm = torch.randperm(b, device="cuda")
    loss_p, df_p, dg_p, _ = run(2.5 * f[perm], 0.5 * g[perm], 1.0)
    assert relerr(loss_p, loss) < 1e-5
    # (re-normalising scaled rows can flip the bf16 rounding of a few operand elements)
    assert relerr(2.5 * df_p, df[perm]) < GRAD_RTOL and relerr(0.5 * dg_p, dg[perm]) < GRAD_RTOL
    # (5) deterministic
    loss_b, df_b, _, _ = run(f, g, 1.0)
    assert torch.equal(loss_b, loss) and torch.equal(df_b, df)


def test_full_size_index_mode_vs_oracle(ops):
    b, d = 8192, 2048
    f, g = orc.synth_embeddings(b, d, seed=1, correlated=True)
    fl, gl = f.cuda().requires_grad_(True), g.cuda().requires_grad_(True)
    t = torch.tensor(orc.T_INIT, device="cuda", requires_grad=True)
    loss, _ = ops.jsd_index_loss(fl, gl, t)
    loss.backward()
    ref = orc.jsd_index(f.cuda().double(), g.cuda().double(), orc.T_INIT)
    rdf, rdg, rdt = orc.jsd_index_grads(f.cuda().double(), g.cuda().double(), orc.T_INIT)
    assert relerr(loss, ref["loss"]) < 1e-5
    assert relerr(fl.grad, rdf) < 1e-4 and relerr(gl.grad, rdg) < 1e-4 and relerr(t.grad, rdt) < 1e-4


def test_cuda_graph_replay_matches_eager(ops):
    from clip_lite_b200.graph import GraphedStep
    f, g = orc.synth_embeddings(512, 128, seed=7, correlated=True)
    f, g = f.cuda(), g.cuda()
    t = torch.tensor(orc.T_INIT, device="cuda", requires_grad=True)
    gs = GraphedStep(lambda a, b, tt: ops.jsd_dense_loss(a, b, tt), f, g, t)
    for seed in (7, 8):
        f2, g2 = orc.synth_embeddings(512, 128, seed=seed, correlated=True)
        loss, df, dg, dt = gs(f2.cuda(), g2.cuda())
        fl, gl = f2.cuda().requires_grad_(True), g2.cuda().requires_grad_(True)
        ref_loss, _ = ops.jsd_dense_loss(fl, gl, t)
        rdf, rdg, rdt = torch.autograd.grad(ref_loss, (fl, gl, t))
        assert torch.equal(loss, ref_loss) and torch.equal(df, rdf) and torch.equal(dg, rdg) and torch.equal(dt, rdt)


def test_stress_slab_65536_columns():
    """BASELINE configs[3] shape per rank: an 8192-row slab against 65536 text rows, D = 512 (the 17 GB
    fp32 score matrix is never formed; Gmat is 1 GB of bf16).  Checked against the oracle evaluated in
    row blocks on the GPU, plus the gradient-orthogonality property."""
    from clip_lite_b200 import kernels as K
    m, n, d, off = 8192, 65536, 512, 3 * 8192
    gen = torch.Generator("cuda").manual_seed(0)
    g = torch.randn(n, d, device="cuda", generator=gen)
    f = 0.6 * g[off:off + m] + 0.8 * torch.randn(m, d, device="cuda", generator=gen)
    t = torch.tensor(orc.T_INIT, device="cuda")
    gamma = torch.tensor(1.0, device="cuda")
    u, _, inv_f, _ = K.normalize_cast_pair(f, g[off:off + m].contiguous())
    v, inv_g = K.normalize_cast(g)
    out4, loss, gmat, gdiag = K.dense_fwd(u, v, t, row_offset=off)
    df, dt = K.dense_backward_image_side(f, v, inv_f, gmat, gdiag, t, gamma, off)
    dv = K.dense_bwd_dv(gmat, u, n, t, gamma)
    # oracle in fp64, 1024 rows at a time
    pos = neg = 0.0
    rdf = torch.empty(m, d, dtype=torch.float64, device="cuda")
    rdv = torch.zeros(n, d, dtype=torch.float64, device="cuda")
    un, vn = u.double(), v.double()
    tau = float(np.exp(orc.T_INIT))
    for r0 in range(0, m, 1024):
        ref = orc.dense_from_unit(un[r0:r0 + 1024], vn, orc.T_INIT, row_offset=off + r0)
        pos += float(ref["pos"]) * 1024 / m
        neg += float(ref["neg"]) * 1024 / m
        rdv += ref["dv_acc"] * (1024 / m)          # dense_from_unit scales by its own slab height
        rdf[r0:r0 + 1024] = ref["du"] * (1024 / m)
    assert relerr(out4[0], torch.tensor(pos)) < LOSS_RTOL and relerr(out4[1], torch.tensor(neg)) < LOSS_RTOL
    assert relerr(dv, rdv) < GRAD_RTOL
    # image-side gradient: project the fp64 dU through the normalisation and compare
    uu = f.double() * inv_f.double()[:, None]
    ref_df = (rdf - uu * (uu * rdf).sum(-1, keepdim=True)) * inv_f.double()[:, None]
    assert relerr(df, ref_df) < GRAD_RTOL
    assert float((f * df).sum(-1).abs().max()) < 1e-3 * float(df.abs().max() * f.norm(dim=-1).max())
    assert torch.isfinite(dt)


def test_graphed_gathered_step_single_process_matches_eager(ops):
    """The segmented CUDA-graph step of the data-parallel path (world size 1 here) equals the eager autograd path."""
    from clip_lite_b200 import parallel
    f, g = orc.synth_embeddings(512, 128, seed=9, correlated=True)
    f, g = f.cuda(), g.cuda()
    t = torch.tensor(orc.T_INIT, device="cuda", requires_grad=True)
    gs = parallel.GraphedGatheredStep(f, g, t)
    for seed in (9, 10):
        f2, g2 = orc.synth_embeddings(512, 128, seed=seed, correlated=True)
        loss, df, dg, dt = gs(f2.cuda(), g2.cuda())
        fl, gl = f2.cuda().requires_grad_(True), g2.cuda().requires_grad_(True)
        ref_loss, _ = parallel.gathered_dense_loss(fl, gl, t)
        rdf, rdg, rdt = torch.autograd.grad(ref_loss, (fl, gl, t))
        assert torch.equal(loss, ref_loss) and torch.equal(df, rdf) and torch.equal(dg, rdg) and torch.equal(dt, rdt)

"""GPU parity tests of every C-ABI entry point against the oracle (oracle/ is the
checker only).  Tolerances: BASELINE.json -- loss <= 1e-3 relative, gradients
<= 1e-2 relative (max-norm per tensor) against the fp32/fp64 restatement; the
stage-level checks on identical (already rounded) operands are held much tighter."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import jsd_oracle as orc

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-3
GRAD_RTOL = 1e-2
T0 = orc.T_INIT


def relerr(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.fixture(scope="module")
def K():
    from clip_lite_b200 import kernels
    return kernels


def dev_t(t=T0):
    return torch.tensor(t, dtype=torch.float32, device="cuda")


# ------------------------------------------------------------------ rowwise kernels
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("rows,d", [(7, 8), (64, 128), (100, 1000), (33, 2048), (5, 30)])
def test_normalize_cast(K, dtype, rows, d):
    x = (torch.randn(rows, d, device="cuda") * 3).to(dtype)
    xn, inv = K.normalize_cast(x)
    ref, n = orc.l2_normalize(x.double())
    assert relerr(inv, 1.0 / n.squeeze(-1)) < 1e-5
    assert (xn.double() - ref).abs().max() < 2 ** -8


def test_normalize_zero_row(K):
    x = torch.zeros(4, 16, device="cuda")
    x[1] = 1.0
    xn, inv = K.normalize_cast(x)
    assert torch.isfinite(xn.float()).all() and (xn[0] == 0).all()
    assert float(inv[0]) == pytest.approx(1e12, rel=1e-5)       # 1 / eps, as F.normalize


# ------------------------------------------------------------------ index mode (reference semantics)
def _csr_inverse(neg_index, n):
    order = torch.argsort(neg_index, stable=True)
    counts = torch.bincount(neg_index, minlength=n)
    ptr = torch.zeros(n + 1, dtype=torch.int64)
    ptr[1:] = torch.cumsum(counts, 0)
    return ptr.int(), order.int()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("b,d", [(1, 8), (2, 8), (3, 5), (16, 128), (128, 100), (64, 2048), (1024, 256)])
@pytest.mark.parametrize("mode", ["shift1", "cluster", "random"])
def test_index_vs_oracle(K, dtype, b, d, mode):
    if mode == "cluster" and (b % 2 or b < 2):
        pytest.skip("cluster mode needs an even batch")
    f, g = orc.synth_embeddings(b, d, seed=b + d, correlated=True)
    f, g = f.to(dtype), g.to(dtype)
    neg = None
    if mode == "cluster":
        neg = orc.cluster_index(b // 2)
    elif mode == "random":
        neg = torch.randint(0, b, (b,), generator=torch.Generator().manual_seed(1))   # not a permutation
    args = {}
    if neg is not None:
        ptr, idx = _csr_inverse(neg, b)
        args = dict(neg_index=neg.int().cuda(), inv_ptr=ptr.cuda(), inv_idx=idx.cuda())
    out4, loss, df, dg, gs = K.index_fwd_bwd(f.cuda(), g.cuda(), dev_t(), **args)
    df, dg = df.float() / gs, dg.float() / gs     # stored gradients carry grad_scale (1 by default)
    assert torch.equal(loss, out4[2])
    # the upstream gradient as a device scalar and a host-side scale: 0.25 * 4 == 1, powers of two are exact
    _, _, df2, dg2, gs2 = K.index_fwd_bwd(f.cuda(), g.cuda(), dev_t(), grad_scale=4.0,
                                          gamma=torch.tensor(0.25, device="cuda"), **args)
    assert gs2 == 4.0 and torch.equal(df2.float(), df * gs) and torch.equal(dg2.float(), dg * gs)
    # forward only (eval / no_grad): same loss, no gradient buffers
    o4, l2, n1, n2, _ = K.index_fwd_bwd(f.cuda(), g.cuda(), dev_t(), want_grad=False, **args)
    assert n1 is None and n2 is None and torch.equal(o4, out4)
    ref = orc.jsd_index(f.double(), g.double(), T0, neg)
    rdf, rdg, rdt = orc.jsd_index_grads(f.double(), g.double(), T0, neg)
    assert relerr(out4[0], ref["pos"]) < 1e-5
    assert relerr(out4[1], ref["neg"]) < 1e-5
    assert relerr(out4[2], ref["loss"]) < 1e-5
    assert relerr(out4[3], rdt) < 1e-4
    gtol = 1e-4 if dtype == torch.float32 else 1e-2
    assert relerr(df, rdf) < gtol
    assert relerr(dg, rdg) < gtol


def test_index_vs_reference_golden(K, golden_dir):
    """Identity-head estimator cases generated from the unmodified reference loss.py."""
    files = sorted(glob.glob(os.path.join(golden_dir, "estimator_*.npz")))
    assert files
    for path in files:
        z = np.load(path, allow_pickle=True)
        if bool(z["ssl"]):
            continue                     # covered through the module in test_gpu_module.py
        f = torch.from_numpy(z["in_image_features"])
        g = torch.from_numpy(z["in_text_features"])
        neg = None
        gf, gg = z["grad_image_features"], z["grad_text_features"]
        if str(z["mode"]) == "cluster":
            half = f.shape[0]
            f = torch.cat((f, torch.from_numpy(z["in_neg_image_features"])))
            g = torch.cat((g, torch.from_numpy(z["in_neg_text_features"])))
            gf = np.concatenate((gf, z["grad_neg_image_features"]))
            gg = np.concatenate((gg, z["grad_neg_text_features"]))
            neg = orc.cluster_index(half)
        args = {}
        if neg is not None:
            ptr, idx = _csr_inverse(neg, f.shape[0])
            args = dict(neg_index=neg.int().cuda(), inv_ptr=ptr.cuda(), inv_idx=idx.cuda())
        out4, _, df, dg, gs = K.index_fwd_bwd(f.cuda(), g.cuda(), dev_t(float(z["t"])), **args)
        df, dg = df / gs, dg / gs
        cross = float(z["out_cross_modal_loss"])
        assert abs(float(out4[2]) - cross) <= LOSS_RTOL * abs(cross), path
        # golden grads are d(total_loss) = 0.9 * d(cross)
        assert relerr(0.9 * df, torch.from_numpy(gf)) < 1e-4, path
        assert relerr(0.9 * dg, torch.from_numpy(gg)) < 1e-4, path
        assert relerr(0.9 * out4[3], torch.from_numpy(z["grad_temperature"])) < 1e-4, path


def test_index_fp16_gradients_survive_loss_scaling():
    """fp16 features with large norms: the unscaled per-element gradient sigma / (B ||f||) is below fp16's
    normal range; the upstream gradient (GradScaler's factor) reaches the kernel as a device scalar and is applied
    in fp32 before the single rounding to fp16 (ops.py:90)."""
    from clip_lite_b200 import ops
    b, d = 4096, 256
    f, g = orc.synth_embeddings(b, d, seed=3, correlated=True)
    f, g = (f * 40).half(), (g * 40).half()
    fl, gl = f.cuda().requires_grad_(True), g.cuda().requires_grad_(True)
    t = dev_t().requires_grad_(True)
    scale = 65536.0
    loss, _ = ops.jsd_index_loss(fl, gl, t)
    (loss * scale).backward()
    rdf, rdg, _ = orc.jsd_index_grads(f.double(), g.double(), T0, None)
    assert float(rdf.abs().max()) < 6.2e-5            # unscaled: fp16-subnormal territory
    assert relerr(fl.grad.double() / scale, rdf) < GRAD_RTOL
    assert relerr(gl.grad.double() / scale, rdg) < GRAD_RTOL


# ------------------------------------------------------------------ tensor-core GEMM (tcgen05 + TMA)
GEMM_SHAPES = [(128, 256, 64), (128, 256, 256), (256, 512, 128), (384, 256, 1024), (300, 520, 200),
               (64, 8, 72), (1000, 1024, 1000),
               (8192, 1024, 1024),      # 256 tiles on 148 SMs: stream-K with split tiles
               (2432, 2048, 520),       # 152 tiles, ragged K
               (19000, 256, 192)]       # 149 tiles of 3 k-chunks: a tile split over several CTAs


def _operand(rows, k, mn_major):
    """Logical [rows, k] bf16 operand; stored transposed ([k, rows padded to 8]) when mn_major."""
    x = torch.randn(rows, k, device="cuda").bfloat16()
    if not mn_major:
        return x, x
    rp = (rows + 7) // 8 * 8
    xt = torch.zeros(k, rp, device="cuda").bfloat16()
    xt[:, :rows] = x.t()
    return x, xt


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (True, False), (False, True), (True, True)])
@pytest.mark.parametrize("m,n,k", GEMM_SHAPES)
def test_gemm_all_operand_layouts(K, m, n, k, a_mn, b_mn):
    a, a_store = _operand(m, k, a_mn)
    b, b_store = _operand(n, k, b_mn)
    ref = a.double() @ b.double().t()
    for stream_k in (True, False):
        c = K.gemm_bf16(a_store, b_store, a_mn_major=a_mn, b_mn_major=b_mn, stream_k=stream_k)[:m, :n]
        assert relerr(c, ref) < 1e-5, f"stream_k={stream_k}"


def test_stream_k_is_repeatable_and_leaves_flags_clear(K):
    a = torch.randn(8192, 1024, device="cuda").bfloat16()
    b = torch.randn(1024, 1024, device="cuda").bfloat16()
    c0 = K.gemm_bf16(a, b)
    for _ in range(3):
        assert torch.equal(K.gemm_bf16(a, b), c0)
    from clip_lite_b200 import _lib
    ws = K.streamk_workspace(a.device)
    torch.cuda.synchronize()
    assert int(ws[:_lib.load().jsd_streamk_flag_bytes()].sum()) == 0


# ------------------------------------------------------------------ dense mode
def _unit_bf16(b, d, seed, correlated=True):
    f, g = orc.synth_embeddings(b, d, seed, correlated)
    u = orc.l2_normalize(f)[0].bfloat16().cuda()
    v = orc.l2_normalize(g)[0].bfloat16().cuda()
    return u, v


@pytest.mark.parametrize("m,n,d,off", [(128, 128, 64, 0), (256, 256, 128, 0), (1024, 1024, 128, 0),
                                        (384, 384, 72, 0), (100, 100, 64, 0), (128, 512, 256, 256),
                                        (200, 1000, 128, 800), (2, 2, 8, 0), (512, 512, 1024, 0)])
def test_dense_fwd_stage(K, m, n, d, off):
    _, v = _unit_bf16(n, d, seed=n + d)
    u = _unit_bf16(n, d, seed=n + d)[0][off:off + m].contiguous()
    out4, loss, gmat, gdiag = K.dense_fwd(u, v, dev_t(), row_offset=off)
    assert torch.equal(loss, out4[2])
    ref = orc.dense_from_unit(u.double(), v.double(), T0, row_offset=off)
    assert relerr(out4[0], ref["pos"]) < 1e-4
    assert relerr(out4[1], ref["neg"]) < 1e-4
    assert relerr(out4[2], ref["loss"]) < 1e-4
    assert float(out4[3]) == 0.0                     # dense dL/dt comes out of the backward
    assert relerr(gdiag, ref["gdiag"]) < 1e-4
    assert (gmat[:, :n].double() - ref["gmat"]).abs().max() < 2 ** -8     # bf16 rounding of sigma in (0, 1)
    rows = torch.arange(m, device="cuda")
    assert (gmat[rows, rows + off] == 0).all()


@pytest.mark.parametrize("m,n,d,off", [(128, 128, 64, 0), (1024, 1024, 128, 0), (384, 384, 72, 0),
                                        (128, 512, 256, 256), (200, 1000, 128, 800), (512, 512, 1024, 0),
                                        (4096, 4096, 1024, 0), (1024, 8192, 1024, 3072)])
def test_dense_bwd_stage(K, m, n, d, off):
    _, v = _unit_bf16(n, d, seed=n + d)
    u = _unit_bf16(n, d, seed=n + d)[0][off:off + m].contiguous()
    t = dev_t()
    gamma = torch.tensor(0.9 * 128.0, device="cuda")
    _, _, gmat, _ = K.dense_fwd(u, v, t, row_offset=off)
    du = K.dense_bwd_du(gmat, v, t, gamma)
    dv = K.dense_bwd_dv(gmat, u, n, t, gamma)
    # the opt-in stream-K schedule must give the same result up to fp32 summation order
    assert relerr(K.dense_bwd_du(gmat, v, t, gamma, stream_k=True), du) < 1e-5
    assert relerr(K.dense_bwd_dv(gmat, u, n, t, gamma, stream_k=True), dv) < 1e-5
    # same bf16 Gmat fed to an fp64 contraction
    scale = float(gamma) * np.exp(T0) / (m * (n - 1))
    ref_du = scale * (gmat[:, :n].double() @ v.double())
    ref_dv = scale * (gmat[:, :n].double().t() @ u.double())
    assert relerr(du, ref_du) < 1e-5
    assert relerr(dv, ref_dv) < 1e-5
    ref = orc.dense_from_unit(u.double(), v.double(), T0, row_offset=off, gamma=float(gamma))
    assert relerr(du, ref["du_acc"]) < 5e-3
    assert relerr(dv, ref["dv_acc"]) < 5e-3


# ------------------------------------------------------------------ fused single-pass kernel (D <= 256)
@pytest.mark.parametrize("b,d", [(128, 64), (256, 128), (200, 64), (1000, 256), (1024, 128), (333, 192),
                                  (2, 64), (129, 128), (4096, 256), (8192, 128)])
def test_dense_fused_stage(K, b, d):
    """Score tile, sigma(S) and both gradient accumulators never leave the SM: loss, gdiag and the UNSCALED
    accumulators sum_j sigma(S_ij) v_j / sum_i sigma(S_ij) u_i against the fp64 restatement on the same bf16
    operands, and against the staged path (forward + two contractions through the bf16 Gmat)."""
    from clip_lite_b200 import _lib
    assert _lib.load().jsd_dense_fused_supported(b, d)
    ns = _lib.load().jsd_dense_fused_splits(b, d)
    assert 1 <= ns <= 8
    u, v = _unit_bf16(b, d, seed=b + d)
    t = dev_t()
    acc = torch.full((2, ns, b, d), float("nan"), device="cuda")
    gdiag = torch.empty(b, device="cuda")
    out4 = torch.empty(4, device="cuda")
    loss = torch.empty((), device="cuda")
    ws = K.dense_workspace(u.device)
    _lib.call("jsd_dense_fused_fwd_bwd", u.data_ptr(), v.data_ptr(), b, d, t.data_ptr(), acc[0].data_ptr(),
              acc[1].data_ptr(), gdiag.data_ptr(), ws.data_ptr(), out4.data_ptr(), loss.data_ptr(), K._stream())
    torch.cuda.synchronize()
    ref = orc.dense_from_unit(u.double(), v.double(), T0)
    assert torch.equal(loss, out4[2])
    assert relerr(out4[0], ref["pos"]) < 1e-4
    assert relerr(out4[1], ref["neg"]) < 1e-4
    assert relerr(out4[2], ref["loss"]) < 1e-4
    assert relerr(gdiag, ref["gdiag"]) < 1e-4
    ref_u = ref["gmat"] @ v.double()                 # gmat: sigma(S) with 0 on the positives
    ref_v = ref["gmat"].t() @ u.double()
    assert torch.isfinite(acc).all()
    acc_s = acc.double().sum(1)                      # the column splits' slices
    assert relerr(acc_s[0], ref_u) < 5e-3            # sigma is rounded to bf16 before the second contraction
    assert relerr(acc_s[1], ref_v) < 5e-3
    if b >= 128:
        # the staged path rounds (nearly) the same sigma values to bf16: the two routes agree far inside the tolerance
        _, _, gmat, gd2 = K.dense_fwd(u, v, t)
        one = torch.ones((), device="cuda")
        scale = float(np.exp(T0)) / (b * (b - 1))
        du = K.dense_bwd_du(gmat, v, t, one) / scale
        dv = K.dense_bwd_dv(gmat, u, b, t, one) / scale
        assert relerr(acc_s[0], du) < 1e-3 and relerr(acc_s[1], dv) < 1e-3
        assert relerr(gdiag, gd2) < 1e-6
    # a second launch on the re-armed workspace gives bit-identical results (deterministic, tickets reset)
    acc2 = torch.empty_like(acc)
    out4b = torch.empty(4, device="cuda")
    _lib.call("jsd_dense_fused_fwd_bwd", u.data_ptr(), v.data_ptr(), b, d, t.data_ptr(), acc2[0].data_ptr(),
              acc2[1].data_ptr(), gdiag.data_ptr(), ws.data_ptr(), out4b.data_ptr(), loss.data_ptr(), K._stream())
    torch.cuda.synchronize()
    assert torch.equal(acc, acc2) and torch.equal(out4, out4b)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("b,d", [(256, 64), (1024, 128), (1000, 256), (640, 192)])
def test_dense_fused_autograd_matches_staged_and_oracle(monkeypatch, dtype, b, d):
    from clip_lite_b200 import ops
    f, g = orc.synth_embeddings(b, d, seed=5, correlated=True)
    f, g = f.to(dtype), g.to(dtype)
    # fp16: a GradScaler-sized upstream gradient keeps the B x D gradients out of fp16's subnormal range
    gamma = 0.7 * (4096.0 if dtype == torch.float16 else 1.0)
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("JSD_FUSED", mode)
        fl, gl = f.cuda().requires_grad_(True), g.cuda().requires_grad_(True)
        t = dev_t().requires_grad_(True)
        loss, _ = ops.jsd_dense_loss(fl, gl, t)
        (gamma * loss).backward()
        res[mode] = (loss.detach(), fl.grad, gl.grad, t.grad)
    rdf, rdg, rdt = orc.jsd_dense_grads(f.double(), g.double(), T0, gamma=gamma)
    ref = orc.jsd_dense(f.double(), g.double(), T0)
    for mode in ("1", "0"):
        loss, df, dg, dt = res[mode]
        assert relerr(loss, ref["loss"]) < LOSS_RTOL
        assert relerr(df, rdf) < GRAD_RTOL and relerr(dg, rdg) < GRAD_RTOL and relerr(dt, rdt) < GRAD_RTOL
    assert relerr(res["1"][0], res["0"][0]) < 1e-5
    tol = {torch.float32: 1e-3, torch.bfloat16: 1e-2, torch.float16: 1e-2}[dtype]
    assert relerr(res["1"][1], res["0"][1]) < tol and relerr(res["1"][2], res["0"][2]) < tol


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("b,d", [(128, 64), (256, 128), (1024, 128), (1024, 1024), (200, 72)])
def test_dense_pipeline_vs_oracle(K, dtype, b, d):
    """normalise -> fwd -> dU/dV -> normalise-bwd, against the fp64 restatement on the same inputs."""
    f, g = orc.synth_embeddings(b, d, seed=3, correlated=True)
    f, g = f.to(dtype).cuda(), g.to(dtype).cuda()
    t = dev_t()
    gamma = torch.tensor(0.9, device="cuda")
    u, inv_f = K.normalize_cast(f)
    v, inv_g = K.normalize_cast(g)
    out4, _, gmat, gdiag = K.dense_fwd(u, v, t)
    du = K.dense_bwd_du(gmat, v, t, gamma)
    dv = K.dense_bwd_dv(gmat, u, b, t, gamma)
    df, dt = K.normalize_bwd(f, inv_f, du, v, 0, gdiag, t, gamma, b, want_dt=True)
    dg = K.normalize_bwd(g, inv_g, dv, u, 0, gdiag, t, gamma, b)
    ref = orc.jsd_dense(f.double(), g.double(), T0)
    rdf, rdg, rdt = orc.jsd_dense_grads(f.double(), g.double(), T0, gamma=0.9)
    assert relerr(out4[2], ref["loss"]) < LOSS_RTOL
    assert relerr(dt, rdt) < GRAD_RTOL
    assert relerr(df, rdf) < GRAD_RTOL
    assert relerr(dg, rdg) < GRAD_RTOL


def test_dense_fwd_loss_only(K):
    u, v = _unit_bf16(256, 128, seed=0)
    a, _, gmat, _ = K.dense_fwd(u, v, dev_t(), want_grad=False)
    b, _, _, _ = K.dense_fwd(u, v, dev_t(), want_grad=True)
    assert gmat is None and torch.equal(a, b)


def test_bad_arguments_fail_loudly(K):
    from clip_lite_b200._lib import JSDLibraryError
    u, v = _unit_bf16(64, 12, seed=0)           # D not a multiple of 8
    with pytest.raises(JSDLibraryError):
        K.dense_fwd(u, v, dev_t())
    with pytest.raises(RuntimeError):
        K.index_fwd_bwd(torch.randn(4, 8), torch.randn(4, 8), dev_t())   # CPU tensors: no fallback


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_fused_forward_backward_calls_equal_the_staged_calls(K, dtype):
    """jsd_dense_forward / jsd_dense_backward are exactly the staged entry points in one FFI crossing."""
    b, d = 512, 256
    f, g = orc.synth_embeddings(b, d, seed=5, correlated=True)
    f, g = f.to(dtype).cuda(), g.to(dtype).cuda()
    t = dev_t()
    gamma = torch.tensor(0.7, device="cuda")
    out4, loss, saved = K.dense_forward(f, g, t)
    df, dg, dt = K.dense_backward(f, g, t, gamma, saved)
    u, inv_f = K.normalize_cast(f)
    v, inv_g = K.normalize_cast(g)
    out4b, lossb, gmat, gdiag = K.dense_fwd(u, v, t)
    du = K.dense_bwd_du(gmat, v, t, gamma)
    dv = K.dense_bwd_dv(gmat, u, b, t, gamma)
    dfb, dtb = K.normalize_bwd(f, inv_f, du, v, 0, gdiag, t, gamma, b, want_dt=True)
    dgb = K.normalize_bwd(g, inv_g, dv, u, 0, gdiag, t, gamma, b)
    assert torch.equal(out4, out4b) and torch.equal(loss, lossb)
    # at this size the fused backward splits the contraction of its underfilled GEMM launches over idle CTA pairs
    # (split-K, slices added in a fixed order): same arithmetic, different fp32 summation order than the staged calls
    tol = 1e-5 if dtype == torch.float32 else 8e-3        # 16-bit outputs: the last fp32 bits can flip one rounding
    assert relerr(df, dfb) < tol and relerr(dg, dgb) < tol and relerr(dt, dtb) < 1e-5
    df2, dg2, dt2 = K.dense_backward(f, g, t, gamma, saved)
    assert torch.equal(df, df2) and torch.equal(dg, dg2) and torch.equal(dt, dt2)      # and it is deterministic
    _, _, rdt = orc.jsd_dense_grads(f.double(), g.double(), T0, gamma=0.7)
    assert relerr(dt, rdt) < GRAD_RTOL


def test_slab_convenience_calls_equal_the_staged_calls(K):
    m, n, d, off = 256, 1024, 128, 512
    f, g = orc.synth_embeddings(n, d, seed=6, correlated=True)
    f, g = f.cuda(), g.cuda()
    t = dev_t()
    gamma = torch.tensor(0.5, device="cuda")
    u_all, v_all, inv_f_all, inv_g_all = K.normalize_cast_pair(f, g)
    u2, inv2 = K.normalize_cast(f)
    assert torch.equal(u_all, u2) and torch.equal(inv_f_all, inv2)
    fl, u, inv_f = f[off:off + m].contiguous(), u_all[off:off + m].contiguous(), inv_f_all[off:off + m].contiguous()
    _, _, gmat, gdiag = K.dense_fwd(u, v_all, t, row_offset=off)
    df, dt = K.dense_backward_image_side(fl, v_all, inv_f, gmat, gdiag, t, gamma, off)
    du = K.dense_bwd_du(gmat, v_all, t, gamma)
    dfb, dtb = K.normalize_bwd(fl, inv_f, du, v_all, off, gdiag, t, gamma, m, want_dt=True)
    assert relerr(df, dfb) < 1e-5 and relerr(dt, dtb) < 1e-5      # split-K in the fused call: other summation order
    df2, dt2 = K.dense_backward_image_side(fl, v_all, inv_f, gmat, gdiag, t, gamma, off)
    assert torch.equal(df, df2) and torch.equal(dt, dt2)

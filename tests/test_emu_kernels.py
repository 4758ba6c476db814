"""CPU execution of the row-wise CUDA kernels (tests/emu): the kernel SOURCE of clip_lite_b200/csrc is compiled by g++
against a shim in which every CUDA thread is an OS thread and barriers / shuffles are real rendezvous, and checked
against fp64 PyTorch.  Test infrastructure only -- the product has no CPU path and never loads this library.

Why: the projection-head tail (jsd_heads.cuh: LayerNorm + F.normalize fused, reference loss.py:36-38 + :94-95) was
written while no GPU was available; this tier checks its indexing, reductions, barriers and arithmetic.  The first
tests run kernels that ARE covered by the GPU tier (normalise, Jacobian, index mode) through the same shim and compare
them with the oracle: they validate the shim itself."""
import ctypes

import pytest
import torch

from oracle import heads_oracle as ho
from oracle import jsd_oracle as orc

from tests import _emu_backend

pytestmark = pytest.mark.skipif(not _emu_backend.available(), reason="needs g++ and the CUDA headers")

_c = ctypes
_P, _I, _L, _F = _c.c_void_p, _c.c_int, _c.c_longlong, _c.c_float


@pytest.fixture(scope="module")
def emu():
    lib = _emu_backend.load()
    lib.emu_ln_normalize_pair.argtypes = [_P, _P, _I, _L, _L, _P, _P, _F, _P, _P, _F, _I, _P, _P, _P, _P, _I]
    lib.emu_ln_normalize_bwd_pair.argtypes = [_P, _P, _I, _L, _L, _P, _P, _P, _P, _P, _P, _P, _P, _L, _L, _F, _P, _L,
                                              _P, _L, _P, _P, _P, _L, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I]
    lib.emu_normalize_cast.argtypes = [_P, _I, _L, _L, _P, _P, _I]
    lib.emu_normalize_bwd.argtypes = [_P, _I, _L, _L, _P, _P, _P, _L, _P, _P, _P, _L, _P, _P, _P, _P, _I]
    lib.emu_index_fwd_bwd.argtypes = [_P, _P, _I, _L, _L, _P, _P, _P, _P, _P, _P, _P, _P, _P, _F, _P, _I]
    return lib


CODE = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}
TOL = {torch.float32: 2e-5, torch.bfloat16: 1.2e-2, torch.float16: 1.5e-3}   # relative to the tensor's largest entry


def ptr(t):
    return None if t is None else t.data_ptr()


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def unit_rows_bf16(n, d, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, d, generator=g)
    return torch.nn.functional.normalize(x, dim=-1).bfloat16()


# ------------------------------------------------------------------ the shim itself, on GPU-tested kernels
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("rows,d,variant", [(19, 100, 2), (9, 64, 1), (11, 256, 0), (3, 1024, 0), (5, 72, 1)])
def test_shim_normalize_cast_matches_oracle(emu, dtype, rows, d, variant):
    x = (torch.randn(rows, d, generator=torch.Generator().manual_seed(rows * d)) * 3).to(dtype)
    xn = torch.empty(rows, d, dtype=torch.bfloat16)
    inv = torch.empty(rows)
    assert emu.emu_normalize_cast(ptr(x), CODE[dtype], rows, d, ptr(xn), ptr(inv), variant) == 0
    u, n = orc.l2_normalize(x.double())
    assert rel(inv, 1.0 / n.squeeze(-1)) < 1e-6
    assert float((xn.double() - u).abs().max()) < 2.0 ** -8      # bf16 rounding of entries <= 1


@pytest.mark.parametrize("rows,d,variant", [(19, 100, 2), (9, 64, 1), (11, 256, 0)])
def test_shim_normalize_bwd_matches_closed_form(emu, rows, d, variant):
    g = torch.Generator().manual_seed(7 + d)
    x = torch.randn(rows, d, generator=g)
    acc = torch.randn(rows, d, generator=g) * 1e-3
    partner = unit_rows_bf16(rows + 4, d, 11)
    gdiag = -torch.rand(rows, generator=g)
    t = torch.tensor([0.7])
    gamma = torch.tensor([1.3])
    inv = (1.0 / x.double().norm(dim=-1)).float()
    dx = torch.empty_like(x)
    rowdot = torch.zeros(rows)
    ticket = torch.zeros(4, dtype=torch.int32)
    dt = torch.zeros(1)
    assert emu.emu_normalize_bwd(ptr(x), 0, rows, d, ptr(inv), ptr(acc), ptr(partner), 2, ptr(gdiag), ptr(t),
                                 ptr(gamma), rows, ptr(dx), ptr(rowdot), ptr(ticket), ptr(dt), variant) == 0
    c = float(gamma) * float(t.exp()) / rows
    du = acc.double() + c * gdiag.double()[:, None] * partner[2:2 + rows].double()
    u = x.double() * inv.double()[:, None]
    want = (du - u * (u * du).sum(-1, keepdim=True)) * inv.double()[:, None]
    assert rel(dx, want) < 2e-5
    assert rel(rowdot, (u * du).sum(-1)) < 2e-5
    assert abs(float(dt) - float((u * du).sum())) < 1e-6 + 2e-5 * float((u * du).sum().abs())
    assert int(ticket[0]) == 0                                   # re-armed by the last block


@pytest.mark.parametrize("b,d,vec,mode", [(19, 100, 0, "roll"), (16, 64, 1, "roll"), (12, 128, 1, "cluster"),
                                          (13, 36, 1, "general")])
def test_shim_index_kernel_matches_oracle(emu, b, d, vec, mode):
    f, g = orc.synth_embeddings(b, d, seed=b, correlated=True)
    f, g = (f * 2).float().contiguous(), (g * 0.5).float().contiguous()
    neg = iptr = iidx = None
    if mode == "cluster":
        half = b // 2
        i = torch.arange(half)
        neg = torch.cat((half + i, (i + 1) % half))
    elif mode == "general":
        neg = torch.randint(0, b, (b,), generator=torch.Generator().manual_seed(3))
    if neg is not None:
        order = torch.argsort(neg, stable=True)
        p = torch.zeros(b + 1, dtype=torch.int64)
        p[1:] = torch.cumsum(torch.bincount(neg, minlength=b), 0)
        neg, iptr, iidx = neg.int(), p.int(), order.int()
    t = torch.tensor([orc.T_INIT])
    coefp, partials = torch.zeros(b), torch.zeros(3 * ((b + 7) // 8))
    out4 = torch.zeros(4)
    df, dg = torch.empty_like(f), torch.empty_like(g)
    assert emu.emu_index_fwd_bwd(ptr(f), ptr(g), 0, b, d, ptr(neg), ptr(iptr), ptr(iidx), ptr(t), ptr(coefp),
                                 ptr(partials), ptr(out4), ptr(df), ptr(dg), 1.0, None, vec) == 0
    ni = None if neg is None else neg.long()
    ref = orc.jsd_index(f.double(), g.double(), orc.T_INIT, ni)
    grads = orc.jsd_index_grads(f.double(), g.double(), orc.T_INIT, ni)
    assert rel(out4[2], ref["loss"]) < 1e-5
    assert rel(df, grads[0]) < 1e-4 and rel(dg, grads[1]) < 1e-4
    assert rel(out4[3], grads[2]) < 1e-4


# ------------------------------------------------------------------ projection-head tail: forward
def ln_unit_reference(x, w, b, eps):
    """(unit rows, mean, rstd, 1/||LN(x)||) by the fp64 oracle (oracle/heads_oracle.py <- loss.py:36-38, :94-95)."""
    u, (mean, rstd, inv) = ho.ln_unit(x.double(), None if w is None else w.double(),
                                      None if b is None else b.double(), eps)
    return u, mean, rstd, inv


def make_ln(d, seed, affine=True):
    g = torch.Generator().manual_seed(seed)
    if not affine:
        return None, None
    return (1.0 + 0.3 * torch.randn(d, generator=g)).contiguous(), (0.2 * torch.randn(d, generator=g)).contiguous()


FWD_SHAPES = [(19, 102, 2), (19, 100, 1), (9, 64, 1), (11, 256, 0), (3, 2048, 0), (4, 2176, 1), (2, 4096, 1), (1, 128, 0), (8, 4, 1),
              (5, 1, 2)]


@pytest.mark.parametrize("out_bf16", [0, 1])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("rows,d,variant", FWD_SHAPES)
def test_ln_normalize_forward(emu, rows, d, variant, dtype, out_bf16):
    g = torch.Generator().manual_seed(rows + d)
    x0 = (torch.randn(rows, d, generator=g) * 2 + 0.5).to(dtype)
    x1 = (torch.randn(rows, d, generator=g) * 0.1 - 3).to(dtype)
    w0, b0 = make_ln(d, 1)
    w1, b1 = make_ln(d, 2, affine=(d % 2 == 0))          # one head without affine parameters now and then
    odt = torch.bfloat16 if out_bf16 else torch.float32
    out0, out1 = torch.empty(rows, d, dtype=odt), torch.empty(rows, d, dtype=odt)
    st0, st1 = torch.empty(3, rows), torch.empty(3, rows)
    got = emu.emu_ln_normalize_pair(ptr(x0), ptr(x1), CODE[dtype], rows, d, ptr(w0), ptr(b0), 1e-5, ptr(w1), ptr(b1),
                                    1e-3, out_bf16, ptr(out0), ptr(out1), ptr(st0), ptr(st1), -1)
    assert got == variant                                 # the variant the product would pick for this shape
    for x, w, b, eps, out, st in ((x0, w0, b0, 1e-5, out0, st0), (x1, w1, b1, 1e-3, out1, st1)):
        if d == 1 and eps == 1e-5:
            continue                                      # LN of one element is all bias: nothing to compare but u = +-1
        u, mean, rstd, inv = ln_unit_reference(x, w, b, eps)
        assert float((out.double() - u).abs().max()) < (2.0 ** -8 if out_bf16 else 3e-6)
        assert rel(st[0], mean) < 1e-5 and rel(st[1], rstd) < 1e-5 and rel(st[2], inv) < 1e-5


@pytest.mark.parametrize("rows,d,natural,forced", [(6, 256, 0, 1), (6, 256, 0, 2), (7, 64, 1, 2)])
def test_ln_normalize_forward_variants_agree(emu, rows, d, natural, forced):
    """The register-resident, the 16-byte and the element-wise forward compute the same thing."""
    x = torch.randn(rows, d, generator=torch.Generator().manual_seed(5))
    w, b = make_ln(d, 9)
    outs = []
    for v in (natural, forced):
        out, st = torch.empty(rows, d), torch.empty(3, rows)
        assert emu.emu_ln_normalize_pair(ptr(x), None, 0, rows, d, ptr(w), ptr(b), 1e-5, None, None, 0.0, 0, ptr(out),
                                         None, ptr(st), None, v) == v
        outs.append((out, st))
    assert float((outs[0][0] - outs[1][0]).abs().max()) < 1e-6
    assert float((outs[0][1] - outs[1][1]).abs().max() / outs[0][1].abs().max()) < 1e-6


def test_ln_normalize_forward_zero_variance_row_is_finite(emu):
    """A constant row: variance 0, rstd = 1/sqrt(eps), LN output = bias, unit row = bias / ||bias||."""
    d = 128
    x = torch.full((3, d), 2.5)
    w, b = make_ln(d, 4)
    out, st = torch.empty(3, d), torch.empty(3, 3)
    emu.emu_ln_normalize_pair(ptr(x), None, 0, 3, d, ptr(w), ptr(b), 1e-5, None, None, 0.0, 0, ptr(out), None, ptr(st),
                              None, -1)
    assert torch.isfinite(out).all() and torch.isfinite(st).all()
    assert float((out.double() - (b / b.norm()).double()).abs().max()) < 1e-5


# ------------------------------------------------------------------ projection-head tail: backward
def ln_bwd_reference(x, w, b, eps, du):
    """(dx, dw, db, <u, dU>) for the upstream gradient dU by the fp64 oracle's closed forms (pinned against
    autograd of the reference's ops in tests/test_oracle.py)."""
    return ho.ln_unit_grads(x.double(), None if w is None else w.double(), None if b is None else b.double(), eps,
                            du.double())


def run_ln_fwd(emu, xs, lns, eps):
    rows, d = xs[0].shape
    stats = [torch.empty(3, rows) for _ in xs]
    outs = [torch.empty(rows, d, dtype=torch.bfloat16) for _ in xs]
    two = len(xs) == 2
    emu.emu_ln_normalize_pair(ptr(xs[0]), ptr(xs[1]) if two else None, CODE[xs[0].dtype], rows, d, ptr(lns[0][0]),
                              ptr(lns[0][1]), eps, ptr(lns[1][0]) if two else None, ptr(lns[1][1]) if two else None,
                              eps, 1, ptr(outs[0]), ptr(outs[1]) if two else None, ptr(stats[0]),
                              ptr(stats[1]) if two else None, -1)
    return outs, stats


BWD_SHAPES = [
    # rows, D, blocks, expected plan (vec * 1000 + threads * 10 + kch), force_scalar
    (19, 102, 3, 1000 + 128 * 10 + 1, 0),     # D % 4 != 0: element-wise, 128 threads
    (19, 100, 19, 4000 + 32 * 10 + 1, 0),     # 25 pieces: one warp, one block per row
    (7, 64, 7, 4000 + 32 * 10 + 1, 0),        # one warp per row
    (9, 1024, 2, 4000 + 256 * 10 + 1, 0),
    (5, 2048, 5, 4000 + 256 * 10 + 2, 0),     # the heads' width
    (3, 2176, 1, 4000 + 256 * 10 + 4, 0),     # 3 pieces per thread -> the 4-piece instantiation, one block for all rows
    (2, 4096, 2, 4000 + 256 * 10 + 4, 0),
    (6, 300, 4, 1000 + 256 * 10 + 2, 1),      # element-wise, 2 columns per thread
    (4, 1000, 3, 1000 + 256 * 10 + 4, 1),
]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("rows,d,blocks,plan,force_scalar", BWD_SHAPES)
def test_ln_normalize_backward_gradient_in(emu, rows, d, blocks, plan, force_scalar, dtype):
    """The generic flavour (index mode and every multi-GPU route): the gradient with respect to the unit rows comes
    in as one fp32 tensor; no positive-pair term, no scaling."""
    g = torch.Generator().manual_seed(rows * 31 + d)
    x = (torch.randn(rows, d, generator=g) * 1.5 + 0.3).to(dtype)
    w, b = make_ln(d, 21)
    _, stats = run_ln_fwd(emu, [x], [(w, b)], 1e-5)
    du = torch.randn(rows, d, generator=g).contiguous()
    ws = torch.full((2 * blocks * 2 * d,), float("nan"))
    dx = torch.empty_like(x)
    dw, db = torch.empty(d), torch.empty(d)
    rowdot, dt = torch.empty(rows), torch.empty(1)
    got = emu.emu_ln_normalize_bwd_pair(ptr(x), None, CODE[dtype], rows, d, ptr(w), ptr(b), None, None, ptr(stats[0]),
                                        None, ptr(du), None, 1, 0, 0.0, None, 0, None, 0, None, None, None, rows,
                                        ptr(ws), ptr(dx), None, ptr(dw), ptr(db), None, None, ptr(rowdot), ptr(dt),
                                        blocks, force_scalar)
    assert got == plan
    rdx, rdw, rdb, rdot = ln_bwd_reference(x, w, b, 1e-5, du.double())
    assert rel(dx, rdx) < TOL[dtype]
    assert rel(dw, rdw) < 5e-5 and rel(db, rdb) < 5e-5
    assert rel(rowdot, rdot) < 5e-5
    assert abs(float(dt) - float(rdot.sum())) < 1e-4 * float(rdot.abs().sum())


@pytest.mark.parametrize("rows,d,blocks,n_slices,fused", [(10, 256, 4, 1, False), (10, 256, 3, 3, True),
                                                           (6, 2048, 6, 2, True), (13, 102, 5, 1, False),
                                                           (13, 100, 2, 4, True)])
def test_ln_normalize_backward_dense_pair(emu, rows, d, blocks, n_slices, fused):
    """The dense flavour, both heads in one launch: accumulator (slices summed in order; the fused kernel's are unscaled
    and get gamma tau acc_scale here) + positive-pair term c * partner, then both Jacobians; row dots of the image
    side summed to gamma dL/dt."""
    g = torch.Generator().manual_seed(rows + 7 * d + n_slices)
    xs = [torch.randn(rows, d, generator=g) * 2, torch.randn(rows, d, generator=g) - 1.0]
    lns = [make_ln(d, 31), make_ln(d, 32)]
    (u16, v16), stats = run_ln_fwd(emu, xs, lns, 1e-5)
    t, gamma = torch.tensor([1.1]), torch.tensor([0.6])
    gdiag = -torch.rand(rows, generator=g)
    stride = rows * d + 8                                 # slices need not be packed
    accs = [(torch.randn(n_slices, stride, generator=g) * (1.0 if fused else 1e-3)).contiguous() for _ in range(2)]
    acc_scale = 1.0 / (rows * (rows - 1)) if fused else 0.0
    ws = torch.full((2 * blocks * 2 * d,), float("nan"))
    dxs = [torch.empty_like(x) for x in xs]
    dws, dbs = [torch.empty(d), torch.empty(d)], [torch.empty(d), torch.empty(d)]
    rowdot, dt = torch.empty(rows), torch.empty(1)
    rc = emu.emu_ln_normalize_bwd_pair(ptr(xs[0]), ptr(xs[1]), 0, rows, d, ptr(lns[0][0]), ptr(lns[0][1]),
                                       ptr(lns[1][0]), ptr(lns[1][1]), ptr(stats[0]), ptr(stats[1]), ptr(accs[0]),
                                       ptr(accs[1]), n_slices, stride, acc_scale, ptr(v16), 0, ptr(u16), 0, ptr(gdiag),
                                       ptr(t), ptr(gamma), rows, ptr(ws), ptr(dxs[0]), ptr(dxs[1]), ptr(dws[0]),
                                       ptr(dbs[0]), ptr(dws[1]), ptr(dbs[1]), ptr(rowdot), ptr(dt), blocks, 0)
    assert rc > 0
    c = float(gamma) * float(t.exp()) / rows
    scale = float(gamma) * float(t.exp()) * acc_scale if fused else 1.0
    for j, (x, (w, b), partner) in enumerate(zip(xs, lns, (v16, u16))):
        acc = accs[j][:, :rows * d].double().sum(0).reshape(rows, d) * scale
        du = acc + c * gdiag.double()[:, None] * partner.double()
        rdx, rdw, rdb, rdot = ln_bwd_reference(x, w, b, 1e-5, du)
        assert rel(dxs[j], rdx) < 5e-5, j
        assert rel(dws[j], rdw) < 5e-5 and rel(dbs[j], rdb) < 5e-5, j
        if j == 0:
            assert rel(rowdot, rdot) < 5e-5
            assert abs(float(dt) - float(rdot.sum())) < 1e-4 * float(rdot.abs().sum())


def test_ln_backward_plan_rejects_oversized_rows(emu):
    """D beyond 4 pieces per thread is refused (the C ABI turns this into an error message), not silently truncated."""
    d, rows = 4100, 2
    x = torch.randn(rows, d)
    st = torch.empty(3, rows)
    du = torch.randn(rows, d)
    ws = torch.empty(2 * 1 * 2 * d)
    dx = torch.empty_like(x)
    rc = emu.emu_ln_normalize_bwd_pair(ptr(x), None, 0, rows, d, None, None, None, None, ptr(st), None, ptr(du), None,
                                       1, 0, 0.0, None, 0, None, 0, None, None, None, rows, ptr(ws), ptr(dx), None,
                                       None, None, None, None, None, None, 1, 0)
    assert rc == -1


# ------------------------------------------------------------------ racecheck on the CPU
def test_head_tail_kernels_are_race_free_under_tsan(tmp_path):
    """compute-sanitizer --tool racecheck needs a GPU; here the same question is put to ThreadSanitizer: with one OS
    thread per CUDA thread and barriers as the only synchronisation, any shared- or global-memory access of the
    LayerNorm / normalise kernels that a __syncthreads / __syncwarp does not order is a data race TSan reports.
    A deliberately broken block reduction (--racy) is the negative control: the check only counts where it fires."""
    import os
    import subprocess
    exe = str(tmp_path / "racecheck")
    cmd = ["g++", "-std=c++20", "-O1", "-g", "-pthread", "-fsanitize=thread", "-Wno-tsan", "-DJSD_HOST_EMU",
           "-I" + _emu_backend.CUDA_INC, "-I" + _emu_backend.EMU_DIR, "-I" + _emu_backend.CSRC,
           os.path.join(_emu_backend.EMU_DIR, "emu_racecheck.cpp"), "-o", exe]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        pytest.skip("ThreadSanitizer build not available: " + res.stderr[-300:])
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=1 exitcode=66")
    racy = subprocess.run([exe, "--racy"], capture_output=True, text=True, env=env, timeout=300)
    if racy.returncode != 66:
        pytest.skip("ThreadSanitizer is not functional in this environment (the negative control did not fire)")
    assert "data race" in racy.stderr
    ok = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=600)
    assert ok.returncode == 0, ok.stderr[-3000:]
    assert "racecheck cases done rc=0" in ok.stdout


def test_head_tail_kernels_are_memcheck_clean_under_asan(tmp_path):
    """The CPU stand-in for compute-sanitizer --tool memcheck: the same cases with exactly-sized heap buffers under
    AddressSanitizer + UBSan (an access one element past a row, a partial buffer or the workspace aborts)."""
    import os
    import subprocess
    exe = str(tmp_path / "memcheck")
    cmd = ["g++", "-std=c++20", "-O1", "-g", "-pthread", "-fsanitize=address,undefined", "-fno-sanitize-recover=all",
           "-DJSD_HOST_EMU", "-I" + _emu_backend.CUDA_INC, "-I" + _emu_backend.EMU_DIR, "-I" + _emu_backend.CSRC,
           os.path.join(_emu_backend.EMU_DIR, "emu_racecheck.cpp"), "-o", exe]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        pytest.skip("AddressSanitizer build not available: " + res.stderr[-300:])
    ok = subprocess.run([exe], capture_output=True, text=True, timeout=600,
                        env=dict(os.environ, ASAN_OPTIONS="detect_leaks=0"))
    assert ok.returncode == 0, (ok.stdout[-500:], ok.stderr[-3000:])
    assert "racecheck cases done rc=0" in ok.stdout


def test_ln_tail_full_size_with_the_production_launch_shape(emu):
    """The heads' real shape -- B = 8192 rows of D = 2048 (the bench's index workload; the module's width) -- with the
    launch the product makes on a B200: register-resident forward, backward with 256 threads x two 16-byte pieces
    and 3 x 148 = 444 blocks walking ~18 rows each, column partials summed over 444 blocks."""
    rows, d, blocks = 8192, 2048, 444
    g = torch.Generator().manual_seed(0)
    x = torch.randn(rows, d, generator=g) * 1.5 + 0.3
    w, b = make_ln(d, 21)
    out, st = torch.empty(rows, d), torch.empty(3, rows)
    assert emu.emu_ln_normalize_pair(ptr(x), None, 0, rows, d, ptr(w), ptr(b), 1e-5, None, None, 0.0, 0, ptr(out), None,
                                     ptr(st), None, -1) == 0
    u, mean, rstd, inv = ln_unit_reference(x, w, b, 1e-5)
    assert float((out.double() - u).abs().max()) < 1e-6
    assert rel(st[0], mean) < 1e-5 and rel(st[1], rstd) < 1e-5 and rel(st[2], inv) < 1e-5
    du = torch.randn(rows, d, generator=g)
    ws = torch.full((2 * blocks * 2 * d,), float("nan"))
    dx, dw, db = torch.empty_like(x), torch.empty(d), torch.empty(d)
    rowdot, dt = torch.empty(rows), torch.empty(1)
    got = emu.emu_ln_normalize_bwd_pair(ptr(x), None, 0, rows, d, ptr(w), ptr(b), None, None, ptr(st), None, ptr(du),
                                        None, 1, 0, 0.0, None, 0, None, 0, None, None, None, rows, ptr(ws), ptr(dx),
                                        None, ptr(dw), ptr(db), None, None, ptr(rowdot), ptr(dt), blocks, 0)
    assert got == 4000 + 256 * 10 + 2
    rdx, rdw, rdb, rdot = ln_bwd_reference(x, w, b, 1e-5, du.double())
    assert rel(dx, rdx) < 2e-5 and rel(dw, rdw) < 1e-5 and rel(db, rdb) < 1e-5
    assert rel(rowdot, rdot) < 1e-5
    assert abs(float(dt) - float(rdot.sum())) < 1e-6 * float(rdot.abs().sum())

"""world_size-2 gloo test of the data-parallel dense path (host logic only: sharding,
all-gather, row offsets, reduce-scatter, gradient convention).  The CUDA kernels are
replaced by the oracle-backed stand-in in tests/_standin_kernels.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import jsd_oracle as orc

B, D, T = 16, 32, 2.2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, results, route="reduce"):
    from tests import _standin_kernels
    from clip_lite_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        parallel.K = _standin_kernels
        f, g = orc.synth_embeddings(B, D, seed=0, correlated=True)
        m = B // world
        fl = f[rank * m:(rank + 1) * m].clone().requires_grad_(True)
        gl = g[rank * m:(rank + 1) * m].clone().requires_grad_(True)
        t = torch.tensor(T, requires_grad=True)
        loss, stats = parallel.gathered_dense_loss(fl, gl, t, route=route)
        (0.5 * loss).backward()
        logged = parallel.global_loss_for_logging(loss)
        results[rank] = (loss.detach(), fl.grad, gl.grad, t.grad, logged, stats)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,route", [(2, "reduce"), (2, "symmetric"), (4, "symmetric")])
def test_gathered_dense_equals_single_process_dense(world, route):
    """Both ways of completing the text-side gradient across ranks -- reduce-scatter of the partials, or gathering
    the image rows as well and recomputing the owned column slab -- give the single-process dense gradients."""
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), results, route), nprocs=world, join=True)
    f, g = orc.synth_embeddings(B, D, seed=0, correlated=True)
    fd, gd = f.double(), g.double()
    full = orc.jsd_dense(fd, gd, T)
    df, dg, dt = orc.jsd_dense_grads(fd, gd, T, gamma=0.5)
    m = B // world
    tot_dt = 0.0
    for r in range(world):
        loss, gf, gg, gt, logged, stats = results[r]
        slab = orc.jsd_dense(fd[r * m:(r + 1) * m], gd, T, row_offset=r * m)
        assert abs(float(loss) - float(slab["loss"])) < 1e-3 * float(slab["loss"])
        assert abs(float(logged) - float(full["loss"])) < 1e-3 * float(full["loss"])
        # every rank back-propagates its own slab loss; DDP's mean over ranks then gives the global gradient
        ref_f = world * df[r * m:(r + 1) * m]
        ref_g = world * dg[r * m:(r + 1) * m]
        assert (gf.double() - ref_f).abs().max() < 1e-2 * ref_f.abs().max()
        assert (gg.double() - ref_g).abs().max() < 1e-2 * ref_g.abs().max()
        tot_dt += float(gt)
        assert stats.shape == (4,)
    assert abs(tot_dt / world - float(dt)) < 1e-2 * abs(float(dt))


def test_single_process_path_needs_no_process_group():
    from tests import _standin_kernels
    from clip_lite_b200 import parallel
    old = parallel.K
    parallel.K = _standin_kernels
    try:
        f, g = orc.synth_embeddings(8, 16, seed=1, correlated=True)
        fl, gl = f.clone().requires_grad_(True), g.clone().requires_grad_(True)
        t = torch.tensor(T, requires_grad=True)
        loss, _ = parallel.gathered_dense_loss(fl, gl, t)
        loss.backward()
        ref = orc.jsd_dense(f.double(), g.double(), T)
        df, dg, dt = orc.jsd_dense_grads(f.double(), g.double(), T)
        assert abs(float(loss) - float(ref["loss"])) < 1e-3 * float(ref["loss"])
        assert (fl.grad.double() - df).abs().max() < 1e-2 * df.abs().max()
        assert (gl.grad.double() - dg).abs().max() < 1e-2 * dg.abs().max()
        assert abs(float(t.grad) - float(dt)) < 1e-2 * abs(float(dt))
    finally:
        parallel.K = old


def _avg_worker(rank, world, port, results):
    from clip_lite_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        comps = {"total_loss": torch.tensor(1.0 + rank), "cross_modal_loss": torch.tensor(0.5 * (rank + 1)),
                 "visual_loss": torch.tensor(0.0), "textual_loss": torch.tensor(float(rank) ** 2)}
        out = parallel.average_loss_components(comps)
        single = parallel.average_loss_components(torch.tensor(10.0 * rank))
        results[rank] = ({k: float(v) for k, v in out.items()}, float(single), out is comps)
    finally:
        dist.destroy_process_group()


def test_average_loss_components_matches_the_reference_semantics():
    """One packed all-reduce gives what utils/distributed.py:141-159 computes with four: the mean over ranks,
    written back in place, identical on every rank."""
    world = 2
    results = mp.Manager().dict()
    mp.spawn(_avg_worker, args=(world, _free_port(), results), nprocs=world, join=True)
    want = {"total_loss": 1.5, "cross_modal_loss": 0.75, "visual_loss": 0.0, "textual_loss": 0.5}
    for r in range(world):
        got, single, same_object = results[r]
        assert got == want and single == 5.0 and same_object
    from clip_lite_b200 import parallel            # without a process group it is the identity
    d = {"total_loss": torch.tensor(2.0)}
    assert parallel.average_loss_components(d) is d and float(d["total_loss"]) == 2.0


def test_unknown_route_is_refused():
    from clip_lite_b200 import parallel
    with pytest.raises(ValueError):
        parallel.gathered_dense_loss(torch.zeros(4, 8), torch.zeros(4, 8), torch.tensor(0.0), route="ring")


# ------------------------------------------------------------------ peer route: autograd wiring on CPU
class _FakeExchange:
    """CPU stand-in for clip_lite_b200.peer.PeerExchange: the same four-method surface the autograd functions
    use, with the push replaced by a gloo all-gather and the kernels by the oracle-backed stand-ins.  It lets the
    WIRING of the peer routes (argument order, role swapping, row offsets, what is saved for backward, the
    gradient convention) be checked without CUDA IPC; the flag protocol itself is covered by the multi-GPU tests."""

    def __init__(self, rows, dim):
        from tests import _standin_kernels as SK
        self.SK = SK
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.rows, self.dim, self.step = rows, dim, 0
        self.v_all = [None, None]
        self.pushes = 0

    def normalize_push(self, f, g, parity):
        u, inv_f = self.SK.normalize_cast(f)
        v, inv_g = self.SK.normalize_cast(g)
        out = torch.empty(self.world * self.rows, self.dim, dtype=v.dtype)
        dist.all_gather_into_tensor(out, v.contiguous())
        self.v_all[parity] = out
        self.pushes += 1
        return u, inv_f, inv_g

    def dense_fwd(self, u, t, parity, want_grad=True):
        return self.SK.dense_fwd(u, self.v_all[parity], t, row_offset=self.rank * self.rows, want_grad=want_grad)

    def dense_backward(self, f, g, t, gamma, parity, u, inv_f, inv_g, gmat, gdiag):
        """What jsd_peer_dense_backward does: dV partial over all text rows, summed across ranks on the owner."""
        n, m, off = self.world * self.rows, self.rows, self.rank * self.rows
        part = self.SK.dense_bwd_dv(gmat, u, n, t, gamma)
        dv = self._sum_rows(part)[off:off + m].contiguous()       # gloo has no reduce-scatter: all-reduce + slice
        df, dt = self.SK.dense_backward_image_side(f, self.v_all[parity], inv_f, gmat, gdiag, t, gamma, off)
        dg = self.SK.normalize_bwd(g, inv_g, dv, u, 0, gdiag, t, gamma, m)
        return df, dg, dt

    @staticmethod
    def _sum_rows(part):
        tot = part.clone()
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        return tot


def _peer_wiring_worker(rank, world, port, results, route):
    from tests import _standin_kernels
    from clip_lite_b200 import peer
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        peer.K = _standin_kernels
        m = B // world
        ex_v, ex_u = _FakeExchange(m, D), _FakeExchange(m, D)
        out = []
        for seed in (0, 1):                      # two steps: both parities
            f, g = orc.synth_embeddings(B, D, seed=seed, correlated=True)
            fl = f[rank * m:(rank + 1) * m].clone().requires_grad_(True)
            gl = g[rank * m:(rank + 1) * m].clone().requires_grad_(True)
            t = torch.tensor(T, requires_grad=True)
            if route == "symmetric":
                loss, stats = peer._PeerSymmetricFn.apply(fl, gl, t, ex_v, ex_u)
            else:
                loss, stats = peer._PeerDenseFn.apply(fl, gl, t, ex_v)
            (0.5 * loss).backward()
            out.append((loss.detach(), fl.grad, gl.grad, t.grad))
        with torch.no_grad():                    # forward only: both pushes still happen on the symmetric route
            before = (ex_v.pushes, ex_u.pushes)
            if route == "symmetric":
                peer._PeerSymmetricFn.apply(fl.detach(), gl.detach(), t.detach(), ex_v, ex_u)
            else:
                peer._PeerDenseFn.apply(fl.detach(), gl.detach(), t.detach(), ex_v)
            pushes = (ex_v.pushes - before[0], ex_u.pushes - before[1])
        results[rank] = (out, pushes, ex_v.step)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,route", [(2, "reduce"), (2, "symmetric"), (4, "symmetric")])
def test_peer_autograd_wiring_matches_single_process_dense(world, route):
    results = mp.Manager().dict()
    mp.spawn(_peer_wiring_worker, args=(world, _free_port(), results, route), nprocs=world, join=True)
    m = B // world
    for si, seed in enumerate((0, 1)):
        f, g = orc.synth_embeddings(B, D, seed=seed, correlated=True)
        fd, gd = f.double(), g.double()
        df, dg, dt = orc.jsd_dense_grads(fd, gd, T, gamma=0.5)
        tot_dt = 0.0
        for r in range(world):
            loss, gf, gg, gt = results[r][0][si]
            slab = orc.jsd_dense(fd[r * m:(r + 1) * m], gd, T, row_offset=r * m)
            assert abs(float(loss) - float(slab["loss"])) < 1e-3 * float(slab["loss"])
            ref_f, ref_g = world * df[r * m:(r + 1) * m], world * dg[r * m:(r + 1) * m]
            assert (gf.double() - ref_f).abs().max() < 1e-2 * ref_f.abs().max()
            assert (gg.double() - ref_g).abs().max() < 1e-2 * ref_g.abs().max()
            tot_dt += float(gt)
        assert abs(tot_dt / world - float(dt)) < 1e-2 * abs(float(dt))
    for r in range(world):
        _, pushes, steps = results[r]
        assert pushes == ((1, 1) if route == "symmetric" else (1, 0))
        assert steps == 3                       # the exchange's step counter advanced once per forward


# ------------------------------------------------------------------ fused head tail in front of the gathered estimator
def _fused_tail_worker(rank, world, port, results):
    from tests import _emu_backend, _standin_kernels
    from clip_lite_b200 import ops, parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mp_ = pytest.MonkeyPatch()
    try:
        _emu_backend.install(mp_)                  # row-wise entry points -> CPU emulation of the kernel source
        parallel.K = _standin_kernels
        xf, xg, lns = _tail_inputs()
        m = B // world
        xfl = xf[rank * m:(rank + 1) * m].clone().requires_grad_(True)
        xgl = xg[rank * m:(rank + 1) * m].clone().requires_grad_(True)
        t = torch.tensor(T, requires_grad=True)
        f, g = ops.ln_normalize_pair(xfl, xgl, lns[0], lns[1])
        loss, _ = parallel.gathered_dense_loss(f, g, t)
        (0.5 * loss).backward()
        results[rank] = (loss.detach(), xfl.grad, xgl.grad, t.grad,
                         [p.grad.clone() for ln in lns for p in (ln.weight, ln.bias)])
    finally:
        mp_.undo()
        dist.destroy_process_group()


def _tail_inputs():
    gen = torch.Generator().manual_seed(5)
    xf = torch.randn(B, D, generator=gen) * 1.5 + 0.2
    xg = 0.5 * xf + torch.randn(B, D, generator=gen)
    lns = []
    for _ in range(2):
        ln = torch.nn.LayerNorm(D)
        with torch.no_grad():
            ln.weight.copy_(1.0 + 0.3 * torch.randn(D, generator=gen))
            ln.bias.copy_(0.2 * torch.randn(D, generator=gen))
        lns.append(ln)
    return xf, xg, lns


def test_fused_head_tail_in_front_of_the_gathered_loss():
    """fused_heads=True with gather=True: every rank runs the LayerNorm + normalise tail on its own rows and hands
    fp32 unit rows to the gathered estimator.  LayerNorm is row-wise, so the sharded run must reproduce the
    single-process gradients of LayerNorm -> dense estimator on the whole batch: input gradients x world (DDP's
    mean convention), LayerNorm parameter gradients summing over the ranks to world x the global gradient."""
    from tests import _emu_backend
    if not _emu_backend.available():
        pytest.skip("needs g++ and the CUDA headers")
    _emu_backend.build()                           # once, before the ranks race for it
    world = 2
    results = mp.Manager().dict()
    mp.spawn(_fused_tail_worker, args=(world, _free_port(), results), nprocs=world, join=True)
    xf, xg, lns = _tail_inputs()
    leaves = [xf.double().requires_grad_(True), xg.double().requires_grad_(True)]
    params = [p.detach().double().requires_grad_(True) for ln in lns for p in (ln.weight, ln.bias)]
    tt = torch.tensor(T, dtype=torch.float64, requires_grad=True)
    f = torch.nn.functional.layer_norm(leaves[0], (D,), params[0], params[1], lns[0].eps)
    g = torch.nn.functional.layer_norm(leaves[1], (D,), params[2], params[3], lns[1].eps)
    (0.5 * orc.jsd_dense(f, g, tt)["loss"]).backward()
    m = B // world
    psum = [torch.zeros_like(p) for p in params]
    tot_dt = 0.0
    for r in range(world):
        loss, gxf, gxg, gt, gp = results[r]
        for got, leaf in ((gxf, leaves[0]), (gxg, leaves[1])):
            ref = world * leaf.grad[r * m:(r + 1) * m]
            assert (got.double() - ref).abs().max() < 1e-2 * ref.abs().max()
        for acc, p in zip(psum, gp):
            acc += p.double()
        tot_dt += float(gt)
    for acc, p in zip(psum, params):
        assert (acc / world - p.grad).abs().max() < 1e-2 * p.grad.abs().max()
    assert abs(tot_dt / world - float(tt.grad)) < 1e-2 * abs(float(tt.grad))

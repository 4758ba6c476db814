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

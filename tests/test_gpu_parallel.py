"""Multi-GPU parity (NCCL): the gathered dense loss on R ranks equals the single-GPU dense
loss on the concatenated batch (loss, dF, dG, dt), with the gradient convention of
SURVEY 8e.  Skipped on boxes with fewer than two GPUs."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import jsd_oracle as orc

pytestmark = pytest.mark.gpu

B, D = 1024, 256


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, results):
    from clip_lite_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        f, g = orc.synth_embeddings(B, D, seed=0, correlated=True)
        m = B // world
        fl = f[rank * m:(rank + 1) * m].cuda().requires_grad_(True)
        gl = g[rank * m:(rank + 1) * m].cuda().requires_grad_(True)
        t = torch.tensor(orc.T_INIT, device="cuda", requires_grad=True)
        loss, _ = parallel.gathered_dense_loss(fl, gl, t)
        (0.5 * loss).backward()
        logged = parallel.global_loss_for_logging(loss)
        torch.cuda.synchronize()
        results[rank] = tuple(x.detach().cpu() for x in (loss, fl.grad, gl.grad, t.grad, logged))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_gathered_dense_matches_single_gpu(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    results = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), results), nprocs=world, join=True)
    f, g = orc.synth_embeddings(B, D, seed=0, correlated=True)
    fd, gd = f.double(), g.double()
    full = orc.jsd_dense(fd, gd, orc.T_INIT)
    df, dg, dt = orc.jsd_dense_grads(fd, gd, orc.T_INIT, gamma=0.5)
    m = B // world
    tot_dt = 0.0
    for r in range(world):
        loss, gf, gg, gt, logged = results[r]
        slab = orc.jsd_dense(fd[r * m:(r + 1) * m], gd, orc.T_INIT, row_offset=r * m)
        assert abs(float(loss) - float(slab["loss"])) < 1e-3 * float(slab["loss"])
        assert abs(float(logged) - float(full["loss"])) < 1e-3 * float(full["loss"])
        ref_f, ref_g = world * df[r * m:(r + 1) * m], world * dg[r * m:(r + 1) * m]
        assert (gf.double() - ref_f).abs().max() < 1e-2 * ref_f.abs().max()
        assert (gg.double() - ref_g).abs().max() < 1e-2 * ref_g.abs().max()
        tot_dt += float(gt)
    assert abs(tot_dt / world - float(dt)) < 1e-2 * abs(float(dt))

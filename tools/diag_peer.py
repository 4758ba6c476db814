#!/usr/bin/env python
"""Stage-by-stage timing of the sharded step (torchrun, N GPUs): each launch bracketed by CUDA events, so a
consumer's time includes its wait for the peers.  Development aid; bench.py is the contract."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clip_lite_b200 import kernels as K, peer  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    batch, dim = int(sys.argv[1]), int(sys.argv[2])
    m = batch // world
    f = torch.randn(m, dim, device="cuda").bfloat16()
    g = torch.randn(m, dim, device="cuda").bfloat16()
    t = torch.tensor(2.6593, device="cuda")
    gamma = torch.ones((), device="cuda")
    ex = peer.get_exchange(m, dim)
    names = ["push", "fwd", "dV", "dU+imgJ", "txtJ(pull)"]
    iters = 30
    acc = [0.0] * len(names)
    tot = 0.0
    for it in range(iters + 5):
        dist.barrier()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
        p = ex.step & 1
        ex.step += 1
        ev[0].record()
        u, inv_f, inv_g = ex.normalize_push(f, g, p); ev[1].record()
        out4, loss, gmat, gdiag = ex.dense_fwd(u, t, p); ev[2].record()
        ex.dense_bwd_dv(gmat, u, t, gamma); ev[3].record()
        df, dt = K.dense_backward_image_side(f, ex.v_all[p], inv_f, gmat, gdiag, t, gamma, rank * m); ev[4].record()
        dg = ex.normalize_bwd_text(g, inv_g, u, gdiag, t, gamma); ev[5].record()
        torch.cuda.synchronize()
        if it >= 5:
            for i in range(len(names)):
                acc[i] += ev[i].elapsed_time(ev[i + 1])
            tot += ev[0].elapsed_time(ev[-1])
    if rank == 0:
        print(f"== peer step B={batch} D={dim} world={world} (eager launches, events per stage)")
        for n, a in zip(names, acc):
            print(f"  {n:12s} {a / iters * 1e3:8.1f} us")
        print(f"  {'total':12s} {tot / iters * 1e3:8.1f} us")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

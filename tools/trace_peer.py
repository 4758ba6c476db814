#!/usr/bin/env python
"""Kernel timeline of graph-replayed peer-exchange steps, per rank (needs the trace build of the library:
build.build_library(trace=True), JSD_LIB=.../libjsd_b200_trace.so).  torchrun, N GPUs:  trace_peer.py B D [route]"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clip_lite_b200 import peer  # noqa: E402
from clip_lite_b200.trace import KernelTrace  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    batch, dim = int(sys.argv[1]), int(sys.argv[2])
    route = sys.argv[3] if len(sys.argv) > 3 else "reduce"
    m = batch // world
    f = torch.randn(m, dim, device="cuda").bfloat16()
    g = torch.randn(m, dim, device="cuda").bfloat16()
    t = torch.tensor(2.6593, device="cuda")
    gs = peer.PeerGraphedStep(f, g, t, route=route)
    for _ in range(10):
        gs()
    torch.cuda.synchronize()
    dist.barrier()
    with KernelTrace() as tr:
        for _ in range(4):
            gs()
        torch.cuda.synchronize()
        text = tr.summary()
    for r in range(world):
        dist.barrier()
        if r == rank and r in (0, world - 1):
            print(f"== rank {rank} of {world}, B={batch} D={dim} route={route}: 4 graph-replayed steps\n{text}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

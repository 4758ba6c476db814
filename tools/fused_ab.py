#!/usr/bin/env python
"""A/B of the single-pass fused kernel (D <= 256: score tile + accumulators resident in TMEM, nothing B x B in HBM)
against the staged path (forward writes the bf16 sigma matrix, two contraction launches read it back).

    python tools/fused_ab.py                 graph-replayed fwd+bwd step times for both routes, several (B, D)
    python tools/fused_ab.py once B D MODE   one eager step (MODE = 1 fused / 0 staged) -- run under
                                             ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clip_lite_b200 import ops  # noqa: E402
from clip_lite_b200.graph import GraphedStep  # noqa: E402


def inputs(b, d):
    gen = torch.Generator("cpu").manual_seed(0)
    f0 = torch.randn(b, d, generator=gen)
    g0 = 0.6 * f0 + 0.8 * torch.randn(b, d, generator=gen)
    f = torch.nn.functional.normalize(f0, dim=-1).bfloat16().cuda()
    g = torch.nn.functional.normalize(g0, dim=-1).bfloat16().cuda()
    return f, g, torch.tensor(2.659260036932778, device="cuda", requires_grad=True)


def step_time(b, d, mode, iters=50):
    os.environ["JSD_FUSED"] = mode
    f, g, t = inputs(b, d)
    gs = GraphedStep(lambda ff, gg, tt: ops.jsd_dense_loss(ff, gg, tt), f, g, t)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(5):
        gs()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gs()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    loss = float(gs.loss.detach())
    return tot / iters * 1e3, loss, gs


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "once":
        b, d, mode = int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
        os.environ["JSD_FUSED"] = mode
        f, g, t = inputs(b, d)
        f.requires_grad_(True)
        g.requires_grad_(True)
        for _ in range(2):
            loss, _ = ops.jsd_dense_loss(f, g, t)
            torch.autograd.grad(loss, (f, g, t))
        torch.cuda.synchronize()
        print("once", b, d, "fused" if mode == "1" else "staged", float(loss))
        return
    print("B      D    fused us  staged us  speed-up   6B^2D TFLOP/s (fused)   B x B bytes avoided per step")
    for b, d in ((1024, 128), (1024, 256), (1024, 64), (4096, 128), (8192, 128), (8192, 256), (16384, 128)):
        tf, lf, _ = step_time(b, d, "1")
        ts, ls, _ = step_time(b, d, "0")
        assert abs(lf - ls) < 1e-4 * abs(ls), (lf, ls)
        print(f"{b:6d} {d:4d} {tf:9.1f} {ts:10.1f} {ts / tf:9.2f}x {6.0 * b * b * d / (tf * 1e-6) / 1e12:12.1f}"
              f" {3 * b * b * 2 / 1e6:24.0f} MB")


if __name__ == "__main__":
    main()

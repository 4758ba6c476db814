#!/usr/bin/env python
"""A/B of the fused projection-head tail (jsd_heads.cuh) against the default route (nn.LayerNorm + the
estimator's own normalisation): eager fwd+bwd times of the estimator INCLUDING the heads' tail, on the pre-LayerNorm
head outputs, plus parity of the two routes.  First thing to run when a GPU is available again:

    python tools/heads_ab.py                     index and dense mode at the heads' width (D = 2048), several B
    python tools/heads_ab.py once MODE B FUSED   one eager step (MODE = index | dense, FUSED = 0 | 1) -- run under
        ncu --set full --clock-control none -k regex:ln_normalize -c 6   (HBM roofline of the two tail kernels:
        algorithmic bytes per element in DESIGN 4.6)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clip_lite_b200 import ops  # noqa: E402

D = 2048


def inputs(b, dtype=torch.float32):
    gen = torch.Generator("cpu").manual_seed(0)
    x0 = torch.randn(b, D, generator=gen)
    xf = (1.5 * x0 + 0.2).to(dtype).cuda().requires_grad_(True)
    xg = (0.9 * x0 + 1.2 * torch.randn(b, D, generator=gen)).to(dtype).cuda().requires_grad_(True)
    lns = [torch.nn.LayerNorm(D).cuda() for _ in range(2)]
    t = torch.tensor(2.659260036932778, device="cuda", requires_grad=True)
    return xf, xg, lns[0], lns[1], t


def step(mode, fused, xf, xg, ln_f, ln_g, t):
    if fused and mode == "dense":
        loss, _ = ops.jsd_dense_loss_ln(xf, xg, ln_f, ln_g, t)
    else:
        f, g = ops.ln_normalize_pair(xf, xg, ln_f, ln_g) if fused else (ln_f(xf.float()), ln_g(xg.float()))
        loss, _ = (ops.jsd_dense_loss if mode == "dense" else ops.jsd_index_loss)(f, g, t)
    params = (xf, xg, ln_f.weight, ln_f.bias, ln_g.weight, ln_g.bias, t)
    return loss, torch.autograd.grad(loss, params)


def timed(mode, fused, b, iters=30):
    args = inputs(b)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        out = step(mode, fused, *args)
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = step(mode, fused, *args)
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1e3, out


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "once":
        mode, b, fused = sys.argv[2], int(sys.argv[3]), sys.argv[4] == "1"
        args = inputs(b)
        for _ in range(2):
            loss, _ = step(mode, fused, *args)
        torch.cuda.synchronize()
        print("once", mode, b, "fused" if fused else "default", float(loss))
        return
    print("mode   B      fused us  default us  speed-up   worst rel. difference of (loss, grads) between the routes")
    for mode, sizes in (("index", (1024, 8192)), ("dense", (1024, 4096, 8192))):
        for b in sizes:
            tf, (lf, gf) = timed(mode, True, b)
            td, (ld, gd) = timed(mode, False, b)
            worst = max([rel(lf, ld)] + [rel(a, c) for a, c in zip(gf, gd)])
            print(f"{mode:6s} {b:6d} {tf:9.1f} {td:11.1f} {td / tf:9.2f}x   {worst:.2e}")


if __name__ == "__main__":
    main()

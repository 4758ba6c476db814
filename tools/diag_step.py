#!/usr/bin/env python
"""Where does the step time go?  (development aid)"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clip_lite_b200 import kernels as K, ops
import bench

b, d = 8192, 1024
f, g = bench.synth(b, d)
f = f.cuda().requires_grad_(True); g = g.cuda().requires_grad_(True)
t = torch.tensor(bench.T_INIT, device="cuda", requires_grad=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
gamma = torch.ones((), device="cuda")

def step_autograd():
    loss = ops.jsd_dense_loss(f, g, t)[0]
    return torch.autograd.grad(loss, (f, g, t))

def step_kernels():
    fd, gd, td = f.detach(), g.detach(), t.detach()
    u, inv_f = K.normalize_cast(fd); v, inv_g = K.normalize_cast(gd)
    out4, _, gmat, gdiag = K.dense_fwd(u, v, td)
    du = K.dense_bwd_du(gmat, v, td, gamma); dv = K.dense_bwd_dv(gmat, u, b, td, gamma)
    K.normalize_bwd(fd, inv_f, du, v, 0, gdiag, td, gamma, b); K.normalize_bwd(gd, inv_g, dv, u, 0, gdiag, td, gamma, b)

def run(fn, n, flush_l2, per_iter_events):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    if per_iter_events:
        st = [torch.cuda.Event(enable_timing=True) for _ in range(n)]; en = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
        h0 = time.perf_counter()
        for i in range(n):
            if flush_l2: flush.zero_()
            st[i].record(); fn(); en[i].record()
        host = (time.perf_counter() - h0) / n
        torch.cuda.synchronize()
        return sum(s.elapsed_time(e) for s, e in zip(st, en)) / n * 1e3, host * 1e6
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    h0 = time.perf_counter(); s.record()
    for i in range(n): fn()
    e.record(); host = (time.perf_counter() - h0) / n
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3, host * 1e6

for name, fn in (("autograd", step_autograd), ("kernels", step_kernels)):
    for n in (20, 200):
        for fl, pe in ((False, False), (False, True), (True, True)):
            gpu, host = run(fn, n, fl, pe)
            print(f"{name:9s} n={n:4d} flush={int(fl)} per_iter_events={int(pe)}  gpu {gpu:7.1f} us/step   host enqueue {host:7.1f} us/step")

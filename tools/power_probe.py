#!/usr/bin/env python
"""Steady-state power / clock of the tensor-core kernels vs cuBLAS (development aid)."""
import os, subprocess, sys, time, threading
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clip_lite_b200 import kernels as K

def sample(stop, out):
    while not stop.is_set():
        r = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown,temperature.gpu", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True).stdout.strip()
        out.append(r)
        time.sleep(0.05)

def run(name, fn, flops, seconds=2.5):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    stop, out = threading.Event(), []
    th = threading.Thread(target=sample, args=(stop, out)); th.start()
    t0 = time.perf_counter(); n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.perf_counter() - t0 < seconds:
        for _ in range(50): fn()
        n += 50
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    stop.set(); th.join()
    us = e0.elapsed_time(e1) / n * 1e3
    tail = out[len(out)//2:]
    clocks = [float(x.split(",")[0]) for x in tail if x]; power = [float(x.split(",")[1]) for x in tail if x]
    print(f"{name:28s} {us:8.1f} us  {flops/(us*1e-6)/1e12:7.1f} TFLOP/s  clk {sum(clocks)/len(clocks):6.0f} MHz  power {sum(power)/len(power):6.0f} W  last: {tail[-1]}")

b, d = 8192, 1024
f = torch.randn(b, d, device="cuda").bfloat16(); g = torch.randn(b, d, device="cuda").bfloat16()
t = torch.tensor(2.6593, device="cuda"); gamma = torch.ones((), device="cuda")
u, _ = K.normalize_cast(f); v, _ = K.normalize_cast(g)
_, _, gmat, _ = K.dense_fwd(u, v, t)
gm = gmat[:, :b].contiguous()
fl = 2.0 * b * b * d
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which == "all":
    run("cuBLAS U@V.T", lambda: torch.matmul(u, v.t()), fl)
    run("cuBLAS Gmat@V", lambda: torch.matmul(gm, v), fl)
    run("jsd fwd (loss only)", lambda: K.dense_fwd(u, v, t, want_grad=False), fl)
    run("jsd fwd (+Gmat)", lambda: K.dense_fwd(u, v, t), fl)
if which in ("all", "grad"):
    run("jsd bwd_du", lambda: K.dense_bwd_du(gmat, v, t, gamma), fl)
    run("jsd bwd_dv", lambda: K.dense_bwd_dv(gmat, u, b, t, gamma), fl)
if which == "layouts":
    m, n, k = 8192, 1024, 8192
    a = torch.randn(m, k, device="cuda").bfloat16(); at = a.t().contiguous()
    bb = torch.randn(n, k, device="cuda").bfloat16(); bt = bb.t().contiguous()
    for a_mn in (False, True):
        for b_mn in (False, True):
            run(f"gemm a_mn={int(a_mn)} b_mn={int(b_mn)}", lambda: K.gemm_bf16(at if a_mn else a, bt if b_mn else bb, a_mn_major=a_mn, b_mn_major=b_mn, stream_k=False), 2.0 * m * n * k)
    run("gemm a_mn=0 b_mn=1 stream-K", lambda: K.gemm_bf16(a, bt, b_mn_major=True, stream_k=True), 2.0 * m * n * k)
    run("cuBLAS a@bt", lambda: torch.matmul(a, bt), 2.0 * m * n * k)
    run("cuBLAS a@bb.T", lambda: torch.matmul(a, bb.t()), 2.0 * m * n * k)

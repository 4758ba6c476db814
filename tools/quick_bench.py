#!/usr/bin/env python
"""Stage-by-stage timing of the dense pipeline (development aid; bench.py is the contract)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clip_lite_b200 import kernels as K  # noqa: E402


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    return ts[len(ts) // 2], ts[0]


def main():
    shapes = [(8192, 1024), (1024, 1024), (1024, 128), (16384, 512)]
    if len(sys.argv) > 2:
        shapes = [(int(sys.argv[1]), int(sys.argv[2]))]
    for b, d in shapes:
        f = torch.randn(b, d, device="cuda").bfloat16()
        g = (0.6 * f.float() + 0.8 * torch.randn(b, d, device="cuda")).bfloat16()
        t = torch.tensor(2.6593, device="cuda")
        gamma = torch.tensor(0.9, device="cuda")
        u, inv_f = K.normalize_cast(f)
        v, inv_g = K.normalize_cast(g)
        out4, _, gmat, gdiag = K.dense_fwd(u, v, t)
        du = K.dense_bwd_du(gmat, v, t, gamma)
        dv = K.dense_bwd_dv(gmat, u, b, t, gamma)
        flops = 2.0 * b * b * d

        def full():
            u, inv_f = K.normalize_cast(f)
            v, inv_g = K.normalize_cast(g)
            out4, _, gmat, gdiag = K.dense_fwd(u, v, t)
            du = K.dense_bwd_du(gmat, v, t, gamma)
            dv = K.dense_bwd_dv(gmat, u, b, t, gamma)
            K.normalize_bwd(f, inv_f, du, v, 0, gdiag, t, gamma, b)
            K.normalize_bwd(g, inv_g, dv, u, 0, gdiag, t, gamma, b)

        stages = {
            "normalize_cast x2": lambda: (K.normalize_cast(f), K.normalize_cast(g)),
            "dense_fwd": lambda: K.dense_fwd(u, v, t),
            "dense_fwd(loss only)": lambda: K.dense_fwd(u, v, t, want_grad=False),
            "bwd_du": lambda: K.dense_bwd_du(gmat, v, t, gamma),
            "bwd_dv": lambda: K.dense_bwd_dv(gmat, u, b, t, gamma),
            "normalize_bwd x2": lambda: (K.normalize_bwd(f, inv_f, du, v, 0, gdiag, t, gamma, b),
                                         K.normalize_bwd(g, inv_g, dv, u, 0, gdiag, t, gamma, b)),
            "FULL fwd+bwd": full,
        }
        print(f"== B={b} D={d}  (2*B*B*D = {flops/1e9:.1f} GFLOP per GEMM)")
        for name, fn in stages.items():
            med, mn = timeit(fn)
            extra = ""
            if name in ("dense_fwd", "dense_fwd(loss only)", "bwd_du", "bwd_dv"):
                extra = f"  {flops/ (med*1e-3)/1e12:8.1f} TFLOP/s"
            if name == "FULL fwd+bwd":
                extra = f"  {3*flops/(med*1e-3)/1e12:8.1f} TFLOP/s (6B^2D)  {b/(med*1e-3)/1e6:.2f} Mpairs/s"
            print(f"  {name:24s} median {med*1e3:9.1f} us   min {mn*1e3:9.1f} us{extra}")
        # cuBLAS reference for the same GEMM shape
        a32 = u
        med, mn = timeit(lambda: torch.matmul(a32, v.t()))
        print(f"  {'torch.matmul U@V.T bf16':24s} median {med*1e3:9.1f} us   min {mn*1e3:9.1f} us  {flops/(med*1e-3)/1e12:8.1f} TFLOP/s")


if __name__ == "__main__" and not (len(sys.argv) > 1 and sys.argv[1] in ("gemm", "index")):
    main()


def gemm_variants():
    """A/B-layout and stream-K variants of the backward-GEMM shape (development aid)."""
    m, n, k = 8192, 1024, 8192
    a = torch.randn(m, k, device="cuda").bfloat16()
    at = a.t().contiguous()
    b = torch.randn(n, k, device="cuda").bfloat16()
    bt = b.t().contiguous()
    flops = 2.0 * m * n * k
    print(f"== GEMM variants M={m} N={n} K={k}")
    for a_mn in (False, True):
        for b_mn in (False, True):
            for sk in (False, True):
                fn = lambda: K.gemm_bf16(at if a_mn else a, bt if b_mn else b, a_mn_major=a_mn, b_mn_major=b_mn, stream_k=sk)
                med, mn = timeit(fn)
                print(f"  a_mn={int(a_mn)} b_mn={int(b_mn)} stream_k={int(sk)}  median {med*1e3:8.1f} us  min {mn*1e3:8.1f} us  {flops/(med*1e-3)/1e12:7.1f} TFLOP/s")
    med, mn = timeit(lambda: torch.matmul(a, bt))
    print(f"  torch.matmul                     median {med*1e3:8.1f} us  min {mn*1e3:8.1f} us  {flops/(med*1e-3)/1e12:7.1f} TFLOP/s")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "gemm":
    gemm_variants()


def index_mode():
    """Reference-semantics (index) kernel vs the HBM roofline."""
    for b, d, dt in ((8192, 2048, torch.float32), (8192, 2048, torch.bfloat16), (1024, 2048, torch.float32), (65536, 512, torch.bfloat16)):
        f = torch.randn(b, d, device="cuda").to(dt)
        g = torch.randn(b, d, device="cuda").to(dt)
        t = torch.tensor(2.6593, device="cuda")
        for _ in range(3):
            K.index_fwd_bwd(f, g, t)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):                      # back to back: device throughput, host overhead amortised
            K.index_fwd_bwd(f, g, t)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        nbytes = 4.0 * b * d * f.element_size()
        print(f"  index B={b} D={d} {str(dt):15s} {us:8.1f} us/call (incl. finalize)  {nbytes/(us*1e-6)/1e9:8.1f} GB/s algorithmic")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "index":
    index_mode()

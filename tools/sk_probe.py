#!/usr/bin/env python
"""Stream-K vs whole tiles on underfilled GRAD shapes (back-to-back launches, host overhead amortised)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clip_lite_b200 import kernels as K  # noqa: E402


from sk_probe_util import loop_time  # noqa: E402


for (m, n, k) in ((1024, 1024, 1024), (1024, 1024, 2048), (1024, 1024, 8192), (2048, 1024, 8192), (4096, 1024, 8192)):
    a = torch.randn(m, k, device="cuda").bfloat16()
    bt = torch.randn(k, n, device="cuda").bfloat16()          # B^T stored [K, N]: the dU layout (A K-major, B MN-major)
    res = {}
    for sk in (False, True):
        res[sk] = loop_time(lambda: K.gemm_bf16(a, bt, a_mn_major=False, b_mn_major=True, stream_k=sk))
    ref = loop_time(lambda: torch.matmul(a, bt))
    print(f"M={m} N={n} K={k}: whole tiles {res[False]:7.1f} us   stream-K {res[True]:7.1f} us   cuBLAS {ref:7.1f} us")

#!/usr/bin/env python
"""BASELINE configs[3] per-rank work on ONE GPU: an 8192-row slab against 65536 text rows, D = 512 (what a rank of an
8-GPU run computes; pre-gathered text rows, no exchange).  Graph-replayed stage times and the 6 M N D rate."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clip_lite_b200 import kernels as K  # noqa: E402
from sk_probe_util import loop_time  # noqa: E402

m, n, d, off = 8192, 65536, 512, 3 * 8192
g_all = torch.randn(n, d, device="cuda").bfloat16()
f = (0.6 * g_all[off:off + m].float() + 0.8 * torch.randn(m, d, device="cuda")).bfloat16()
g = g_all[off:off + m].contiguous()
t = torch.tensor(2.6593, device="cuda")
gamma = torch.tensor(1.0, device="cuda")
v_all, _ = K.normalize_cast(g_all)
u, v, inv_f, inv_g = K.normalize_cast_pair(f, g)
out4, loss, gmat, gdiag = K.dense_fwd(u, v_all, t, row_offset=off)
print(f"slab {m} x {n}, D={d}: Gmat {gmat.numel() * 2 / 1e9:.2f} GB bf16 (fp32 scores would be {m * n * 4 / 1e9:.1f} GB); "
      f"loss {float(loss):.5f}")
stages = {
    "normalize pair": lambda: K.normalize_cast_pair(f, g),
    "fwd slab": lambda: K.dense_fwd(u, v_all, t, row_offset=off),
    "dV partial": lambda: K.dense_bwd_dv(gmat, u, n, t, gamma),
    "dU + image J": lambda: K.dense_backward_image_side(f, v_all, inv_f, gmat, gdiag, t, gamma, off),
}
tot = 0.0
for name, fn in stages.items():
    us = loop_time(fn, iters=5)
    tot += us
    print(f"  {name:16s} {us:9.1f} us")
flops = 6.0 * m * n * d
print(f"  sum              {tot:9.1f} us  -> {flops / (tot * 1e-6) / 1e12:.0f} TFLOP/s of 6 M N D "
      f"({flops / 1e12:.3f} TFLOP per rank and step), {8 * m / (tot * 1e-6) / 1e6:.1f} M pairs/s for 8 such ranks "
      f"before exchange")

#!/usr/bin/env python
"""One rank's row slab on ONE GPU (pre-gathered text rows, no exchange): per-stage times of the shapes a rank sees
at 2/4/8 GPUs for global B = 8192, D = 1024.  Development aid (split-K A/B: JSD_SPLITK=0)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clip_lite_b200 import kernels as K  # noqa: E402
from sk_probe_util import loop_time  # noqa: E402


def main():
    n, d = 8192, 1024
    t = torch.tensor(2.6593, device="cuda")
    gamma = torch.tensor(1.0, device="cuda")
    g_all = torch.randn(n, d, device="cuda").bfloat16()
    v_all, _ = K.normalize_cast(g_all)
    for world in (2, 4, 8):
        m = n // world
        f = torch.randn(m, d, device="cuda").bfloat16()
        g = g_all[:m].contiguous()
        u, v, inv_f, inv_g = K.normalize_cast_pair(f, g)
        out4, loss, gmat, gdiag = K.dense_fwd(u, v_all, t, row_offset=0)
        stages = {
            "normalize pair": lambda: K.normalize_cast_pair(f, g),
            "fwd slab": lambda: K.dense_fwd(u, v_all, t, row_offset=0),
            "dV partial": lambda: K.dense_bwd_dv(gmat, u, n, t, gamma),
            "dU + image J": lambda: K.dense_backward_image_side(f, v_all, inv_f, gmat, gdiag, t, gamma, 0),
        }
        print(f"== rows/rank {m} (world {world}), N={n}, D={d}  JSD_SPLITK={os.environ.get('JSD_SPLITK', '1')}")
        tot = 0.0
        for name, fn in stages.items():
            us = loop_time(fn, iters=20)
            tot += us
            print(f"  {name:16s} {us:8.1f} us  (graph replay of 20 back-to-back launches, warm L2)")
        print(f"  {'sum':16s} {tot:8.1f} us")


if __name__ == "__main__":
    main()

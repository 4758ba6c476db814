#!/usr/bin/env python
"""One small step of every library path for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):

    compute-sanitizer --tool racecheck python tools/sanitize_step.py [dense|index|score|heads|all]
    compute-sanitizer --tool memcheck  python -m torch.distributed.run --nproc-per-node 2 ... tools/sanitize_step.py peer

Sizes are small (the tools slow kernels down 10-100x) but cover edge tiles (B not a multiple of 256), both CTA-group
sizes, split-K, the helper-stream overlap and the ticketed last-CTA reductions.  Development aid; the logs go to
profiles/sanitizer_*.log."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clip_lite_b200 import ops  # noqa: E402


def dense(b, d, dtype=torch.bfloat16):
    f = torch.randn(b, d, device="cuda").to(dtype).requires_grad_(True)
    g = torch.randn(b, d, device="cuda").to(dtype).requires_grad_(True)
    t = torch.tensor(2.6593, device="cuda", requires_grad=True)
    loss, _ = ops.jsd_dense_loss(f, g, t)
    loss.backward()
    torch.cuda.synchronize()
    print(f"dense B={b} D={d} loss={float(loss):.5f} |dF|={float(f.grad.float().norm()):.4e} dt={float(t.grad):.4e}")


def index(b, d, dtype=torch.float32):
    f = torch.randn(b, d, device="cuda").to(dtype).requires_grad_(True)
    g = torch.randn(b, d, device="cuda").to(dtype).requires_grad_(True)
    t = torch.tensor(2.6593, device="cuda", requires_grad=True)
    loss, _ = ops.jsd_index_loss(f, g, t)
    loss.backward()
    loss2, _ = ops.jsd_index_loss(f, g, t, ops.NegativeIndex.cluster(b // 2))
    loss2.backward()
    torch.cuda.synchronize()
    print(f"index B={b} D={d} loss={float(loss):.5f} cluster={float(loss2):.5f}")


def score():
    from clip_lite_b200 import retrieval
    img = torch.randn(96, 128, device="cuda")
    txt = torch.randn(288, 128, device="cuda")
    rows = [list(range(3 * i, 3 * i + 3)) for i in range(96)]
    cols = torch.arange(288, dtype=torch.int32) // 3
    r1, r2 = retrieval.retrieval_ranks(img, txt, rows, cols, normalize=True)
    val, col = retrieval.score_argmax(img, txt, normalize=True)
    torch.cuda.synchronize()
    print("score ranks", int(r1.sum()), int(r2.sum()), int(col.sum()))


def heads(b, d, mode, dtype=torch.float32):
    """Projection-head tail (jsd_heads.cuh): LayerNorm + normalise fused, forward and backward.  NOT part of "all":
    these kernels were written after the round-2 logs under profiles/ were taken; run `heads` explicitly."""
    xf = torch.randn(b, d, device="cuda").to(dtype).requires_grad_(True)
    xg = torch.randn(b, d, device="cuda").to(dtype).requires_grad_(True)
    ln_f, ln_g = torch.nn.LayerNorm(d).cuda(), torch.nn.LayerNorm(d).cuda()
    t = torch.tensor(2.6593, device="cuda", requires_grad=True)
    if mode == "dense":
        loss, _ = ops.jsd_dense_loss_ln(xf, xg, ln_f, ln_g, t)
    else:
        f, g = ops.ln_normalize_pair(xf, xg, ln_f, ln_g)
        loss, _ = ops.jsd_index_loss(f, g, t)
    loss.backward()
    torch.cuda.synchronize()
    print(f"heads {mode} B={b} D={d} loss={float(loss.detach()):.5f} |dX|={float(xf.grad.float().norm()):.4e} "
          f"|dw|={float(ln_f.weight.grad.norm()):.4e}")


def peer():
    import torch.distributed as dist
    from clip_lite_b200 import peer as P
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    b, d = 512, 256
    m = b // world
    t = torch.tensor(2.6593, device="cuda", requires_grad=True)
    for step in range(3):
        f = torch.randn(m, d, device="cuda").bfloat16().requires_grad_(True)
        g = torch.randn(m, d, device="cuda").bfloat16().requires_grad_(True)
        loss, _ = P.peer_dense_loss(f, g, t)
        loss.backward()
    torch.cuda.synchronize()
    print(f"peer rank {rank}/{world} loss={float(loss):.5f}")
    dist.barrier()
    dist.destroy_process_group()


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    torch.manual_seed(0)
    if what in ("dense", "all"):
        dense(512, 256)             # CTA pairs, whole tiles
        dense(200, 64)              # edge tiles, ragged rows
        dense(128, 128, torch.float32)   # single CTA (CG = 1)
        dense(1024, 128)            # split-K backward
    if what in ("index", "all"):
        index(300, 200)
        index(256, 512, torch.bfloat16)
    if what in ("score", "all"):
        score()
    if what == "heads":
        heads(300, 2048, "index")                     # register-resident forward, two 16-byte pieces per thread backward
        heads(96, 102, "index", torch.bfloat16)       # element-wise variants
        heads(256, 128, "dense")                      # fused single-pass kernel's unscaled accumulator slices
        heads(200, 2176, "dense")                     # staged path, 4-piece backward, generic forward
    if what == "peer":
        peer()


if __name__ == "__main__":
    main()

"""Data-parallel dense estimator: each rank scores its rows against the global batch.

Not in the reference (its loss is per-rank); this is the north-star
path of BASELINE.json.  One process per GPU, ``torch.distributed`` (NCCL over
NVLink/NVSwitch) for the two exchange steps the path really has:

  forward   V_r --all-gather--> V_all [B, D] bf16; rank r computes the row slab
            S_r = tau U_r V_all^T  (positives on column r*M + i) and its local
            loss L_r; nothing else crosses ranks.
  backward  dU_r = tau G_r V_all is complete locally;  dV_all^(r) = tau G_r^T U_r is
            a partial over all text rows --reduce-scatter(sum)--> dV_r.

Gradient convention: every rank back-propagates its OWN slab loss
L_r; the reduce-scatter sums the partials, so DDP's later mean over ranks of the
parameter gradients yields the gradient of mean_r L_r, the global loss.

All kernel work goes through ``clip_lite_b200.kernels`` (module attribute ``K``
so that the CPU gloo tests can substitute an oracle-backed stand-in for the
collective/sharding logic; the product never does).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from . import kernels as K


def _world(group) -> tuple:
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def _all_gather_rows(x: torch.Tensor, world: int, group) -> torch.Tensor:
    out = torch.empty(world * x.shape[0], x.shape[1], dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x.contiguous(), group=group)
    return out


ROUTES = ("reduce", "symmetric")


class _GatheredDenseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f, g, t, group, route="reduce"):
        rank, world = _world(group)
        need_grad = any(ctx.needs_input_grad)
        if f.dim() != 2 or f.shape != g.shape:
            raise ValueError(f"features must both be [B_local, D]; got {tuple(f.shape)} and {tuple(g.shape)}")
        m = f.shape[0]
        with torch.autocast(f.device.type, enabled=False):
            dt = torch.promote_types(f.dtype, g.dtype)
            fc, gc = f.to(dt).contiguous(), g.to(dt).contiguous()
            u, v, inv_f, inv_g = K.normalize_cast_pair(fc, gc)
            v_all = _all_gather_rows(v, world, group) if world > 1 else v
            out4, loss, gmat, gdiag = K.dense_fwd(u, v_all, t, row_offset=rank * m, want_grad=need_grad)
        if need_grad:
            ctx.save_for_backward(fc, gc, t, u, v_all, inv_f, inv_g, gmat, gdiag)
        ctx.group, ctx.rank, ctx.world, ctx.route = group, rank, world, route
        ctx.dtypes = (f.dtype, g.dtype, t.dtype)
        ctx.mark_non_differentiable(out4)
        return loss, out4

    @staticmethod
    def backward(ctx, grad_loss, _grad_stats):
        fc, gc, t, u, v_all, inv_f, inv_g, gmat, gdiag = ctx.saved_tensors
        m, n = fc.shape[0], v_all.shape[0]
        if ctx.route == "symmetric" and ctx.world > 1:
            return _GatheredDenseFn._backward_symmetric(ctx, grad_loss)
        with torch.autocast(fc.device.type, enabled=False):
            gamma = grad_loss.float()
            # text side first: its reduce-scatter runs on NCCL's stream while the image side computes
            dv_partial = K.dense_bwd_dv(gmat, u, n, t, gamma)                  # [N, D], partial over ranks
            work = None
            if ctx.world > 1:
                dv = torch.empty(m, dv_partial.shape[1], dtype=dv_partial.dtype, device=dv_partial.device)
                work = dist.reduce_scatter_tensor(dv, dv_partial, op=dist.ReduceOp.SUM, group=ctx.group,
                                                  async_op=True)
            else:
                dv = dv_partial
            # image side: dU_r is complete on this rank; dL_r/dt = sum_i <u_i, dU_i> falls out of it
            df, dt = K.dense_backward_image_side(fc, v_all, inv_f, gmat, gdiag, t, gamma, ctx.rank * m)
            if work is not None:
                work.wait()                                                    # compute stream waits for the reduction
            dg = K.normalize_bwd(gc, inv_g, dv, u, 0, gdiag, t, gamma, m)
        fd, gd, td = ctx.dtypes
        return df.to(fd), dg.to(gd), dt.to(td), None, None

    @staticmethod
    def _backward_symmetric(ctx, grad_loss):
        """No gradient traffic: the image rows are gathered as well and this rank recomputes the COLUMN slab it
        owns, sigma(tau V_r U_all^T) = (G[:, own columns])^T, with the forward kernel (roles of the two modalities
        swapped, same row offset); dV_r = scale G[:, own]^T U_all is then the image-side backward with the roles
        swapped.  Costs one more slab forward, saves the reduce-scatter of a [B, D] fp32 partial."""
        fc, gc, t, u, v_all, inv_f, inv_g, gmat, gdiag = ctx.saved_tensors
        m = fc.shape[0]
        off = ctx.rank * m
        with torch.autocast(fc.device.type, enabled=False):
            gamma = grad_loss.float()
            u_all = _all_gather_rows(u, ctx.world, ctx.group)
            df, dt = K.dense_backward_image_side(fc, v_all, inv_f, gmat, gdiag, t, gamma, off)
            v = v_all[off:off + m]
            _, _, gmat_t, gdiag_t = K.dense_fwd(v, u_all, t, row_offset=off, want_grad=True)
            dg, _ = K.dense_backward_image_side(gc, u_all, inv_g, gmat_t, gdiag_t, t, gamma, off)
        fd, gd, td = ctx.dtypes
        return df.to(fd), dg.to(gd), dt.to(td), None, None


def gathered_dense_loss(f: torch.Tensor, g: torch.Tensor, t: torch.Tensor, group: Optional[object] = None,
                        route: str = "reduce"):
    """(L_r, stats) for this rank's rows against the all-gathered text batch.
    f, g: [B_local, D] projected features (same B_local on every rank).

    route  how the text-side gradient is completed across ranks:
           "reduce"     every rank's partial over all text rows is reduce-scattered (sum) to the owners;
           "symmetric"  the image rows are all-gathered too and every rank recomputes its own column slab: one
                        more slab forward, no gradient traffic, a single exchange phase per step.  Requires the
                        same upstream gradient d(total)/d(L_r) on every rank (true under DDP / GradScaler)."""
    if route not in ROUTES:
        raise ValueError(f"route must be one of {ROUTES}, got {route!r}")
    _, world = _world(group)
    if world * f.shape[0] < 2:
        raise ValueError("the dense estimator needs at least two rows in the global batch")
    return _GatheredDenseFn.apply(f, g, t, group, route)


def global_loss_for_logging(local_loss: torch.Tensor, group: Optional[object] = None) -> torch.Tensor:
    """mean_r L_r (detached): the value to log; never back-propagate it (see module docstring)."""
    out = local_loss.detach().clone()
    _, world = _world(group)
    if world > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
        out /= world
    return out


def average_loss_components(components, group: Optional[object] = None):
    """Fused replacement for the reference's ``average_across_processes`` on the loss dictionary
    (utils/distributed.py:141-159 reduces each of the four 0-dim tensors of model.py:103-111 with its own
    all-reduce): the values are packed into one vector, reduced by ONE collective and written back in place.
    Accepts the dictionary of 0-dim tensors (same device) or a single tensor; returns its argument."""
    _, world = _world(group)
    single = isinstance(components, torch.Tensor)
    items = [("", components)] if single else list(components.items())
    if world > 1 and items:
        packed = torch.stack([v.detach().reshape(()).float() for _, v in items])
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
        packed /= world
        with torch.no_grad():
            for i, (_, v) in enumerate(items):
                v.copy_(packed[i].to(v.dtype))
    return components


class GraphedGatheredStep:
    """Forward + backward of the gathered dense loss with the library kernels replayed from CUDA graphs.

    NCCL collectives are not captured; the step is cut at the two exchange points into four graph
    segments with the collectives launched eagerly in between:

        [normalise F, G] -> all-gather V -> [forward slab, dV partial] -> reduce-scatter dV (async)
                                            [dU, image-side Jacobian, dL/dt] -> wait -> [text-side Jacobian]

    so a step costs four graph launches and two NCCL calls on the host instead of ~20 eager launches.
    The upstream gradient is a device scalar (``gamma``, default 1) that may be updated between replays.
    Returns the same static tensors on every call: (loss, dF, dG, dt).
    """

    def __init__(self, f: torch.Tensor, g: torch.Tensor, t: torch.Tensor, group=None, warmup: int = 2):
        self.group = group
        self.rank, self.world = _world(group)
        dev = f.device
        m, d = f.shape
        self.f = f.detach().clone()
        self.g = g.detach().clone()
        self.t = t.detach()
        self.gamma = torch.ones((), dtype=torch.float32, device=dev)
        self.v_all = torch.empty(self.world * m, d, dtype=torch.bfloat16, device=dev)
        self.dv = torch.empty(m, d, dtype=torch.float32, device=dev)
        n = self.world * m

        def seg1():
            return K.normalize_cast_pair(self.f, self.g)

        def seg2(u):
            out4, loss, gmat, gdiag = K.dense_fwd(u, self.v_all, self.t, row_offset=self.rank * m)
            return loss, gmat, gdiag, K.dense_bwd_dv(gmat, u, n, self.t, self.gamma)

        def seg3(inv_f, gmat, gdiag):
            return K.dense_backward_image_side(self.f, self.v_all, inv_f, gmat, gdiag, self.t, self.gamma,
                                               self.rank * m)

        def seg4(u, inv_g, gdiag):
            return K.normalize_bwd(self.g, inv_g, self.dv, u, 0, gdiag, self.t, self.gamma, m)

        def eager_once(capture: bool):
            graphs = []

            def run(fn, *a):
                if not capture:
                    return fn(*a)
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    out = fn(*a)
                gr.replay()                      # capture does not execute: run it once for the next segment's inputs
                graphs.append(gr)
                return out

            u, v, inv_f, inv_g = run(seg1)
            self._gather(v)
            loss, gmat, gdiag, dv_partial = run(seg2, u)
            work = self._scatter(dv_partial)
            df, dt = run(seg3, inv_f, gmat, gdiag)
            if work is not None:
                work.wait()
            dg = run(seg4, u, inv_g, gdiag)
            return graphs, (v, dv_partial), (loss, df, dg, dt)

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                eager_once(False)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graphs, (self.v, self.dv_partial), self.outputs = eager_once(True)
        torch.cuda.synchronize()

    def _gather(self, v):
        if self.world > 1:
            dist.all_gather_into_tensor(self.v_all, v, group=self.group)
        else:
            self.v_all.copy_(v)

    def _scatter(self, dv_partial):
        if self.world > 1:
            return dist.reduce_scatter_tensor(self.dv, dv_partial, op=dist.ReduceOp.SUM, group=self.group,
                                              async_op=True)
        self.dv.copy_(dv_partial)
        return None

    def __call__(self, f: torch.Tensor = None, g: torch.Tensor = None):
        if f is not None:
            self.f.copy_(f, non_blocking=True)
        if g is not None:
            self.g.copy_(g, non_blocking=True)
        g1, g2, g3, g4 = self.graphs
        g1.replay()
        self._gather(self.v)
        g2.replay()
        work = self._scatter(self.dv_partial)
        g3.replay()
        if work is not None:
            work.wait()
        g4.replay()
        return self.outputs

"""Peer-memory exchange of the sharded dense estimator (one process per GPU, NVLink / NVSwitch).

The sharded path has two exchange steps: every rank needs all text rows (all-gather of V) and
every rank's dV partial has to be summed on the rank owning the text rows (reduce-scatter).  Here both are
fused into the kernels on either side of them instead of being NCCL collectives:

  normalise + gather   the normalise kernel stores each bf16 text row into EVERY rank's gathered V buffer
                       (peer stores over NVLink), one destination after the other with one flag each; the
                       forward starts on its own column block and every tile waits only for the ranks whose
                       rows it loads, so the exchange runs underneath the forward.
  dV + reduce          ``partials="bf16"`` (default): the dV contraction's epilogue pushes every output tile as
                       bf16 by TMA store into the OWNER's slot buffer (one flag per owner) -- the traffic hides
                       behind the contraction; the owner's text-side Jacobian waits for all flags and sums the
                       world slots in rank order (deterministic).  ``partials="fp32"``: the contraction leaves an
                       fp32 partial in its own peer-mapped buffer and the owners pull their rows over NVLink.

A step is therefore five kernel launches and no collective call, which also makes the whole step
capturable in a CUDA graph (``PeerGraphedStep``).  ``torch.distributed`` is used once, at set-up, to
exchange the CUDA IPC handles of the buffers.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch
import torch.distributed as dist

from . import _lib
from . import kernels as K


class _DevBuffer:
    """A cudaMalloc'ed (IPC-exportable) device buffer exposed to torch through __cuda_array_interface__."""

    def __init__(self, nbytes: int):
        ptr = ctypes.c_void_p()
        _lib.call("jsd_peer_alloc", nbytes, ctypes.byref(ptr))
        self.ptr, self.nbytes = ptr.value, nbytes

    def free(self):
        if self.ptr:
            _lib.call("jsd_peer_free", self.ptr)
            self.ptr = 0

    def handle(self) -> bytes:
        buf = ctypes.create_string_buffer(_lib.PEER_HANDLE_BYTES)
        _lib.call("jsd_peer_export", self.ptr, buf)
        return buf.raw

    def tensor(self, shape, dtype) -> torch.Tensor:
        typestr = {torch.bfloat16: "<u2", torch.float32: "<f4", torch.int32: "<i4"}[dtype]
        holder = type("_Cai", (), {})()
        holder.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (self.ptr, False),
                                           "version": 2}
        holder._keepalive = self
        t = torch.as_tensor(holder, device=torch.device("cuda", torch.cuda.current_device()))
        return t.view(dtype) if t.dtype != dtype else t


def _open(handle: bytes) -> int:
    ptr = ctypes.c_void_p()
    _lib.call("jsd_peer_open", ctypes.create_string_buffer(handle, _lib.PEER_HANDLE_BYTES), ctypes.byref(ptr))
    return ptr.value


class PeerExchange:
    """Buffers, flags and peer mappings for a fixed (rows per rank, D) on a process group of <= 8 GPUs of one
    NVSwitch domain.  Collective constructor: every rank of ``group`` must create it at the same time."""

    def __init__(self, rows: int, dim: int, group=None, partials: Optional[str] = None):
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerExchange needs an initialised torch.distributed process group")
        partials = partials or os.environ.get("JSD_PEER_PARTIALS", "bf16")
        if partials not in ("bf16", "fp32"):
            raise ValueError(f"partials must be 'bf16' or 'fp32', got {partials!r}")
        if partials == "bf16" and rows % 32:
            partials = "fp32"              # the pushed boxes are 32 rows tall and must not straddle two owners
        self.partials = partials
        self._bf16 = int(partials == "bf16")
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > _lib.MAX_PEERS:
            raise ValueError(f"peer exchange supports up to {_lib.MAX_PEERS} GPUs, got {self.world}")
        self.rows, self.dim = int(rows), int(dim)
        n = self.world * self.rows
        lib = _lib.load()
        self._v = [_DevBuffer(n * self.dim * 2) for _ in range(2)]
        self._stage = _DevBuffer(n * self.dim * 4)
        self._flags = _DevBuffer(lib.jsd_peer_flag_bytes())
        mine = [b.handle() for b in (*self._v, self._stage, self._flags)]
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)
        self._opened = []
        ctx = _lib.PeerCtx()
        ctx.rank, ctx.world, ctx.rows, ctx.dim = self.rank, self.world, self.rows, self.dim
        for q in range(self.world):
            if q == self.rank:
                ptrs = [b.ptr for b in (*self._v, self._stage, self._flags)]
            else:
                ptrs = [_open(h) for h in everyone[q]]
                self._opened += ptrs
            ctx.v_all[0][q], ctx.v_all[1][q], ctx.stage[q], ctx.flags[q] = ptrs
        self.ctx = ctx
        self._ctx_ptr = ctypes.addressof(ctx)
        self.v_all = [b.tensor((n, self.dim), torch.bfloat16) for b in self._v]
        self.step = 0
        torch.cuda.synchronize()
        dist.barrier(group=group)          # every rank has mapped every buffer before anyone launches

    def close(self):
        """Collective: unmap the peers' buffers, then free this rank's (nobody may still be reading them)."""
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        for p in self._opened:
            _lib.call("jsd_peer_close", p)
        self._opened = []
        dist.barrier(group=self.group)
        self.v_all = []
        for b in (*self._v, self._stage, self._flags):
            b.free()
        for key in [k for k, v in _EXCHANGES.items() if v is self]:
            del _EXCHANGES[key]

    @staticmethod
    def wait_error() -> Optional[str]:
        """Readable report of an expired peer wait of this process (None if there is none).  The kernels trap
        after `jsd_peer_set_timeout` seconds without progress, which poisons the CUDA context; the reason is kept
        in host memory and can still be read here (e.g. from an except handler around the failing sync)."""
        kind, index, target = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        if not _lib.load().jsd_peer_wait_error(ctypes.byref(kind), ctypes.byref(index), ctypes.byref(target)):
            return None
        what = {1: "the forward kernel waited for the text rows of", 2: "the text-side Jacobian waited for the "
                "gradient partial of"}.get(kind.value, "a kernel waited for")
        return (f"peer exchange timed out: {what} rank {index.value} (step count {target.value}) -- that rank never "
                f"published it; the ranks must stay in lock step within the wait limit (jsd_peer_set_timeout)")

    # ------------------------------------------------------------------ the four fused launches
    def normalize_push(self, f: torch.Tensor, g: torch.Tensor, parity: int):
        """(U, inv_f, inv_g); V rows land in every rank's v_all[parity]."""
        m, d = f.shape
        if (m, d) != (self.rows, self.dim) or g.shape != f.shape or g.dtype != f.dtype:
            raise ValueError(f"expected two [{self.rows}, {self.dim}] tensors of one dtype")
        u = torch.empty(m, d, dtype=torch.bfloat16, device=f.device)
        inv = torch.empty(2, m, dtype=torch.float32, device=f.device)
        _lib.call("jsd_peer_normalize_push", f.data_ptr(), g.data_ptr(), K._code(f), self._ctx_ptr, parity,
                  u.data_ptr(), inv[0].data_ptr(), inv[1].data_ptr(), K._stream())
        return u, inv[0], inv[1]

    def dense_fwd(self, u: torch.Tensor, t: torch.Tensor, parity: int, want_grad: bool = True):
        m, n = self.rows, self.rows * self.world
        tt = K._scalar(t, "temperature")
        small = torch.empty(m + 8, dtype=torch.float32, device=u.device)
        gdiag, out4, loss = small[:m], small[m:m + 4], small[m + 4]
        gmat, ldg = None, 0
        if want_grad:
            ldg = K.round_up(n, 64)
            gmat = torch.empty(m, ldg, dtype=torch.bfloat16, device=u.device)
        ws = K.dense_workspace(u.device)
        _lib.call("jsd_peer_dense_fwd", u.data_ptr(), self._ctx_ptr, parity, tt.data_ptr(), K._ptr(gmat), ldg,
                  gdiag.data_ptr(), ws.data_ptr(), out4.data_ptr(), loss.data_ptr(), K._stream())
        return out4, loss, gmat, gdiag

    def dense_bwd_dv(self, gmat: torch.Tensor, u: torch.Tensor, t: torch.Tensor, gamma: torch.Tensor):
        tt, gg = K._scalar(t, "temperature"), K._scalar(gamma, "gamma")
        _lib.call("jsd_peer_dense_bwd_dv", gmat.data_ptr(), gmat.shape[1], u.data_ptr(), self._ctx_ptr,
                  tt.data_ptr(), gg.data_ptr(), self._bf16, K._stream())

    def dense_backward(self, f, g, t, gamma, parity: int, u, inv_f, inv_g, gmat, gdiag):
        """Whole backward of a step in one library call.  Returns (dF, dG, dt)."""
        m, d = f.shape
        tt, gg = K._scalar(t, "temperature"), K._scalar(gamma, "gamma")
        acc = torch.empty(m, d, dtype=torch.float32, device=f.device)
        small = torch.empty(m + 1, dtype=torch.float32, device=f.device)
        df, dg = torch.empty_like(f), torch.empty_like(g)
        ws = K.dense_workspace(f.device)
        _lib.call("jsd_peer_dense_backward", f.data_ptr(), g.data_ptr(), K._code(f), self._ctx_ptr, parity,
                  u.data_ptr(), inv_f.data_ptr(), inv_g.data_ptr(), gmat.data_ptr(), gmat.shape[1], gdiag.data_ptr(),
                  tt.data_ptr(), gg.data_ptr(), self._bf16, acc.data_ptr(), small.data_ptr(), ws.data_ptr(),
                  K.streamk_workspace(f.device).data_ptr(), df.data_ptr(), dg.data_ptr(), small[m:].data_ptr(),
                  K._stream())
        return df, dg, small[m]

    def normalize_bwd_text(self, g: torch.Tensor, inv_g, u, gdiag, t, gamma) -> torch.Tensor:
        tt, gg = K._scalar(t, "temperature"), K._scalar(gamma, "gamma")
        dg = torch.empty_like(g)
        _lib.call("jsd_peer_normalize_bwd_text", g.data_ptr(), K._code(g), self._ctx_ptr, inv_g.data_ptr(),
                  u.data_ptr(), gdiag.data_ptr(), tt.data_ptr(), gg.data_ptr(), self._bf16, dg.data_ptr(), K._stream())
        return dg


_EXCHANGES = {}


def get_exchange(rows: int, dim: int, group=None, tag: str = "text", partials: Optional[str] = None) -> PeerExchange:
    """Cached PeerExchange per (group, rows, D, tag, partials); the first call is collective.  ``tag`` names
    independent exchanges of the same shape (the symmetric route gathers the image rows through a second one).
    ``partials`` ("bf16" / "fp32", default: environment JSD_PEER_PARTIALS or "bf16") must agree on all ranks."""
    partials = partials or os.environ.get("JSD_PEER_PARTIALS", "bf16")
    key = (id(group) if group is not None else 0, rows, dim, torch.cuda.current_device(), tag, partials)
    ex = _EXCHANGES.get(key)
    if ex is None:
        ex = _EXCHANGES[key] = PeerExchange(rows, dim, group, partials=partials)
    return ex


class _PeerDenseFn(torch.autograd.Function):
    """Same contract as parallel._GatheredDenseFn (rank r back-propagates its own slab loss L_r)."""

    @staticmethod
    def forward(ctx, f, g, t, ex: PeerExchange):
        need_grad = any(ctx.needs_input_grad)
        with torch.autocast(f.device.type, enabled=False):
            dt = torch.promote_types(f.dtype, g.dtype)
            fc, gc = f.to(dt).contiguous(), g.to(dt).contiguous()
            parity = ex.step & 1
            ex.step += 1
            u, inv_f, inv_g = ex.normalize_push(fc, gc, parity)
            out4, loss, gmat, gdiag = ex.dense_fwd(u, t, parity, want_grad=need_grad)
        if need_grad:
            ctx.save_for_backward(fc, gc, t, u, inv_f, inv_g, gmat, gdiag)
        ctx.ex, ctx.parity, ctx.step = ex, parity, ex.step
        ctx.dtypes = (f.dtype, g.dtype, t.dtype)
        ctx.mark_non_differentiable(out4)
        return loss, out4

    @staticmethod
    def backward(ctx, grad_loss, _grad_stats):
        ex = ctx.ex
        if ex.step != ctx.step:
            raise RuntimeError("peer exchange: backward of a step after a newer forward (one step in flight at a time)")
        fc, gc, t, u, inv_f, inv_g, gmat, gdiag = ctx.saved_tensors
        with torch.autocast(fc.device.type, enabled=False):
            gamma = grad_loss.float()
            df, dg, dt = ex.dense_backward(fc, gc, t, gamma, ctx.parity, u, inv_f, inv_g, gmat, gdiag)
        fd, gd, td = ctx.dtypes
        return df.to(fd), dg.to(gd), dt.to(td), None


class _PeerSymmetricFn(torch.autograd.Function):
    """Peer route without gradient traffic (see parallel.gathered_dense_loss, route="symmetric"): the text rows AND
    the image rows are pushed into every rank's gathered buffers (two exchanges), the forward computes this rank's
    row slab sigma(tau U_r V_all^T) and its column slab sigma(tau V_r U_all^T) (the same kernel with the roles of
    the modalities swapped), and the backward is two purely local image-side backwards.  All waiting on peers
    happens in the forward; nothing is pulled or reduced afterwards.  Built from the entry points the "reduce"
    route uses -- not yet timed on hardware (written after the round's GPU budget was spent)."""

    @staticmethod
    def forward(ctx, f, g, t, ex_v: PeerExchange, ex_u: PeerExchange):
        need_grad = any(ctx.needs_input_grad)
        with torch.autocast(f.device.type, enabled=False):
            dt = torch.promote_types(f.dtype, g.dtype)
            fc, gc = f.to(dt).contiguous(), g.to(dt).contiguous()
            parity = ex_v.step & 1
            ex_v.step += 1
            ex_u.step = ex_v.step
            # both pushes always happen (collective semantics must not depend on who needs gradients)
            u, inv_f, inv_g = ex_v.normalize_push(fc, gc, parity)      # U local, text rows -> every rank
            v, _, _ = ex_u.normalize_push(gc, fc, parity)              # V local, image rows -> every rank
            out4, loss, gmat, gdiag = ex_v.dense_fwd(u, t, parity, want_grad=need_grad)
            gmat_t = gdiag_t = None
            if need_grad:
                _, _, gmat_t, gdiag_t = ex_u.dense_fwd(v, t, parity, want_grad=True)
        if need_grad:
            ctx.save_for_backward(fc, gc, t, inv_f, inv_g, gmat, gdiag, gmat_t, gdiag_t)
        ctx.ex_v, ctx.ex_u, ctx.parity, ctx.step = ex_v, ex_u, parity, ex_v.step
        ctx.dtypes = (f.dtype, g.dtype, t.dtype)
        ctx.mark_non_differentiable(out4)
        return loss, out4

    @staticmethod
    def backward(ctx, grad_loss, _grad_stats):
        ex_v, ex_u = ctx.ex_v, ctx.ex_u
        if ex_v.step != ctx.step:
            raise RuntimeError("peer exchange: backward of a step after a newer forward (one step in flight at a time)")
        fc, gc, t, inv_f, inv_g, gmat, gdiag, gmat_t, gdiag_t = ctx.saved_tensors
        off = ex_v.rank * ex_v.rows
        with torch.autocast(fc.device.type, enabled=False):
            gamma = grad_loss.float()
            df, dt = K.dense_backward_image_side(fc, ex_v.v_all[ctx.parity], inv_f, gmat, gdiag, t, gamma, off)
            dg, _ = K.dense_backward_image_side(gc, ex_u.v_all[ctx.parity], inv_g, gmat_t, gdiag_t, t, gamma, off)
        fd, gd, td = ctx.dtypes
        return df.to(fd), dg.to(gd), dt.to(td), None, None


def peer_dense_loss(f: torch.Tensor, g: torch.Tensor, t: torch.Tensor, group=None,
                    exchange: Optional[PeerExchange] = None, route: str = "reduce", partials: Optional[str] = None):
    """(L_r, stats) for this rank's rows against the text rows of every rank, exchanged over peer memory.
    route="reduce" (default, measured): dV partials are pulled and summed by the owners; route="symmetric": the
    image rows are exchanged as well and every rank recomputes its own column slab (no gradient traffic; needs
    the same upstream gradient on every rank).  partials: "bf16" (default: the dV contraction pushes bf16 tiles into
    the owners' memory while it runs) or "fp32" (exact to fp32, the owners pull) -- the same on every rank."""
    if route == "symmetric":
        ex_v = exchange if exchange is not None else get_exchange(f.shape[0], f.shape[1], group)
        ex_u = get_exchange(f.shape[0], f.shape[1], group, tag="image")
        return _PeerSymmetricFn.apply(f, g, t, ex_v, ex_u)
    if route != "reduce":
        raise ValueError(f"route must be 'reduce' or 'symmetric', got {route!r}")
    ex = exchange if exchange is not None else get_exchange(f.shape[0], f.shape[1], group, partials=partials)
    return _PeerDenseFn.apply(f, g, t, ex)


class PeerGraphedStep:
    """Forward + backward of the peer-exchange loss replayed from CUDA graphs: ONE graph launch per step (there
    is no collective call to keep outside the graph).  Two graphs are captured, one per parity of the
    double-buffered gathered V, and replayed alternately.  Returns static tensors (loss, dF, dG, dt)."""

    def __init__(self, f: torch.Tensor, g: torch.Tensor, t: torch.Tensor, group=None, warmup: int = 2,
                 route: str = "reduce", partials: Optional[str] = None):
        if route not in ("reduce", "symmetric"):
            raise ValueError(f"route must be 'reduce' or 'symmetric', got {route!r}")
        self.ex = get_exchange(f.shape[0], f.shape[1], group, partials=partials)
        ex_u = get_exchange(f.shape[0], f.shape[1], group, tag="image") if route == "symmetric" else None
        self.f = f.detach().clone()
        self.g = g.detach().clone()
        self.t = t.detach()
        self.gamma = torch.ones((), dtype=torch.float32, device=f.device)
        ex = self.ex

        def step_reduce(parity):
            u, inv_f, inv_g = ex.normalize_push(self.f, self.g, parity)
            out4, loss, gmat, gdiag = ex.dense_fwd(u, self.t, parity)
            df, dg, dt = ex.dense_backward(self.f, self.g, self.t, self.gamma, parity, u, inv_f, inv_g, gmat, gdiag)
            return loss, df, dg, dt

        def step_symmetric(parity):
            off = ex.rank * ex.rows
            u, inv_f, inv_g = ex.normalize_push(self.f, self.g, parity)
            v, _, _ = ex_u.normalize_push(self.g, self.f, parity)
            out4, loss, gmat, gdiag = ex.dense_fwd(u, self.t, parity)
            _, _, gmat_t, gdiag_t = ex_u.dense_fwd(v, self.t, parity)
            df, dt = K.dense_backward_image_side(self.f, ex.v_all[parity], inv_f, gmat, gdiag, self.t, self.gamma, off)
            dg, _ = K.dense_backward_image_side(self.g, ex_u.v_all[parity], inv_g, gmat_t, gdiag_t, self.t,
                                                self.gamma, off)
            return loss, df, dg, dt

        step = step_symmetric if route == "symmetric" else step_reduce

        p0 = ex.step & 1                   # pushes below come in (p0, p0 ^ 1) pairs: the exchange's parity is kept
        side = torch.cuda.Stream(device=f.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                step(p0)
                step(p0 ^ 1)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graphs, self.outputs = [], []
        for parity in (p0, p0 ^ 1):
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                out = step(parity)
            gr.replay()                    # capture does not execute: keep the ranks' push counters in step
            self.graphs.append(gr)
            self.outputs.append(out)
        torch.cuda.synchronize()
        self.p0 = p0                       # graphs[i] is the one for parity p0 ^ i

    def __call__(self, f: torch.Tensor = None, g: torch.Tensor = None):
        if f is not None:
            self.f.copy_(f, non_blocking=True)
        if g is not None:
            self.g.copy_(g, non_blocking=True)
        i = (self.ex.step & 1) ^ self.p0   # follows the exchange's step count, so eager calls may be interleaved
        self.ex.step += 1
        self.graphs[i].replay()
        return self.outputs[i]

"""Device-side event trace of the library's kernels (development aid).

nsys is not available on the GPU boxes and ncu serialises launches, which hides exactly what matters for the
peer-exchange path: when each kernel of a graph replay starts, how long it waits for its peers' flags, when it
ends.  The trace code is compiled in only into the instrumented build: ``build.build_library(trace=True)`` writes
``csrc/libjsd_b200_trace.so``; run with ``JSD_LIB=<that path>`` (first used in round 2: the nine peer-exchange
timelines under ``profiles/trace_peer_*`` come from it, via ``tools/trace_peer.py``).
``with KernelTrace() as tr: ...; tr.events()`` installs a buffer into which one thread of every kernel
stamps ``%globaltimer``; works inside CUDA-graph replays (enable it BEFORE capturing or replaying, not during a
capture) and on every rank of a multi-GPU run (each rank traces its own device; the timers of different GPUs are
only loosely aligned, so compare intervals, not absolute times, across ranks).
"""
from __future__ import annotations

from typing import List, Tuple

import torch

from . import _lib

KERNELS = {1: "normalise", 2: "forward GEMM", 3: "backward GEMM", 4: "Jacobian", 5: "index", 6: "score",
           7: "normalise+push"}
EVENTS = {0: "start", 1: "peers in", 2: "end"}


class KernelTrace:
    def __init__(self, capacity: int = 1 << 16, device=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.buf = torch.zeros(capacity, dtype=torch.int64, device=self.device)
        self.count = torch.zeros(1, dtype=torch.int32, device=self.device)

    def __enter__(self) -> "KernelTrace":
        torch.cuda.synchronize(self.device)
        with torch.cuda.device(self.device):
            _lib.call("jsd_trace_enable", self.buf.data_ptr(), self.buf.numel(), self.count.data_ptr())
        return self

    def __exit__(self, *exc) -> None:
        torch.cuda.synchronize(self.device)
        with torch.cuda.device(self.device):
            _lib.call("jsd_trace_enable", None, 0, None)

    def reset(self) -> None:
        torch.cuda.synchronize(self.device)
        self.count.zero_()

    def events(self) -> List[Tuple[str, str, int]]:
        """[(kernel, event, t_ns)] in time order (t_ns = the device's %globaltimer, 56 bits)."""
        torch.cuda.synchronize(self.device)
        n = min(int(self.count.item()), self.buf.numel())
        out = []
        for w in self.buf[:n].tolist():
            w &= (1 << 64) - 1
            out.append((KERNELS.get(w >> 60, str(w >> 60)), EVENTS.get((w >> 56) & 0xF, "?"), w & ((1 << 56) - 1)))
        return sorted(out, key=lambda e: e[2])

    def summary(self) -> str:
        ev = self.events()
        if not ev:
            return "(no events)"
        t0 = ev[0][2]
        return "\n".join(f"{(t - t0) / 1e3:10.1f} us  {k:16s} {e}" for k, e, t in ev)

#!/usr/bin/env python
"""Benchmark of the JSD-loss hot path (BASELINE.json metric: fwd+bwd image-text pairs/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

One "step" = one forward + backward of the dense JSD estimator over one batch of
synthetic embeddings (N(0,1) features, text correlated 0.6 with the image,
pre-normalised, bf16, resident in HBM), through the public API
(clip_lite_b200.ops.jsd_dense_loss, or parallel.gathered_dense_loss for N > 1).

Workloads
  dense_b8192_d1024 (default)  global batch 8192, D = 1024 -- the configuration the north-star
                               target is quoted on; for N > 1 the same global batch is sharded
                               B/N rows per GPU with all-gather + reduce-scatter ("strong")
  weak1024_d1024               1024 rows per GPU, global batch 1024*N (BASELINE configs[2])
  dense_b1024_d1024 / dense_b1024_d128   BASELINE configs[1] (launch-latency-bound)
  stress_b65536_d512           BASELINE configs[3] (global batch 65536, D = 512; sharded B/N rows per GPU)
  index_b8192_d2048            the reference's OWN estimator (one rolled negative per row, loss.py:204-222; fp32,
                               D = 2048 = the reference's projection width): HBM-bound fused kernel, roofline
                               bound "hbm"; per-rank loss as in the reference (8192 rows per GPU, no exchange)

Prints ONE JSON line on rank 0 (see the task contract): value = device-resident throughput,
e2e = the same through host (pinned) buffers with H2D/D2H inside the timed region, roofline for
the tcgen05 GEMM family, cpu_baseline = the oracle restatement timed on the host cores.
`--impl reference` times the reference's CPU path: the unmodified reference loss.py (Identity heads) when
/root/reference is present and the workload has the reference's semantics (index_*), otherwise the PyTorch fp32
restatement under oracle/ (the reference tree cannot travel to the GPU box; its loss has no dense mode).
Every line carries "parity": loss / dF / dG / dt of the timed configuration against oracle.jsd_dense(_grads)
(fp64, evaluated on the device, max over ranks) next to the BASELINE tolerances.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "jsd_loss_fwd_bwd_pairs_per_s"
UNIT = "pairs/s"
WORKLOADS = {
    #  name                : (global batch at N=1, D, rows-per-gpu fixed?)
    "dense_b8192_d1024": dict(batch=8192, dim=1024, weak=False, mode="dense", dtype="bf16"),
    "weak1024_d1024": dict(batch=1024, dim=1024, weak=True, mode="dense", dtype="bf16"),
    "dense_b1024_d1024": dict(batch=1024, dim=1024, weak=False, mode="dense", dtype="bf16"),
    "dense_b1024_d128": dict(batch=1024, dim=128, weak=False, mode="dense", dtype="bf16"),
    "stress_b65536_d512": dict(batch=65536, dim=512, weak=False, mode="dense", dtype="bf16"),
    "index_b8192_d2048": dict(batch=8192, dim=2048, weak=True, mode="index", dtype="f32"),
}
LOSS_RTOL, GRAD_RTOL = 1e-3, 1e-2   # BASELINE.json / BASELINE.md section 2


def workload_config(args, wl):
    """The `config` object of the JSON line: a function of (--workload, --gpus) only, so that both arms
    (--impl b200 / reference) print the same one."""
    batch = wl["batch"] * (args.gpus if wl["weak"] else 1)
    return {"workload": args.workload, "global_batch": batch, "dim": wl["dim"],
            "neg_mode": "dense (all off-diagonal pairs)" if wl["mode"] == "dense" else "shift1 (reference: text batch rolled by one)",
            "input_dtype": wl["dtype"], "n_gpus": args.gpus,
            "inputs": "N(0,1) features, text = 0.6 img + 0.8 noise, unit rows (SURVEY 8d), seed 0"}
T_INIT = 2.659260036932778   # log(1/0.07), loss.py:82


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            p = json.load(open(path))
            return float(p["bf16_tflops"]), float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst)"
        except Exception:
            pass
    return 1590.0, 6650.0, "fallback (B200_PROFILING.md)"


def synth(batch, dim, seed=0):
    """SURVEY 8d synthetic inputs: correlated pairs, pre-normalised, bf16."""
    gen = torch.Generator("cpu").manual_seed(seed)
    f0 = torch.randn(batch, dim, generator=gen)
    g0 = torch.randn(batch, dim, generator=gen)
    g = 0.6 * f0 + 0.8 * g0
    f = torch.nn.functional.normalize(f0, dim=-1).bfloat16()
    g = torch.nn.functional.normalize(g, dim=-1).bfloat16()
    return f, g


class ClockSampler:
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    REASONS = {
        0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
        0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
        0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting",
    }

    def __init__(self, index):
        self.samples, self.power, self.bits, self.max_mhz, self._stop = [], [], 0, None, threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                self.bits |= int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                self.power.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                try:
                    self.bits |= int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                except Exception:
                    pass
            time.sleep(0.0005)       # the default timed region is ~7 ms: poll as fast as NVML answers

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        s = sorted(self.samples)
        reasons = [n for b, n in self.REASONS.items() if self.bits & b and n != "gpu_idle"]
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(s), "power_w_max": (max(self.power) if self.power else None)}


# ------------------------------------------------------------------ CPU (reference) arm
def cpu_dense_step(f, g, t, row_offset=0):
    """fwd + backward() of the fp32 restatement, exactly how the reference runs its loss (autograd)."""
    from oracle import jsd_oracle as orc
    fl = f.detach().clone().requires_grad_(True)
    gl = g.detach().clone().requires_grad_(True)
    tt = torch.tensor(t, dtype=torch.float32, requires_grad=True)
    out = orc.jsd_dense(fl, gl, tt, row_offset=row_offset)
    out["loss"].backward()
    return float(out["loss"].detach())


def _cpu_index_stepper():
    """(step(f, g), kind): one fwd + backward() of the reference's own estimator on (B, D) embeddings.  With the
    reference tree present (build container) this is the UNMODIFIED reference loss.py with Identity projection
    heads (kind "reference"); on the GPU box, where /root/reference does not exist, the oracle's
    restatement of the same lines (loss.py:94-105,204-222,254; kind "port")."""
    from oracle import jsd_oracle as orc
    from oracle import reference_loader as rl
    if rl.reference_available():
        mod = rl.reference_estimator_module()

        def step(f, g):
            fl = f.detach().clone().requires_grad_(True)
            gl = g.detach().clone().requires_grad_(True)
            mod.zero_grad(set_to_none=True)
            with rl.cuda_calls_neutralised():
                out = mod(image_features=fl, text_features=gl)
            out["cross_modal_loss"].backward()
            return float(out["cross_modal_loss"].detach())
        return step, "reference"

    def step(f, g):
        fl = f.detach().clone().requires_grad_(True)
        gl = g.detach().clone().requires_grad_(True)
        tt = torch.tensor(T_INIT, dtype=torch.float32, requires_grad=True)
        out = orc.jsd_index(fl, gl, tt)
        out["loss"].backward()
        return float(out["loss"].detach())
    return step, "port"


def time_cpu(wl, batch, dim, budget_s, steps=None, warmup=1):
    """Times the CPU path on a bounded sample of the workload.  Dense: a row slab of `rows` image rows against
    all `batch` text rows (work per row is constant, so pairs/s = rows / t is the whole-problem rate).  Index
    (reference semantics): the first `rows` pairs (work per pair is constant)."""
    torch.set_num_threads(os.cpu_count() or 1)
    f, g = synth(batch, dim)
    f, g = f.float(), g.float()
    if wl["mode"] == "index":
        stepper, kind = _cpu_index_stepper()
        run = lambda r: stepper(f[:r], g[:r])
        rows, what = batch, "index-mode (one rolled negative per row) fwd+backward() of {rows} pairs"
    else:
        kind = "port"
        run = lambda r: cpu_dense_step(f[:r], g, T_INIT)
        rows, what = min(batch, 256), "dense fwd+backward() of a {rows}x{batch} row slab"
    run(rows)                                                # first call pays one-off thread-pool start-up
    t0 = time.perf_counter()
    run(rows)
    probe = max(time.perf_counter() - t0, 1e-4)
    n_steps = steps if steps is not None else 3
    per_step = budget_s / (n_steps + warmup)
    rows = int(min(batch, max(64, rows * per_step / probe)))
    rows = max(64, rows // 64 * 64)
    for _ in range(warmup):
        run(rows)
    times = []
    for _ in range(n_steps):
        t0 = time.perf_counter()
        run(rows)
        times.append(time.perf_counter() - t0)
    mean = sum(times) / len(times)
    src = {"reference": "the unmodified reference loss.py (Identity heads)", "port": "oracle/ restatement"}[kind]
    return {"value": rows / mean, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
            "sample": what.format(rows=rows, batch=batch) + f", D={dim}, fp32 torch CPU ({src}), "
                      f"{n_steps} steps after {warmup} warm-up (mean {mean*1e3:.1f} ms/step)"}, mean, rows


_orig_time_cpu = time_cpu


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = workload_config(args, wl)
    batch, dim = cfg["global_batch"], wl["dim"]
    base, mean, rows = time_cpu(wl, batch, dim, budget_s=150.0, steps=args.steps, warmup=max(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": mean * 1e3, "higher_is_better": True,
        "scaling": "weak" if wl["weak"] else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "details": {"device": "host CPU", "sample_rows": rows},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ parity of the timed configuration
def _rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def parity_dense(step_out, f_all, g_all, rank, rows, dev, world, dist):
    """Loss / dF / dG / dt of this rank's step against oracle.jsd_dense(_grads) in fp64 ON THE DEVICE (checker
    only), inputs = the very bf16 features the step consumed, up-cast.  dF and the loss need this rank's ROW slab
    of the score matrix, dG its COLUMN slab (= the row slab of the transposed problem), so the check scales to
    BASELINE configs[3] (8192 x 65536 fp64 per slab).  Errors are the max over ranks."""
    from oracle import jsd_oracle as orc
    loss, df, dg, dt = step_out
    lo, hi = rank * rows, (rank + 1) * rows
    f64, g64 = f_all.to(dev).double(), g_all.to(dev).double()
    ref = orc.jsd_dense(f64[lo:hi], g64, T_INIT, row_offset=lo)
    rdf, _, rdt = orc.jsd_dense_grads(f64[lo:hi], g64, T_INIT, row_offset=lo)
    rdg, _, _ = orc.jsd_dense_grads(g64[lo:hi], f64, T_INIT, row_offset=lo)     # transposed problem: column slab
    errs = torch.tensor([_rel(loss, ref["loss"]), _rel(df, rdf), _rel(dg, rdg), _rel(dt, rdt)],
                        device=dev, dtype=torch.float64)
    del f64, g64, rdf, rdg
    if world > 1:
        dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    e = [float(x) for x in errs]
    return {"loss_rel": e[0], "dF_rel": e[1], "dG_rel": e[2], "dt_rel": e[3],
            "tol": {"loss": LOSS_RTOL, "grad": GRAD_RTOL},
            "ok": bool(e[0] <= LOSS_RTOL and max(e[1:]) <= GRAD_RTOL),
            "against": "oracle.jsd_dense / jsd_dense_grads, fp64 on the device, same bf16 inputs up-cast; "
                       "max-norm relative error per tensor, max over ranks"}


def parity_index(step_out, f, g):
    from oracle import jsd_oracle as orc
    loss, df, dg, dt = step_out
    f64, g64 = f.detach().double(), g.detach().double()
    ref = orc.jsd_index(f64, g64, T_INIT)
    rdf, rdg, rdt = orc.jsd_index_grads(f64, g64, T_INIT)
    e = [_rel(loss, ref["loss"]), _rel(df, rdf), _rel(dg, rdg), _rel(dt, rdt)]
    return {"loss_rel": e[0], "dF_rel": e[1], "dG_rel": e[2], "dt_rel": e[3],
            "tol": {"loss": LOSS_RTOL, "grad": GRAD_RTOL},
            "ok": bool(e[0] <= LOSS_RTOL and max(e[1:]) <= GRAD_RTOL),
            "against": "oracle.jsd_index / jsd_index_grads (= reference loss.py:204-222, golden-pinned), fp64 on "
                       "the device, same inputs; max-norm relative error per tensor"}


# ------------------------------------------------------------------ B200 arm
def run_b200(args, wl):
    import torch.distributed as dist
    from clip_lite_b200 import kernels as K
    from clip_lite_b200 import ops, parallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = workload_config(args, wl)
    batch, dim = cfg["global_batch"], wl["dim"]
    index_mode = wl["mode"] == "index"
    if batch % world:
        raise SystemExit("global batch must divide by the number of GPUs")
    rows = batch // world
    f_all, g_all = synth(batch, dim)
    if wl["dtype"] == "f32":
        f_all, g_all = f_all.float(), g_all.float()
    f_host = f_all[rank * rows:(rank + 1) * rows].contiguous().pin_memory()
    g_host = g_all[rank * rows:(rank + 1) * rows].contiguous().pin_memory()
    f_dev = f_host.to(dev).requires_grad_(True)
    g_dev = g_host.to(dev).requires_grad_(True)
    t_dev = torch.tensor(T_INIT, device=dev, requires_grad=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    # N > 1, dense: the two exchange steps either fused into the kernels over NVLink peer memory (default) or as NCCL
    # collectives (--exchange nccl); a peer set-up failure (no CUDA IPC on the box) is reported and NCCL used.
    # Index mode has the reference's per-rank loss: every rank works on its own rows, nothing is exchanged.
    exchange, partials = "none", None
    if world > 1 and not index_mode:
        exchange = args.exchange
        if exchange == "peer":
            try:
                from clip_lite_b200 import peer
                partials = peer.get_exchange(rows, dim).partials
            except Exception as exc:
                print(f"[bench] peer exchange unavailable ({type(exc).__name__}: {exc}); NCCL collectives",
                      file=sys.stderr)
                exchange = "nccl"
            ok = torch.tensor(1 if exchange == "peer" else 0, device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            exchange = "peer" if int(ok) == 1 else "nccl"

    def loss_fn(f, g):
        if index_mode:
            return ops.jsd_index_loss(f, g, t_dev)[0]
        if exchange == "peer":
            return peer.peer_dense_loss(f, g, t_dev, route=args.route)[0]
        if world > 1:
            return parallel.gathered_dense_loss(f, g, t_dev, route=args.route)[0]
        return ops.jsd_dense_loss(f, g, t_dev)[0]

    def eager_step():
        loss = loss_fn(f_dev, g_dev)
        return (loss,) + torch.autograd.grad(loss, (f_dev, g_dev, t_dev))

    # The step is a fixed kernel sequence: replay it as one CUDA graph (clip_lite_b200.graph.GraphedStep)
    # unless --cuda-graph 0; the eager autograd path is what the e2e leg below measures.
    step, graphed = eager_step, False
    if args.cuda_graph:
        try:
            if world == 1 or index_mode:
                from clip_lite_b200.graph import GraphedStep
                gs = GraphedStep(lambda f, g, t: loss_fn(f, g), f_dev, g_dev, t_dev)
            elif exchange == "peer":   # no collective call in the step: one graph launch per step
                gs = peer.PeerGraphedStep(f_dev, g_dev, t_dev, route=args.route)
            elif args.route != "reduce":
                raise RuntimeError("the segmented graph step of the NCCL exchange implements the reduce route only")
            else:      # NCCL is not captured: graph segments between the two eagerly launched collectives
                gs = parallel.GraphedGatheredStep(f_dev, g_dev, t_dev)
            step, graphed = gs, True
        except Exception as exc:                      # capture not possible here: fall back to eager launches
            print(f"[bench] CUDA-graph capture failed ({type(exc).__name__}: {exc}); eager launches", file=sys.stderr)
            torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, flush_l2=True):
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        barrier()
        for i in range(steps):
            if flush_l2:
                flush.zero_()
            starts[i].record()
            fn()
            ends[i].record()
        barrier()
        total_ms = sum(s.elapsed_time(e) for s, e in zip(starts, ends))
        tt = torch.tensor(total_ms, device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt)

    for _ in range(max(args.warmup, 3)):
        out = step()
    torch.cuda.synchronize()
    loss_value = float(out[0].detach())

    # ---- parity of exactly what is timed (the step's own outputs) against the oracle, before the clock starts
    parity = None
    if not args.no_parity:
        try:
            if index_mode:
                parity = parity_index(out, f_dev, g_dev)
                if world > 1:
                    worst = torch.tensor([parity[k] for k in ("loss_rel", "dF_rel", "dG_rel", "dt_rel")], device=dev,
                                         dtype=torch.float64)
                    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
                    for k, v in zip(("loss_rel", "dF_rel", "dG_rel", "dt_rel"), worst.tolist()):
                        parity[k] = v
                    parity["ok"] = bool(parity["loss_rel"] <= LOSS_RTOL and
                                        max(parity["dF_rel"], parity["dG_rel"], parity["dt_rel"]) <= GRAD_RTOL)
            else:
                parity = parity_dense(out, f_all, g_all, rank, rows, dev, world, dist)
        except Exception as exc:
            parity = {"ok": None, "error": f"{type(exc).__name__}: {exc}"}
            if world > 1:
                raise
        torch.cuda.empty_cache()

    with ClockSampler(local_rank) as clocks:
        total_ms = timed(step, args.steps)
    ms_per_step = total_ms / args.steps
    value = batch / (ms_per_step * 1e-3)

    # ---- end to end through host buffers (pinned): every step copies its inputs host -> device and reads
    # the loss back.  As in a real input pipeline the H2D copy of step i+1 (copy stream, double-buffered
    # device staging) overlaps the compute of step i; all copies are inside the timed region.
    loss_host = torch.empty(2, dtype=torch.float32).pin_memory()
    copy_stream = torch.cuda.Stream(device=dev)
    stage = [(torch.empty_like(f_dev.detach()), torch.empty_like(g_dev.detach())) for _ in range(2)]
    h2d_done = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def issue_h2d(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])          # the step that last used this slot has finished
            stage[slot][0].copy_(f_host, non_blocking=True)
            stage[slot][1].copy_(g_host, non_blocking=True)
            h2d_done[slot].record(copy_stream)

    def e2e_run(steps):
        main = torch.cuda.current_stream()
        for ev in consumed:
            ev.record(main)
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        start.record(main)
        issue_h2d(0)
        for i in range(steps):
            slot = i & 1
            if i + 1 < steps:
                issue_h2d(slot ^ 1)
            main.wait_event(h2d_done[slot])
            if graphed and world > 1 and not index_mode:
                # sharded step: eager launches (~0.44 ms of host work) would exceed the 0.3 ms H2D copy
                loss = step(stage[slot][0], stage[slot][1])[0]
            else:
                # plain autograd API on the freshly copied tensors (one GPU: the copy engine, not launch
                # overhead, bounds this leg at the headline size)
                f = stage[slot][0].requires_grad_(True)
                g = stage[slot][1].requires_grad_(True)
                loss = loss_fn(f, g)
                torch.autograd.grad(loss, (f, g, t_dev))
                stage[slot][0].requires_grad_(False)
                stage[slot][1].requires_grad_(False)
            consumed[slot].record(main)
            loss_host[slot:slot + 1].copy_(loss.detach().reshape(1), non_blocking=True)
        end.record(main)
        barrier()
        tt = torch.tensor(start.elapsed_time(end), device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt)

    e2e_run(3)
    e2e_steps = max(5, min(args.steps, 100))
    e2e_ms = e2e_run(e2e_steps) / e2e_steps
    esize = f_host.element_size()
    e2e = {"value": batch / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": 2 * rows * dim * esize, "d2h_bytes_per_step": 4, "input_dtype": wl["dtype"],
           "note": f"inputs ({wl['dtype']} features, the dtype named in config.input_dtype) from pinned host memory "
                   "every step (H2D of step i+1 overlaps step i on a copy stream); the loss is read back to the host "
                   "every step, gradients stay on the device as in training"}

    peak_tf, peak_gbs, peak_src = load_peaks()
    n_meas = max(5, min(args.steps, 50))
    slab_nocomm_ms = None
    if index_mode:
        # ---- roofline of the fused index kernel (HBM-bound): read F, G once, write dF, dG once
        # (the fused call that also writes the gradients; an autograd step makes a loss-only call in the forward
        #  -- reads F, G: 2 B D e bytes -- and this call, with the upstream gradient, in the backward: 6 B D e)
        with torch.no_grad():
            fd, gd, tc = f_dev.detach(), g_dev.detach(), t_dev.detach()
            for _ in range(3):
                K.index_fwd_bwd(fd, gd, tc)
            k_ms = timed(lambda: K.index_fwd_bwd(fd, gd, tc), n_meas) / n_meas
            f_ms = timed(lambda: K.index_fwd_bwd(fd, gd, tc, want_grad=False), n_meas) / n_meas
        alg_bytes = 4.0 * rows * dim * esize
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
                    "traffic": _traffic(args.workload), "peak_source": peak_src,
                    "kernel": "jsd_index_kernel, gradient-writing call (+ 1-block finalize)",
                    "algorithmic_bytes_per_launch": alg_bytes,
                    "launch_ms": {"index_fwd_bwd": k_ms, "index_fwd_only": f_ms},
                    "step_algorithmic_bytes": 6.0 * rows * dim * esize,
                    "step_frac_of_peak": 6.0 * rows * dim * esize / (ms_per_step * 1e-3) / 1e9 / peak_gbs}
        launches_per_step = 4             # two calls (forward / backward), index kernel + finalize each
        tflops = None
    else:
        # ---- roofline of the dominant kernel family (the three tcgen05 GEMM launches of a step),
        # timed live with CUDA events around each launch on the launching stream
        with torch.no_grad():
            u, inv_f = K.normalize_cast(f_dev.detach())
            v, inv_g = K.normalize_cast(g_dev.detach())
            if world > 1:
                v_all = torch.empty(batch, dim, dtype=torch.bfloat16, device=dev)
                dist.all_gather_into_tensor(v_all, v)
            else:
                v_all = v
            gamma = torch.ones((), device=dev)
            t_c = t_dev.detach()
            fused = world == 1 and K.fused_supported(rows, dim)
            if fused:
                # D <= 256: ONE tensor-core launch computes the loss and both gradient contractions
                from clip_lite_b200 import _lib as L
                ns = L.load().jsd_dense_fused_splits(rows, dim)
                acc = torch.empty(2, ns, rows, dim, dtype=torch.float32, device=dev)
                gd_ = torch.empty(rows, dtype=torch.float32, device=dev)
                o4 = torch.empty(5, dtype=torch.float32, device=dev)
                wsf = K.dense_workspace(dev)
                stages = {"fused_fwd_bwd": lambda: L.call(
                    "jsd_dense_fused_fwd_bwd", u.data_ptr(), v.data_ptr(), rows, dim, t_c.data_ptr(),
                    acc[0].data_ptr(), acc[1].data_ptr(), gd_.data_ptr(), wsf.data_ptr(), o4.data_ptr(),
                    o4[4:].data_ptr(), torch.cuda.current_stream().cuda_stream)}
            else:
                _, _, gmat, _ = K.dense_fwd(u, v_all, t_c, row_offset=rank * rows)
                stages = {
                    "fwd": lambda: K.dense_fwd(u, v_all, t_c, row_offset=rank * rows),
                    "bwd_du": lambda: K.dense_bwd_du(gmat, v_all, t_c, gamma),
                    "bwd_dv": lambda: K.dense_bwd_dv(gmat, u, batch, t_c, gamma),
                }
            stage_ms = {}
            for name, fn in stages.items():
                for _ in range(3):
                    fn()
                stage_ms[name] = timed(fn, n_meas) / n_meas
        # ---- N > 1: the same row slab with pre-gathered operands and NO exchange (the numerator of
        # the weak-scaling efficiency E(R) = t_slab_nocomm / t_step); replayed from a CUDA graph like the step itself
        if world > 1:
            f_c, g_c = f_dev.detach(), g_dev.detach()

            def local_slab():
                uu, vv, i_f, i_g = K.normalize_cast_pair(f_c, g_c)
                _, _, gm, gd = K.dense_fwd(uu, v_all, t_c, row_offset=rank * rows)
                dv_part = K.dense_bwd_dv(gm, uu, batch, t_c, gamma)
                K.dense_backward_image_side(f_c, v_all, i_f, gm, gd, t_c, gamma, rank * rows)
                K.normalize_bwd(g_c, i_g, dv_part[rank * rows:(rank + 1) * rows], uu, 0, gd, t_c, gamma, rows)

            try:
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side), torch.no_grad():
                    local_slab()
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr), torch.no_grad():
                    local_slab()
                for _ in range(3):
                    gr.replay()
                slab_nocomm_ms = timed(gr.replay, n_meas) / n_meas
            except Exception as exc:
                print(f"[bench] slab-without-exchange timing failed ({type(exc).__name__}: {exc})", file=sys.stderr)
                torch.cuda.synchronize()
        flops_per_launch = 2.0 * rows * batch * dim                       # S = U V^T, dU = G V, dV = G^T U
        gemm_ms = sum(stage_ms.values())
        achieved = 3 * flops_per_launch / (gemm_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                    "traffic": _traffic(args.workload) if world == 1 else None, "peak_source": peak_src,
                    "kernel": ("jsd_fused_kernel (tcgen05, ONE launch/step: score tiles, sigma and both gradient "
                               "accumulators stay on the SM; 3 x 2 B^2 D algorithmic FLOPs, 8 B^2 D executed)") if fused
                              else "jsd_gemm_kernel (tcgen05, 3 launches/step: fwd, dU, dV)",
                    "algorithmic_flops_per_launch": (3 if fused else 1) * flops_per_launch,
                    "launch_ms": stage_ms,
                    "step_frac_of_peak": 3 * flops_per_launch / (ms_per_step * 1e-3) / 1e12 / peak_tf}
        # library launches per step: normalise pair (+push), forward (+ loss), dU, dV, image Jacobian (helper
        # stream), text Jacobian (+ dL/dt); with JSD_OVERLAP=0 one GPU runs both Jacobians in one launch;
        # the symmetric route runs 2 pushes, 2 forwards, 2 contractions, 2 Jacobians
        launches_per_step = 5 if (world == 1 and os.environ.get("JSD_OVERLAP", "1") == "0") else 6
        if world > 1 and args.route == "symmetric":
            launches_per_step = 8
        if fused:
            launches_per_step = 3         # normalise pair, fused kernel, both Jacobians (+ dL/dt) in one launch
        tflops = 6.0 * rows * batch * dim / (ms_per_step * 1e-3) / 1e12

    line = None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu, _, _ = time_cpu(wl, batch, dim, budget_s=20.0)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak" if wl["weak"] else "strong", "vs_baseline": None, "dtype": wl["dtype"],
            "data": "synthetic", "config": cfg,
            "details": {"rows_per_gpu": rows, "parallelism": f"dp{world}", "cuda_graph": graphed,
                        "exchange": exchange, "route": args.route if exchange != "none" else "none",
                        "grad_partials": partials,
                        "l2": "flushed between timed iterations (256 MiB memset outside the event brackets)",
                        "inputs_resident": f"{wl['dtype']} unit rows resident in HBM before the timed region"},
            "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
            "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "loss": loss_value,
        }
        if tflops is not None:
            line["tflops_6B2D"] = tflops
        if slab_nocomm_ms is not None:
            line["slab_nocomm"] = {"ms_per_step": slab_nocomm_ms,
                                   "note": "this rank's row slab with pre-gathered text rows and no exchange, same "
                                           "kernels, graph replay (SURVEY 8d: E(R) = t_slab_nocomm / t_step)"}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


def _traffic(workload):
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            return json.load(open(tpath)).get(workload, None)
        except Exception:
            return None
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="dense_b8192_d1024")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of the timed step's outputs")
    ap.add_argument("--cuda-graph", type=int, default=1, help="replay the step as one CUDA graph (default 1)")
    ap.add_argument("--route", choices=["reduce", "symmetric"], default="reduce",
                    help="N > 1: text-side gradient by reducing the ranks' partials (measured default) or by also "
                         "exchanging the image rows and recomputing the owned column slab (no gradient traffic)")
    ap.add_argument("--exchange", choices=["peer", "nccl"], default="peer",
                    help="N > 1: exchange fused into the kernels over NVLink peer memory, or NCCL collectives")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_b200(args, wl)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Turn the ncu reports brought back in gpurun_out/ into the small text summaries committed under
profiles/ (per round), plus profiles/traffic.json which bench.py reads for roofline.traffic."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__cluster_size",
    "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum",
]


def ncu_raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    return rows[0], rows[1], rows[2:]


def short(name):
    name = name.replace("void ", "").replace("jsd::", "")
    return name.split("(")[0]


def launches():
    path = os.path.join(SRC, "launches.csv")
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    seq = [(short(r[ki]), float(r[vi].replace(",", "")) / 1000.0) for r in data if len(r) > vi]
    # one steady-state step = the launches from one normalise pre-pass to the next
    starts = [i for i, (n, _) in enumerate(seq) if n.startswith("normalize_cast")]
    lines = ["# ncu --metrics gpu__time_duration.sum --clock-control none (python bench.py --steps 3 --warmup 3)",
             "# per-launch device time, serialised and cold-cache: compare SHARES, not absolutes", ""]
    steps = [(a, b) for a, b in zip(starts, starts[1:]) if any(n.startswith("jsd_gemm_kernel<0") for n, _ in seq[a:b])]
    if len(steps) >= 3:
        a, b = steps[len(steps) // 2]
        step = [(n, t) for n, t in seq[a:b] if "FillFunctor<unsigned char>" not in n]   # drop bench.py's L2 flush
        total = sum(t for _, t in step)
        lines.append(f"## one step ({len(step)} launches, {total:.1f} us of kernel time)")
        for n, t in step:
            lines.append(f"{t:9.1f} us  {100 * t / total:5.1f} %  {n}")
        mine = sum(t for n, t in step if n.startswith(("jsd_", "normalize", "finalize", "scale_scalar")))
        gemm = sum(t for n, t in step if n.startswith("jsd_gemm_kernel"))
        lines += ["", f"tcgen05 GEMM family share of the step: {100 * gemm / total:.1f} %",
                  f"library kernels (libjsd_b200.so) share of the step: {100 * mine / total:.1f} %"]
    agg = {}
    for n, t in seq:
        c, s = agg.get(n, (0, 0.0))
        agg[n] = (c + 1, s + t)
    lines += ["", "## all launches of the run, aggregated", f"{'count':>6} {'mean us':>10} {'total us':>10}  kernel"]
    for n, (c, s) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{c:6d} {s / c:10.1f} {s:10.1f}  {n}")
    open(os.path.join(OUT, f"launches_{TAG}.txt"), "w").write("\n".join(lines) + "\n")


def full(rep, name):
    hdr, units, data = ncu_raw(os.path.join(SRC, rep))
    ki = hdr.index("Kernel Name")
    lines = [f"# ncu --set full --clock-control none ({rep}); one block per captured launch", ""]
    traffic = {}
    for d in data:
        kn = short(d[ki])
        lines.append(kn)
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                lines.append(f"    {m:68s} {d[i]:>16s} {units[i]}")
        scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        traffic.setdefault(kn, []).append(float(d[ir]) * scale[units[ir]] + float(d[iw]) * scale[units[iw]])
        lines.append("")
    open(os.path.join(OUT, f"{name}_{TAG}.txt"), "w").write("\n".join(lines))
    return {k: sum(v) / len(v) for k, v in traffic.items()}


def main():
    os.makedirs(OUT, exist_ok=True)
    launches()
    t = full("prof_step.ncu-rep", "ncu_dense_step")
    if os.path.exists(os.path.join(SRC, "prof_index.ncu-rep")):
        full("prof_index.ncu-rep", "ncu_index")
    gemm = {k: v for k, v in t.items() if k.startswith("jsd_gemm_kernel")}
    if gemm:
        per_launch = sum(gemm.values()) / len(gemm)
        json.dump({"dense_b8192_d1024": per_launch,
                   "_detail_bytes_per_launch": gemm,
                   "_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch of the tcgen05 GEMM family "
                            f"(mean of fwd, dU, dV), ncu --set full, {TAG}"},
                  open(os.path.join(OUT, "traffic.json"), "w"), indent=1)
    bj = os.path.join(SRC, "bench_n1.json")
    if os.path.exists(bj):
        open(os.path.join(OUT, f"bench_n1_{TAG}.json"), "w").write(open(bj).read())
    print(open(os.path.join(OUT, f"launches_{TAG}.txt")).read()[:3000])


if __name__ == "__main__":
    main()

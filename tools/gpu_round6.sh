#!/bin/bash
# per-kernel times of one step (ncu launch list) + bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_module.py -m gpu -x -q --timeout 300 --timeout-method=thread -p no:cacheprovider 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_run.log 2>&1; echo "ncu launches exit $?"
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/launches.csv")))
hi=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
h=rows[hi]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
for r in rows[hi+1:][-40:]:
    print("%9.1f us  %s" % (float(r[vi].replace(",",""))/1000.0, r[ki][:90]))
PY
timeout 300 python bench.py --no-cpu-baseline --steps 100 > gpurun_out/bench6_n1.json 2>gpurun_out/bench6_n1.err; echo "bench n1 exit $?"; cat gpurun_out/bench6_n1.json

#!/bin/bash
# round-2 evidence (1 GPU): full test tier, smoke, bench lines, ncu launch list + full captures
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_gpu.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; cut -c1-400 gpurun_out/bench_n1.json
for w in dense_b1024_d1024 dense_b1024_d128 index_b8192_d2048; do timeout 300 python bench.py --workload $w --steps 100 > gpurun_out/bench_$w.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_$w.json')); print('$w', 'ms/step', round(d['ms_per_step'],4), 'parity', d['parity']['ok'], 'roofline', round(d['roofline']['frac'],3), d['roofline']['launch_ms'], 'cpu', d['cpu_baseline']['value'] if d['cpu_baseline'] else None)"; done
timeout 120 python bench.py --impl reference --steps 2 --warmup 1 | tail -1 | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/ncu_launch_run.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"jsd_gemm|normalize_bwd|normalize_cast" -s 12 -c 7 -o gpurun_out/prof_step -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/ncu_full_run.log 2>&1; echo "ncu full exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"jsd_fused" -s 1 -c 1 -o gpurun_out/prof_fused -f python tools/fused_ab.py once 8192 128 1 > gpurun_out/ncu_fused_run.log 2>&1; echo "ncu fused exit $?"
timeout 600 ncu --set full --clock-control none -k regex:"jsd_index" -s 2 -c 2 -o gpurun_out/prof_index -f python tools/quick_bench.py index > gpurun_out/ncu_index_run.log 2>&1; echo "ncu index exit $?"
for cfg in "1024 128" "8192 128"; do for mode in 1 0; do
  timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ab_${cfg// /x}_m$mode.csv python tools/fused_ab.py once $cfg $mode > /dev/null 2>&1
done; done; echo "ab metrics done"
ls -la gpurun_out/*.ncu-rep gpurun_out/ab_*.csv | cut -c1-150

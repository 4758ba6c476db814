#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 400 -p no:cacheprovider > gpurun_out/r2t_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/r2t_pytest_gpu.log | cut -c1-200
timeout 100 python tools/quick_bench.py index 2>&1 | grep "index B" | tee gpurun_out/r2t_index.log
timeout 200 python bench.py --workload index_b8192_d2048 --steps 50 --no-cpu-baseline > gpurun_out/bench_index_b8192_d2048.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_index_b8192_d2048.json')); print('index bench ms/step', round(d['ms_per_step'],4), 'parity', d['parity']['ok'], 'frac', round(d['roofline']['frac'],3), d['roofline']['launch_ms'])"

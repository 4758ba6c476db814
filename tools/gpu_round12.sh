#!/bin/bash
# split-K for underfilled backward launches: tests, slab shapes (1 GPU), small-batch benches
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
JSD_SPLITK=0 timeout 200 python tools/slab_bench.py
JSD_SPLITK=1 timeout 200 python tools/slab_bench.py
for sp in 0 1; do
for wl in dense_b1024_d1024 dense_b1024_d128; do
JSD_SPLITK=$sp timeout 300 python bench.py --no-cpu-baseline --steps 200 --workload $wl > gpurun_out/bench12_$wl.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench12_$wl.json')); print('splitk=$sp $wl', d['ms_per_step'], d['value'])"
done
done
timeout 300 python bench.py --no-cpu-baseline --steps 200 > gpurun_out/bench12_n1.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench12_n1.json')); print('headline', d['ms_per_step'], d['value'], d['roofline']['step_frac_of_peak'])"

#!/bin/bash
# A/B: helper-stream overlap of the image-side Jacobian with the dV contraction
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_module.py -m gpu -x -q --timeout 300 --timeout-method=thread -p no:cacheprovider 2>&1 | tail -3
for ov in 0 1 0 1; do
JSD_OVERLAP=$ov timeout 300 python bench.py --no-cpu-baseline --steps 200 > gpurun_out/bench8_ov$ov.json 2>gpurun_out/bench8_ov$ov.err; echo "overlap=$ov exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench8_ov$ov.json')); print(d['ms_per_step'], d['value'], d['roofline']['step_frac_of_peak'], d['e2e']['ms_per_step'], d['clocks'])"
done
for wl in dense_b1024_d1024 dense_b1024_d128; do
timeout 300 python bench.py --no-cpu-baseline --steps 200 --workload $wl > gpurun_out/bench8_$wl.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench8_$wl.json')); print('$wl', d['ms_per_step'], d['value'])"
done

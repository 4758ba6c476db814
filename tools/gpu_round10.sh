#!/bin/bash
# A/B: polynomial reciprocal (default build) vs MUFU.RCP (libjsd_b200_scalar.so) in the forward epilogue
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_module.py -m gpu -x -q --timeout 300 --timeout-method=thread -p no:cacheprovider 2>&1 | tail -3
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep smoke
for rep in 1 2; do
for lib in clip_lite_b200/csrc/libjsd_b200_scalar.so clip_lite_b200/csrc/libjsd_b200.so; do
echo "== $lib"
JSD_LIB=$PWD/$lib timeout 300 python tools/quick_bench.py 8192 1024 2>&1 | grep -E "dense_fwd|FULL"
JSD_LIB=$PWD/$lib timeout 300 python bench.py --no-cpu-baseline --steps 200 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['ms_per_step'], d['roofline']['launch_ms'], d['roofline']['step_frac_of_peak'])"
done
done

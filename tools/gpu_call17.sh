#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q --timeout 300 -p no:cacheprovider -k "fused or index" > gpurun_out/r2n_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2n_pytest.log | cut -c1-250
for lib in clip_lite_b200/csrc/libjsd_b200_idx_*.so clip_lite_b200/csrc/libjsd_b200.so; do
  echo "== $lib"; JSD_LIB=$PWD/$lib timeout 120 python tools/quick_bench.py index 2>&1 | grep "index B"
done | tee gpurun_out/r2n_index_variants.log

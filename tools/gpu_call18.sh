#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_module.py -m gpu -x -q --timeout 300 -p no:cacheprovider > gpurun_out/r2o_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2o_pytest.log | cut -c1-250
echo "== tanh sigmoid (default)"; timeout 300 python tools/fused_ab.py 2>&1 | tail -8 | tee gpurun_out/r2o_fused_ab_tanh.log
echo "== ex2 + rcp sigmoid (JSD_FUSED_TANH=0 build)"; JSD_LIB=$PWD/clip_lite_b200/csrc/libjsd_b200_notanh.so timeout 300 python tools/fused_ab.py 2>&1 | tail -8 | tee gpurun_out/r2o_fused_ab_notanh.log

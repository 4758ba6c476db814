#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q --timeout 300 -p no:cacheprovider -k "fused" > gpurun_out/r2m_pytest_fused.log 2>&1; echo "pytest fused exit $?"; tail -25 gpurun_out/r2m_pytest_fused.log | cut -c1-250
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 -p no:cacheprovider > gpurun_out/r2m_pytest_gpu.log 2>&1; echo "pytest all exit $?"; tail -3 gpurun_out/r2m_pytest_gpu.log | cut -c1-250
timeout 300 python tools/fused_ab.py > gpurun_out/r2m_fused_ab.log 2>&1; echo "ab exit $?"; tail -12 gpurun_out/r2m_fused_ab.log

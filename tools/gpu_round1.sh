#!/bin/bash
# first GPU contact: each test group in its own process (a trapped kernel poisons the context)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for grp in "normalize or transpose" "index" "gemm_k_major" "gemm_a_mn_major" "dense_fwd" "dense_bwd" "pipeline or bad_arguments"; do
  name=$(echo "$grp" | tr ' ' '_')
  timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "$grp" --timeout 180 --timeout-method=thread -p no:cacheprovider > gpurun_out/t_${name}.log 2>&1
  echo "== $grp -> exit $?"; tail -n 15 gpurun_out/t_${name}.log
done
timeout 600 python tools/quick_bench.py > gpurun_out/quick_bench.log 2>&1; echo "bench exit $?"; cat gpurun_out/quick_bench.log

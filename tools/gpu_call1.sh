#!/bin/bash
# round 2, call 1 (1 GPU): per-stage slab timings + compute-sanitizer evidence
mkdir -p gpurun_out
nvidia-smi -L
timeout 300 python tools/slab_bench.py > gpurun_out/r2_slab_bench.log 2>&1; echo "slab exit $?"; cat gpurun_out/r2_slab_bench.log
timeout 300 python tools/quick_bench.py 1024 1024 > gpurun_out/r2_quick_1024.log 2>&1; tail -12 gpurun_out/r2_quick_1024.log
timeout 300 python tools/quick_bench.py 1024 128 >> gpurun_out/r2_quick_1024.log 2>&1; tail -12 gpurun_out/r2_quick_1024.log
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_step.py all > gpurun_out/sanitizer_${tool}.log 2>&1
  echo "sanitizer $tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|dense B|index B|score ranks|Error|hazard" gpurun_out/sanitizer_${tool}.log | head -20
done

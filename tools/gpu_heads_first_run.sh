#!/bin/bash
# First hardware session of the projection-head tail (jsd_heads.cuh) -- written when no GPU minutes were
# left; run as  gpurun --timeout 1500 -- 'bash tools/gpu_heads_first_run.sh'.  Everything lands in gpurun_out/.
#   1. its GPU parity tests (119 cases), then the rest of the GPU tier
#   2. fused tail vs default route: step times + parity of the two routes (tools/heads_ab.py)
#   3. compute-sanitizer memcheck / racecheck / synccheck on small steps of every variant
#   4. ncu: launch list of one A/B run, full capture of the two tail kernels (HBM roofline, DESIGN 4.6 bytes/element)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_gpu_heads.py -q --timeout 600 -p no:cacheprovider > gpurun_out/heads_pytest.log 2>&1
echo "heads pytest exit $?"; tail -3 gpurun_out/heads_pytest.log | cut -c1-200
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 -p no:cacheprovider --deselect tests/test_zz_gpu_heads.py > gpurun_out/heads_pytest_rest.log 2>&1
echo "rest of the GPU tier exit $?"; tail -2 gpurun_out/heads_pytest_rest.log | cut -c1-200
timeout 600 python tools/heads_ab.py > gpurun_out/heads_ab.log 2>&1; echo "heads_ab exit $?"; cat gpurun_out/heads_ab.log
for tool in memcheck racecheck synccheck; do
  timeout 300 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_step.py heads > gpurun_out/sanitizer_heads_${tool}.log 2>&1
  echo "sanitizer heads $tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard" gpurun_out/sanitizer_heads_${tool}.log | head -5
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/heads_launches.csv \
  python tools/heads_ab.py once dense 8192 1 > gpurun_out/heads_ncu_list.log 2>&1; echo "ncu list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ln_normalize -c 6 -o gpurun_out/heads_tail \
  python tools/heads_ab.py once dense 8192 1 > gpurun_out/heads_ncu_full.log 2>&1; echo "ncu full exit $?"

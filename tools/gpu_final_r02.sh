#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 -p no:cacheprovider > gpurun_out/r2q_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/r2q_pytest_gpu.log | cut -c1-200
for w in dense_b1024_d1024; do for pv in 1 0; do JSD_PAIRED=$pv timeout 300 python bench.py --workload $w --steps 100 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w paired=$pv', 'ms/step', round(d['ms_per_step'],4), 'parity', d['parity']['ok'], d['roofline']['launch_ms'])"; done; done
timeout 300 python bench.py --steps 50 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('headline ms/step', round(d['ms_per_step'],4), 'parity', d['parity']['ok'])"
for tool in memcheck racecheck synccheck; do
  timeout 400 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_step.py all > gpurun_out/sanitizer_${tool}.log 2>&1
  echo "sanitizer $tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard" gpurun_out/sanitizer_${tool}.log | head -5
done
for tool in memcheck racecheck; do
  timeout 400 compute-sanitizer --tool $tool --print-limit 20 --target-processes all python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29533 tools/sanitize_step.py peer > gpurun_out/sanitizer_peer_${tool}.log 2>&1
  echo "sanitizer peer $tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard|peer rank" gpurun_out/sanitizer_peer_${tool}.log | head -5
done

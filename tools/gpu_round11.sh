#!/bin/bash
# 2-GPU tier: stream-K for underfilled launches; parity tests; slab shapes; bench N=2
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
JSD_STREAMK=0 timeout 200 python tools/slab_bench.py
JSD_STREAMK=1 timeout 200 python tools/slab_bench.py
for wl in dense_b1024_d1024 dense_b1024_d128; do
timeout 300 python bench.py --no-cpu-baseline --steps 200 --workload $wl > gpurun_out/bench11_$wl.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench11_$wl.json')); print('$wl', d['ms_per_step'], d['value'])"
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 > gpurun_out/bench11_n2.json 2>gpurun_out/bench11_n2.err; echo "bench n2 exit $?"; cat gpurun_out/bench11_n2.json; grep -v "^W\|^\*\*\*\|OMP_NUM\|^$" gpurun_out/bench11_n2.err | tail -8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --workload weak1024_d1024 2>/dev/null > gpurun_out/bench11_weak_n2.json; cat gpurun_out/bench11_weak_n2.json

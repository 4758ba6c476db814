#!/bin/bash
# 2-GPU: peer backward with the pull next to the dU contraction
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parallel.py -m gpu -x -q --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu2.log
for rep in 1 2; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 > gpurun_out/bench13_n2.json 2>gpurun_out/bench13_n2.err; echo "bench n2 exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench13_n2.json')); print('n2 peer', d['ms_per_step'], d['value'], d['slab_nocomm'])"
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --workload weak1024_d1024 2>/dev/null > gpurun_out/bench13_weak_n2.json; python -c "
import json; d=json.load(open('gpurun_out/bench13_weak_n2.json')); print('weak n2', d['ms_per_step'], d['value'], d['slab_nocomm'])"

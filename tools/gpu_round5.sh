#!/bin/bash
# 2-GPU tier: NCCL + peer-exchange parity tests, bench at N=1 and N=2 (peer and nccl exchange)
mkdir -p gpurun_out
nvidia-smi -L | head -3
nvidia-smi topo -m 2>/dev/null | head -6
timeout 600 python -m pytest tests/test_gpu_parallel.py -m gpu -x -q --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/pytest_gpu2.log
timeout 300 python bench.py --no-cpu-baseline --steps 100 > gpurun_out/bench5_n1.json 2>gpurun_out/bench5_n1.err; echo "bench n1 exit $?"; cat gpurun_out/bench5_n1.json
for ex in peer nccl; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --exchange $ex > gpurun_out/bench5_n2_$ex.json 2>gpurun_out/bench5_n2_$ex.err; echo "bench n2 $ex exit $?"; cat gpurun_out/bench5_n2_$ex.json; grep -v "^W\|^\*\*\*\|^$" gpurun_out/bench5_n2_$ex.err | tail -8
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --workload weak1024_d1024 > gpurun_out/bench5_weak_n2.json 2>/dev/null; cat gpurun_out/bench5_weak_n2.json

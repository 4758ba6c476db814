#!/bin/bash
# N-GPU tier (N = $1): parity tests of both exchange paths, bench at N with the peer and the NCCL exchange
N=${1:-4}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_parallel.py -m gpu -x -q --timeout 600 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu_n$N.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu_n$N.log
for ex in peer nccl; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --exchange $ex > gpurun_out/benchN_n${N}_$ex.json 2>gpurun_out/benchN_n${N}_$ex.err; echo "bench n$N $ex exit $?"; cat gpurun_out/benchN_n${N}_$ex.json; grep -v "^W\|^\*\*\*\|OMP_NUM\|^$" gpurun_out/benchN_n${N}_$ex.err | tail -6
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 100 --workload weak1024_d1024 2>/dev/null > gpurun_out/benchN_weak_n$N.json; cat gpurun_out/benchN_weak_n$N.json

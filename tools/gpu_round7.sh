#!/bin/bash
# 2-GPU tier: parity tests, stage timing of the peer step, bench N=1 / N=2 (peer, nccl)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parallel.py -m gpu -x -q --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu2.log
for shape in "8192 1024" "2048 1024" "16384 512"; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 tools/diag_peer.py $shape 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM\|^$"
done
timeout 300 python bench.py --no-cpu-baseline --steps 100 > gpurun_out/bench7_n1.json 2>gpurun_out/bench7_n1.err; echo "bench n1 exit $?"; cat gpurun_out/bench7_n1.json
for ex in peer nccl; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --exchange $ex > gpurun_out/bench7_n2_$ex.json 2>gpurun_out/bench7_n2_$ex.err; echo "bench n2 $ex exit $?"; cat gpurun_out/bench7_n2_$ex.json; grep -v "^W\|^\*\*\*\|OMP_NUM\|^$" gpurun_out/bench7_n2_$ex.err | tail -8
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --workload weak1024_d1024 2>/dev/null > gpurun_out/bench7_weak_n2.json; cat gpurun_out/bench7_weak_n2.json

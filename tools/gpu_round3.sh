#!/bin/bash
# 2-GPU tier: NCCL parity test + bench at N=1,2
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_parallel.py -m gpu -q --timeout 600 --timeout-method=thread -p no:cacheprovider 2>&1 | tail -8
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench2_n1.json 2>gpurun_out/bench2_n1.err; echo "bench n1 exit $?"; cat gpurun_out/bench2_n1.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/bench2_n2.json 2>gpurun_out/bench2_n2.err; echo "bench n2 exit $?"; cat gpurun_out/bench2_n2.json; tail -5 gpurun_out/bench2_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload weak1024_d1024 > gpurun_out/bench2_weak_n2.json 2>/dev/null; cat gpurun_out/bench2_weak_n2.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1

#!/bin/bash
# 2-GPU confirmation of HEAD: peer parity (bf16 partials, both shapes) + the default bench line
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parallel.py -m gpu -x -q --timeout 250 -p no:cacheprovider -s -k "test_sharded_dense_matches_single_gpu and 2-peer-bf16" > gpurun_out/r2r_pytest_par_n2.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|rror|worst gradient" gpurun_out/r2r_pytest_par_n2.log | tail -4
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 2>gpurun_out/r2r_bench_n2.err | grep '^{' > gpurun_out/r2r_bench_n2.json
python -c "
import json; d=json.load(open('gpurun_out/r2r_bench_n2.json')); print('bench n2: ms/step', round(d['ms_per_step'],4), 'slab', d['slab_nocomm']['ms_per_step'], 'parity', d['parity']['ok'], 'details', d['details']['exchange'], d['details']['grad_partials'])"

#!/bin/bash
# round 2, call 2 (N GPUs): first run of the symmetric route + device-side kernel timeline of the peer step
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
JSD_TEST_SYMMETRIC=1 timeout 600 python -m pytest tests/test_gpu_parallel.py -m gpu -x -q --timeout 500 -p no:cacheprovider -k "$N" > gpurun_out/r2_pytest_par_n$N.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/r2_pytest_par_n$N.log
for r in reduce symmetric; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --route $r > gpurun_out/r2_bench_n${N}_$r.json 2>gpurun_out/r2_bench_n${N}_$r.err; echo "bench n$N $r exit $?"; cut -c1-600 gpurun_out/r2_bench_n${N}_$r.json; grep -v "^W\|^\*\*\*\|OMP_NUM\|^$" gpurun_out/r2_bench_n${N}_$r.err | tail -6
JSD_LIB=$PWD/clip_lite_b200/csrc/libjsd_b200_trace.so timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/trace_peer.py 8192 1024 $r > gpurun_out/r2_trace_n${N}_$r.log 2>&1; echo "trace $r exit $?"
done
grep -v "^W\|^\*\*\*\|OMP_NUM" gpurun_out/r2_trace_n${N}_reduce.log | head -120

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 -p no:cacheprovider > gpurun_out/r2k_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2k_pytest_gpu.log | cut -c1-200
JSD_PEER_GATHER=kernel timeout 600 python -m pytest tests/test_gpu_parallel.py -m gpu -x -q --timeout 500 -p no:cacheprovider -s -k "1-peer" > gpurun_out/r2k_pytest_w1_kernel.log 2>&1; echo "pytest(kernel gather, world 1) exit $?"; grep -E "passed|failed|rror|worst gradient" gpurun_out/r2k_pytest_w1_kernel.log | tail -8
timeout 600 python bench.py > gpurun_out/r2k_bench_n1.json 2> gpurun_out/r2k_bench_n1.err; echo "bench exit $?"; cut -c1-1500 gpurun_out/r2k_bench_n1.json
timeout 300 python bench.py --workload index_b8192_d2048 --steps 50 > gpurun_out/r2k_bench_index.json 2> gpurun_out/r2k_bench_index.err; echo "bench index exit $?"; cut -c1-2500 gpurun_out/r2k_bench_index.json; tail -3 gpurun_out/r2k_bench_index.err
for w in dense_b1024_d1024 dense_b1024_d128; do timeout 300 python bench.py --workload $w --steps 100 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w', 'ms/step', round(d['ms_per_step'],4), 'parity', d['parity']['ok'], d['roofline']['launch_ms'])"; done

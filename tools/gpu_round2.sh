#!/bin/bash
# full GPU tier: tests, smoke, bench, ncu launch list + full capture of the tcgen05 kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -n 12 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 5 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; cat gpurun_out/bench_n1.json; tail -n 3 gpurun_out/bench_n1.err
for wl in dense_b1024_d1024 dense_b1024_d128; do
  timeout 300 python bench.py --workload $wl --no-cpu-baseline > gpurun_out/bench_$wl.json 2>/dev/null; cat gpurun_out/bench_$wl.json
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_run.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:jsd_gemm -s 9 -c 3 -o gpurun_out/prof_gemm -f python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_run.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out

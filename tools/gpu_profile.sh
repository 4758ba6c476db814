#!/bin/bash
# ncu evidence for profiles/: bench line, launch list of a short bench run, full capture of one step's kernels
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; cat gpurun_out/bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_run.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"jsd_gemm|normalize_bwd|normalize_cast" -s 12 -c 7 -o gpurun_out/prof_step -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_run.log 2>&1; echo "ncu full exit $?"
timeout 120 python bench.py --impl reference --steps 2 --warmup 1 | tail -1 | cut -c1-400
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/smi_after.csv
ls -la gpurun_out | tail -8

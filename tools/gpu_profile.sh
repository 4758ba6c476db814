#!/bin/bash
# ncu evidence for profiles/: launch list of a short bench run + full capture of the tcgen05 kernels
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; cat gpurun_out/bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_run.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"jsd_gemm|normalize_bwd|normalize_cast|jsd_index" -s 16 -c 8 -o gpurun_out/prof_step -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_run.log 2>&1; echo "ncu full exit $?"
timeout 300 ncu --set full --clock-control none -k regex:"jsd_index" -c 2 -o gpurun_out/prof_index -f python -c "
import torch, sys
sys.path.insert(0,'.')
from clip_lite_b200 import ops
f=torch.randn(8192,2048,device='cuda'); g=torch.randn(8192,2048,device='cuda'); t=torch.tensor(2.659,device='cuda')
for _ in range(3): ops.jsd_index_loss(f,g,t)
torch.cuda.synchronize()
" > gpurun_out/ncu_index_run.log 2>&1; echo "ncu index exit $?"
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/smi_after.csv
ls -la gpurun_out

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "gemm or stream_k" --timeout 300 --timeout-method=thread -p no:cacheprovider 2>&1 | tail -3
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"jsd_index" -c 2 -o gpurun_out/prof_index2 -f python -c "
import torch, sys
sys.path.insert(0,'.')
from clip_lite_b200 import kernels as K
t=torch.tensor(2.659,device='cuda')
for dt in (torch.float32, torch.bfloat16):
    f=torch.randn(8192,2048,device='cuda').to(dt); g=torch.randn(8192,2048,device='cuda').to(dt)
    K.index_fwd_bwd(f,g,t)
torch.cuda.synchronize()
" > gpurun_out/ncu_index2_run.log 2>&1; echo "ncu exit $?"

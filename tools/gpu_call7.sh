#!/bin/bash
N=${1:-2}; TAG=${2:-r2e}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parallel.py -m gpu -x -q --timeout 500 -p no:cacheprovider -s -k "test_sharded_dense_matches_single_gpu and ${N}-peer" > gpurun_out/${TAG}_pytest_par_n$N.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|rror|worst gradient" gpurun_out/${TAG}_pytest_par_n$N.log | tail -8
JSD_PEER_GATHER=kernel timeout 600 python -m pytest tests/test_gpu_parallel.py -m gpu -x -q --timeout 500 -p no:cacheprovider -s -k "test_sharded_dense_matches_single_gpu and ${N}-peer-bf16" > gpurun_out/${TAG}_pytest_par_n${N}_kernel.log 2>&1; echo "pytest(kernel gather) exit $?"; grep -E "passed|failed|rror|worst gradient" gpurun_out/${TAG}_pytest_par_n${N}_kernel.log | tail -4
for mode in fwd kernel; do
JSD_PEER_GATHER=$mode JSD_PEER_PARTIALS=bf16 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --no-cpu-baseline 2>gpurun_out/${TAG}_bench_n${N}_$mode.err | grep '^{' > gpurun_out/${TAG}_bench_n${N}_$mode.json
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench_n${N}_$mode.json"))
print("bench n$N gather=$mode: ms/step", round(d["ms_per_step"], 4), "slab_nocomm", d.get("slab_nocomm", {}).get("ms_per_step"), "parity", (d.get("parity") or {}).get("ok"))
PY
done
JSD_PEER_PARTIALS=bf16 JSD_LIB=$PWD/clip_lite_b200/csrc/libjsd_b200_trace.so timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/trace_peer.py 8192 1024 reduce > gpurun_out/${TAG}_trace_n${N}.log 2>&1; echo "trace exit $?"
grep -v "^W\|^\*\*\*\|OMP_NUM" gpurun_out/${TAG}_trace_n${N}.log | sed -n 1,22p

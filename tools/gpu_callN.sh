#!/bin/bash
# N-GPU tier: parity of every exchange x route at world N, bench A/B of the peer route variants, kernel timeline
N=${1:-2}
TAG=${2:-r2}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_parallel.py -m gpu -x -q --timeout 600 -p no:cacheprovider -s -k "test_sharded_dense_matches_single_gpu and (${N}- or 1-)" > gpurun_out/${TAG}_pytest_par_n$N.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|rror|worst gradient" gpurun_out/${TAG}_pytest_par_n$N.log | tail -14
run_bench() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --no-cpu-baseline 2>gpurun_out/${TAG}_bench_n${N}_$name.err | grep '^{' > gpurun_out/${TAG}_bench_n${N}_$name.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench_n${N}_$name.json"))
    print("bench n$N $name: ms/step", round(d["ms_per_step"], 4), "slab_nocomm", d.get("slab_nocomm", {}).get("ms_per_step"), "parity", d.get("parity"), "e2e", round(d["e2e"]["ms_per_step"], 4))
except Exception as e:
    print("bench n$N $name FAILED", e)
PY
  grep -v "^W\|^\*\*\*\|OMP_NUM\|^$" gpurun_out/${TAG}_bench_n${N}_$name.err | tail -4
}
run_bench bf16 JSD_PEER_PARTIALS=bf16
run_bench fp32 JSD_PEER_PARTIALS=fp32
run_bench bf16_waitall JSD_PEER_PARTIALS=bf16 JSD_PEER_WAIT_ALL=1
for part in bf16 fp32; do
JSD_PEER_PARTIALS=$part JSD_LIB=$PWD/clip_lite_b200/csrc/libjsd_b200_trace.so timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/trace_peer.py 8192 1024 reduce > gpurun_out/${TAG}_trace_n${N}_$part.log 2>&1; echo "trace $part exit $?"
done
grep -v "^W\|^\*\*\*\|OMP_NUM" gpurun_out/${TAG}_trace_n${N}_bf16.log | head -34

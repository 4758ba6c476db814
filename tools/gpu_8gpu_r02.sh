#!/bin/bash
# 8-GPU box: parity (world 8, peer bf16) + bench N=8 / N=4 + kernel timeline at N=8
TAG=${1:-r2l}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parallel.py -m gpu -x -q --timeout 300 -p no:cacheprovider -s -k "test_sharded_dense_matches_single_gpu and 8-peer-bf16" > gpurun_out/${TAG}_pytest_par_n8.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|rror|worst gradient" gpurun_out/${TAG}_pytest_par_n8.log | tail -4
run_bench() {  # N name extra-args -- env...
  N=$1; name=$2; extra=$3; shift 3
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --no-cpu-baseline $extra 2>gpurun_out/${TAG}_bench_n${N}_$name.err | grep '^{' > gpurun_out/${TAG}_bench_n${N}_$name.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench_n${N}_$name.json"))
    p = d.get("parity") or {}
    print("bench n$N $name: ms/step", round(d["ms_per_step"], 4), "slab_nocomm", d.get("slab_nocomm", {}).get("ms_per_step"), "parity ok", p.get("ok"), {k: round(v, 5) for k, v in p.items() if k.endswith("_rel")}, "e2e", round(d["e2e"]["ms_per_step"], 4))
except Exception as e:
    print("bench n$N $name FAILED", e)
PY
  grep -v "^W\|^\*\*\*\|OMP_NUM\|^$" gpurun_out/${TAG}_bench_n${N}_$name.err | tail -3
}
run_bench 8 default "" JSD_PEER_PARTIALS=bf16
run_bench 4 default "" JSD_PEER_PARTIALS=bf16
run_bench 8 unpaired "" JSD_PEER_PARTIALS=bf16 JSD_PAIRED=0
JSD_PEER_PARTIALS=bf16 JSD_LIB=$PWD/clip_lite_b200/csrc/libjsd_b200_trace.so timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/trace_peer.py 8192 1024 reduce > gpurun_out/${TAG}_trace_n8_bf16.log 2>&1; echo "trace exit $?"
grep -v "^W\|^\*\*\*\|OMP_NUM" gpurun_out/${TAG}_trace_n8_bf16.log | head -32

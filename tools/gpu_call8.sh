#!/bin/bash
# 8-GPU tier: peer parity at world 8, bench A/B (bf16 / fp32 partials), N=4 on the same box, configs[3], kernel timeline
TAG=${1:-r2d}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_parallel.py -m gpu -x -q --timeout 500 -p no:cacheprovider -s -k "test_sharded_dense_matches_single_gpu and 8-peer and not symmetric" > gpurun_out/${TAG}_pytest_par_n8.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|rror|worst gradient" gpurun_out/${TAG}_pytest_par_n8.log | tail -8
run_bench() {  # N name extra-args -- env...
  N=$1; name=$2; extra=$3; shift 3
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --no-cpu-baseline $extra 2>gpurun_out/${TAG}_bench_n${N}_$name.err | grep '^{' > gpurun_out/${TAG}_bench_n${N}_$name.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench_n${N}_$name.json"))
    p = d.get("parity") or {}
    print("bench n$N $name: ms/step", round(d["ms_per_step"], 4), "slab_nocomm", d.get("slab_nocomm", {}).get("ms_per_step"), "parity ok", p.get("ok"), {k: (round(v, 5) if isinstance(v, float) else v) for k, v in p.items() if k.endswith("_rel")}, "e2e", round(d["e2e"]["ms_per_step"], 4), "launch_ms", d["roofline"]["launch_ms"])
except Exception as e:
    print("bench n$N $name FAILED", e)
PY
  grep -v "^W\|^\*\*\*\|OMP_NUM\|^$" gpurun_out/${TAG}_bench_n${N}_$name.err | tail -3
}
run_bench 8 bf16 "" JSD_PEER_PARTIALS=bf16
run_bench 8 fp32 "" JSD_PEER_PARTIALS=fp32
run_bench 4 bf16 "" JSD_PEER_PARTIALS=bf16
run_bench 8 stress "--workload stress_b65536_d512 --steps 20" JSD_PEER_PARTIALS=bf16
JSD_PEER_PARTIALS=bf16 JSD_LIB=$PWD/clip_lite_b200/csrc/libjsd_b200_trace.so timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/trace_peer.py 8192 1024 reduce > gpurun_out/${TAG}_trace_n8_bf16.log 2>&1; echo "trace exit $?"
grep -v "^W\|^\*\*\*\|OMP_NUM" gpurun_out/${TAG}_trace_n8_bf16.log | head -60

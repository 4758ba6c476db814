#!/bin/bash
# last call of the round: full gpu test tier, smoke, then the profile artefacts
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 --timeout-method=thread -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; grep "smoke" gpurun_out/smoke.log | tail -3
bash tools/gpu_profile.sh

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 -p no:cacheprovider -s > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|error|worst gradient" gpurun_out/r2_pytest_gpu.log | tail -20; tail -30 gpurun_out/r2_pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4

#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu_latest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu_latest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3

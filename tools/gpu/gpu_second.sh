#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu2.log
timeout 1500 python tools/sweep.py > gpurun_out/sweep.jsonl 2> gpurun_out/sweep.err; echo "sweep rc=$?"; tail -3 gpurun_out/sweep.err; wc -l gpurun_out/sweep.jsonl

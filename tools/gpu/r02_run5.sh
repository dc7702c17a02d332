#!/bin/bash
# round 2, run 5 (1 GPU): whole GPU suite on the code with asynchronous snapshots + automatic persistent mode restricted to
# y-slabs; smoke; default bench + reference arm
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 2400 python -m pytest tests -m gpu -q --durations=8 > $O/pytest_gpu_all_run5.log 2>&1; echo "pytest all rc=$?"; tail -25 $O/pytest_gpu_all_run5.log
timeout 300 python __graft_entry__.py smoke > $O/smoke_run5.log 2>&1; echo "smoke rc=$?"; tail -3 $O/smoke_run5.log
timeout 600 python bench.py > $O/bench_run5.json 2> $O/bench_run5.err; echo "bench rc=$?"; cut -c1-600 $O/bench_run5.json; tail -3 $O/bench_run5.err

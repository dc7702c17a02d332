#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu3.log
timeout 2400 python tools/sweep.py --variants 0,2,3,4,5,6,8,12,13,14,15,16,22,23,24,25,26,102,103,104,105,106,112,113,114,115 --steps 30 > gpurun_out/sweep2.jsonl 2> gpurun_out/sweep2.err; echo "sweep rc=$?"; tail -3 gpurun_out/sweep2.err; wc -l gpurun_out/sweep2.jsonl
for ar in exact fast; do timeout 600 python bench.py --steps 10 --warmup 3 --arith $ar --no-cpu > gpurun_out/bench3_$ar.json 2> gpurun_out/bench3_$ar.err; echo "bench $ar rc=$?"; cut -c1-400 gpurun_out/bench3_$ar.json; done

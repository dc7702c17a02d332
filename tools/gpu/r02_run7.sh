#!/bin/bash
# round 2, run 7 (1 GPU): micro-benchmark of the velocity-change reduction's memory layout; Linearized* tests
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 120 tools/microbench/rmw_reduce 4096 > $O/microbench_rmw_reduce.jsonl 2>&1; cat $O/microbench_rmw_reduce.jsonl
timeout 120 tools/microbench/rmw_reduce 4000 >> $O/microbench_rmw_reduce.jsonl 2>&1; tail -3 $O/microbench_rmw_reduce.jsonl


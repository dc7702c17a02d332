#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu4.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu4.log
timeout 1200 python tools/sweep.py --variants 0 --steps 50 > gpurun_out/sweep3.jsonl 2> gpurun_out/sweep3.err; echo "sweep rc=$?"; tail -3 gpurun_out/sweep3.err; wc -l gpurun_out/sweep3.jsonl
for dt in f64 f32; do for ar in exact fast; do timeout 600 python bench.py --steps 10 --warmup 3 --arith $ar --dtype $dt --no-cpu > gpurun_out/bench4_${dt}_$ar.json 2> gpurun_out/bench4_${dt}_$ar.err; echo "bench $dt $ar rc=$?"; cut -c1-260 gpurun_out/bench4_${dt}_$ar.json; done; done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 30 -c 1 -o gpurun_out/prof4_f64_exact python bench.py --steps 1 --warmup 1 --inner 20 --no-e2e --no-cpu > gpurun_out/ncu4a.log 2>&1; echo "ncu rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 30 -c 1 -o gpurun_out/prof4_f64_fast python bench.py --steps 1 --warmup 1 --inner 20 --no-e2e --no-cpu --arith fast > gpurun_out/ncu4b.log 2>&1; echo "ncu rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 30 -c 1 -o gpurun_out/prof4_f32_fast python bench.py --steps 1 --warmup 1 --inner 20 --no-e2e --no-cpu --arith fast --dtype f32 > gpurun_out/ncu4c.log 2>&1; echo "ncu rc=$?"

#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu5.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu5.log
timeout 900 python tools/sweep.py --dtypes f32 --ariths fast --models SRT,TRT --variants 0,99 --steps 50 > gpurun_out/sweep5.jsonl 2> gpurun_out/sweep5.err; echo "sweep rc=$?"; tail -3 gpurun_out/sweep5.err; cat gpurun_out/sweep5.jsonl | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 --dtype f32 --no-cpu > gpurun_out/bench5_f32_fast.json 2> gpurun_out/bench5_f32_fast.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench5_f32_fast.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 30 -c 1 -o gpurun_out/prof5_f32_x2 python bench.py --steps 1 --warmup 1 --inner 20 --no-e2e --no-cpu --dtype f32 > gpurun_out/ncu5.log 2>&1; echo "ncu rc=$?"

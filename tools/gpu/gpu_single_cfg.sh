#!/bin/bash
# 1-GPU points of the BASELINE configs (+ Float32) and the default bench line with the CPU baseline
mkdir -p gpurun_out
run() { name=$1; shift; timeout 1200 python bench.py --gpus 1 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  echo "$name rc=$? $(python -c "
import json
try:
    d=json.loads(open('gpurun_out/$name.json').read().strip().splitlines()[-1]); print(round(d['value']),'MLUPS frac',round(d['roofline']['frac'],3),'ms',round(d['ms_per_step'],2),'e2e',d['e2e'] and round(d['e2e']['value']),'cpu',d['cpu_baseline'] and (round(d['cpu_baseline']['value']),d['cpu_baseline']['cores']))
except Exception as e: print('ERR',e)
")"; tail -2 gpurun_out/$name.err | cut -c1-200; }
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu7.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu7.log
run s1_C2_default --steps 10 --warmup 3
run s1_C2_exact --steps 10 --warmup 3 --arith exact --no-cpu
run s1_C2_f32 --steps 10 --warmup 3 --dtype f32 --no-cpu
run s1_C3 --config C3 --steps 5 --warmup 3 --no-cpu
run s1_C4 --config C4 --steps 5 --warmup 3 --no-cpu
run s1_C5w --config C5w --steps 3 --warmup 3 --no-cpu
run s1_C5s --config C5s --steps 3 --warmup 3 --no-cpu
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s1_reference.json 2> gpurun_out/s1_reference.err; echo "reference rc=$?"; cut -c1-200 gpurun_out/s1_reference.json

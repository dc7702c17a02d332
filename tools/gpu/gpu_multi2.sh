#!/bin/bash
# 2-GPU validation of the side-stream overlap + config presets
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_multi2.log 2>&1; echo "pytest multi rc=$?"; tail -5 gpurun_out/pytest_multi2.log
run() { # name nproc args...
  name=$1; n=$2; shift 2
  if [ $n -eq 1 ]; then timeout 900 python bench.py --gpus 1 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; fi
  echo "$name rc=$? $(python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/$name.json').read().strip().splitlines()[-1]); print(round(d['value']),'MLUPS frac',round(d['roofline']['frac'],3),'ms',round(d['ms_per_step'],2),'e2e',d['e2e'] and round(d['e2e']['value']))
except Exception as e: print('ERR',e)
")"; tail -2 gpurun_out/$name.err | cut -c1-300
}
run m2_C2_n1 1 --steps 10 --warmup 3 --no-cpu
run m2_C2_n2 2 --steps 10 --warmup 3 --no-cpu
run m2_C3_n1 1 --config C3 --steps 5 --warmup 3 --no-cpu
run m2_C3_n2 2 --config C3 --steps 5 --warmup 3 --no-cpu
run m2_C4_n2 2 --config C4 --steps 3 --warmup 3 --no-cpu
run m2_C5w_n2 2 --config C5w --steps 3 --warmup 3 --no-cpu

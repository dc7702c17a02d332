#!/bin/bash
# 2-GPU validation of the peer-memory halo exchange: parity tests, then C2/C3 at N=2 with both halo paths
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_multi.py -x -q --durations=5 > gpurun_out/pytest_p2p.log 2>&1; echo "pytest multi rc=$?"; tail -12 gpurun_out/pytest_p2p.log
run() { # name nproc args...
  name=$1; n=$2; shift 2
  if [ $n -eq 1 ]; then timeout 600 python bench.py --gpus 1 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; fi
  echo "$name rc=$? $(python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/$name.json').read().strip().splitlines()[-1]); print(round(d['value']),'MLUPS frac',round(d['roofline']['frac'],3),'ms',round(d['ms_per_step'],2),'e2e',d['e2e'] and round(d['e2e']['value']), d['config']['halo_exchange'])
except Exception as e: print('ERR',e)
")"; tail -2 gpurun_out/$name.err | cut -c1-300
}
N=${1:-2}
run p2p_C2_n${N}_p2p $N --steps 5 --warmup 3 --no-cpu --no-e2e

run p2p_C3_n${N}_p2p $N --config C3 --steps 5 --warmup 3 --no-cpu --no-e2e
run p2p_C3_n${N}_nccl $N --config C3 --steps 5 --warmup 3 --no-cpu --no-e2e --p2p 0
run p2p_C4_n${N}_p2p $N --config C4 --steps 3 --warmup 3 --no-cpu --no-e2e

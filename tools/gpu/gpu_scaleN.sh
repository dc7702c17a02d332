#!/bin/bash
# N-GPU box (N = $1): slab parity at world 3..N, then the scaling points of C2 (weak) and C3 (strong) on the peer-memory path
N=${1:-4}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -k "more_slabs or one_process or agree_on" --durations=3 > gpurun_out/pytest_slabs_n$N.log 2>&1; echo "pytest more_slabs rc=$?"; tail -6 gpurun_out/pytest_slabs_n$N.log
run() { # name nproc args...
  name=$1; n=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  echo "$name rc=$? $(python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/$name.json').read().strip().splitlines()[-1]); print(round(d['value']),'MLUPS frac',round(d['roofline']['frac'],3),'ms',round(d['ms_per_step'],3),'e2e',d['e2e'] and round(d['e2e']['value']), d['config']['halo_exchange'][:12])
except Exception as e: print('ERR',e)
")"; tail -2 gpurun_out/$name.err | cut -c1-300
}
S="--steps 10 --warmup 3 --no-cpu"
run sN_C3_n${N}_p2p $N --config C3 $S --no-e2e
run sN_C2_n${N}_p2p $N $S
if [ "$2" == "more" ]; then
run sN_C3_n${N}_nccl $N --config C3 $S --no-e2e --p2p 0
run sN_C4_n${N}_p2p $N --config C4 --steps 3 --warmup 3 --no-cpu --no-e2e
fi

#!/bin/bash
# 2-GPU box: full GPU test-suite, then launch-bound small-slab benches (graphs / peer memory / NCCL) and the default bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/pytest_gpu_latest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu_latest.log
run() { # name nproc args...
  name=$1; n=$2; shift 2
  if [ $n -eq 1 ]; then timeout 600 python bench.py --gpus 1 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; fi
  echo "$name rc=$? $(python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/$name.json').read().strip().splitlines()[-1]); print(round(d['value']),'MLUPS frac',round(d['roofline']['frac'],3),'ms',round(d['ms_per_step'],3),'wall',round(d['config']['wall_ms_per_step'],3),'e2e',d['e2e'] and round(d['e2e']['value']), d['config']['halo_exchange'][:12], 'graphs', d['config']['cuda_graphs'])
except Exception as e: print('ERR',e)
")"; tail -2 gpurun_out/$name.err | cut -c1-300
}
S="--steps 10 --warmup 3 --no-cpu --no-e2e"
run g_1k_n1_graph 1 --nx 1024 --ny 1024 $S
run g_1k_n1_plain 1 --nx 1024 --ny 1024 $S --graph 0
run g_256_n1_graph 1 --nx 256 --ny 256 $S
run g_256_n1_plain 1 --nx 256 --ny 256 $S --graph 0
run g_1k_n2_p2p_graph 2 --nx 1024 --ny 1024 $S
run g_1k_n2_p2p_plain 2 --nx 1024 --ny 1024 $S --graph 0
run g_1k_n2_nccl 2 --nx 1024 --ny 1024 $S --p2p 0
run g_C3_n2_graph 2 --config C3 $S
run g_C2_n1_default 1 --steps 5 --warmup 3

#!/bin/bash
# round 2, run 16 (2 GPUs): merged boundary + interior launch (overlap = 2): parity tests, then launch-bound slabs
# (1024 x 1024 per GPU) and C3 with overlap 1 / 2 / 0
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 900 python -m pytest tests/test_gpu_multi.py -q -k "merged or two_slabs_equal_single_domain" --durations=3 > $O/pytest_multi_run16.log 2>&1; echo "pytest multi rc=$?"; tail -8 $O/pytest_multi_run16.log
run() { # name nproc args...
  name=$1; n=$2; shift 2
  if [ $n -eq 1 ]; then timeout 300 python bench.py --gpus 1 "$@" > $O/$name.json 2> $O/$name.err
  else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@" > $O/$name.json 2> $O/$name.err; fi
  echo "$name rc=$? $(python -c "
import json,sys
try:
    d=json.loads(open('$O/$name.json').read().strip().splitlines()[-1]); print(round(d['value']),'MLUPS ms',round(d['ms_per_step'],3),'launches',d['gpu_launches'])
except Exception as e: print('ERR',e)
")"; tail -2 $O/$name.err | cut -c1-300
}
for ov in 1 2 0; do
  run r16_C3q_n2_ov$ov 2 --config C3 --also-shrink 4 --steps 5 --warmup 3 --no-cpu --no-parity --no-e2e --overlap $ov
  run r16_C3_n2_ov$ov 2 --config C3 --steps 5 --warmup 3 --no-cpu --no-parity --no-e2e --overlap $ov
done
run r16_C3q_n1 1 --config C3 --also-shrink 4 --steps 5 --warmup 3 --no-cpu --no-parity --no-e2e
run r16_C2_n2_ov2 2 --steps 5 --warmup 3 --no-cpu --no-e2e --no-also --overlap 2

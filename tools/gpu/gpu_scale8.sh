#!/bin/bash
# 8-GPU pass: slab parity tests + BASELINE configs C2/C3/C4/C5 at N = 2, 4, 8 (N = 4 and N = 2 run concurrently on disjoint GPUs)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_multi8.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/pytest_multi8.log
launch() { # name nproc devices port args...
  name=$1; n=$2; devs=$3; port=$4; shift 4
  CUDA_VISIBLE_DEVICES=$devs timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --no-cpu "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  echo "$name rc=$? $(python -c "
import json
try:
    d=json.loads(open('gpurun_out/$name.json').read().strip().splitlines()[-1]); print(round(d['value']),'MLUPS frac',round(d['roofline']['frac'],3),'ms',round(d['ms_per_step'],2))
except Exception as e: print('ERR',e)
")"
}
for cfg in C2 C3 C4 C5s; do
  st=5; [ $cfg = C5s ] && st=3
  launch s8_${cfg}_n8 8 0,1,2,3,4,5,6,7 29511 --config $cfg --steps $st --warmup 3
  launch s8_${cfg}_n4 4 0,1,2,3 29512 --config $cfg --steps $st --warmup 3 &
  launch s8_${cfg}_n2 2 4,5 29513 --config $cfg --steps $st --warmup 3 &
  wait
done
launch s8_C5w_n8 8 0,1,2,3,4,5,6,7 29511 --config C5w --steps 3 --warmup 3
grep -l "Error\|error" gpurun_out/s8_*.err | head

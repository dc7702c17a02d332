#!/bin/bash
# round 2, run 10 (2 GPUs): bench line at N = 1 and N = 2 (parity check on both halo paths, strong-scaling `also` configs with
# efficiency against the N = 1 figures of the same box); Float32 slab tests
mkdir -p gpurun_out/r02
O=gpurun_out/r02
run() { # name nproc args...
  name=$1; n=$2; shift 2
  if [ $n -eq 1 ]; then timeout 400 python bench.py --gpus 1 "$@" > $O/$name.json 2> $O/$name.err
  else timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@" > $O/$name.json 2> $O/$name.err; fi
  echo "$name rc=$? $(python -c "
import json,sys
try:
    d=json.loads(open('$O/$name.json').read().strip().splitlines()[-1]); print(round(d['value']),'MLUPS frac',round(d['roofline']['frac'],3),'ms',round(d['ms_per_step'],2),'e2e',d['e2e'] and round(d['e2e']['value']),'parity',d.get('parity_check') and (d['parity_check']['ok'],d['parity_check']['bit_identical'],d['parity_check']['max_rel_err']),'also',[(e['preset'],round(e.get('value',0)),e.get('efficiency_vs_n1'),e.get('skipped')) for e in (d.get('also') or [])])
except Exception as e: print('ERR',e)
")"; tail -2 $O/$name.err | cut -c1-300
}
run r10_C2_n1 1 --steps 10 --warmup 3 --no-cpu
run r10_C2_n2 2 --steps 10 --warmup 3
timeout 600 python -m pytest tests/test_zz_gpu_multi_f32.py -q > $O/pytest_multi_f32_run10.log 2>&1; echo "pytest f32 multi rc=$?"; tail -4 $O/pytest_multi_f32_run10.log

#!/bin/bash
# round 2, run 23 (2 GPUs): final multi-GPU check of the current library: slab tests (all), bench N = 1 / 2 as the driver runs it
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 1500 python -m pytest tests/test_gpu_multi.py tests/test_zz_gpu_multi_f32.py -q --durations=3 > $O/pytest_multi_run23.log 2>&1; echo "pytest multi rc=$?"; tail -6 $O/pytest_multi_run23.log
run() { # name nproc args...
  name=$1; n=$2; shift 2
  if [ $n -eq 1 ]; then timeout 400 python bench.py --gpus 1 "$@" > $O/$name.json 2> $O/$name.err
  else timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@" > $O/$name.json 2> $O/$name.err; fi
  echo "$name rc=$? $(python -c "
import json,sys
try:
    d=json.loads(open('$O/$name.json').read().strip().splitlines()[-1]); print(round(d['value']),'MLUPS frac',round(d['roofline']['frac'],3),'ms',round(d['ms_per_step'],2),'e2e',d['e2e'] and round(d['e2e']['value']),'parity',d.get('parity_check') and (d['parity_check']['ok'],d['parity_check']['bit_identical']),'also',[(e['preset'],round(e.get('value',0)),e.get('efficiency_vs_n1'),e.get('launches_per_lattice_step'),e.get('skipped')) for e in (d.get('also') or [])])
except Exception as e: print('ERR',e)
")"; tail -2 $O/$name.err | cut -c1-300
}
run r23_C2_n1 1 --steps 10 --warmup 3 --no-cpu
run r23_C2_n2 2 --steps 10 --warmup 3

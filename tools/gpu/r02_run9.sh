#!/bin/bash
# round 2, run 9 (2 GPUs): multi-GPU parity tests (both halo paths, persistent kernel across GPUs, Float32 slabs), the bench
# line at N = 1 and N = 2 with the post-timing parity check and the strong-scaling `also` configs, launch-bound slabs with the
# persistent kernel on / off
mkdir -p gpurun_out/r02
O=gpurun_out/r02
nvidia-smi -L | head -3
timeout 1500 python -m pytest tests/test_gpu_multi.py tests/test_zz_gpu_multi_f32.py -q --durations=5 > $O/pytest_multi_run9.log 2>&1; echo "pytest multi rc=$?"; tail -12 $O/pytest_multi_run9.log
run() { # name nproc args...
  name=$1; n=$2; shift 2
  if [ $n -eq 1 ]; then timeout 900 python bench.py --gpus 1 "$@" > $O/$name.json 2> $O/$name.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@" > $O/$name.json 2> $O/$name.err; fi
  echo "$name rc=$? $(python -c "
import json,sys
try:
    d=json.loads(open('$O/$name.json').read().strip().splitlines()[-1]); print(round(d['value']),'MLUPS frac',round(d['roofline']['frac'],3),'ms',round(d['ms_per_step'],2),'e2e',d['e2e'] and round(d['e2e']['value']),'parity',d.get('parity_check') and d['parity_check']['ok'],'also',[(e['preset'],round(e.get('value',0)),e.get('efficiency_vs_n1'),e.get('skipped')) for e in (d.get('also') or [])])
except Exception as e: print('ERR',e)
")"; tail -2 $O/$name.err | cut -c1-300
}
run r9_C2_n1 1 --steps 10 --warmup 3
run r9_C2_n2 2 --steps 10 --warmup 3
# launch-bound slabs: 1024 x 1024 per GPU (C3 shrunk 4x), persistent kernel off / on / automatic
run r9_C3q_n2_p0 2 --config C3 --also-shrink 4 --steps 5 --warmup 3 --no-cpu --no-parity --no-e2e --persistent 0
run r9_C3q_n2_p1 2 --config C3 --also-shrink 4 --steps 5 --warmup 3 --no-cpu --no-parity --no-e2e --persistent 1
run r9_C3q_n2_p2 2 --config C3 --also-shrink 4 --steps 5 --warmup 3 --no-cpu --no-parity --no-e2e --persistent 2
run r9_C3q_n1 1 --config C3 --also-shrink 4 --steps 5 --warmup 3 --no-cpu --no-parity --no-e2e
timeout 120 python tools/profile_case.py --lattice D2Q9 --diag > $O/diag_D2Q9_v6.json 2>&1; cat $O/diag_D2Q9_v6.json

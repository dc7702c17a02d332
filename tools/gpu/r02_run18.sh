#!/bin/bash
# round 2, run 18 (1 GPU): L2 prefetch distance sweep for the latency-bound fused kernels; lbm_moments through the copy pipeline
mkdir -p gpurun_out/r02
O=gpurun_out/r02
: > $O/prefetch_sweep.jsonl
for case in "D2Q37 TRT f32" "D2Q37 MRT f32" "D2Q17 TRT f32" "D2Q21 TRT f32" "D2Q9 MRT f32" "D2Q37 MRT f64" "D2Q17 MRT f64" "D2Q9 TRT f64"; do set -- $case
  for pf in 0 16 64 256; do
    timeout 60 python tools/profile_case.py --lattice $1 --model $2 --dtype $3 --prefetch $pf --sustain 0.3 >> $O/prefetch_sweep.jsonl 2>> $O/prefetch_sweep.err
  done
done
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/r02/prefetch_sweep.jsonl') if l.startswith('{')]
seen={}
for d in rows: seen.setdefault((d['lattice'],d['model'],d['dtype']),[]).append((d['prefetch'],d.get('frac')))
for k,v in seen.items(): print(k, v)
PY
timeout 120 python tools/profile_case.py --lattice D2Q9 --diag 2>&1 | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1])['diag']; print({k:(v['host_ms'],v['frac']) for k,v in d.items()})"

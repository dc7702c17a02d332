#!/bin/bash
# round 2, run 29 (1 GPU): L2 prefetch issued by one warp per CTA: Float32 wide lattices (distance sweep), Float64 check
mkdir -p gpurun_out/r02
O=gpurun_out/r02
: > $O/prefetch_sweep_v3.jsonl
for case in "D2Q37 TRT f32" "D2Q17 TRT f32" "D2Q21 MRT f32" "D2Q9 TRT f32"; do set -- $case
  for pf in 0 4 8 16 32; do
    timeout 60 python tools/profile_case.py --lattice $1 --model $2 --dtype $3 --prefetch $pf --sustain 0.3 >> $O/prefetch_sweep_v3.jsonl 2>> $O/prefetch_sweep.err
  done
done
for case in "D2Q9 TRT f64" "D2Q17 MRT f64" "D2Q37 TRT f64"; do set -- $case
  for pf in -1 0; do
    timeout 60 python tools/profile_case.py --lattice $1 --model $2 --dtype $3 --prefetch $pf --sustain 0.3 >> $O/prefetch_sweep_v3.jsonl 2>> $O/prefetch_sweep.err
  done
done
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/r02/prefetch_sweep_v3.jsonl') if l.startswith('{')]
seen={}
for d in rows: seen.setdefault((d['lattice'],d['model'],d['dtype']),[]).append((d['prefetch'],d.get('frac')))
for k,v in seen.items(): print(k, v)
PY

#!/bin/bash
# round 2, run 19 (1 GPU): finer L2 prefetch distance sweep
mkdir -p gpurun_out/r02
O=gpurun_out/r02
: > $O/prefetch_sweep_v2.jsonl
for case in "D2Q9 TRT f64" "D2Q9 MRT f64" "D2Q13 TRT f64" "D2Q17 TRT f64" "D2Q17 MRT f64" "D2Q21 TRT f64" "D2Q37 TRT f64" "D2Q37 MRT f64" "D2Q37 TRT f32" "D2Q17 TRT f32"; do set -- $case
  for pf in 0 2 4 8 12 16 24 32; do
    timeout 60 python tools/profile_case.py --lattice $1 --model $2 --dtype $3 --prefetch $pf --sustain 0.4 >> $O/prefetch_sweep_v2.jsonl 2>> $O/prefetch_sweep.err
  done
done
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/r02/prefetch_sweep_v2.jsonl') if l.startswith('{')]
seen={}
for d in rows: seen.setdefault((d['lattice'],d['model'],d['dtype']),[]).append((d['prefetch'],d.get('frac')))
for k,v in seen.items(): print(k, v)
PY

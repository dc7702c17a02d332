#!/bin/bash
# round 2, run 14 (1 GPU): first run of the TMA-staged fused kernel: identity with the register kernel, then a sweep of its
# configurations against the register kernel on D2Q9 / D2Q17 / D2Q37, Float32 and Float64
mkdir -p gpurun_out/r02
O=gpurun_out/r02
: > $O/tma_sweep_v2.jsonl
for lat in D2Q9 D2Q37; do for dt in f32 f64; do
  timeout 60 python tools/profile_case.py --lattice $lat --model TRT --dtype $dt --tma 1 --check --n 512 --steps 10 --walls 2>&1 | tail -1 | cut -c1-400
done; done
timeout 60 python tools/profile_case.py --lattice D2Q17 --model MRT --dtype f32 --tma 1 --check --n 300 --ny 77 --steps 10 2>&1 | tail -1 | cut -c1-400
timeout 60 python tools/profile_case.py --lattice D2Q37 --model SRT --dtype f64 --arith exact --tma 1 --check --n 1000 --ny 64 --steps 10 --walls 2>&1 | tail -1 | cut -c1-400
for lat in D2Q37 D2Q17 D2Q9; do for dt in f32 f64; do for m in TRT; do
  timeout 60 python tools/profile_case.py --lattice $lat --model $m --dtype $dt --tma 0 --sustain 0.3 >> $O/tma_sweep_v2.jsonl 2>> $O/tma_sweep_v2.err
  for cfg in 122 123 124 125 126 132 133 134 142 143; do
    timeout 60 python tools/profile_case.py --lattice $lat --model $m --dtype $dt --tma 1 --tma-cfg $cfg --sustain 0.3 >> $O/tma_sweep_v2.jsonl 2>> $O/tma_sweep_v2.err
  done
done; done; done
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/r02/tma_sweep_v2.jsonl') if l.startswith('{')]
key=lambda d:(d['lattice'],d['dtype'],d['model'])
seen={}
for d in rows: seen.setdefault(key(d),[]).append((d['tma'],d['tma_cfg'],d.get('frac'),d.get('launches_per_batch')))
for k,v in seen.items(): print(k, v)
PY
tail -5 $O/tma_sweep_v2.err

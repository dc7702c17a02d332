#!/bin/bash
# round 2, run 21 (1 GPU): automatic L2 prefetch distance in nodes: bench (C2 + the strong configs incl. 32768^2) and other row widths
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 600 python bench.py --no-cpu > $O/bench_run21.json 2> $O/bench_run21.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('$O/bench_run21.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['roofline']['frac'], [(e['preset'], round(e.get('value',0)), e.get('frac_of_hbm_peak_per_gpu')) for e in d['also']])"; tail -2 $O/bench_run21.err
for args in "--lattice D2Q9 --n 1024 --ny 8192" "--lattice D2Q9 --n 16384 --ny 4096" "--lattice D2Q9 --n 2048" "--lattice D2Q9 --n 512" "--lattice D2Q37 --n 8192 --ny 2048" "--lattice D2Q37 --n 1024 --ny 4096" "--lattice D2Q17 --model MRT --n 4096 --ny 2048"; do
  for pf in -1 0; do timeout 120 python tools/profile_case.py $args --prefetch $pf --sustain 0.3 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['lattice'],d['model'],d['n'],d['ny'],'pf',d['prefetch'],'frac',d['frac'])"; done
done

#!/bin/bash
# round 2, run 25 (1 GPU): k_errors with two rows in flight per thread (LBM_ERRORS_MINB=2) vs the default; the reference's
# BenchmarkTools suite through the drop-in at scale 2 (what the reference times) and at scale 128, with the CPU restatement
mkdir -p gpurun_out/r02
O=gpurun_out/r02
for mb in 3 2; do
  LBM_ERRORS_MINB=$mb timeout 120 python tools/profile_case.py --lattice D2Q9 --diag 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read())['diag']; print('minb $mb', {k:(v['device_ms'],v['frac']) for k,v in d.items() if 'errors' in k or 'process' in k})"
done
timeout 600 python tools/bench_suite.py --scale 2 --reps 30 --cpu > $O/bench_suite_scale2_r02.jsonl 2> $O/bench_suite.err; wc -l $O/bench_suite_scale2_r02.jsonl
timeout 600 python tools/bench_suite.py --scale 128 --reps 5 --cpu --suites simulation,collision_models > $O/bench_suite_scale128_r02.jsonl 2>> $O/bench_suite.err; wc -l $O/bench_suite_scale128_r02.jsonl; tail -3 $O/bench_suite.err
python - <<'PY'
import json,collections
for f in ('gpurun_out/r02/bench_suite_scale2_r02.jsonl','gpurun_out/r02/bench_suite_scale128_r02.jsonl'):
    rows=[json.loads(l) for l in open(f) if l.startswith('{')]
    by=collections.defaultdict(dict)
    for r in rows: by[(r['suite'],r['lattice'],r['case'],r['nodes'])][r['impl']]=r['median_us']
    n=0
    for k,v in by.items():
        if k[1] in ('D2Q9',) : print(f, k, v); n+=1
        if n>14: break
PY

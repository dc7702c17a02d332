#!/bin/bash
# round 2, run 20 (1 GPU): whole GPU suite on the current library (automatic L2 prefetch, merged launch, TMA option, copy
# pipeline), smoke, default bench + reference arm, ncu launch list of the bench, production sweep (fast mode)
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 2400 python -m pytest tests -m gpu -q --durations=5 > $O/pytest_gpu_all_run20.log 2>&1; echo "pytest all rc=$?"; tail -10 $O/pytest_gpu_all_run20.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_run20.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke_run20.log
timeout 600 python bench.py > $O/bench_run20.json 2> $O/bench_run20.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('$O/bench_run20.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['e2e'].get('serial_value'), d['roofline']['frac'], d['roofline']['traffic'], d['cpu_baseline'] and d['cpu_baseline']['value'], [(e['preset'], round(e.get('value',0))) for e in d['also']], d['clocks'])"; tail -2 $O/bench_run20.err
timeout 600 python bench.py --impl reference --steps 3 > $O/bench_run20_ref.json 2>> $O/bench_run20.err; cut -c1-200 $O/bench_run20_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_run20.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-parity --no-also > /dev/null 2>&1; grep -c . $O/launches_run20.csv
: > $O/sweep_production_run20.jsonl
for lat in D2Q9 D2Q13 D2Q17 D2Q21 D2Q37; do for m in SRT TRT MRT; do for dt in f64 f32; do
  timeout 60 python tools/profile_case.py --lattice $lat --model $m --dtype $dt --sustain 0.3 >> $O/sweep_production_run20.jsonl 2>> $O/sweep_run20.err
done; done; done
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/r02/sweep_production_run20.jsonl') if l.startswith('{')]
for d in rows: print(d['lattice'], d['model'], d['dtype'], d['frac'])
PY

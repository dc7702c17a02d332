#!/bin/bash
# round 2, run 27 (1 GPU): final state -- whole GPU suite, smoke, default bench + reference arm
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 2400 python -m pytest tests -m gpu -x -q --durations=5 > $O/pytest_gpu_all_run27.log 2>&1; echo "pytest all rc=$?"; tail -9 $O/pytest_gpu_all_run27.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_run27.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke_run27.log
timeout 600 python bench.py > $O/bench_run27.json 2> $O/bench_run27.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('$O/bench_run27.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['e2e'].get('serial_value'), d['roofline']['frac'], d['roofline']['traffic'], d['cpu_baseline'] and d['cpu_baseline']['value'], [(e['preset'], round(e.get('value',0))) for e in d['also']], d['clocks'], d['parity_check']['ok'])"; tail -2 $O/bench_run27.err
timeout 600 python bench.py --impl reference > $O/bench_run27_ref.json 2>> $O/bench_run27.err; cut -c1-200 $O/bench_run27_ref.json

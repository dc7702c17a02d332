#!/bin/bash
# round 2, run 12 (1 GPU): copy pipeline for pageable host arrays (array-level operators, upload / download), k_errors at 3 / 4
# CTAs per SM, bench line with the live ncu traffic measurement, tests that touch the copies
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 900 python -m pytest tests -m gpu -q -x -k "upload or download or stream or collide or apply or rows or snapshot or c_consumer or contract or float32 or f32" > $O/pytest_run12.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_run12.log
timeout 600 python tools/bench_array_ops.py > $O/array_ops_v1.jsonl 2> $O/array_ops_v1.err; cat $O/array_ops_v1.jsonl; tail -3 $O/array_ops_v1.err
LBM_COPY_THREADS=1 timeout 600 python tools/bench_array_ops.py --sizes 2048 > $O/array_ops_v1_t1.jsonl 2>> $O/array_ops_v1.err; cat $O/array_ops_v1_t1.jsonl
LBM_COPY_THREADS=4 timeout 600 python tools/bench_array_ops.py --sizes 2048 > $O/array_ops_v1_t4.jsonl 2>> $O/array_ops_v1.err; cat $O/array_ops_v1_t4.jsonl
for mb in 3 4; do
  LBM_ERRORS_MINB=$mb timeout 120 python tools/profile_case.py --lattice D2Q9 --diag > $O/diag_D2Q9_minb$mb.json 2>&1; python -c "
import json; d=json.load(open('$O/diag_D2Q9_minb$mb.json'))['diag']; print('minb $mb', {k:(v['device_ms'],v['frac']) for k,v in d.items() if 'errors' in k or 'process' in k})"
done
timeout 600 python bench.py > $O/bench_run12.json 2> $O/bench_run12.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('$O/bench_run12.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['roofline'], [ (e['preset'], round(e.get('value',0))) for e in d['also']])"; tail -3 $O/bench_run12.err
timeout 600 python bench.py --impl reference --steps 3 > $O/bench_run12_ref.json 2>> $O/bench_run12.err; cut -c1-300 $O/bench_run12_ref.json

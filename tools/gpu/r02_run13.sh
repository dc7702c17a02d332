#!/bin/bash
# round 2, run 13 (1 GPU): asynchronous copies (test + bench e2e with two jobs in flight), default bench line
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "asynchronous or contract or snapshot" > $O/pytest_run13.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_run13.log
timeout 600 python bench.py > $O/bench_run13.json 2> $O/bench_run13.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('$O/bench_run13.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e'], d['roofline']['frac'], d['roofline']['traffic'])"; tail -3 $O/bench_run13.err

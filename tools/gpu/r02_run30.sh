#!/bin/bash
# round 2, run 30 (1 GPU): tests that exercise the fused kernel at production geometry with the one-warp L2 prefetch
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 900 python -m pytest tests -m gpu -x -q -k "prefetch or production or large_grid or c_consumer or contract or graph_replay or tma" > $O/pytest_run30.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_run30.log
timeout 300 python bench.py --no-cpu --no-also > $O/bench_run30.json 2> $O/bench_run30.err; python -c "
import json; d=json.loads(open('$O/bench_run30.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['parity_check']['ok'])"

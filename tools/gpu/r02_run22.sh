#!/bin/bash
# round 2, run 22 (1 GPU): packed two-node Float32 kernel extended to MRT (D2Q4..D2Q13): tests that run Float32, timings
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 1200 python -m pytest tests -m gpu -q -x -k "f32 or float32 or Float32 or F32 or fused or production or figure or linearized" > $O/pytest_run22.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_run22.log
for lat in D2Q9 D2Q13 D2Q5; do for v in 0 99; do
  timeout 60 python tools/profile_case.py --lattice $lat --model MRT --dtype f32 --variant $v --sustain 0.3 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['lattice'],d['model'],d['dtype'],'variant',d['variant'],'frac',d['frac'],'mlups',d['mlups'])"
done; done

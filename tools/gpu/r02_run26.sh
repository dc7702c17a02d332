#!/bin/bash
# round 2, run 26 (1 GPU): cache operator of the population stores (tuning build of D2Q9 / D2Q37 fast)
for case in "D2Q9 TRT f64" "D2Q37 TRT f64" "D2Q9 TRT f32" "D2Q37 TRT f32"; do set -- $case
  for st in 0 1 2 3; do v=0; [ "$3" = f32 ] && v=99
    timeout 60 python tools/profile_case.py --lattice $1 --model $2 --dtype $3 --store $st --variant $v --sustain 0.4 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['lattice'],d['model'],d['dtype'],'store',d['store'],'frac',d['frac'])"
  done
done

#!/bin/bash
# round 2, run 8 (1 GPU): reducing diagnostics after the instruction-count pass (one reciprocal, expected-field squares on the
# host, 3 CTAs per SM, dependent in-place stores): tests that touch them, event timings, ncu of k_errors
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 1200 python -m pytest tests -m gpu -q -x -k "reduc or error or process or moments or couette or figure or golden or batch or stop or converg or mei or linearized" > $O/pytest_run8.log 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest_run8.log
for lat in D2Q9 D2Q37; do
  timeout 120 python tools/profile_case.py --lattice $lat --diag > $O/diag_${lat}_v5.json 2>&1; cat $O/diag_${lat}_v5.json
done
timeout 120 python tools/profile_case.py --lattice D2Q9 --diag --dtype f32 > $O/diag_D2Q9_f32_v5.json 2>&1; cat $O/diag_D2Q9_f32_v5.json
timeout 120 python tools/profile_case.py --lattice D2Q9 --diag --walls > $O/diag_D2Q9_walls_v5.json 2>&1; cat $O/diag_D2Q9_walls_v5.json
prof() { # name, kernel regex, skip, args...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o /tmp/ncu_$name python tools/profile_case.py "$@" > $O/ncu_$name.log 2>&1
  echo "ncu $name rc=$?"
  ncu -i /tmp/ncu_$name.ncu-rep --page raw --csv > $O/ncu_$name.raw.csv 2>/dev/null
  ncu -i /tmp/ncu_$name.ncu-rep --page details --csv > $O/ncu_$name.details.csv 2>/dev/null
}
prof d2q9_errors_v5 k_errors 2 --lattice D2Q9 --diag
prof d2q9_reduce_vc_v5 k_reduce 8 --lattice D2Q9 --diag

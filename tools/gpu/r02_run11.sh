#!/bin/bash
# round 2, run 11 (1 GPU): whole GPU suite on the current library; reducing diagnostics event-timed (D2Q9 / D2Q37, f64 / f32,
# with walls); ncu of k_errors
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 2400 python -m pytest tests -m gpu -q --durations=5 > $O/pytest_gpu_all_run11.log 2>&1; echo "pytest all rc=$?"; tail -12 $O/pytest_gpu_all_run11.log
for lat in D2Q9 D2Q37; do for dt in f64 f32; do
  timeout 120 python tools/profile_case.py --lattice $lat --diag --dtype $dt > $O/diag_${lat}_${dt}_v7.json 2>&1; cat $O/diag_${lat}_${dt}_v7.json
done; done
timeout 120 python tools/profile_case.py --lattice D2Q9 --diag --walls > $O/diag_D2Q9_walls_v7.json 2>&1; cat $O/diag_D2Q9_walls_v7.json
prof() { # name, kernel regex, skip, args...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o /tmp/ncu_$name python tools/profile_case.py "$@" > $O/ncu_$name.log 2>&1
  echo "ncu $name rc=$?"
  ncu -i /tmp/ncu_$name.ncu-rep --page raw --csv > $O/ncu_$name.raw.csv 2>/dev/null
  ncu -i /tmp/ncu_$name.ncu-rep --page details --csv > $O/ncu_$name.details.csv 2>/dev/null
}
prof d2q9_errors_v7 k_errors 2 --lattice D2Q9 --diag

#!/bin/bash
# round 2, run 28 (1 GPU): ncu --set full of the flagship kernel (D2Q9 TRT f64 4096^2, automatic L2 prefetch) and of the
# packed Float32 MRT kernel
mkdir -p gpurun_out/r02
O=gpurun_out/r02
prof() { # name, kernel regex, skip, args...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o /tmp/ncu_$name python tools/profile_case.py "$@" > $O/ncu_$name.log 2>&1
  echo "ncu $name rc=$?"
  ncu -i /tmp/ncu_$name.ncu-rep --page raw --csv > $O/ncu_$name.raw.csv 2>/dev/null
  ncu -i /tmp/ncu_$name.ncu-rep --page details --csv > $O/ncu_$name.details.csv 2>/dev/null
}
prof d2q9_trt_f64_final k_step 6 --lattice D2Q9 --model TRT --dtype f64 --steps 4
prof d2q9_trt_f64_nopf k_step 6 --lattice D2Q9 --model TRT --dtype f64 --steps 4 --prefetch 0
prof d2q9_mrt_f32_x2_final k_step_x2 6 --lattice D2Q9 --model MRT --dtype f32 --steps 4

#!/bin/bash
# round 2, run 2 (1 GPU): batched small problems (tests + the reference's 902 500-solve sweep), C2 full-size parity,
# ncu captures (raw/details CSV only: the .ncu-rep files exceed the 64 MiB return limit)
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 1200 python -m pytest tests/test_gpu_batch.py -m gpu -x -q --durations=5 > $O/pytest_batch.log 2>&1; echo "pytest batch rc=$?"; tail -25 $O/pytest_batch.log
timeout 600 python -m pytest tests/test_gpu_production_geometry.py -m gpu -q -k "c2 or c3 or c4 or stride" > $O/pytest_production2.log 2>&1; echo "pytest production rc=$?"; tail -15 $O/pytest_production2.log
timeout 600 python tools/sweep_trt_magic.py --step 0.1 2>&1 | tail -2
timeout 900 python tools/sweep_trt_magic.py --out $O/trt_magic_sweep.json 2>&1 | tail -2
timeout 900 python tools/sweep_trt_magic.py --arith fast 2>&1 | tail -2
timeout 900 python tools/sweep_trt_magic.py --dtype f32 --arith fast 2>&1 | tail -2
prof() { # name, kernel regex, skip, args...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 300 ncu --set full --clock-control none -k regex:$rx -s $skip -c 1 -f -o /tmp/ncu_$name python tools/profile_case.py "$@" > $O/ncu_$name.log 2>&1
  echo "ncu $name rc=$?"
  ncu -i /tmp/ncu_$name.ncu-rep --page raw --csv > $O/ncu_$name.raw.csv 2>/dev/null
  ncu -i /tmp/ncu_$name.ncu-rep --page details --csv > $O/ncu_$name.details.csv 2>/dev/null
}
prof d2q37_mrt_f64 k_step 5 --lattice D2Q37 --model MRT --dtype f64 --steps 4
prof d2q37_mrt_f32 k_step 5 --lattice D2Q37 --model MRT --dtype f32 --steps 4
prof d2q37_trt_f32 k_step 5 --lattice D2Q37 --model TRT --dtype f32 --steps 4
prof d2q17_trt_f32 k_step 5 --lattice D2Q17 --model TRT --dtype f32 --steps 4
prof d2q21_mrt_f32 k_step 5 --lattice D2Q21 --model MRT --dtype f32 --steps 4
prof d2q9_mrt_f32 k_step 5 --lattice D2Q9 --model MRT --dtype f32 --steps 4
prof d2q9_trt_f32x2 k_step 5 --lattice D2Q9 --model TRT --dtype f32 --steps 4
prof d2q9_errors k_errors 1 --lattice D2Q9 --diag
prof d2q9_reduce "k_reduce.*KParams" 2 --lattice D2Q9 --diag
prof d2q9_moments k_moments 1 --lattice D2Q9 --diag
du -sh $O; ls $O | head -60

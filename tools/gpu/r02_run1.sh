#!/bin/bash
# round 2, run 1 (1 GPU): production-geometry parity tests, default bench with parity_check, sustained sweep of the
# kernels the round-1 verdict lists below 0.9 of peak, ncu captures of those and of the diagnostics kernels
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 1500 python -m pytest tests/test_gpu_production_geometry.py -m gpu -x -q --durations=8 > $O/pytest_production.log 2>&1; echo "pytest rc=$?"; tail -15 $O/pytest_production.log
timeout 600 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; cut -c1-2500 $O/bench_n1.json; tail -3 $O/bench_n1.err
# sustained (>= 0.4 s) fast-mode numbers for every <lattice, model, dtype>
for lat in D2Q9 D2Q13 D2Q17 D2Q21 D2Q37; do for m in SRT TRT MRT; do for dt in f64 f32; do
  timeout 120 python tools/profile_case.py --lattice $lat --model $m --dtype $dt --sustain 0.4 >> $O/sweep_fast.jsonl 2>> $O/sweep.err
done; done; done
cat $O/sweep_fast.jsonl | cut -c1-220
timeout 120 python tools/profile_case.py --lattice D2Q9 --diag > $O/diag_d2q9.json 2>&1; cat $O/diag_d2q9.json
prof() { # name, kernel regex, args...
  local name=$1 rx=$2; shift 2
  timeout 300 ncu --set full --clock-control none -k regex:$rx -s 5 -c 1 -f -o $O/ncu_$name python tools/profile_case.py "$@" > $O/ncu_$name.log 2>&1
  echo "ncu $name rc=$?"
  ncu -i $O/ncu_$name.ncu-rep --page raw --csv > $O/ncu_$name.raw.csv 2>/dev/null
  ncu -i $O/ncu_$name.ncu-rep --page details --csv > $O/ncu_$name.details.csv 2>/dev/null
}
prof d2q37_mrt_f64 k_step --lattice D2Q37 --model MRT --dtype f64 --steps 4
prof d2q37_mrt_f32 k_step --lattice D2Q37 --model MRT --dtype f32 --steps 4
prof d2q37_trt_f32 k_step --lattice D2Q37 --model TRT --dtype f32 --steps 4
prof d2q17_trt_f32 k_step --lattice D2Q17 --model TRT --dtype f32 --steps 4
prof d2q21_mrt_f32 k_step --lattice D2Q21 --model MRT --dtype f32 --steps 4
prof d2q9_trt_f32x2 k_step --lattice D2Q9 --model TRT --dtype f32 --steps 4
prof d2q9_errors k_errors --lattice D2Q9 --diag
prof d2q9_reduce "k_reduce<" --lattice D2Q9 --diag
prof d2q9_moments k_moments --lattice D2Q9 --diag
ls -la $O | head -50

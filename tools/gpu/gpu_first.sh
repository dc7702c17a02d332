#!/bin/bash
# first GPU pass: smoke, parity tests, bench (exact+fast), ncu launch list + one full capture
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" 
tail -5 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log
for ar in exact fast; do
  timeout 600 python bench.py --steps 5 --warmup 3 --arith $ar > gpurun_out/bench_$ar.json 2> gpurun_out/bench_$ar.err; echo "bench $ar rc=$?"; cat gpurun_out/bench_$ar.json; tail -3 gpurun_out/bench_$ar.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --inner 20 --no-e2e --no-cpu > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 30 -c 2 -o gpurun_out/prof_exact python bench.py --steps 1 --warmup 1 --inner 20 --no-e2e --no-cpu > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 30 -c 2 -o gpurun_out/prof_fast python bench.py --steps 1 --warmup 1 --inner 20 --no-e2e --no-cpu --arith fast > gpurun_out/ncu_full_fast.log 2>&1; echo "ncu full fast rc=$?"
ls -la gpurun_out

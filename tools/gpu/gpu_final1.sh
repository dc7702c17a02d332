#!/bin/bash
# 1-GPU: full GPU test-suite, smoke, default bench (+ reference arm), ncu launch list and one full capture of the fused kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/pytest_gpu_latest.log 2>&1; echo "pytest rc=$?"; tail -10 gpurun_out/pytest_gpu_latest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/final_bench_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_reference.json 2> gpurun_out/final_reference.err; echo "ref rc=$?"; cut -c1-600 gpurun_out/final_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 1 --inner 50 --no-e2e --no-cpu > gpurun_out/final_launches.log 2>&1; echo "ncu list rc=$?"; tail -3 gpurun_out/final_launches.csv | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 30 -c 1 -o gpurun_out/final_prof_f64_fast python bench.py --steps 1 --warmup 1 --inner 20 --no-e2e --no-cpu > gpurun_out/final_ncu.log 2>&1; echo "ncu full rc=$?"; ls -la gpurun_out/final_prof_f64_fast.ncu-rep

#!/bin/bash
# round 2, run 4 (1 GPU): device-side initialize (all strategies), process! sums on the device, Linearized* problems, C consumer;
# whole GPU suite; diagnostics kernels device-timed + ncu launch list; default bench line
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_abi.py -m gpu -x -q --durations=5 -k "initialisation or process_sums or linearized or couette or c_consumer" > $O/pytest_run4_new.log 2>&1; echo "pytest new rc=$?"; tail -12 $O/pytest_run4_new.log
timeout 120 python tools/profile_case.py --lattice D2Q9 --diag > $O/diag_d2q9_v3.json 2>&1; cat $O/diag_d2q9_v3.json
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/diag_launches.csv python tools/profile_case.py --lattice D2Q9 --diag > /dev/null 2>&1; grep -c . $O/diag_launches.csv
timeout 2400 python -m pytest tests -m gpu -x -q --durations=8 > $O/pytest_gpu_all_run4.log 2>&1; echo "pytest all rc=$?"; tail -15 $O/pytest_gpu_all_run4.log
timeout 600 python bench.py > $O/bench_run4.json 2> $O/bench_run4.err; echo "bench rc=$?"; cut -c1-1500 $O/bench_run4.json; tail -3 $O/bench_run4.err

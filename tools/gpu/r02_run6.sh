#!/bin/bash
# round 2, run 6 (1 GPU): whole GPU suite (asynchronous snapshots, reworked reducing diagnostics), smoke, diagnostics kernels
# event-timed + ncu full captures of k_errors / k_reduce<velocity change>, default bench
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 2400 python -m pytest tests -m gpu -q --durations=8 > $O/pytest_gpu_all_run6.log 2>&1; echo "pytest all rc=$?"; tail -25 $O/pytest_gpu_all_run6.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_run6.log 2>&1; echo "smoke rc=$?"; tail -3 $O/smoke_run6.log
for lat in D2Q9 D2Q37; do
  timeout 120 python tools/profile_case.py --lattice $lat --diag > $O/diag_${lat}_v4.json 2>&1; cat $O/diag_${lat}_v4.json
  timeout 120 python tools/profile_case.py --lattice $lat --diag --dtype f32 > $O/diag_${lat}_f32_v4.json 2>&1; cat $O/diag_${lat}_f32_v4.json
done
prof() { # name, kernel regex, skip, args...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o /tmp/ncu_$name python tools/profile_case.py "$@" > $O/ncu_$name.log 2>&1
  echo "ncu $name rc=$?"
  ncu -i /tmp/ncu_$name.ncu-rep --page raw --csv > $O/ncu_$name.raw.csv 2>/dev/null
  ncu -i /tmp/ncu_$name.ncu-rep --page details --csv > $O/ncu_$name.details.csv 2>/dev/null
}
prof d2q9_errors_v4 k_errors 2 --lattice D2Q9 --diag
prof d2q9_reduce_vc_v4 k_reduce 8 --lattice D2Q9 --diag
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/diag_launches_v4.csv python tools/profile_case.py --lattice D2Q9 --diag > /dev/null 2>&1; grep -c . $O/diag_launches_v4.csv
timeout 600 python bench.py > $O/bench_run6.json 2> $O/bench_run6.err; echo "bench rc=$?"; cut -c1-400 $O/bench_run6.json; tail -3 $O/bench_run6.err

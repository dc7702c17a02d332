#!/bin/bash
# round 2, run 3 (1 GPU): persistent multi-step kernel (tests + launch-bound sizes vs graphs vs plain), whole GPU suite,
# MRT after pair folding, diagnostics kernels
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 1500 python -m pytest tests/test_gpu_persistent.py -m gpu -x -q --durations=5 > $O/pytest_persistent.log 2>&1; echo "pytest persistent rc=$?"; tail -12 $O/pytest_persistent.log
for n in 128 256 512 1024 2048; do for mode in "--persistent 1" "--persistent 0 --graph 1" "--persistent 0 --graph 0"; do
  timeout 120 python tools/profile_case.py --lattice D2Q9 --model TRT --dtype f64 --n $n --steps 400 --sustain 0.3 $mode >> $O/persistent_sizes.jsonl 2>> $O/persistent.err
done; done
timeout 120 python tools/profile_case.py --lattice D2Q9 --model TRT --dtype f64 --n 1024 --ny 8192 --steps 200 --sustain 0.3 --persistent 1 >> $O/persistent_sizes.jsonl 2>> $O/persistent.err
timeout 120 python tools/profile_case.py --lattice D2Q9 --model TRT --dtype f64 --n 1024 --ny 8192 --steps 200 --sustain 0.3 --persistent 0 >> $O/persistent_sizes.jsonl 2>> $O/persistent.err
timeout 120 python tools/profile_case.py --lattice D2Q9 --model TRT --dtype f32 --n 1024 --steps 400 --sustain 0.3 --persistent 1 >> $O/persistent_sizes.jsonl 2>> $O/persistent.err
timeout 120 python tools/profile_case.py --lattice D2Q9 --model TRT --dtype f32 --n 1024 --steps 400 --sustain 0.3 --persistent 0 >> $O/persistent_sizes.jsonl 2>> $O/persistent.err
timeout 120 python tools/profile_case.py --lattice D2Q37 --model TRT --dtype f64 --n 1024 --steps 100 --sustain 0.3 --persistent 1 >> $O/persistent_sizes.jsonl 2>> $O/persistent.err
timeout 120 python tools/profile_case.py --lattice D2Q37 --model TRT --dtype f64 --n 1024 --steps 100 --sustain 0.3 --persistent 0 >> $O/persistent_sizes.jsonl 2>> $O/persistent.err
cut -c1-330 $O/persistent_sizes.jsonl; tail -3 $O/persistent.err
for lat in D2Q9 D2Q13 D2Q17 D2Q21 D2Q37; do for dt in f64 f32; do
  timeout 120 python tools/profile_case.py --lattice $lat --model MRT --dtype $dt --sustain 0.4 >> $O/sweep_mrt_folded.jsonl 2>> $O/sweep.err
done; done
cut -c1-260 $O/sweep_mrt_folded.jsonl
timeout 120 python tools/profile_case.py --lattice D2Q9 --diag > $O/diag_d2q9_v2.json 2>&1; cat $O/diag_d2q9_v2.json
timeout 2400 python -m pytest tests -m gpu -x -q --durations=8 > $O/pytest_gpu_all.log 2>&1; echo "pytest all rc=$?"; tail -15 $O/pytest_gpu_all.log

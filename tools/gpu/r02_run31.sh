#!/bin/bash
# round 2, run 31 (1 GPU): smoke of the final library (copy pipeline after the exception-safety change)
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 120 python tools/bench_array_ops.py --sizes 2048 --reps 3 2>&1 | tail -1 | cut -c1-400
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "asynchronous or stream_matches or collide_matches" 2>&1 | tail -2

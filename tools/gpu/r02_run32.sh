#!/bin/bash
# round 2, run 32 (1 GPU): last sanity of the rebuilt library: smoke, reductions, error norms, bench contract
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "reductions or error_norms or process_sums or contract or couette" 2>&1 | tail -2

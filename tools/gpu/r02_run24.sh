#!/bin/bash
# round 2, run 24 (2 GPUs): L2 prefetch also in the peer-memory instances of the fused kernel (merged launch): parity + C3
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 600 python -m pytest tests/test_gpu_multi.py -q -k "merged or stress" > $O/pytest_multi_run24.log 2>&1; echo "pytest multi rc=$?"; tail -3 $O/pytest_multi_run24.log
for cfg in "C3 1" "C3 4"; do set -- $cfg
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --config $1 --also-shrink $2 --steps 5 --warmup 3 --no-cpu --no-parity --no-e2e > $O/r24_$1_$2_n2.json 2> $O/r24.err; python -c "
import json; d=json.loads(open('$O/r24_$1_$2_n2.json').read().strip().splitlines()[-1]); print('$1 shrink $2', round(d['value']), d['ms_per_step'], d['gpu_launches'])"
done

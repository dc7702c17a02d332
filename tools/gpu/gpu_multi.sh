#!/bin/bash
# multi-GPU pass (run with gpurun --gpus N): slab parity tests + weak-scaling bench at 1..N
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L | head -8
timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -5 gpurun_out/pytest_multi.log
for n in 1 2 4 8; do
  if [ $n -le $N ]; then
    if [ $n -eq 1 ]; then
      timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
    else
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 --no-cpu > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
    fi
    echo "bench N=$n rc=$?"; grep -o '"value": [0-9.]*' gpurun_out/scale_$n.json | head -2; tail -2 gpurun_out/scale_$n.err
  fi
done

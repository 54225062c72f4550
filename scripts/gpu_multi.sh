#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/gpus.txt
echo "== multi tests"; timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -30 | tee gpurun_out/pytest_multi.log
N=$(nvidia-smi -L | wc -l)
echo "== bench c3 N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --config c3 --steps 2 --warmup 3 --e2e-steps 1 2>&1 | grep -v "^W\|^\[W\|OMP_NUM" | tail -3 | tee gpurun_out/bench_c3_n$N.log

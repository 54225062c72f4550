#!/bin/bash
# First GPU round: tests, smoke, benches.  Everything logs into gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
free -g | head -2 >> gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
echo "== pytest" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench c1"; timeout 300 python bench.py --config c1 --steps 3 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_c1.log
echo "== bench c2"; timeout 600 python bench.py --config c2 --steps 2 --warmup 3 --no-cpu 2>&1 | tail -3 | tee gpurun_out/bench_c2.log
echo "== bench c3"; timeout 900 python bench.py --config c3 --steps 1 --warmup 3 --no-cpu --e2e-steps 1 2>&1 | tail -3 | tee gpurun_out/bench_c3.log

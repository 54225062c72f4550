#!/bin/bash
# round 2, session A: tests + first bench of the fused Lanczos step
mkdir -p gpurun_out
(nproc; free -g; lscpu | head -20; nvidia-smi --query-gpu=name,memory.total --format=csv) > gpurun_out/r2_box.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
SNAPB200_DEBUG=1 python bench.py --config c3s --steps 3 --warmup 2 --no-cpu --no-e2e > gpurun_out/r2a_c3s.json 2> gpurun_out/r2a_c3s.err
tail -c 1500 gpurun_out/r2a_c3s.json
python bench.py --config c3 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2a_c3.json 2> gpurun_out/r2a_c3.err
tail -c 2500 gpurun_out/r2a_c3.json
SNAPB200_SYNC_RATIO=0 python bench.py --config c3 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2a_c3_spec.json 2> gpurun_out/r2a_c3_spec.err
SNAPB200_SYNC_RATIO=1e300 python bench.py --config c3 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2a_c3_sync.json 2> gpurun_out/r2a_c3_sync.err

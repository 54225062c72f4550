#!/bin/bash
mkdir -p gpurun_out
echo "== pytest tiled"; timeout 900 python -m pytest tests/test_gpu_tiled.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_tiled.log
for MODE in auto+matched; do
for B in 4 8; do
echo "== bench c3 block=$B spmm=$MODE"; timeout 900 python bench.py --config c3 --steps 2 --warmup 3 --no-cpu --no-e2e --block $B --spmm $MODE 2>&1 | tail -1 | tee gpurun_out/bench_c3_b${B}_$MODE.log | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['frac'], d['roofline']['ms_pass1'], d['roofline']['ms_pass2'], d['timing'], d['solver'])"
done; done

#!/bin/bash
mkdir -p gpurun_out
echo "== pytest tiled"; timeout 900 python -m pytest tests/test_gpu_tiled.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -30 | tee gpurun_out/pytest_tiled.log
for MODE in auto auto+matched; do
echo "== bench c3 block=4 spmm=$MODE"; timeout 900 python bench.py --config c3 --steps 2 --warmup 3 --no-cpu --no-e2e --block 4 --spmm $MODE 2>&1 | tail -1 | tee gpurun_out/bench_c3_b4_$MODE.log | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['frac'], d['roofline']['ms_pass1'], d['roofline']['ms_pass2'], d['timing'], d['solver'])"
done
echo "== launch list c3 b4 matched"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c3_b4_matched.csv python scripts/profile_op.py c3 auto+matched solve4 > gpurun_out/launches_c3.out 2>&1
tail -2 gpurun_out/launches_c3.out

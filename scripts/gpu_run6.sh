#!/bin/bash
mkdir -p gpurun_out
echo "== pytest tiled"; timeout 900 python -m pytest tests/test_gpu_tiled.py -m gpu -x -q 2>&1 | tail -30 | tee gpurun_out/pytest_tiled.log
echo "== sanitizer"; timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_tiled.py -m gpu -x -q -k "operator and False-4" 2>&1 | tail -8 | tee gpurun_out/sanitizer.log
for B in 8 4; do
echo "== bench c3 block=$B"; timeout 900 python bench.py --config c3 --steps 2 --warmup 3 --no-cpu --no-e2e --block $B 2>&1 | tail -1 | tee gpurun_out/bench_c3_b$B.log | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['frac'], d['roofline']['ms_pass1'], d['roofline']['ms_pass2'], d['timing'], d['solver'])"
done
echo "== ncu b4"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sell_spmm -c 2 -o gpurun_out/prof_sell_c3_b4 -f python scripts/profile_op.py c3 tiled op4 > gpurun_out/prof_sell_c3_b4.out 2>&1
tail -2 gpurun_out/prof_sell_c3_b4.out

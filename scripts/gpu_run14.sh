#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
echo "== sanitizer padded"; timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_tiled.py -m gpu -x -q -k "operator and False-4" 2>&1 | tail -6 | tee gpurun_out/sanitizer.log
for B in 4 8; do
echo "== bench c3 padded b=$B"; timeout 900 python bench.py --config c3 --steps 3 --warmup 3 --no-cpu --no-e2e --block $B 2>&1 | tail -1 | tee gpurun_out/bench_c3_padded_b$B.log | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['solver']; print(round(d['ms_per_step'],1), round(d['roofline']['frac'],4), round(d['roofline']['ms_pass1'],2), round(d['roofline']['ms_pass2'],2), d['timing']['step_wall_ms'], {k:round(s[k],1) for k in ('ms_transpose','ms_prepare','ms_format','ms_eigsh','ms_spmm')}, s['n_ops'])"
done

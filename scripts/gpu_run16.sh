#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
echo "== bench c3"; timeout 900 python bench.py --config c3 --steps 3 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | tee gpurun_out/bench_c3_v18.log | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['solver']; print(round(d['ms_per_step'],1), round(d['roofline']['frac'],4), d['timing']['step_wall_ms'], {k:round(s[k],1) for k in ('ms_transpose','ms_prepare','ms_format','ms_eigsh','ms_spmm')}, s['n_ops'])"
echo "== bench c2"; timeout 900 python bench.py --config c2 --steps 3 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['solver']; print(round(d['ms_per_step'],1), round(d['roofline']['frac'],4), d['timing']['step_wall_ms'], {k:round(s[k],1) for k in ('ms_transpose','ms_prepare','ms_format','ms_eigsh','ms_spmm')}, s['n_ops'])"
echo "== launches c3"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3200 --csv --log-file gpurun_out/launches_c3_v18.csv python bench.py --config c3 --steps 1 --warmup 0 --no-cpu --no-e2e > gpurun_out/launches_c3_v18.out 2>&1
tail -2 gpurun_out/launches_c3_v18.out | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_emit_kernel -c 1 -f -o gpurun_out/prof_temit python scripts/profile_op.py c2 auto op4 > gpurun_out/prof_temit.out 2>&1

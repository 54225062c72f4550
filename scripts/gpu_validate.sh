#!/bin/bash
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
echo "== bench c3 (default flags)"; timeout 1200 python bench.py 2>gpurun_out/bench_c3_full.err | tail -1 | tee gpurun_out/bench_c3_full.log | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['solver']; print(round(d['ms_per_step'],1), round(d['roofline']['frac'],4), d['timing']['step_wall_ms'], {k:round(s[k],1) for k in ('ms_transpose','ms_prepare','ms_format','ms_eigsh','ms_spmm')}, s['n_ops'], 'e2e', d['e2e'], 'cpu', d['cpu_baseline'])"
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 1 --warmup 0 2>/dev/null | tail -1 | tee gpurun_out/bench_ref.log | cut -c1-400
echo "== bench c2"; timeout 900 python bench.py --config c2 --steps 3 --warmup 3 --no-cpu --no-e2e 2>/dev/null | tail -1 | tee gpurun_out/bench_c2.log | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['solver']; print(round(d['ms_per_step'],1), round(d['roofline']['frac'],4))"
# launch list: ncu --metrics gpu__time_duration.sum --clock-control none -c 3200 --csv --log-file gpurun_out/launches.csv python bench.py --config c3 --steps 1 --warmup 0 --no-cpu --no-e2e

#!/bin/bash
mkdir -p gpurun_out
echo "== pytest tiled"; timeout 1200 python -m pytest tests/test_gpu_tiled.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
echo "== bench c3 default"; timeout 900 python bench.py --config c3 --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_c3_default.log | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['solver']; print(round(d['ms_per_step'],1), 'e2e', round(d['e2e']['ms_per_step'],1), round(d['roofline']['frac'],4), round(d['roofline']['ms_pass1'],2), round(d['roofline']['ms_pass2'],2), d['timing']['last_step_call_ms'], {k:round(s[k],1) for k in ('ms_transpose','ms_prepare','ms_format','ms_eigsh','ms_spmm','ms_pool')}, s['n_ops'], d['cpu_baseline']['value'])"
echo "== bench c2 default"; timeout 900 python bench.py --config c2 --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | tee gpurun_out/bench_c2_default.log | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['solver']; print(round(d['ms_per_step'],1), 'e2e', round(d['e2e']['ms_per_step'],1), round(d['roofline']['frac'],4), {k:round(s[k],1) for k in ('ms_transpose','ms_prepare','ms_format','ms_eigsh','ms_spmm')}, s['n_ops'])"
echo "== bench c1 default"; timeout 900 python bench.py --config c1 --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_c1_default.log | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['solver']; print(round(d['ms_per_step'],1), 'e2e', round(d['e2e']['ms_per_step'],1), round(d['roofline']['frac'],4), {k:round(s[k],1) for k in ('ms_transpose','ms_prepare','ms_format','ms_eigsh','ms_spmm','ms_host')}, s['n_ops'], d['cpu_baseline']['value'])"
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 1 --warmup 0 2>&1 | tail -1 | cut -c1-400

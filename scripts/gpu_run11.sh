#!/bin/bash
mkdir -p gpurun_out
echo "== pytest all"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
for MODE in auto auto+matched; do
echo "== bench c3 default block spmm=$MODE"; timeout 900 python bench.py --config c3 --steps 3 --warmup 3 --no-cpu --e2e-steps 1 --spmm $MODE 2>&1 | tail -1 | tee gpurun_out/bench_c3_$MODE.log | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['solver']; print(round(d['ms_per_step'],1), 'e2e', round(d['e2e']['ms_per_step'],1), round(d['roofline']['frac'],4), round(d['roofline']['ms_pass1'],2), round(d['roofline']['ms_pass2'],2), d['timing']['last_step_call_ms'], {k:round(s[k],1) for k in ('ms_transpose','ms_prepare','ms_format','ms_eigsh','ms_spmm','ms_pool')}, s['n_ops'], s['pool_mallocs'], d['clocks'])"
done
echo "== launch list c3 default"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c3_default.csv python scripts/profile_op.py c3 auto solve4 > gpurun_out/launches_c3.out 2>&1
tail -1 gpurun_out/launches_c3.out | cut -c1-300

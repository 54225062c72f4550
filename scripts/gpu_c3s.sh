#!/bin/bash
mkdir -p gpurun_out
echo "== tiled tests"; timeout 900 python -m pytest tests/test_gpu_tiled.py -m gpu -x -q 2>&1 | tail -3
for cfg in c3s c3; do
timeout 600 python bench.py --config $cfg --steps 3 --warmup 3 --no-cpu --no-e2e 2>/dev/null | tail -1 | tee gpurun_out/bench_$cfg.log | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['solver']; r=d['roofline']; print('$cfg', round(d['ms_per_step'],2), round(r['frac'],4), 'p1',round(r['ms_pass1'],3),'p2',round(r['ms_pass2'],3), {k:round(s[k],2) for k in ('ms_transpose','ms_prepare','ms_format','ms_eigsh','ms_spmm','ms_ortho','ms_host','ms_comm')}, s['n_ops'])"
done

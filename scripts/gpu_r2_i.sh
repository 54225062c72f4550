#!/bin/bash
# sweep of the per-chunk charge of the SpMM range split on a 1/8 shard and on C3
mkdir -p gpurun_out
for cc in 0 1 2 3 5 8 12; do
SNAPB200_CHUNK_COST=$cc python - <<P
import json,subprocess,sys,os
out=subprocess.run([sys.executable,'bench.py','--config','c3s','--steps','2','--warmup','1','--no-cpu','--no-e2e','--op-iters','10'],capture_output=True,text=True).stdout.strip().splitlines()[-1]
d=json.loads(out); r=d['roofline']; print('c3s cost',os.environ['SNAPB200_CHUNK_COST'], 'p1 %.4f p2 %.4f frac %.4f step %.2f'%(r['ms_pass1'],r['ms_pass2'],r['frac'],d['ms_per_step']))
P
done
for cc in 1 3 6; do
SNAPB200_CHUNK_COST=$cc python - <<P
import json,subprocess,sys,os
out=subprocess.run([sys.executable,'bench.py','--config','c3','--steps','2','--warmup','1','--no-cpu','--no-e2e','--op-iters','5'],capture_output=True,text=True).stdout.strip().splitlines()[-1]
d=json.loads(out); r=d['roofline']; print('c3 cost',os.environ['SNAPB200_CHUNK_COST'], 'p1 %.4f p2 %.4f frac %.4f step %.2f'%(r['ms_pass1'],r['ms_pass2'],r['frac'],d['ms_per_step']))
P
done
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "int64 or deferred" 2>&1 | tail -2

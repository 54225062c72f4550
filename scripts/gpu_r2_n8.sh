#!/bin/bash
# round 2, 8-GPU session: C3 scaling point, shard-invariance tests (incl. C3 N-rank vs 1-rank), C5, C4 (+ Nystrom comparator)
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
nvidia-smi -L > gpurun_out/r2n8_gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
F='^W\|^\[W\|OMP_NUM\|^\*\*\*\|Setting OMP'
echo "== bench c3 N=$N"
timeout 400 $TR --master-port 29541 bench.py --gpus $N --config c3 --steps 10 --warmup 3 --no-cpu 2> gpurun_out/r2n8_c3.err | grep -v "$F" | tail -1 > gpurun_out/r2n8_c3_n$N.json
python - <<P
import json
d=json.loads(open('gpurun_out/r2n8_c3_n$N.json').read()); print(d['ms_per_step'], d['value'], d['solver'], d['e2e'].get('ms_per_step') if d['e2e'] else None, d['roofline']['frac'])
P
echo "== C3 N-rank vs 1-rank"
SNAPB200_MULTI_CONFIG=c3 timeout 400 $TR --master-port 29542 tests/_multi_worker.py tiled 2>&1 | grep -v "$F" | tail -3 | tee gpurun_out/r2n8_c3_shard_invariance.log
echo "== multi tests"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2n8_pytest_multi.log
echo "== bench c5 N=$N"
timeout 400 $TR --master-port 29543 bench.py --gpus $N --config c5 --steps 3 --warmup 2 2> gpurun_out/r2n8_c5.err | grep -v "$F" | tail -1 > gpurun_out/r2n8_c5_n$N.json
tail -c 900 gpurun_out/r2n8_c5_n$N.json; echo
echo "== bench c4 N=$N"
timeout 600 $TR --master-port 29544 bench.py --gpus $N --config c4 --steps 3 --warmup 2 --no-cpu --nystrom 100000 2> gpurun_out/r2n8_c4.err | grep -v "$F" | tail -1 > gpurun_out/r2n8_c4_n$N.json
tail -c 2500 gpurun_out/r2n8_c4_n$N.json; echo
tail -3 gpurun_out/r2n8_c4.err

#!/bin/bash
# round 2, second 8-GPU session (final build): C3 at N=8 and N=4, C5, C4 (+ Nystrom comparator)
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
F='^W\|^\[W\|OMP_NUM\|^\*\*\*\|Setting OMP'
run() {  # run <nproc> <port> <outfile> <bench args...>
  np=$1; port=$2; out=$3; shift 3
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $port bench.py --gpus $np "$@" 2> gpurun_out/$out.err | grep -v "$F" | tail -1 > gpurun_out/$out.json
  python - <<P
import json
try:
    d=json.loads(open('gpurun_out/$out.json').read()); print('$out', d['ms_per_step'], d['value'], d.get('solver'), (d.get('e2e') or {}).get('ms_per_step'))
except Exception as e: print('$out failed', e)
P
}
run $N 29541 r2n8b_c3_n$N --config c3 --steps 10 --warmup 3 --no-cpu
run 4 29542 r2n8b_c3_n4 --config c3 --steps 10 --warmup 3 --no-cpu --no-e2e
run 2 29543 r2n8b_c3_n2 --config c3 --steps 10 --warmup 3 --no-cpu --no-e2e
run $N 29544 r2n8b_c5_n$N --config c5 --steps 3 --warmup 2
run $N 29545 r2n8b_c4_n$N --config c4 --steps 3 --warmup 2 --no-cpu --no-e2e --nystrom 100000
tail -2 gpurun_out/r2n8b_c4_n$N.err

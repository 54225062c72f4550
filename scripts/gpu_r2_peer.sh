#!/bin/bash
# fused small all-reduce over peer memory: correctness (shard invariance tests) and step time, against NCCL
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
F='^W\|^\[W\|OMP_NUM\|^\*\*\*\|Setting OMP'
echo "== multi tests (peer path)"; timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2p_pytest_multi.log
for mode in peer nccl; do
  if [ $mode = nccl ]; then export SNAPB200_NO_PEER=1; else unset SNAPB200_NO_PEER; fi
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$N bench.py --gpus $N --config c3 --steps 10 --warmup 3 --no-cpu --no-e2e 2> gpurun_out/r2p_$mode.err | grep -v "$F" | tail -1 > gpurun_out/r2p_c3_n${N}_$mode.json
  python - <<P
import json
d=json.loads(open('gpurun_out/r2p_c3_n${N}_$mode.json').read()); print('$mode', d['ms_per_step'], {k:d['solver'][k] for k in ('ms_eigsh','ms_spmm','ms_ortho','ms_comm','fused_allreduce','n_ops')}, d['evals_head'])
P
done
tail -3 gpurun_out/r2p_peer.err

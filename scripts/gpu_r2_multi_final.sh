#!/bin/bash
# final multi-GPU check: shard-invariance tests (operator, multi-view, Nystrom, pp.knn) and one bench line with the e2e leg
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
F='^W\|^\[W\|OMP_NUM\|^\*\*\*\|Setting OMP'
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2w_pytest_multi.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2957$N bench.py --gpus $N --config c3 --steps 5 --warmup 3 --no-cpu 2> gpurun_out/r2w_bench.err | grep -v "$F" | tail -1 > gpurun_out/r2w_c3_n${N}.json
python - <<P
import json
d=json.loads(open('gpurun_out/r2w_c3_n${N}.json').read()); print(d['n_gpus'], d['ms_per_step'], d['e2e'] and {k:d['e2e'][k] for k in ('ms_per_step','ms_load','h2d_bytes_per_step','host_threads')}, d['evals_head'])
P
tail -2 gpurun_out/r2w_bench.err

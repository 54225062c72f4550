#!/bin/bash
# where does the index transfer spend its time?  host probe, then the e2e line (SNAPB200_DEBUG prints the staging team's accounting)
mkdir -p gpurun_out
scripts/_build/host_bw_probe 2000 0 2>&1 | tee gpurun_out/r2k_host_bw_probe.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "delta or int64 or blockwise" 2>&1 | tail -3
for mode in delta plain; do
  export SNAPB200_DEBUG=1; unset SNAPB200_NO_DELTA SNAPB200_THREADS
  [ $mode = plain ] && export SNAPB200_NO_DELTA=1
  timeout 600 python bench.py --gpus 1 --steps 1 --warmup 1 --no-cpu --e2e-steps 3 > gpurun_out/r2k_$mode.json 2> gpurun_out/r2k_$mode.err
  grep "stage_indices" gpurun_out/r2k_$mode.err | tail -4
  python -c "
import json,sys
d=json.loads(open('gpurun_out/r2k_$mode.json').read().strip().splitlines()[-1]); e=d['e2e']; print('$mode', e['ms_per_step'], e['ms_load'], e['h2d_bytes_per_step'])"
done

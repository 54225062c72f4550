#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_tiled.py tests/test_gpu_parity.py -m gpu -q -x -k "transpose or config2 or tiled_operator" > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -4 gpurun_out/r2d_pytest.log
SNAPB200_TRANSPOSE=bucketed SNAPB200_DEBUG=1 python bench.py --config c3 --steps 3 --warmup 2 --no-cpu --no-e2e > gpurun_out/r2d_c3.json 2> gpurun_out/r2d_c3.err
grep "bucketed" gpurun_out/r2d_c3.err | tail -1
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2d_c3.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['solver'])
P
SNAPB200_DEBUG=1 python bench.py --config c3s --steps 2 --warmup 2 --no-cpu --no-e2e 2>&1 >/dev/null | grep bucketed | tail -1
ncu --set full --clock-control none --import-source on -k regex:tr_ -c 4 -o gpurun_out/r2d_prof_tr3 python bench.py --config c3s --steps 1 --warmup 1 --no-cpu --no-e2e --op-iters 1 > gpurun_out/r2d_ncu.out 2>&1

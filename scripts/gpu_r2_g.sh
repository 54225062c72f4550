#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_tiled.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
tail -4 gpurun_out/r2g_pytest.log
SNAPB200_DEBUG=1 python bench.py --config c3 --steps 3 --warmup 2 --no-cpu --no-e2e > gpurun_out/r2g_c3.json 2> gpurun_out/r2g_c3.err
grep "bucketed" gpurun_out/r2g_c3.err | tail -1
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2g_c3.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['solver'])
P
python bench.py --config c3s --steps 3 --warmup 2 --no-cpu --no-e2e > gpurun_out/r2g_c3s.json 2> /dev/null
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2g_c3s.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['solver'])
P

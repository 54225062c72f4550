#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -6 gpurun_out/r2f_pytest.log
python bench.py --config c3 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2f_c3.json 2> gpurun_out/r2f_c3.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2f_c3.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['solver'])
P
python bench.py --config c3s --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2f_c3s.json 2> gpurun_out/r2f_c3s.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2f_c3s.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['solver'])
P
SNAPB200_TRANSPOSE=bucketed python bench.py --config c3s --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2f_c3s_bucketed.json 2> /dev/null
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2f_c3s_bucketed.json').read().strip().splitlines()[-1]); print('bucketed c3s', d['ms_per_step'], d['solver']['ms_transpose'])
P

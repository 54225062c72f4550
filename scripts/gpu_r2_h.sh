#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
tail -6 gpurun_out/r2h_pytest.log
python bench.py --config c3 --steps 10 --warmup 3 > gpurun_out/r2h_c3.json 2> gpurun_out/r2h_c3.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2h_c3.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']); print(d['cpu_baseline']['value'], d['roofline']['frac'])
P
tail -3 gpurun_out/r2h_c3.err

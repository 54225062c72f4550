#!/bin/bash
# round 2, session B: tests (multi-view on device, staged ingest), C3 bench with the plugin-call e2e
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -15 gpurun_out/r2b_pytest.log
python bench.py --config c3 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2b_c3.json 2> gpurun_out/r2b_c3.err
tail -c 3000 gpurun_out/r2b_c3.json; tail -5 gpurun_out/r2b_c3.err
SNAPB200_THREADS=8 python bench.py --config c3 --steps 1 --warmup 1 --no-cpu > gpurun_out/r2b_c3_t8.json 2> gpurun_out/r2b_c3_t8.err
python bench.py --config c2 --steps 3 --warmup 2 --no-cpu > gpurun_out/r2b_c2.json 2> gpurun_out/r2b_c2.err

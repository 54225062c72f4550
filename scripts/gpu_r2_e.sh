#!/bin/bash
# round 2, session E: whole GPU suite, C3 with the plugin-call e2e, Nystrom comparator (C2), multi-view (c5s)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -15 gpurun_out/r2e_pytest.log
python bench.py --config c3 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2e_c3.json 2> gpurun_out/r2e_c3.err
tail -c 600 gpurun_out/r2e_c3.json; tail -3 gpurun_out/r2e_c3.err
python bench.py --config c2 --steps 3 --warmup 2 --no-cpu --nystrom 20000 > gpurun_out/r2e_c2_nys.json 2> gpurun_out/r2e_c2_nys.err
tail -3 gpurun_out/r2e_c2_nys.err
python bench.py --config c5s --steps 3 --warmup 2 > gpurun_out/r2e_c5s.json 2> gpurun_out/r2e_c5s.err
tail -c 1200 gpurun_out/r2e_c5s.json; tail -3 gpurun_out/r2e_c5s.err

#!/bin/bash
# round 2, session C: bucketed transpose
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -12 gpurun_out/r2c_pytest.log
python bench.py --config c3 --steps 4 --warmup 3 --no-cpu > gpurun_out/r2c_c3.json 2> gpurun_out/r2c_c3.err
tail -c 1500 gpurun_out/r2c_c3.json; tail -3 gpurun_out/r2c_c3.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c_launches_c3.csv python bench.py --config c3 --steps 1 --warmup 1 --no-cpu --no-e2e --op-iters 1 > gpurun_out/r2c_launches.out 2>&1
python bench.py --config c3s --steps 3 --warmup 2 --no-cpu --no-e2e > gpurun_out/r2c_c3s.json 2> gpurun_out/r2c_c3s.err
SNAPB200_TR_CTAS=4 python bench.py --config c3 --steps 2 --warmup 2 --no-cpu --no-e2e > gpurun_out/r2c_c3_ctas4.json 2> gpurun_out/r2c_c3_ctas4.err
SNAPB200_TR_CTAS=1 python bench.py --config c3 --steps 2 --warmup 2 --no-cpu --no-e2e > gpurun_out/r2c_c3_ctas1.json 2> gpurun_out/r2c_c3_ctas1.err

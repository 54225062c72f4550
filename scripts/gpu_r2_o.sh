#!/bin/bash
# kNN: where does the time go?  bare filter loop (probe 1) vs the full kernel, and a full ncu capture at 400k
mkdir -p gpurun_out
SNAPB200_KNN_PROBE=1 timeout 600 python scripts/bench_knn.py --n 1000000 --steps 1 --warmup 1 --no-cpu > gpurun_out/r2o_probe1_1m.json 2> gpurun_out/r2o_probe1.err; python -c "
import json; d=json.loads(open('gpurun_out/r2o_probe1_1m.json').read()); print('probe1 1M ms', d['ms_per_step'], d['roofline']['frac'])"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_scan -c 1 -o gpurun_out/r2o_knn_scan_400k python scripts/bench_knn.py --n 400000 --steps 1 --warmup 0 --no-cpu > gpurun_out/r2o_ncu.log 2>&1; tail -2 gpurun_out/r2o_ncu.log

#!/bin/bash
# kNN after staging the float64 rows: parity, timing, one full ncu capture of the scan kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_knn.py -m gpu -q -x 2>&1 | tail -4
timeout 600 python scripts/bench_knn.py --n 1000000 --steps 2 --no-cpu > gpurun_out/r2m_knn_1m.json 2> gpurun_out/r2m_knn_1m.err; tail -c 900 gpurun_out/r2m_knn_1m.json; tail -3 gpurun_out/r2m_knn_1m.err
timeout 600 python scripts/bench_knn.py --n 200000 --steps 2 --no-cpu > gpurun_out/r2m_knn_200k.json 2> gpurun_out/r2m_knn_200k.err; tail -c 500 gpurun_out/r2m_knn_200k.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn_scan -c 1 -o gpurun_out/r2m_knn_scan_200k python scripts/bench_knn.py --n 200000 --steps 1 --warmup 0 --no-cpu > gpurun_out/r2m_ncu.log 2>&1; tail -2 gpurun_out/r2m_ncu.log

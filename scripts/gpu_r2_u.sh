#!/bin/bash
# kNN: unroll 4 vs 8 at 1M, then the evidence for profiles/: launch list and one full ncu capture (200k points)
mkdir -p gpurun_out
for u in 4 8; do
SNAPB200_KNN_UNROLL=$u timeout 600 python scripts/bench_knn.py --n 1000000 --steps 2 --no-cpu > gpurun_out/r2u_knn_1m_u$u.json 2> gpurun_out/r2u_knn_1m_u$u.err; python -c "
import json; d=json.loads(open('gpurun_out/r2u_knn_1m_u$u.json').read()); print('unroll $u', 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'])"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2u_knn_launches.csv python scripts/bench_knn.py --n 200000 --steps 1 --warmup 0 --no-cpu > /dev/null 2>&1
grep -c knn gpurun_out/r2u_knn_launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn_scan -c 1 -o gpurun_out/r2u_knn_scan_200k python scripts/bench_knn.py --n 200000 --steps 1 --warmup 0 --no-cpu > gpurun_out/r2u_ncu.log 2>&1; tail -1 gpurun_out/r2u_ncu.log

#!/bin/bash
# kNN cycle accounting (probe 4) at 1M and 200k; bare filter loop (probe 1) for reference
mkdir -p gpurun_out
for n in 1000000 200000; do
SNAPB200_KNN_PROBE=4 timeout 600 python scripts/bench_knn.py --n $n --steps 1 --warmup 0 --no-cpu > gpurun_out/r2s_probe4_$n.json 2> gpurun_out/r2s_probe4_$n.err
grep "knn probe" gpurun_out/r2s_probe4_$n.err; python -c "
import json; d=json.loads(open('gpurun_out/r2s_probe4_$n.json').read()); print($n, 'probe4 ms', d['ms_per_step'])"
done
SNAPB200_KNN_PROBE=1 timeout 600 python scripts/bench_knn.py --n 1000000 --steps 1 --warmup 0 --no-cpu > gpurun_out/r2s_probe1.json 2> gpurun_out/r2s_probe1.err; python -c "
import json; d=json.loads(open('gpurun_out/r2s_probe1.json').read()); print('probe1 ms', d['ms_per_step'])"

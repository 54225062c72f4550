#!/bin/bash
# kNN after the sorted-list insertion: parity + timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_knn.py -m gpu -q -x 2>&1 | tail -3
for n in 1000000 200000; do
timeout 600 python scripts/bench_knn.py --n $n --steps 2 --no-cpu > gpurun_out/r2q_knn_$n.json 2> gpurun_out/r2q_knn_$n.err; python -c "
import json; d=json.loads(open('gpurun_out/r2q_knn_$n.json').read()); print($n, 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e ms', d['e2e']['ms_per_step'])"
done

#!/bin/bash
# kNN with producer / consumer warps: memcheck, parity, timing
mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python - > gpurun_out/r2r_memcheck.log 2>&1 <<'PY'
import numpy as np
from snapatac2_b200 import Engine
e = Engine(0)
rng = np.random.default_rng(0)
for n, d, k in [(300, 30, 50), (1000, 7, 10), (129, 64, 74), (2, 2, 5), (700, 30, 100)]:
    P = rng.normal(size=(n, d))
    i, dd = e.knn(P, k)
    print(n, d, k, i.shape, float(dd.max()))
PY
echo "memcheck rc=$?"; tail -3 gpurun_out/r2r_memcheck.log
timeout 900 python -m pytest tests/test_knn.py -m gpu -q -x 2>&1 | tail -4
timeout 600 python scripts/bench_knn.py --n 1000000 --steps 2 --no-cpu > gpurun_out/r2r_knn_1m.json 2> gpurun_out/r2r_knn_1m.err; tail -c 700 gpurun_out/r2r_knn_1m.json; tail -3 gpurun_out/r2r_knn_1m.err
timeout 600 python scripts/bench_knn.py --n 200000 --steps 2 --no-cpu > gpurun_out/r2r_knn_200k.json 2> gpurun_out/r2r_knn_200k.err; tail -c 400 gpurun_out/r2r_knn_200k.json

#!/bin/bash
# first run of the kNN kernels: memcheck on small cases, then the parity tests
mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python - > gpurun_out/r2l_memcheck.log 2>&1 <<'PY'
import numpy as np
from snapatac2_b200 import Engine
e = Engine(0)
rng = np.random.default_rng(0)
for n, d, k in [(300, 30, 50), (1000, 7, 10), (129, 64, 84), (2, 2, 5)]:
    P = rng.normal(size=(n, d))
    i, dd = e.knn(P, k)
    print(n, d, k, i.shape, float(dd.max()))
PY
echo "memcheck rc=$?"; tail -5 gpurun_out/r2l_memcheck.log
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python - > gpurun_out/r2l_racecheck.log 2>&1 <<'PY'
import numpy as np
from snapatac2_b200 import Engine
e = Engine(0)
rng = np.random.default_rng(0)
P = rng.normal(size=(600, 30))
i, dd = e.knn(P, 50)
print(i.shape)
PY
echo "racecheck rc=$?"; tail -5 gpurun_out/r2l_racecheck.log
timeout 900 python -m pytest tests/test_knn.py -m gpu -q -x 2>&1 | tail -25
timeout 600 python scripts/bench_knn.py --n 200000 --steps 2 --cpu-queries 2000 > gpurun_out/r2l_knn_200k.json 2> gpurun_out/r2l_knn_200k.err; tail -c 1500 gpurun_out/r2l_knn_200k.json; tail -3 gpurun_out/r2l_knn_200k.err
timeout 900 python scripts/bench_knn.py --n 1000000 --steps 2 --cpu-queries 4000 > gpurun_out/r2l_knn_1m.json 2> gpurun_out/r2l_knn_1m.err; tail -c 1500 gpurun_out/r2l_knn_1m.json; tail -3 gpurun_out/r2l_knn_1m.err

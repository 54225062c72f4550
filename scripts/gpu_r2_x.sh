#!/bin/bash
# tensor-core (mma.sync tf32 x 3) variant of the kNN filter: parity, then timing against the FFMA filter
mkdir -p gpurun_out
export SNAPB200_KNN_MMA=1
timeout 300 python -m pytest tests/test_knn.py -m gpu -q -x 2>&1 | tail -3
python - <<'PY'
import numpy as np
from snapatac2_b200 import Engine
e = Engine(0); P = np.random.default_rng(0).normal(size=(3000, 30)); e.knn(P, 50); print("knn_mma stat:", e.stats()["knn_mma"])
PY
timeout 300 python scripts/bench_knn.py --n 1000000 --steps 2 --no-cpu > gpurun_out/r2x_knn_1m_mma.json 2> gpurun_out/r2x_knn_1m_mma.err; python -c "
import json; d=json.loads(open('gpurun_out/r2x_knn_1m_mma.json').read()); print('mma 1M ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])"
SNAPB200_KNN_PROBE=4 timeout 300 python scripts/bench_knn.py --n 1000000 --steps 1 --warmup 0 --no-cpu 2>&1 >/dev/null | grep "knn probe" | tee gpurun_out/r2x_probe4_mma.log

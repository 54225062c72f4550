#!/bin/bash
# Nystrom path: parity tests, then its timing on C2 with 20000 landmarks
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "nystrom" 2>&1 | tail -5
timeout 900 python scripts/bench_nystrom.py c2 20000 2>&1 | tail -1 | tee gpurun_out/bench_nystrom_c2.json

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tiled.py -m gpu -x -q -k "nystrom or projection" 2>&1 | tail -25

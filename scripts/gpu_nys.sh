#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/bench_nystrom.py c2 20000 2>&1 | tail -1 | tee gpurun_out/bench_nystrom_c2.json

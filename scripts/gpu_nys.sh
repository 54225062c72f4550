#!/bin/bash
timeout 90 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "nystrom_golden" 2>&1 | tail -8

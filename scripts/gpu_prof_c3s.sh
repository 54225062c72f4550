#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sell_spmm_kernel -c 2 -f -o gpurun_out/prof_spmm_c3s python scripts/profile_op.py c3s auto op4 > gpurun_out/prof_spmm_c3s.out 2>&1
tail -2 gpurun_out/prof_spmm_c3s.out

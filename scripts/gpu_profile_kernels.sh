#!/bin/bash
# ncu --set full captures of the prepare kernels (C2) and the SpMM pair (C3)
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:sell_fill_kernel -c 2 -f -o gpurun_out/prof_fill_v10 python scripts/profile_op.py c2 auto op4 > gpurun_out/prof_fill_v10.out 2>&1
timeout 600 $NCU -k regex:bitmap_emit_kernel --launch-skip 20 -c 1 -f -o gpurun_out/prof_emit_v10 python scripts/profile_op.py c2 auto op4 > gpurun_out/prof_emit_v10.out 2>&1
timeout 600 $NCU -k regex:bitmap_set_kernel --launch-skip 20 -c 1 -f -o gpurun_out/prof_set_v10 python scripts/profile_op.py c2 auto op4 > gpurun_out/prof_set_v10.out 2>&1
timeout 600 $NCU -k regex:col_count_kernel -c 1 -f -o gpurun_out/prof_count_v10 python scripts/profile_op.py c2 auto op4 > gpurun_out/prof_count_v10.out 2>&1
timeout 900 $NCU -k regex:sell_spmm_kernel -c 2 -f -o gpurun_out/prof_spmm_c3_v10 python scripts/profile_op.py c3 auto op4 > gpurun_out/prof_spmm_c3_v10.out 2>&1
ls -la gpurun_out/*.ncu-rep
tail -3 gpurun_out/prof_spmm_c3_v10.out

#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sell_fill -c 2 -o gpurun_out/prof_fill_c3_b4 -f python scripts/profile_op.py c3 auto+matched op4 > gpurun_out/prof_fill.out 2>&1
tail -3 gpurun_out/prof_fill.out

#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_emit_kernel -c 1 -f -o gpurun_out/prof_temit python scripts/profile_op.py c2 auto op4 > gpurun_out/prof_temit.out 2>&1
tail -2 gpurun_out/prof_temit.out

#!/bin/bash
mkdir -p gpurun_out
echo "== plain timing c2 tiled solve"; timeout 300 python scripts/profile_op.py c2 tiled solve 2>&1 | tail -3
echo "== plain timing c2 csr solve"; timeout 300 python scripts/profile_op.py c2 csr solve 2>&1 | tail -3
echo "== launch list c2 tiled (solve)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c2_tiled.csv python scripts/profile_op.py c2 tiled solve > gpurun_out/launches_c2_tiled.out 2>&1
tail -2 gpurun_out/launches_c2_tiled.out
echo "== ncu full on tiled kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sell_spmm8 -c 2 -o gpurun_out/prof_sell_c2 -f python scripts/profile_op.py c2 tiled op > gpurun_out/prof_sell_c2.out 2>&1
tail -2 gpurun_out/prof_sell_c2.out
ls -la gpurun_out/

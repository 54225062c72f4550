#!/bin/bash
mkdir -p gpurun_out
echo "== pytest tiled"; timeout 600 python -m pytest tests/test_gpu_tiled.py -m gpu -x -q 2>&1 | tail -30 | tee gpurun_out/pytest_tiled.log
echo "== sanitizer"; timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_tiled.py -m gpu -x -q -k "operator and False" 2>&1 | tail -12 | tee gpurun_out/sanitizer.log
echo "== pytest all"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.log
echo "== bench c2"; timeout 600 python bench.py --config c2 --steps 3 --warmup 3 --no-cpu 2>&1 | tail -2 | tee gpurun_out/bench_c2.log
echo "== bench c3"; timeout 900 python bench.py --config c3 --steps 2 --warmup 3 --no-cpu --e2e-steps 1 2>&1 | tail -2 | tee gpurun_out/bench_c3.log
echo "== ncu full tiled c3"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sell_spmm8 -c 2 -o gpurun_out/prof_sell_c3 -f python scripts/profile_op.py c3 tiled op > gpurun_out/prof_sell_c3.out 2>&1
tail -2 gpurun_out/prof_sell_c3.out

#!/bin/bash
# what the driver runs at round end, on one GPU: GPU tests, smoke, default bench (both arms); then the pp.knn bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2v_pytest.log
tail -4 gpurun_out/r2v_pytest.log
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r2v_smoke.log
( time python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err
tail -c 900 gpurun_out/r2v_bench.json; tail -4 gpurun_out/r2v_bench.err
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r2v_ref.json 2> gpurun_out/r2v_ref.err
tail -c 400 gpurun_out/r2v_ref.json; tail -4 gpurun_out/r2v_ref.err
timeout 600 python scripts/bench_knn.py --n 1000000 --steps 3 --cpu-queries 4000 > gpurun_out/r2v_knn_1m.json 2> gpurun_out/r2v_knn_1m.err; tail -c 1200 gpurun_out/r2v_knn_1m.json; tail -2 gpurun_out/r2v_knn_1m.err

#!/bin/bash
# delta-encoded index transfer: parity of the two transfer modes, the ingest tests, then the e2e line both ways
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "delta or int64 or blockwise or reference_executed or deferred or config2" > gpurun_out/r2j_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log; tail -5 gpurun_out/r2j_pytest.log
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2j_bench_delta.json 2> gpurun_out/r2j_bench_delta.err
tail -c 900 gpurun_out/r2j_bench_delta.json; tail -3 gpurun_out/r2j_bench_delta.err
SNAPB200_NO_DELTA=1 timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2j_bench_plain.json 2> gpurun_out/r2j_bench_plain.err
tail -c 900 gpurun_out/r2j_bench_plain.json; tail -3 gpurun_out/r2j_bench_plain.err

#!/bin/bash
# round 2 profiling session (1 GPU): ncu captures of the final kernels, launch list, final C3 line with the CPU baseline
mkdir -p gpurun_out
B="python bench.py --no-cpu --no-e2e --op-iters 1"
ncu --set full --clock-control none --import-source on -k "regex:tile_emit|sell_fill|sell_spmv64|sell_spmm" -c 8 -o gpurun_out/r02_prof_prepare_spmm_c3 $B --config c3 --steps 1 --warmup 0 > gpurun_out/r02_prof_a.out 2>&1
ncu --set full --clock-control none --import-source on -k "regex:gram_kernel|project_chol|tall_gemm|chol_append" -s 60 -c 5 -o gpurun_out/r02_prof_dense_c3 $B --config c3 --steps 1 --warmup 0 > gpurun_out/r02_prof_b.out 2>&1
ncu --set full --clock-control none --import-source on -k "regex:sell_spmm" -c 2 -o gpurun_out/r02_prof_spmm_c3s $B --config c3s --steps 1 --warmup 0 > gpurun_out/r02_prof_c.out 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_c3.csv $B --config c3 --steps 1 --warmup 1 > gpurun_out/r02_launches.out 2>&1
ls -la gpurun_out/r02_prof_*.ncu-rep
python bench.py --config c3 --steps 10 --warmup 3 > gpurun_out/r02_bench_c3_n1.json 2> gpurun_out/r02_bench_c3_n1.err
tail -c 1800 gpurun_out/r02_bench_c3_n1.json; tail -3 gpurun_out/r02_bench_c3_n1.err
python bench.py --impl reference --config c3 --steps 1 --warmup 0 > gpurun_out/r02_bench_ref_c3.json 2> gpurun_out/r02_bench_ref_c3.err
tail -c 1200 gpurun_out/r02_bench_ref_c3.json

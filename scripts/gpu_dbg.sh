#!/bin/bash
mkdir -p gpurun_out
export PYTHONFAULTHANDLER=1 RANK=0 LOCAL_RANK=0 WORLD_SIZE=1 MASTER_ADDR=127.0.0.1 MASTER_PORT=29577
which gdb cuda-gdb
timeout 300 python tests/_multi_worker.py csr > gpurun_out/dbg_w1.log 2>&1; echo "rc=$?"
grep -n "MULTI_OK\|Fatal\|File " gpurun_out/dbg_w1.log | head
if which gdb >/dev/null; then G=gdb; else G=cuda-gdb; fi
timeout 600 $G -batch -ex run -ex bt --args python tests/_multi_worker.py csr > gpurun_out/dbg_gdb.log 2>&1
grep -n "SIGSEGV\|^#" gpurun_out/dbg_gdb.log | head -30

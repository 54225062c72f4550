"""Small driver for ncu: generate a config, prepare, run a few operator applications / one solve."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from snapatac2_b200 import Engine, synth
import importlib.util
spec_b = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(__file__), "..", "bench.py"))
bench = importlib.util.module_from_spec(spec_b); spec_b.loader.exec_module(bench)

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
mode = sys.argv[2] if len(sys.argv) > 2 else "auto"
what = sys.argv[3] if len(sys.argv) > 3 else "op"
n, m, nnz_row, K, k = bench.CONFIGS[cfg]
spec = synth.make_spec(n, m, nnz_row, K, seed=0)
eng = Engine(0)
eng.set_spmm_mode(mode)
blk = 4 if what.endswith("4") else 8
eng.set_block(blk)
eng.generate(spec)
t0 = time.perf_counter(); eng.prepare(want_outputs=False); t1 = time.perf_counter()
print("prepare wall ms", 1e3 * (t1 - t0), {k2: v for k2, v in eng.stats().items() if k2.startswith("ms_")})
if what.startswith("op"):
    print("operator_time", eng.operator_time(b=blk, iters=2, flush_l2=True))
else:
    t0 = time.perf_counter(); ev, _ = eng.eigsh(k); t1 = time.perf_counter()
    print("eigsh wall ms", 1e3 * (t1 - t0), eng.stats())
eng.close()

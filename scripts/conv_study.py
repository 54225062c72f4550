import sys, os, time
sys.path.insert(0, '/root/repo' if os.path.exists('/root/repo/bench.py') else os.getcwd())
import numpy as np
from snapatac2_b200 import Engine, synth
import importlib.util
spec_b = importlib.util.spec_from_file_location("bench", "bench.py"); bench = importlib.util.module_from_spec(spec_b); spec_b.loader.exec_module(bench)
cfg = sys.argv[1]
n, m, nnz_row, K, k = bench.CONFIGS[cfg]
spec = synth.make_spec(n, m, nnz_row, K, seed=0)
eng = Engine(0); eng.set_spmm_mode("csr"); eng.generate(spec); eng.prepare(want_outputs=False)
ref = None
for block in (8, 4, 16):
    for tol in (1e-5, 1e-4):
        ev, U = eng.eigsh(k, block=block, tol=tol)
        st = eng.stats()
        if ref is None: ref = (ev.copy(), U.copy())
        rel = np.max(np.abs(ev - ref[0]) / np.abs(ref[0])); cos = np.abs(np.sum(U * ref[1], axis=0)).min()
        print(f"{cfg} block={block} tol={tol:g} n_ops={st['n_ops']} restarts={st['n_restarts']} res={st['max_residual']:.2e} rel={rel:.1e} mincos={cos:.7f} ms_spmm={st['ms_spmm']:.1f} ms_ortho={st['ms_ortho']:.1f} ms_host={st['ms_host']:.1f} ms_eigsh={st['ms_eigsh']:.1f}", flush=True)

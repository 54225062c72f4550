"""Timing of the Nystrom path (tl.spectral with sample_size) on a synthetic config, one GPU.
Prints one JSON line: cells/s of the whole call (host CSR in, embedding out) and its split."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.sparse as sp
from snapatac2_b200 import Engine, MiniAnnData, synth, tl
import importlib.util
spec_b = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(__file__), "..", "bench.py"))
bench = importlib.util.module_from_spec(spec_b); spec_b.loader.exec_module(bench)

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
sample = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
n, m, nnz_row, K, k = bench.CONFIGS[cfg]
spec = synth.make_spec(n, m, nnz_row, K, seed=0)
eng = Engine(0)
eng.generate(spec)
X = eng.export_csr()
ad = MiniAnnData(X)
times = []
for it in range(3):
    t0 = time.perf_counter()
    ev, emb = tl.spectral(ad, n_comps=k, features=None, sample_size=sample, inplace=False, engine=eng)
    times.append(time.perf_counter() - t0)
# the full path on the same input, for comparison
t0 = time.perf_counter(); ev_f, emb_f = tl.spectral(ad, n_comps=k, features=None, inplace=False, engine=eng); t_full = time.perf_counter() - t0
t0 = time.perf_counter(); ev_f, emb_f = tl.spectral(ad, n_comps=k, features=None, inplace=False, engine=eng); t_full = time.perf_counter() - t0
# agreement of the two embeddings as subspaces (columns 1..k-1; column 0 is the trivial component)
a = np.linalg.qr(np.real(emb[:, :k]))[0]; b = np.linalg.qr(emb_f[:, :k])[0]
cosines = np.linalg.svd(a.T @ b, compute_uv=False)
print(json.dumps({"metric": "snap.tl.spectral(sample_size) cells/s", "config": cfg, "n_cells": n, "sample_size": sample,
                  "value": n / min(times), "unit": "cells/s", "s_per_call": times, "full_path_s": t_full,
                  "principal_cosines_vs_full_path": [round(float(c), 4) for c in cosines[:8]],
                  "note": "host scipy CSR in, embedding out (includes the 2 x H2D of the matrix and of the landmark rows)"}))

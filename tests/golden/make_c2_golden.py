"""BASELINE.json configs[1] (100k cells x 500k bins, ~5k nnz/cell, n_comps=30) through the CPU oracle.

Run from the repo root (about 15 minutes and ~12 GB on a CPU box):

    python tests/golden/make_c2_golden.py

The full eigenvector matrix (100k x 30 f64 = 24 MB) is not committed; the fixture keeps the
eigenvalues, every 5th degree, every 25th IDF weight and the eigenvector rows of 4000 sampled cells,
which is what tests/test_gpu_parity.py::test_config2_parity_against_oracle compares (a wrong or
rotated eigenvector shows on a 4000-row restriction as surely as on all rows).
"""

import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

import oracle  # noqa: E402
from snapatac2_b200 import synth  # noqa: E402

OUT = Path(__file__).resolve().parent


def _gen(a):
    n, m, nnz_row, K, r0, r1 = a
    return synth.generate_csr(synth.make_spec(n, m, nnz_row, K, seed=0), r0, r1, dtype=np.float64)


def main():
    n, m, nnz_row, K, k = 100_000, 500_000, 5_000, 48, 30
    spec = synth.make_spec(n, m, nnz_row, K, seed=0)
    t0 = time.time()
    from concurrent.futures import ProcessPoolExecutor
    import scipy.sparse as sp
    with ProcessPoolExecutor(max_workers=6) as ex:
        parts = list(ex.map(_gen, [(n, m, nnz_row, K, r0, min(n, r0 + 2500)) for r0 in range(0, n, 2500)]))
    X = sp.vstack(parts, format="csr")
    del parts
    print(f"generated {X.shape} nnz={X.nnz} in {time.time() - t0:.0f} s", flush=True)
    counter = [0]
    t0 = time.time()
    ev, evec, w, deg = oracle.spectral_embedding(X, None, k, 0, counter=counter, return_parts=True)
    print(f"oracle: {time.time() - t0:.0f} s, {counter[0]} mat-vecs, evals[:4]={ev[:4]}", flush=True)
    rows = np.sort(np.random.RandomState(5).choice(n, 4000, replace=False))
    np.savez_compressed(OUT / "c2_100kx500k.npz", evals=ev, rows=rows.astype(np.int64), evecs_rows=evec[rows],
                        degree_5=deg[::5], idf_25=w[::25], nnz=np.int64(X.nnz), matvecs=np.int64(counter[0]),
                        seconds=np.float64(time.time() - t0))


if __name__ == "__main__":
    main()

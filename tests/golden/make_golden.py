"""Generate the golden fixtures in this directory.

Run from the repo root:  python tests/golden/make_golden.py

The reference cannot be built here (Rust/PyO3, no cargo) and ships no golden
vectors for this path, so the fixtures are outputs of the CPU
oracle (``oracle/reference_restatement.py`` -- same scipy ARPACK call as the
reference) on small seeded inputs, cross-checked against the independent dense
``eigh`` statement before being written.  Each .npz holds the CSR input and the
oracle's IDF weights, degrees, eigenvalues and eigenvectors.
"""

import sys
from pathlib import Path

import numpy as np
import scipy.sparse as sp

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

import oracle  # noqa: E402
from snapatac2_b200 import synth  # noqa: E402

OUT = Path(__file__).resolve().parent


def check_and_save(name, X, k, features=None, feature_weights=None):
    X = sp.csr_matrix(X)
    X.sort_indices()
    ev, evec, w, deg = oracle.spectral_embedding(X, features, k, 0, feature_weights, return_parts=True)
    ev_d, evec_d, w_d, deg_d = oracle.dense_check(X, k, features, feature_weights)
    assert np.allclose(w, w_d, rtol=1e-12, atol=0)
    assert np.allclose(deg, deg_d, rtol=1e-10, atol=0)
    assert np.allclose(ev, ev_d, rtol=1e-9, atol=1e-13), (ev, ev_d)
    cos = np.abs(np.sum(evec * evec_d, axis=0))
    gaps = np.minimum(np.abs(np.diff(ev, prepend=np.inf)), np.abs(np.diff(ev, append=-np.inf)))
    assert np.all(cos[gaps > 1e-6] > 1 - 1e-8), cos
    extra = {}
    if features is not None:
        extra["features"] = np.asarray(features)
    if feature_weights is not None:
        extra["feature_weights"] = np.asarray(feature_weights, dtype=np.float64)
    np.savez_compressed(OUT / f"{name}.npz", indptr=X.indptr.astype(np.int64), indices=X.indices.astype(np.int32),
                        data=X.data, shape=np.asarray(X.shape, dtype=np.int64), k=np.int64(k),
                        evals=ev, evecs=evec, idf=w, degree=deg, **extra)
    print(f"{name}: shape={X.shape} nnz={X.nnz} k={k} evals[:4]={ev[:4]} min gap={gaps.min():.2e}")


def main():
    # 1. binarised planted-cluster tile matrix (the synthetic recipe, small)
    spec = synth.make_spec(600, 4000, 220, n_clusters=12, seed=3)
    check_and_save("tile_600x4000", synth.generate_csr(spec), 8)

    # 2. integer count matrix (values path), 300 x 1000, counts 1..4
    rng = np.random.default_rng(7)
    spec = synth.make_spec(300, 1000, 90, n_clusters=6, seed=5)
    X = synth.generate_csr(spec, dtype=np.float64)
    X.data = rng.integers(1, 5, size=X.nnz).astype(np.float64)
    check_and_save("counts_300x1000", X, 5)

    # 3. the shape of the reference's own test_reproducibility input
    #    (tests/test_tools.py:95-110): dense-as-CSR 50 x 100, values in [1, 100]
    rng = np.random.default_rng(11)
    check_and_save("dense_50x100", sp.csr_matrix(rng.uniform(1, 100, size=(50, 100))), 30)

    # 4. feature mask + user feature weights
    spec = synth.make_spec(400, 3000, 150, n_clusters=8, seed=9)
    X = synth.generate_csr(spec)
    rng = np.random.default_rng(13)
    mask = rng.random(3000) < 0.7
    fw = rng.uniform(0.5, 2.0, size=int(mask.sum()))
    check_and_save("masked_400x3000", X, 6, features=mask, feature_weights=fw)


def nystrom_golden():
    """5. Nystrom path (embedding.rs:61-129, 194-267) with a fixed landmark list, cross-checked
    against a dense statement: eigh of the landmark operator, dense products, the same per-chunk
    degree normalisation."""
    spec = synth.make_spec(500, 3000, 160, n_clusters=6, seed=21)
    X = sp.csr_matrix(synth.generate_csr(spec, dtype=np.float64))
    rng = np.random.default_rng(17)
    lm = rng.choice(500, 200, replace=False)
    k, chunk = 6, 150
    ev, q, w, d = oracle.spectral_embedding_nystrom(X, None, k, lm, chunk, return_parts=True)
    # dense statement
    xh = oracle.normalize(X, w).toarray()
    seed = xh[lm]
    s = seed @ seed.T
    np.fill_diagonal(s, 0.0)
    deg = s.sum(axis=1)
    a = s / np.sqrt(np.outer(deg, deg))
    lam, vec = np.linalg.eigh(a)
    pick = np.argsort(-np.abs(lam))[:k]
    pick = pick[np.argsort(-lam[pick])]
    lam, vec = lam[pick], vec[:, pick]
    assert np.allclose(deg, d, rtol=1e-10)
    assert np.allclose(lam, ev, rtol=1e-9)
    u = vec / np.sqrt(deg)[:, None] / lam[None, :]
    proj = seed.T @ u
    qd = []
    for i in range(0, 500, chunk):
        qc = xh[i:i + chunk] @ proj
        t = qc.sum(axis=0) * lam
        dd = qc @ t
        dd[dd <= 0] = np.min(dd[dd > 0])
        qd.append(qc / np.sqrt(dd)[:, None])
    qd = np.vstack(qd)
    cos = np.abs(np.sum(q * qd, axis=0)) / (np.linalg.norm(q, axis=0) * np.linalg.norm(qd, axis=0))
    assert np.all(cos > 1 - 1e-8), cos
    assert np.allclose(np.abs(q), np.abs(qd), rtol=1e-6, atol=1e-10)
    np.savez_compressed(OUT / "nystrom_500x3000.npz", indptr=X.indptr.astype(np.int64), indices=X.indices.astype(np.int32),
                        data=X.data, shape=np.asarray(X.shape, dtype=np.int64), k=np.int64(k), chunk_size=np.int64(chunk),
                        landmarks=lm.astype(np.int64), evals=ev, q=q, idf=w, degree=d)
    print(f"nystrom_500x3000: evals={ev} min|cos| vs dense={cos.min():.12f}")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "nystrom":
        nystrom_golden()
    else:
        main()
        nystrom_golden()

"""GPU (-m gpu): the CUDA path through the C ABI against the CPU oracle, the
committed golden vectors, and size-independent properties.

Tolerances are north_star's: IDF weights and degrees 1e-5 relative,
eigenvalues 1e-4 relative, eigenvectors |cos| >= 0.999 per component (subspace
comparison inside near-degenerate clusters)."""

import numpy as np
import pytest
import scipy.sparse as sp

import oracle
from conftest import load_golden, eigvec_agreement
from snapatac2_b200 import synth, tl, MiniAnnData

pytestmark = pytest.mark.gpu

TOL_VEC = 1e-5      # idf, degree
TOL_EVAL = 1e-4
MIN_COS = 0.999


def _check_against(z_evals, z_evecs, evals, evecs, skip_cluster_tail=False):
    np.testing.assert_allclose(evals, z_evals, rtol=TOL_EVAL, atol=1e-9)
    cos = eigvec_agreement(z_evals, z_evecs, evecs)
    assert cos.min() >= MIN_COS, cos


def test_dense_tensor_core_kernels_selftest(engine):
    # DMMA Gram / projection / rotation against plain fp64 loops
    for n, ncq, p in ((4099, 136, 30), (257, 8, 8), (1000, 64, 64), (33, 48, 5)):
        assert engine.dense_selftest(n, ncq, p) < 1e-11


def test_fused_orthogonalisation_chain_selftest(engine):
    # gram_ext -> project_out -> gram_ext -> project_chol_apply -> chol_append, as the eigensolver chains
    # them: orthonormal basis from random blocks, max |Q^T Q - I| at fp32 storage precision
    for n, ncols, block in ((5000, 64, 4), (4099, 96, 8), (3001, 64, 16), (70001, 168, 4)):
        assert engine.ortho_selftest(n, ncols, block) < 5e-6


def test_generator_is_bit_identical_to_host(engine):
    spec = synth.make_spec(700, 30000, 900, n_clusters=16, seed=6)
    engine.generate(spec)
    dev = engine.export_csr()
    host = synth.generate_csr(spec)
    assert np.array_equal(dev.indptr, host.indptr)
    assert np.array_equal(dev.indices, host.indices)
    # a row shard regenerates exactly its rows
    engine.generate(spec, row0=200, n_local=300)
    part = engine.export_csr()
    ref = host[200:500]
    assert np.array_equal(part.indptr, ref.indptr) and np.array_equal(part.indices, ref.indices)


@pytest.mark.parametrize("name", ["tile_600x4000", "counts_300x1000", "masked_400x3000"])
def test_golden_vectors(engine, name):
    X, z = load_golden(name)
    feats = z["features"] if "features" in z else None
    fw = z["feature_weights"] if "feature_weights" in z else None
    evals, evecs, idf, deg = tl.spectral_embedding(engine, X, feats, int(z["k"]), 0, fw, return_parts=True)
    np.testing.assert_allclose(idf, z["idf"], rtol=TOL_VEC)
    np.testing.assert_allclose(deg, z["degree"], rtol=TOL_VEC)
    _check_against(z["evals"], z["evecs"], evals, evecs)
    assert evals[0] == pytest.approx(1.0, abs=1e-6)         # trivial pair kept
    assert np.all(np.diff(evals) <= 0)                      # argsort()[::-1]


@pytest.mark.parametrize("name", ["tile_600x4000", "counts_300x1000", "masked_400x3000"])
def test_reference_executed_vectors(engine, name):
    """CUDA path against vectors produced by running the reference's own Python code
    (SpectralMatrixFree.fit/_eigen, tools/_embedding.py:434-481; tests/golden/make_ref_golden.py)."""
    from conftest import GOLDEN
    X, z = load_golden(name)
    r = np.load(GOLDEN / f"{name}_ref.npz")
    feats = z["features"] if "features" in z else None
    fw = z["feature_weights"] if "feature_weights" in z else None
    evals, evecs, idf, deg = tl.spectral_embedding(engine, X, feats, int(z["k"]), 0, fw, return_parts=True)
    np.testing.assert_allclose(idf, r["weights"], rtol=TOL_VEC)
    np.testing.assert_allclose(deg, r["degree"], rtol=TOL_VEC)
    _check_against(r["evals"], r["evecs"], evals, evecs)
    # wrapper post-processing against the reference's spectral() run unmodified
    ad = MiniAnnData(X)
    ev_w, emb_w = tl.spectral(ad, n_comps=int(z["k"]), features=feats, feature_weights=fw, inplace=False, engine=engine)
    np.testing.assert_allclose(ev_w, r["wrapped_evals"], rtol=TOL_EVAL)
    cos = eigvec_agreement(r["wrapped_evals"], r["wrapped_evecs"], emb_w)
    assert cos.min() >= MIN_COS, cos
    scale = np.linalg.norm(emb_w, axis=0) / np.linalg.norm(r["wrapped_evecs"], axis=0)
    np.testing.assert_allclose(scale, 1.0, rtol=1e-4)


def test_golden_dense_50x100_exhausts_krylov_space(engine):
    # the reference's own test input shape (tests/test_tools.py:95-110): n=50, k=30
    X, z = load_golden("dense_50x100")
    evals, evecs, idf, deg = tl.spectral_embedding(engine, X, None, int(z["k"]), 0, return_parts=True)
    np.testing.assert_allclose(idf, z["idf"], rtol=TOL_VEC)
    np.testing.assert_allclose(deg, z["degree"], rtol=TOL_VEC)
    np.testing.assert_allclose(evals, z["evals"], rtol=TOL_EVAL, atol=1e-7)
    # eigenvalues here sit in one tight cluster: compare the invariant subspace
    qa, _ = np.linalg.qr(evecs)
    qb, _ = np.linalg.qr(z["evecs"])
    assert np.linalg.svd(qa.T @ qb, compute_uv=False).min() > 0.99


def test_operator_matches_oracle(engine):
    spec = synth.make_spec(900, 20000, 500, n_clusters=14, seed=12)
    X = synth.generate_csr(spec, dtype=np.float64)
    engine.load_csr(X)
    engine.set_feature_weights(None)
    idf, deg = engine.prepare()
    mat = sp.csr_matrix(X)
    w = oracle.idf(mat)
    xt, dinv, _, degree = oracle.operator_pieces(oracle.normalize(mat, w))
    np.testing.assert_allclose(idf, w, rtol=TOL_VEC)
    np.testing.assert_allclose(deg, degree, rtol=TOL_VEC)
    rng = np.random.default_rng(0)
    for b in (4, 8, 16):
        V = rng.standard_normal((900, b)).astype(np.float32)
        Y = engine.operator_apply(V)
        want = xt @ (xt.T @ V.astype(np.float64)) - dinv[:, None] * V
        err = np.abs(Y - want).max() / np.abs(want).max()
        assert err < 2e-5, (b, err)
    # linearity (size-independent property)
    V1 = rng.standard_normal((900, 8)).astype(np.float32)
    V2 = rng.standard_normal((900, 8)).astype(np.float32)
    lhs = engine.operator_apply(V1 + V2)
    rhs = engine.operator_apply(V1) + engine.operator_apply(V2)
    assert np.abs(lhs - rhs).max() / np.abs(rhs).max() < 1e-5


def test_config1_parity_against_oracle(engine):
    # BASELINE.json configs[0]: 5k x 100k, ~3k nnz/cell, n_comps=30
    spec = synth.make_spec(5000, 100000, 3000, n_clusters=48, seed=0)
    engine.generate(spec)
    X = engine.export_csr()
    assert np.array_equal(X.indices, synth.generate_csr(spec).indices)
    ev_o, evec_o, w_o, deg_o = oracle.spectral_embedding(X, None, 30, 0, return_parts=True)
    engine.set_feature_weights(None)
    idf, deg = engine.prepare()
    np.testing.assert_allclose(idf, w_o, rtol=TOL_VEC)
    np.testing.assert_allclose(deg, deg_o, rtol=TOL_VEC)
    for block in (8, 4, 16):
        evals, evecs = engine.eigsh(30, seed=0, block=block)
        _check_against(ev_o, evec_o, evals, evecs)
        st = engine.stats()
        assert st["n_ops"] < 60 and st["max_residual"] < 1e-6


def test_config2_parity_against_oracle(engine):
    """BASELINE.json configs[1]: 100k cells x 500k bins, ~5k nnz/cell, n_comps=30 (4.9e8 stored entries; the
    production kernels: bucketed transpose, tiled SpMM).  The oracle side ran on a CPU box for six minutes
    (tests/golden/make_c2_golden.py, 78 ARPACK mat-vecs); the fixture keeps the eigenvalues, every 5th degree,
    every 25th IDF weight and the eigenvector rows of 4000 sampled cells."""
    from conftest import GOLDEN
    z = np.load(GOLDEN / "c2_100kx500k.npz")
    spec = synth.make_spec(100_000, 500_000, 5_000, n_clusters=48, seed=0)
    engine.generate(spec)
    assert engine.shape()[2] == int(z["nnz"])
    engine.set_feature_weights(None)
    idf, deg = engine.prepare()
    np.testing.assert_allclose(idf[::25], z["idf_25"], rtol=TOL_VEC)
    np.testing.assert_allclose(deg[::5], z["degree_5"], rtol=TOL_VEC)
    evals, evecs = engine.eigsh(30, seed=0)
    assert engine.stats()["spmm_tiled"] == 1
    np.testing.assert_allclose(evals, z["evals"], rtol=TOL_EVAL)
    cos = eigvec_agreement(z["evals"], z["evecs_rows"], evecs[z["rows"]])
    assert cos.min() >= MIN_COS, cos


def test_thick_restart_small_basis(engine):
    X, z = load_golden("tile_600x4000")
    engine.load_csr(X)
    engine.set_feature_weights(None)
    engine.prepare(want_outputs=False)
    evals, evecs = engine.eigsh(int(z["k"]), seed=1, max_basis=48)     # forces restarts
    assert engine.stats()["n_restarts"] >= 1
    _check_against(z["evals"], z["evecs"], evals, evecs)


def test_bitwise_reproducible(engine):
    # the reference's only pinned property for this path (tests/test_tools.py:104-108)
    X, z = load_golden("tile_600x4000")
    ad = MiniAnnData(X)
    runs = [tl.spectral(ad, n_comps=8, features=None, random_state=0, inplace=False, engine=engine)[1]
            for _ in range(3)]
    for r in runs[1:]:
        np.testing.assert_array_equal(r, runs[0])


def test_wrapper_side_effects_and_weighting(engine):
    X, z = load_golden("counts_300x1000")
    ad = MiniAnnData(X)
    assert tl.spectral(ad, n_comps=5, features=None, engine=engine) is None
    ev = ad.uns["spectral_eigenvalue"]
    emb = ad.obsm["X_spectral"]
    assert ev.dtype == np.float64 and emb.dtype == np.float64 and emb.shape == (300, len(ev))
    keep = z["evals"] > 0
    np.testing.assert_allclose(ev, z["evals"][keep], rtol=TOL_EVAL)
    want = z["evecs"][:, keep] * np.sqrt(z["evals"][keep])
    cos = np.abs(np.sum(emb * want, axis=0)) / (np.linalg.norm(emb, axis=0) * np.linalg.norm(want, axis=0))
    assert cos.min() > MIN_COS
    # integer index array as `features` (preprocessing/_basic.py:1069) == boolean mask
    idx = np.flatnonzero(np.arange(1000) % 3 != 0)
    a = tl.spectral(ad, n_comps=4, features=idx, weighted_by_sd=False, inplace=False, engine=engine)
    mask = np.zeros(1000, bool)
    mask[idx] = True
    b = tl.spectral(ad, n_comps=4, features=mask, weighted_by_sd=False, inplace=False, engine=engine)
    np.testing.assert_array_equal(a[0], b[0])
    ev_o, _ = oracle.spectral_embedding(X, mask, 4, 0)
    np.testing.assert_allclose(a[0], ev_o, rtol=TOL_EVAL)


def test_degenerate_rows_are_reported(engine):
    X = sp.csr_matrix(np.array([[1, 1, 0, 0], [0, 0, 0, 0], [0, 1, 1, 0], [1, 0, 1, 1.0]]))
    with pytest.raises(RuntimeError, match="empty row or a non-positive degree"):
        tl.spectral_embedding(engine, X, None, 2, 0)
    # on_degenerate="nan": the reference's outcome (NaN spectrum; weighted_by_sd keeps no component of it)
    ad = MiniAnnData(X)
    ev, emb = tl.spectral(ad, n_comps=2, features=None, weighted_by_sd=False, inplace=False, engine=engine, on_degenerate="nan")
    assert ev.shape == (2,) and emb.shape == (4, 2) and np.isnan(ev).all() and np.isnan(emb).all()
    ev, emb = tl.spectral(ad, n_comps=2, features=None, inplace=False, engine=engine, on_degenerate="nan")
    assert ev.shape == (0,) and emb.shape == (4, 0)


def test_int64_indices_and_count_dtypes(engine):
    X, z = load_golden("counts_300x1000")
    engine.load_arrays(X.indptr.astype(np.int64), X.indices.astype(np.int64), X.data.astype(np.uint32), 300, 1000)
    engine.set_feature_weights(None)
    idf, deg = engine.prepare()
    np.testing.assert_allclose(idf, z["idf"], rtol=TOL_VEC)
    np.testing.assert_allclose(deg, z["degree"], rtol=TOL_VEC)


@pytest.mark.parametrize("dtype", [np.int32, np.int64])
def test_delta_encoded_index_transfer_equals_plain_transfer(engine, monkeypatch, dtype):
    """csrc/ingest.cu ships the column indices as 16-bit differences and rebuilds them with a kernel; the
    plain int32 transfer (SNAPB200_NO_DELTA) must leave the same matrix on the device: identical IDF,
    degrees and spectrum, on matrices whose gaps sit on either side of the 16-bit limit."""
    def spread(n, m, vocab, per_row, seed):
        # rows draw `per_row` columns from a vocabulary of `vocab` columns scattered over [0, m): wide gaps between
        # consecutive stored columns, yet every column is shared by many rows (no degenerate cells)
        rng = np.random.default_rng(seed)
        cols = np.sort(rng.choice(m, size=vocab, replace=False))
        pick = np.sort(np.argpartition(rng.random((n, vocab)), per_row, axis=1)[:, :per_row], axis=1)
        return sp.csr_matrix((np.ones(n * per_row, np.float32), cols[pick].ravel(), np.arange(n + 1) * per_row), shape=(n, m))

    cases = [spread(3000, 3_000_000, 5000, 40, 8),                                         # gaps ~75k: most entries are markers
             synth.generate_csr(synth.make_spec(6000, 200_000, 300, n_clusters=4, seed=9)),  # gaps ~700: markers at row / tile starts
             spread(2500, 1_500_000, 3000, 60, 10)]                                        # gaps ~25k: both kinds mixed
    for ci, X in enumerate(cases):
        out = {}
        for mode in ("delta", "plain"):
            if mode == "plain":
                monkeypatch.setenv("SNAPB200_NO_DELTA", "1")
            else:
                monkeypatch.delenv("SNAPB200_NO_DELTA", raising=False)
            engine.load_arrays(X.indptr.astype(np.int64), X.indices.astype(dtype), X.data.astype(np.float32), *X.shape)
            shipped = engine.stats()["bytes_h2d_indices"]
            if mode == "plain":
                assert shipped == 4 * X.nnz
            else:       # 2 bytes per entry + 4 per marker (row starts, tile starts, gaps >= 0xFFFF) + chunk header
                assert 2 * X.nnz < shipped != 4 * X.nnz
                if ci == 1:
                    assert shipped < 2.1 * X.nnz
            engine.set_feature_weights(None)
            idf, deg = engine.prepare()
            out[mode] = [idf.copy(), deg.copy()]
            if ci == 1:         # (the scattered matrices have a flat spectrum: nothing to learn from solving them)
                out[mode] += [a.copy() for a in engine.eigsh(6, seed=1)]
        for a, b in zip(out["delta"], out["plain"]):
            np.testing.assert_array_equal(a, b)
    monkeypatch.delenv("SNAPB200_NO_DELTA", raising=False)
    # a negative / too large index is still refused
    X = cases[1]
    bad = X.indices.astype(np.int64)
    bad[1234] = -3
    with pytest.raises(RuntimeError):
        engine.load_arrays(X.indptr.astype(np.int64), bad, X.data.astype(np.float32), *X.shape)


def _two_views(n, seed):
    s1 = synth.make_spec(n, 6000, 250, n_clusters=10, seed=seed)
    s2 = synth.make_spec(n, 900, 60, n_clusters=10, seed=seed)          # same planted labels (keyed by seed,row)
    atac = synth.generate_csr(s1, dtype=np.float64)
    rna = synth.generate_csr(s2, dtype=np.float64)
    rng = np.random.default_rng(seed)
    rna.data = 1.0 + rng.poisson(0.5, size=rna.nnz)                      # integer counts (values path)
    return atac, rna


def test_multi_spectral_matches_oracle(engine):
    # embedding.rs:388-452; n <= 2000 so the Frobenius normaliser uses every row (no RNG involved)
    atac, rna = _two_views(1500, 4)
    ev_o, evec_o = oracle.multi_spectral_embedding([atac, rna], [None, None], [1.0, 1.0], 8, 0)
    evals, evecs = tl.multi_spectral_embedding(engine, [atac, rna], [None, None], [1.0, 1.0], 8, 0)
    _check_against(ev_o, evec_o, evals, evecs)
    # the intermediate quantities: view normalisers (both readings of the frobenius_norm snippet) and degrees
    for container in ("csr_matrix", "csr_array"):
        ev_o, evec_o, norms_o, deg_o = oracle.multi_spectral_embedding([atac, rna], [None, None], [1.0, 3.0], 8, 0,
                                                                       return_parts=True, container=container)
        evals, evecs, idf, deg, norms = tl.multi_spectral_embedding(engine, [atac, rna], [None, None], [1.0, 3.0], 8, 0,
                                                                    return_parts=True, container=container)
        np.testing.assert_allclose(norms, norms_o, rtol=TOL_VEC)
        np.testing.assert_allclose(deg, deg_o, rtol=TOL_VEC)
        np.testing.assert_allclose(idf, np.concatenate([oracle.idf(atac), oracle.idf(rna)]), rtol=TOL_VEC)
        _check_against(ev_o, evec_o, evals, evecs)
    # wrapper: weights, weighted_by_sd, no write into the AnnData objects
    a1, a2 = MiniAnnData(atac), MiniAnnData(rna)
    ev_w, emb_w = tl.multi_spectral([a1, a2], n_comps=6, features=None, weights=[2.0, 1.0], engine=engine)
    ev_r, emb_r = oracle.multi_spectral([a1, a2], n_comps=6, features=None, weights=[2.0, 1.0])
    np.testing.assert_allclose(ev_w, ev_r, rtol=TOL_EVAL)
    cos = np.abs(np.sum(emb_w * emb_r, axis=0)) / (np.linalg.norm(emb_w, axis=0) * np.linalg.norm(emb_r, axis=0))
    assert cos.min() > MIN_COS and not a1.obsm and not a1.uns


def test_multi_spectral_sampled_normaliser(engine):
    # n > 2000: the 2000-row sample is passed explicitly to both sides (the reference's Rust RNG
    # cannot be reproduced outside Rust -- SURVEY.md H8)
    atac, rna = _two_views(2600, 9)
    rows = np.sort(np.random.default_rng(1).choice(2600, 2000, replace=False))
    ev_o, evec_o = oracle.multi_spectral_embedding([atac, rna], [None, None], [1.0, 1.0], 6, 0, sample_rows=rows)
    evals, evecs = tl.multi_spectral_embedding(engine, [atac, rna], [None, None], [1.0, 1.0], 6, 0, sample_rows=rows)
    _check_against(ev_o, evec_o, evals, evecs)


def test_second_context_lifetime(engine):
    """Two contexts in one process share the caching allocator: destroying one (and its stream)
    must leave the other usable, and blocks it parked reusable (regression: a parked block kept
    the destroyed stream and the next release recorded an event on it)."""
    from snapatac2_b200 import Engine
    spec = synth.make_spec(3000, 20000, 300, n_clusters=24, seed=11)
    engine.generate(spec)
    engine.set_feature_weights(None)
    idf0, deg0 = engine.prepare()
    other = Engine(engine.device)
    other.generate(spec)
    idf1, deg1 = other.prepare()
    ev1, _ = other.eigsh(8, seed=0)
    other.close()
    np.testing.assert_array_equal(idf0, idf1)
    np.testing.assert_array_equal(deg0, deg1)
    ev0, _ = engine.eigsh(8, seed=0)      # reuses blocks the closed context parked
    np.testing.assert_array_equal(ev0, ev1)
    third = Engine(engine.device)
    third.generate(spec)
    third.close()


@pytest.mark.parametrize("valued", [False, True])
def test_nystrom_matches_oracle_given_landmarks(engine, valued):
    """spectral_embedding_nystrom (embedding.rs:61-129, 194-267) with the landmark draw fixed:
    IDF over all cells, landmark degrees, eigenvalues and the extended vectors against the oracle;
    then the public sample_size path end to end (its own landmark draw) for shape and sanity."""
    spec = synth.make_spec(6000, 30000, 400, n_clusters=16, seed=5)
    engine.generate(spec)
    X = engine.export_csr().astype(np.float64)
    feats = None
    if valued:
        rng = np.random.default_rng(8)
        X.data = rng.integers(1, 4, size=X.nnz).astype(np.float64)
        feats = rng.random(X.shape[1]) < 0.8
    lm = np.random.default_rng(1).choice(6000, 1500, replace=False)
    k = 12
    ev_o, q_o, w_o, d_o = oracle.spectral_embedding_nystrom(X, feats, k, lm, 2000, return_parts=True)
    ev, q, w, d, lm2 = tl.spectral_embedding_nystrom(engine, X, feats, k, 1500, False, 2000, landmarks=lm,
                                                     return_parts=True)
    np.testing.assert_array_equal(lm2, lm)
    np.testing.assert_allclose(w, w_o, rtol=1e-5)
    np.testing.assert_allclose(d, d_o, rtol=1e-5)
    np.testing.assert_allclose(ev, ev_o, rtol=1e-4)
    assert q.shape == q_o.shape
    assert eigvec_agreement(ev_o, q_o, q).min() >= 0.999
    # row scaling (the per-chunk degree normalisation) agrees too, not only directions
    np.testing.assert_allclose(np.linalg.norm(q, axis=0), np.linalg.norm(q_o, axis=0), rtol=1e-3)
    # chunk-streamed variant (one chunk_size block of cells on the device at a time): same numbers
    ev_s, q_s = tl.spectral_embedding_nystrom(engine, X, feats, k, 1500, False, 2000, landmarks=lm, stream=True)
    np.testing.assert_allclose(ev_s, ev, rtol=1e-9)
    assert np.abs(q_s - q).max() <= 1e-4 * np.abs(q).max()

    ad = MiniAnnData(sp.csr_matrix(X))
    evals, emb = tl.spectral(ad, n_comps=k, features=feats, sample_size=1500, chunk_size=2000, inplace=False,
                             engine=engine)
    assert emb.shape[0] == 6000 and emb.shape[1] == evals.shape[0] <= k
    assert np.all(np.isfinite(emb)) and np.all(np.diff(evals) <= 1e-12)


def test_unsorted_or_duplicate_rows_are_rejected(engine):
    """The C ABI wants strictly increasing column indices per row; the raw entry point checks it
    (the scipy-level loader canonicalises instead)."""
    indptr = np.array([0, 3, 5], dtype=np.int64)
    good = np.array([1, 4, 7, 0, 2], dtype=np.int32)
    engine.load_arrays(indptr, good, None, 2, 8)
    for bad in (np.array([4, 1, 7, 0, 2], dtype=np.int32), np.array([1, 4, 4, 0, 2], dtype=np.int32)):
        with pytest.raises(RuntimeError, match="strictly increasing"):
            engine.load_arrays(indptr, bad, None, 2, 8)
    with pytest.raises(RuntimeError, match="out of range"):
        engine.load_arrays(indptr, np.array([1, 4, 8, 0, 2], dtype=np.int32), None, 2, 8)
    # scipy-level loader: duplicates are summed, order restored
    X = sp.csr_matrix((np.ones(6), np.array([4, 1, 1, 7, 2, 0]), np.array([0, 4, 6])), shape=(2, 8))
    engine.load_csr(X)
    assert engine.shape() == (2, 8, 5)


def test_nystrom_streamed_degrees(engine):
    """compute_degrees (embedding.rs:328-360) one block of cells at a time (two streaming passes through the
    fp32 SpMM kernels) against the fp64 degrees of prepare() on the whole matrix."""
    spec = synth.make_spec(5000, 20000, 300, n_clusters=10, seed=13)
    X = synth.generate_csr(spec, dtype=np.float32)
    engine.load_csr(X)
    engine.set_feature_weights(None)
    idf, deg = engine.prepare()
    got = tl._streamed_degrees(engine.device, lambda i: sp.csr_matrix(X[i:i + 1200]), 5000, 1200, idf)
    np.testing.assert_allclose(got, deg, rtol=2e-4)
    # degree-weighted landmarks run end to end in both modes
    ad = MiniAnnData(sp.csr_matrix(X))
    for stream in (False, True):
        v, q = tl.spectral_embedding_nystrom(engine, ad.X, None, 6, 1200, True, 1200, stream=stream)
        assert q.shape == (5000, 6) and np.all(np.isfinite(q)) and v[0] == pytest.approx(1.0, abs=1e-3)


class _Backed:
    """Stand-in for a backed AnnData element: rows come in chunks, never as one matrix."""
    def __init__(self, X):
        self._X, self.shape = X, X.shape

    def chunked(self, chunk_size):
        for i in range(0, self.shape[0], chunk_size):
            yield self._X[i:i + chunk_size], i, min(i + chunk_size, self.shape[0])


class _Sliceable:
    def __init__(self, X):
        self._X, self.shape = X, X.shape

    def __getitem__(self, key):
        return self._X[key]


@pytest.mark.parametrize("name", ["tile_600x4000", "counts_300x1000"])
def test_blockwise_ingest_equals_in_memory(engine, name):
    """adata.X as a backed element (chunked()), a row-sliceable lazy array and a one-shot generator of CSR
    blocks: assembled on the device block by block, bitwise the same matrix and the same embedding."""
    X, z = load_golden(name)
    X = sp.csr_matrix(X)
    engine.load_csr(X)
    ref = engine.export_csr()
    want = tl.spectral(MiniAnnData(X), n_comps=int(z["k"]), features=None, inplace=False, engine=engine)
    sources = {
        "chunked": lambda: _Backed(X),
        "sliceable": lambda: _Sliceable(X),
        "generator": lambda: (X[i:i + 97] for i in range(0, X.shape[0], 97)),
        "int64+mixed": lambda: (sp.csr_matrix((b.data.astype(np.float64), b.indices.astype(np.int64), b.indptr.astype(np.int64)),
                                              shape=b.shape) for b in (X[:50], X[50:51], X[51:51], X[51:])),
    }
    for tag, make in sources.items():
        ad = MiniAnnData(X[:1])                       # placeholder matrix: X is replaced by the block source
        ad.X = make()
        ad.__class__ = type("BackedLike", (MiniAnnData,), {"n_obs": property(lambda self: X.shape[0]),
                                                           "n_vars": property(lambda self: X.shape[1]),
                                                           "shape": property(lambda self: X.shape)})
        got = tl.spectral(ad, n_comps=int(z["k"]), features=None, inplace=False, engine=engine, chunk_size=128)
        dev = engine.export_csr()
        assert np.array_equal(dev.indptr, ref.indptr) and np.array_equal(dev.indices, ref.indices), tag
        np.testing.assert_array_equal(dev.data, ref.data)
        np.testing.assert_array_equal(got[0], want[0])
        np.testing.assert_array_equal(got[1], want[1])


def test_deferred_value_scan(engine):
    """tl.spectral loads the pattern first and scans X.data in the background: a matrix whose values are all 1
    (any dtype) stays pattern-only; one with real counts is recomputed with its values -- same numbers as the
    synchronous load either way, also with a feature mask."""
    X, z = load_golden("counts_300x1000")
    k = int(z["k"])
    ev, evec, idf, deg = tl.spectral_embedding(engine, X, None, k, 0, return_parts=True)
    np.testing.assert_allclose(deg, z["degree"], rtol=TOL_VEC)                  # the counts were used
    _check_against(z["evals"], z["evecs"], ev, evec)
    mask = np.arange(1000) % 4 != 0
    ev_m, _ = tl.spectral_embedding(engine, X, mask, k, 0)
    ev_o, _ = oracle.spectral_embedding(X, mask, k, 0)
    np.testing.assert_allclose(ev_m, ev_o, rtol=TOL_EVAL)
    # binarised copy: float64 ones -> pattern-only kernels, no value array on the device
    B = X.copy()
    B.data[:] = 1.0
    ev_b, evec_b, _, deg_b = tl.spectral_embedding(engine, B, None, k, 0, return_parts=True)
    assert engine.values_all_ones()
    ev_ob, evec_ob, _, deg_ob = oracle.spectral_embedding(B, None, k, 0, return_parts=True)
    np.testing.assert_allclose(deg_b, deg_ob, rtol=TOL_VEC)
    _check_against(ev_ob, evec_ob, ev_b, evec_b)
    # the low-level switch: verdict and late shipment of the values
    engine.load_csr(X, defer_value_scan=True)
    assert not engine.values_all_ones()
    engine.load_values()
    engine.set_feature_weights(None)
    _, deg2 = engine.prepare()
    np.testing.assert_allclose(deg2, z["degree"], rtol=TOL_VEC)


def test_row_gather_between_contexts(engine):
    from snapatac2_b200 import Engine
    spec = synth.make_spec(900, 5000, 120, n_clusters=6, seed=3)
    X = synth.generate_csr(spec, dtype=np.float32)
    X.data = (1 + (X.indices % 4)).astype(np.float32)
    engine.load_csr(X)
    rows = np.array([5, 899, 17, 400, 17 + 1, 0])
    other = Engine(engine.device)
    try:
        engine.gather_rows_into(rows, other)
        got = other.export_csr()
    finally:
        other.close()
    want = X[rows]
    assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
    np.testing.assert_array_equal(got.data, want.data)


def test_nystrom_golden_fixture(engine):
    X, z = load_golden("nystrom_500x3000")
    k, chunk, lm = int(z["k"]), int(z["chunk_size"]), z["landmarks"]
    ev, q, w, d, _ = tl.spectral_embedding_nystrom(engine, X, None, k, lm.size, False, chunk, landmarks=lm,
                                                   return_parts=True)
    np.testing.assert_allclose(w, z["idf"], rtol=TOL_VEC)
    np.testing.assert_allclose(d, z["degree"], rtol=TOL_VEC)
    np.testing.assert_allclose(ev, z["evals"], rtol=TOL_EVAL)
    assert eigvec_agreement(z["evals"], z["q"], q).min() >= MIN_COS

"""GPU (-m gpu): the shared-memory tiled (sliced-ELL) SpMM path against the
oracle operator and against the CSR-gather path, several column tiles in both
passes, binarised and count-valued inputs, block widths 4 and 8."""

import numpy as np
import pytest
import scipy.sparse as sp

import oracle
from conftest import eigvec_agreement
from snapatac2_b200 import synth

pytestmark = pytest.mark.gpu


def _oracle_operator(X):
    mat = sp.csr_matrix(X, dtype=np.float64)
    w = oracle.idf(mat)
    xt, dinv, _, deg = oracle.operator_pieces(oracle.normalize(mat, w))
    return xt, dinv, w, deg


@pytest.mark.parametrize("valued,block", [(False, 8), (True, 8), (False, 4), (True, 4)])
def test_tiled_operator_matches_oracle_and_csr_path(engine, valued, block):
    # 27000 cells -> 3 (b=4) / 5 (b=8) cell tiles in pass 1; 30000 bins -> 3 / 5 feature tiles in pass 2
    spec = synth.make_spec(27000, 30000, 200, n_clusters=20, seed=17)
    engine.generate(spec)
    X = engine.export_csr().astype(np.float64)
    if valued:
        rng = np.random.default_rng(5)
        X.data = rng.integers(1, 6, size=X.nnz).astype(np.float64)
    xt, dinv, w, deg = _oracle_operator(X)
    rng = np.random.default_rng(1)
    V = rng.standard_normal((27000, block)).astype(np.float32)
    want = xt @ (xt.T @ V.astype(np.float64)) - dinv[:, None] * V
    got = {}
    engine.set_block(block)
    try:
        for mode in ("csr", "tiled"):
            engine.set_spmm_mode(mode)
            engine.load_csr(X, binarized=not valued)
            engine.set_feature_weights(None)
            idf, degree = engine.prepare()
            np.testing.assert_allclose(idf, w, rtol=1e-5)
            np.testing.assert_allclose(degree, deg, rtol=1e-5)
            Y = engine.operator_apply(V)
            assert engine.stats()["spmm_tiled"] == (0 if mode == "csr" else 1)
            err = np.abs(Y - want).max() / np.abs(want).max()
            assert err < 2e-5, (mode, err)
            got[mode] = Y
            # bitwise repeatable (no atomics anywhere on the path)
            np.testing.assert_array_equal(engine.operator_apply(V), Y)
        assert np.abs(got["csr"] - got["tiled"]).max() / np.abs(want).max() < 1e-5
    finally:
        engine.set_spmm_mode("auto")
        engine.set_block(4)


@pytest.mark.parametrize("valued,block", [(False, 4), (True, 4), (False, 8)])
def test_tiled_long_and_single_class_segments(engine, valued, block):
    """Format-build edge cases: segments longer than one 256-entry piece (dense rows, features
    present in most cells), segments whose columns all fall into one shared-memory bank class
    (every rotation slot but one class stays empty: the overflow list and its re-walk), empty
    segments, and a ragged last window."""
    rng = np.random.default_rng(3)
    n, m = 9001, 26003
    rows, cols = [], []
    for i in range(n):
        kind = i % 7
        if kind == 0:      # long: ~1500 entries in the first feature tile
            c = rng.choice(6000, size=1500, replace=False)
        elif kind == 1:    # one bank class only: multiples of 8 (and of 4)
            c = 8 * rng.choice(m // 8, size=300, replace=False)
        elif kind == 2:    # very short
            c = rng.choice(m, size=3, replace=False)
        else:
            c = rng.choice(m, size=120, replace=False)
        rows.append(np.full(c.size, i))
        cols.append(c)
    # features seen by (almost) every cell, and one seen by every 8th cell only
    for j in (5, 17, 12290):
        keep = rng.random(n) < 0.9
        rows.append(np.nonzero(keep)[0]); cols.append(np.full(int(keep.sum()), j))
    rows.append(np.arange(0, n, 8)); cols.append(np.full(len(range(0, n, 8)), 26000))
    r, cidx = np.concatenate(rows), np.concatenate(cols)
    X = sp.csr_matrix((np.ones(r.size), (r, cidx)), shape=(n, m))
    X.sum_duplicates()
    X.data[:] = 1.0
    if valued:
        X.data = rng.integers(1, 5, size=X.nnz).astype(np.float64)
    xt, dinv, w, deg = _oracle_operator(X)
    V = rng.standard_normal((n, block)).astype(np.float32)
    want = xt @ (xt.T @ V.astype(np.float64)) - dinv[:, None] * V
    engine.set_block(block)
    engine.set_spmm_mode("tiled")
    try:
        engine.load_csr(X, binarized=not valued)
        engine.set_feature_weights(None)
        idf, degree = engine.prepare()
        np.testing.assert_allclose(idf, w, rtol=1e-5)
        np.testing.assert_allclose(degree, deg, rtol=1e-5)
        Y = engine.operator_apply(V)
        assert engine.stats()["spmm_tiled"] == 1
        assert np.abs(Y - want).max() / np.abs(want).max() < 2e-5
        np.testing.assert_array_equal(engine.operator_apply(V), Y)
    finally:
        engine.set_spmm_mode("auto")
        engine.set_block(4)


@pytest.mark.parametrize("block", [8, 4])
def test_tiled_eigsh_parity_config1(engine, block):
    spec = synth.make_spec(5000, 100000, 3000, n_clusters=48, seed=0)
    engine.set_spmm_mode("tiled")
    engine.set_block(block)
    try:
        engine.generate(spec)
        X = engine.export_csr()
        ev_o, evec_o, w_o, deg_o = oracle.spectral_embedding(X, None, 30, 0, return_parts=True)
        engine.set_feature_weights(None)
        idf, deg = engine.prepare()
        np.testing.assert_allclose(idf, w_o, rtol=1e-5)
        np.testing.assert_allclose(deg, deg_o, rtol=1e-5)
        evals, evecs = engine.eigsh(30, seed=0)
        st = engine.stats()
        assert st["spmm_tiled"] == 1 and st["block"] == block
        np.testing.assert_allclose(evals, ev_o, rtol=1e-4)
        assert eigvec_agreement(ev_o, evec_o, evecs).min() >= 0.999
    finally:
        engine.set_spmm_mode("auto")
        engine.set_block(4)


@pytest.mark.parametrize("valued,block", [(False, 4), (True, 8)])
def test_projection_products_tiled_and_csr(engine, valued, block):
    """snapb200_project (the Nystrom extension's two products) through the tiled copies and through
    the CSR kernels against scipy: Xhat @ M and Xhat.T @ U for k not a multiple of the block width."""
    spec = synth.make_spec(27000, 30000, 200, n_clusters=20, seed=23)
    engine.generate(spec)
    X = engine.export_csr().astype(np.float64)
    rng = np.random.default_rng(4)
    if valued:
        X.data = rng.integers(1, 6, size=X.nnz).astype(np.float64)
    w = oracle.idf(sp.csr_matrix(X))
    xhat = oracle.normalize(sp.csr_matrix(X), w)
    k = 13
    M = rng.standard_normal((30000, k)).astype(np.float32)
    U = rng.standard_normal((27000, k)).astype(np.float32)
    want = xhat @ M.astype(np.float64)
    want_t = xhat.T @ U.astype(np.float64)
    engine.set_block(block)
    try:
        for mode in ("csr", "tiled"):
            engine.set_spmm_mode(mode)
            engine.load_csr(X, binarized=not valued)
            engine.set_feature_weights(None)
            w_dev, rho = engine.prepare_projection()          # no transpose needed for Xhat @ M
            np.testing.assert_allclose(w_dev, w, rtol=1e-5)
            got = engine.project(M)
            assert np.abs(got - want).max() / np.abs(want).max() < 2e-5, mode
            engine.prepare()
            got_t = engine.project_t(U)
            assert np.abs(got_t - want_t).max() / np.abs(want_t).max() < 2e-5, mode
            np.testing.assert_array_equal(engine.project(M), got)   # after the full prepare as well
    finally:
        engine.set_spmm_mode("auto")
        engine.set_block(4)


def test_bucketed_transpose_equals_legacy_bitmap_transpose(engine, monkeypatch):
    """The bucketed transpose (row-run tables + per-bucket counting sort) and the first version
    (shared-memory bitmap per 1024 x 1024 block) must produce the same feature-major copy: segments
    in ascending cell order, hence bitwise identical operator results -- binarised and valued,
    ragged rows, a shard that is not a multiple of the tile height."""
    rng = np.random.default_rng(5)
    for n, m, nnz_row, valued in ((13000, 70000, 400, False), (2500, 3000, 200, True), (300, 1500, 40, False)):
        spec = synth.make_spec(n, m, nnz_row, n_clusters=12, seed=n)
        X = synth.generate_csr(spec, dtype=np.float32)
        if valued:
            X.data = rng.integers(1, 6, size=X.nnz).astype(np.float32)
        V = rng.standard_normal((n, 4)).astype(np.float32)
        outs = []
        for mode in ("bitmap", "bucketed"):
            monkeypatch.setenv("SNAPB200_TRANSPOSE", mode)
            engine.set_spmm_mode("tiled")
            engine.load_csr(X)
            engine.set_feature_weights(None)
            idf, deg = engine.prepare()
            outs.append((idf, deg, engine.operator_apply(V)))
        engine.set_spmm_mode("auto")
        np.testing.assert_array_equal(outs[0][0], outs[1][0])
        np.testing.assert_array_equal(outs[0][1], outs[1][1])
        np.testing.assert_array_equal(outs[0][2], outs[1][2])


def test_bucketed_transpose_dense_rows_and_hot_features(engine, monkeypatch):
    """Rows longer than the scatter pass's shared-memory staging (the bucketed transpose must hand over to the
    bitmap one) and a feature present in every cell (one bucket far above the average): same operator as the
    CSR-gather path."""
    rng = np.random.default_rng(3)
    n, m = 700, 120000
    X = sp.random(n, m, density=0.002, format="csr", random_state=7, dtype=np.float32)
    X.data[:] = 1.0
    dense_row = sp.csr_matrix(np.ones((1, m), dtype=np.float32))       # 120000 entries > staging capacity
    hot = sp.csr_matrix((np.ones(n + 1, np.float32), (np.arange(n + 1), np.full(n + 1, 77))), shape=(n + 1, m))
    Y = sp.csr_matrix(sp.vstack([X, dense_row]).maximum(hot))
    Y.sum_duplicates()
    Y.data[:] = 1.0
    V = rng.standard_normal((n + 1, 4)).astype(np.float32)
    res = {}
    for mode, tr in (("csr", "bitmap"), ("tiled", "bucketed"), ("tiled", "bitmap")):
        monkeypatch.setenv("SNAPB200_TRANSPOSE", tr)
        engine.set_spmm_mode(mode)
        engine.load_csr(Y)
        engine.set_feature_weights(None)
        engine.prepare(want_outputs=False)
        res[(mode, tr)] = engine.operator_apply(V)
    engine.set_spmm_mode("auto")
    np.testing.assert_array_equal(res[("tiled", "bucketed")], res[("tiled", "bitmap")])
    ref = res[("csr", "bitmap")]
    assert np.abs(res[("tiled", "bucketed")] - ref).max() <= 2e-5 * np.abs(ref).max()

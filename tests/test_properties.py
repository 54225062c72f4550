"""CPU: property tests (hypothesis) of the oracle and the host-side logic -- the
size-independent facts the GPU parity tests lean on."""

import numpy as np
import scipy.sparse as sp
from hypothesis import given, settings, strategies as st

import oracle
from snapatac2_b200 import dist, synth, tl


def _random_csr(rng, n, m, density, counts):
    X = sp.random(n, m, density=density, format="csr", random_state=rng, dtype=np.float64)
    X.data = np.floor(1 + 3 * X.data) if counts else np.ones_like(X.data)
    # no empty rows (the reference produces NaN there)
    for i in np.flatnonzero(np.diff(X.indptr) == 0):
        X = X.tolil(); X[i, rng.integers(0, m)] = 1.0; X = X.tocsr()
    X.sort_indices()
    return sp.csr_matrix(X)


@settings(max_examples=25, deadline=None)
@given(st.integers(0, 2**31 - 1), st.integers(8, 40), st.integers(12, 60), st.booleans())
def test_operator_pieces_are_a_normalised_similarity(seed, n, m, counts):
    """rows of normalize() have unit norm; degrees equal row sums of S - I; the trivial pair
    (lambda = 1, sqrt(d)) is an eigenpair of D^-1/2 (S - I) D^-1/2 (embedding.rs:139-171)."""
    rng = np.random.default_rng(seed)
    X = _random_csr(rng, n, m, 0.3, counts)
    w = oracle.idf(X)
    xhat = oracle.normalize(X, w)
    norms = np.sqrt(np.asarray(xhat.multiply(xhat).sum(axis=1)).ravel())
    ok = norms > 0          # a row whose only features have zero IDF weight normalises to NaN upstream too
    if not ok.all():
        return
    np.testing.assert_allclose(norms, 1.0, rtol=1e-12)
    xt, dinv, col_sum, degree = oracle.operator_pieces(xhat)
    s = (xhat @ xhat.T).toarray()
    np.fill_diagonal(s, 0.0)
    np.testing.assert_allclose(degree, s.sum(axis=1), rtol=1e-9, atol=1e-12)
    if np.all(degree > 0):
        a = s / np.sqrt(np.outer(degree, degree))
        u1 = np.sqrt(degree)
        np.testing.assert_allclose(a @ u1, u1, rtol=1e-9, atol=1e-12)
        # the factored operator the kernels apply equals the dense one
        v = rng.standard_normal(n)
        np.testing.assert_allclose(xt @ (xt.T @ v) - dinv * v, a @ v, rtol=1e-9, atol=1e-12)


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 5000), st.integers(1, 9))
def test_row_splits_cover_all_rows(n, parts):
    b = dist.equal_row_splits(n, parts)
    assert b[0] == 0 and b[-1] == n and np.all(np.diff(b) >= 0) and len(b) == parts + 1
    indptr = np.concatenate([[0], np.cumsum(np.random.default_rng(n).integers(0, 50, size=n))])
    bb = dist.balanced_row_splits(indptr, parts)
    assert bb[0] == 0 and bb[-1] == n and np.all(np.diff(bb) >= 0) and len(bb) == parts + 1


@settings(max_examples=30, deadline=None)
@given(st.integers(0, 10**6), st.integers(3, 30), st.data())
def test_feature_mask_round_trip(seed, n_vars, data):
    rng = np.random.default_rng(seed)
    k = data.draw(st.integers(1, n_vars))
    idx = rng.choice(n_vars, size=k, replace=False)
    fw = rng.uniform(0.1, 2.0, size=k)
    mask, fw_sorted = tl._feature_mask(idx, n_vars, fw)
    assert mask.sum() == k and set(np.flatnonzero(mask)) == set(idx.tolist())
    # weights follow their features into ascending column order
    np.testing.assert_array_equal(fw_sorted, fw[np.argsort(idx, kind="stable")])


@settings(max_examples=10, deadline=None)
@given(st.integers(0, 1000), st.integers(2, 5))
def test_generator_rows_do_not_depend_on_the_shard(seed, parts):
    spec = synth.make_spec(90, 700, 30, n_clusters=4, seed=seed)
    full = synth.generate_csr(spec)
    b = dist.equal_row_splits(spec.n, parts)
    blocks = [synth.generate_csr(spec, int(b[i]), int(b[i + 1])) for i in range(parts)]
    assert (sp.vstack(blocks, format="csr") != full).nnz == 0

"""CPU: the oracle against its pins (dense eigh, the reference's second
statement of the algorithm, the committed golden vectors) and the reference's
documented edge cases."""

import numpy as np
import pytest
import scipy.sparse as sp

import oracle
from conftest import load_golden, eigvec_agreement
from snapatac2_b200 import synth, MiniAnnData

GOLDEN_CASES = ["tile_600x4000", "counts_300x1000", "dense_50x100", "masked_400x3000"]


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_reproduces_golden(name):
    X, z = load_golden(name)
    feats = z["features"] if "features" in z else None
    fw = z["feature_weights"] if "feature_weights" in z else None
    ev, evec, w, deg = oracle.spectral_embedding(X, feats, int(z["k"]), 0, fw, return_parts=True)
    np.testing.assert_allclose(w, z["idf"], rtol=1e-12)
    np.testing.assert_allclose(deg, z["degree"], rtol=1e-10)
    np.testing.assert_allclose(ev, z["evals"], rtol=1e-8, atol=1e-12)
    assert eigvec_agreement(z["evals"], z["evecs"], evec).min() > 1 - 1e-6


def test_oracle_matches_dense_and_twin():
    spec = synth.make_spec(350, 2500, 120, n_clusters=7, seed=21)
    X = synth.generate_csr(spec, dtype=np.float64)
    k = 6
    ev, evec, w, deg = oracle.spectral_embedding(X, None, k, 0, return_parts=True)
    ev_d, evec_d, w_d, deg_d = oracle.dense_check(X, k)
    np.testing.assert_allclose(w, w_d, rtol=1e-13)
    np.testing.assert_allclose(deg, deg_d, rtol=1e-11)
    np.testing.assert_allclose(ev, ev_d, rtol=1e-10)
    assert np.abs(np.sum(evec * evec_d, axis=0)).min() > 1 - 1e-9
    # the reference's own SpectralMatrixFree.fit/_eigen on the IDF-weighted matrix
    ev_t, evec_t, deg_t = oracle.matrix_free_twin(X, k, feature_weights=w)
    np.testing.assert_allclose(deg, deg_t, rtol=1e-11)
    np.testing.assert_allclose(ev, ev_t, rtol=1e-10)
    assert ev[0] == pytest.approx(1.0, abs=1e-12)          # trivial pair kept as component 0


def test_idf_edge_cases():
    # df == 0 -> ln(n); df == n -> ln(n/(n-1))   (embedding.rs:277-282)
    X = sp.csr_matrix(np.array([[1, 0, 1, 0], [1, 0, 0, 0], [1, 0, 1, 1.0]]))
    w = oracle.idf(X)
    np.testing.assert_allclose(w, [np.log(3 / 2), np.log(3), np.log(3 / 2), np.log(3)])
    # all columns same document frequency -> all ones   (:273-274)
    X = sp.csr_matrix(np.array([[1, 0], [0, 1.0]]))
    np.testing.assert_array_equal(oracle.idf(X), [1.0, 1.0])
    # explicit zeros are counted as stored entries       (:271)
    X = sp.csr_matrix((np.array([0.0, 1.0, 1.0]), np.array([0, 1, 0]), np.array([0, 2, 3])), shape=(2, 3))
    np.testing.assert_allclose(oracle.idf(X), [np.log(2 / 1), np.log(2 / 1), np.log(2)])


def test_empty_row_gives_nan_like_reference():
    X = sp.csr_matrix(np.array([[1, 1, 0], [0, 0, 0], [0, 1, 1.0]]))
    xhat = oracle.normalize(X, np.ones(3))
    assert np.isnan(xhat.toarray()[1]).all() or xhat[1].nnz == 0
    _, dinv, _, deg = oracle.operator_pieces(sp.csr_matrix(np.array([[1, 0], [0, 1.0]])))
    assert (deg <= 0).all()        # isolated cells: degree 0 -> dinv inf (embedding.rs:146)


def test_wrapper_semantics():
    spec = synth.make_spec(200, 1500, 80, n_clusters=5, seed=2)
    ad = MiniAnnData(synth.generate_csr(spec, dtype=np.float64))
    with pytest.raises(NameError):
        oracle.spectral(ad)                        # no var['selected']  (_embedding.py:229)
    out = oracle.spectral(ad, n_comps=5, features=None, inplace=False)
    assert out[0].shape == (5,) and out[1].shape == (200, 5)
    assert oracle.spectral(ad, n_comps=5, features=None) is None
    assert ad.obsm["X_spectral"].shape == (200, 5) and ad.uns["spectral_eigenvalue"].shape == (5,)
    # weighted_by_sd scales by sqrt(eval) (:286-289)
    ev, evec = oracle.spectral(ad, n_comps=5, features=None, weighted_by_sd=False, inplace=False)
    np.testing.assert_allclose(np.abs(ad.obsm["X_spectral"]), np.abs(evec * np.sqrt(ev)), rtol=1e-6, atol=1e-9)
    # n_comps clamp uses min(n_vars-1, n_obs-1)   (:231)
    small = MiniAnnData(sp.csr_matrix(np.random.default_rng(0).uniform(1, 2, size=(12, 40))))
    ev, evec = oracle.spectral(small, n_comps=30, features=None, weighted_by_sd=False, inplace=False)
    assert ev.shape == (11,)


def test_multi_view_restatement_runs():
    s1 = synth.make_spec(150, 900, 60, n_clusters=4, seed=1)
    s2 = synth.make_spec(150, 300, 30, n_clusters=4, seed=1)
    a, b = synth.generate_csr(s1, dtype=np.float64), synth.generate_csr(s2, dtype=np.float64)
    ev, evec = oracle.multi_spectral_embedding([a, b], [None, None], [1.0, 1.0], 4, 0)
    assert ev.shape == (4,) and evec.shape == (150, 4) and ev[0] == pytest.approx(1.0, abs=1e-10)


def test_nystrom_golden_and_its_structure():
    """Oracle restatement of the Nystrom path against the committed fixture (itself cross-checked
    with a dense eigh statement by make_golden.py), plus two structural properties: landmark cells
    are extended exactly like any other cell, and the result depends on chunk_size only through the
    per-chunk degree normalisation."""
    X, z = load_golden("nystrom_500x3000")
    k, chunk, lm = int(z["k"]), int(z["chunk_size"]), z["landmarks"]
    ev, q, w, d = oracle.spectral_embedding_nystrom(X, None, k, lm, chunk, return_parts=True)
    np.testing.assert_allclose(w, z["idf"], rtol=1e-12)
    np.testing.assert_allclose(d, z["degree"], rtol=1e-10)
    np.testing.assert_allclose(ev, z["evals"], rtol=1e-9)
    assert eigvec_agreement(z["evals"], z["q"], q).min() > 1 - 1e-8
    # one chunk = all rows: directions per row group change only by the chunk-wise scaling
    ev1, q1 = oracle.spectral_embedding_nystrom(X, None, k, lm, X.shape[0])
    ratio = q1[:chunk] / q[:chunk]
    np.testing.assert_allclose(ratio, np.repeat(ratio[:, :1], k, axis=1), rtol=1e-8)
    # orthogonalize: orthonormal columns, eigenvalues sorted descending
    evo, qo = oracle.orthogonalize(ev, q)
    np.testing.assert_allclose(np.real(qo).T @ np.real(qo), np.eye(k), atol=1e-8)
    assert np.all(np.diff(np.real(evo)) <= 1e-12)


# ---------------------------------------------------------------------------------------------
# Pins against vectors produced by EXECUTING the reference's own Python code
# (tests/golden/make_ref_golden.py -> *_ref.npz; see its docstring for what ran)
# ---------------------------------------------------------------------------------------------
def load_ref(name):
    from conftest import GOLDEN
    return np.load(GOLDEN / f"{name}_ref.npz")


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_equals_reference_executed_twin(name):
    """oracle.spectral_embedding == SpectralMatrixFree.fit/_eigen of the reference (run unmodified)."""
    X, z = load_golden(name)
    r = load_ref(name)
    feats = z["features"] if "features" in z else None
    fw = z["feature_weights"] if "feature_weights" in z else None
    ev, evec, w, deg = oracle.spectral_embedding(X, feats, int(z["k"]), 0, fw, return_parts=True)
    np.testing.assert_allclose(w, r["weights"], rtol=1e-12)          # the weights the reference run was fed
    np.testing.assert_allclose(deg, r["degree"], rtol=1e-10)
    np.testing.assert_allclose(ev, r["evals"], rtol=1e-10, atol=1e-13)
    assert eigvec_agreement(r["evals"], r["evecs"], evec).min() > 1 - 1e-8
    # and the committed oracle fixture agrees with the reference run as well
    np.testing.assert_allclose(z["evals"], r["evals"], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(z["degree"], r["degree"], rtol=1e-10)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_idf_closed_form_of_fixture_weights(name):
    """The one piece only Rust computes (embedding.rs:269-286): ln(n/df) with the 0 -> 1 and
    n -> n-1 clamps and the all-equal -> ones branch, recomputed independently of oracle.idf."""
    X, z = load_golden(name)
    if "feature_weights" in z:
        pytest.skip("fixture uses user weights")
    n, m = X.shape
    df = np.zeros(m)
    for j in X.indices:
        df[j] += 1
    if np.all(df == df[0]):
        want = np.ones(m)
    else:
        want = np.array([np.log(n / (1.0 if d == 0 else (n - 1.0 if d == n else d))) for d in df])
    np.testing.assert_allclose(z["idf"], want, rtol=1e-13)


def test_wrapper_postprocessing_equals_reference_executed_wrapper():
    """oracle.spectral's post-processing == the reference's spectral() run unmodified (its native
    call answered from the recorded eigenpairs)."""
    for name in GOLDEN_CASES:
        X, z = load_golden(name)
        r = load_ref(name)
        ev, evec = r["evals"], r["evecs"]
        keep = [i for i in range(ev.shape[0]) if ev[i] > 0]
        np.testing.assert_allclose(r["wrapped_evals"], ev[keep], rtol=0)
        np.testing.assert_allclose(r["wrapped_evecs"], evec[:, keep] * np.sqrt(ev[keep]), rtol=1e-15)


def test_orthogonalize_equals_reference_executed():
    _, z = load_golden("nystrom_500x3000")
    r = load_ref("nystrom_500x3000")
    ev, q = oracle.orthogonalize(np.array(z["evals"]), np.array(z["q"]))
    np.testing.assert_allclose(np.real(ev), np.real(r["evals"]), rtol=1e-10)
    cos = np.abs(np.sum(np.real(q) * np.real(r["evecs"]), axis=0))
    assert cos.min() > 1 - 1e-9
    from snapatac2_b200 import tl
    ev2, q2 = tl.orthogonalize(np.array(z["evals"]), np.array(z["q"]))
    np.testing.assert_allclose(np.real(ev2), np.real(r["evals"]), rtol=1e-10)
    assert np.abs(np.sum(np.real(q2) * np.real(r["evecs"]), axis=0)).min() > 1 - 1e-9


def test_frobenius_snippet_both_containers():
    """embedding.rs:456-460 executed on csr_matrix (np.power -> matrix power) and csr_array
    (element-wise): the oracle restates the snippet verbatim and reproduces both; closed forms."""
    from conftest import GOLDEN
    r = np.load(GOLDEN / "frobenius_snippet_ref.npz")
    assert "np.power(X @ X.T, 2).sum()" in str(r["snippet"])
    for tag in ("a", "b"):
        xhat = sp.csr_matrix((r[f"{tag}_data"], r[f"{tag}_indices"], r[f"{tag}_indptr"]), shape=tuple(r[f"{tag}_shape"]))
        n = xhat.shape[0]
        got_m = oracle.reference_restatement._frobenius_norm(xhat, "csr_matrix")
        got_a = oracle.reference_restatement._frobenius_norm(xhat, "csr_array")
        np.testing.assert_allclose(got_m, np.sqrt(float(r[f"{tag}_sum_csr_matrix"]) - n), rtol=1e-12)
        np.testing.assert_allclose(got_a, np.sqrt(float(r[f"{tag}_sum_csr_array"]) - n), rtol=1e-12)
        S = (xhat @ xhat.T).toarray()
        np.testing.assert_allclose(float(r[f"{tag}_sum_csr_matrix"]), np.sum(S.sum(axis=1) ** 2), rtol=1e-12)
        np.testing.assert_allclose(float(r[f"{tag}_sum_csr_array"]), np.sum(S * S), rtol=1e-12)
        assert abs(got_m - got_a) > 1.0      # the two readings really differ

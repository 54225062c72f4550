"""pp.knn -- the consumer next to the path: oracle pins (CPU) and CUDA parity (-m gpu).

Bar: neighbour indices identical, distances BIT-identical (the CUDA path sums the squared differences in the
reference's order without fused multiply-add, see csrc/knn.cu)."""

import numpy as np
import pytest
import scipy.sparse as sp

import oracle
from conftest import GOLDEN
from snapatac2_b200 import MiniAnnData, pp


def _golden():
    return np.load(GOLDEN / "knn_400x12.npz")


def _csr(z, tag, n=400):
    return sp.csr_matrix((z[f"{tag}_data"], z[f"{tag}_indices"], z[f"{tag}_indptr"]), shape=(n, n))


def _same_graph(a, b):
    a, b = a.tocsr(), b.tocsr()
    assert a.shape == b.shape
    np.testing.assert_array_equal(a.indptr, b.indptr)
    np.testing.assert_array_equal(a.indices, b.indices)
    np.testing.assert_array_equal(a.data, b.data)          # bit for bit


def _blobs(n, d, seed, spread=(0.05, 1.0), n_centres=8):
    rng = np.random.default_rng(seed)
    centres = rng.normal(scale=3.0, size=(n_centres, d))
    lab = rng.integers(0, n_centres, size=n)
    return centres[lab] + rng.normal(size=(n, d)) * rng.uniform(*spread, size=n_centres)[lab][:, None]


# ------------------------------------------------------------------------------------------------ CPU

def test_oracle_graph_of_the_fixture_equals_the_reference_executed_wrapper_output():
    z = _golden()
    P = z["points"]
    _same_graph(oracle.knn.nearest_neighbour_graph(P, 10), _csr(z, "k10"))
    _same_graph(oracle.knn.nearest_neighbour_graph(P[:, :5], 10), _csr(z, "k10_dims5"))
    _same_graph(oracle.knn.nearest_neighbour_graph(P[:, [0, 3, 7]], 10), _csr(z, "k10_dimslist"))
    _same_graph(oracle.knn.nearest_neighbour_graph(P[:, ::-1], 7), _csr(z, "k7_other"))
    _same_graph(_csr(z, "k10"), _csr(z, "k10_ndarray"))
    np.testing.assert_array_equal(z["k450_row_lengths"], 399)


def test_oracle_search_agrees_with_an_independent_kd_tree():
    """scipy's cKDTree is an exact kd-tree like the reference's crate: same neighbours, distances equal up to
    the summation order (cKDTree's own) -- and bit-equal once re-evaluated with the crate's fold."""
    from scipy.spatial import cKDTree
    for n, d, k, seed in [(700, 30, 25, 1), (300, 3, 50, 2), (64, 1, 5, 3)]:
        P = _blobs(n, d, seed)
        A = oracle.knn.nearest_neighbour_graph(P, k)
        dist, idx = cKDTree(P).query(P, k=k + 1)
        for i in range(n):
            mine = A.indices[A.indptr[i]:A.indptr[i + 1]]
            theirs = np.setdiff1d(idx[i], [i])[:k] if i in idx[i] else idx[i][:k]
            assert set(mine) == set(theirs)
        np.testing.assert_allclose(np.sort(A.data.reshape(n, k), axis=1), dist[:, 1:], rtol=1e-13)
        _same_graph(A, oracle.knn.nearest_neighbour_graph_kdtree(P, k))


def test_oracle_distance_is_the_crates_left_to_right_fold():
    rng = np.random.default_rng(0)
    a, b = rng.normal(size=30), rng.normal(size=30)
    acc = 0.0
    for x, y in zip(a, b):
        acc = acc + (x - y) * (x - y)
    assert oracle.knn.squared_euclidean(a, b[None, :])[0] == acc


def test_duplicates_and_ties_keep_the_smaller_index():
    P = np.zeros((6, 2))
    P[3:] = [[1, 0], [0, 1], [-1, 0]]            # points 0,1,2 coincide; 3,4,5 at distance 1 from them
    A = oracle.knn.nearest_neighbour_graph(P, 3)
    np.testing.assert_array_equal(A.indices[A.indptr[0]:A.indptr[1]], [1, 2, 3])     # 3 wins the tie against 4, 5
    np.testing.assert_array_equal(A.data[A.indptr[0]:A.indptr[1]], [0, 0, 1])


def test_wrapper_argument_handling_needs_no_gpu():
    z = _golden()
    with pytest.raises(ValueError, match=str(z["empty_error"])):
        pp.knn(np.zeros((0, 4)))
    with pytest.raises(ValueError, match="method must be one of"):
        pp.knn(z["points"], method="nope")
    assert str(z["method_error"]).startswith("method must be one of")
    with pytest.raises(KeyError):
        pp.knn(MiniAnnData(np.ones((4, 3))))      # no X_spectral yet


class _OracleBackedEngine:
    """Stands where snapatac2_b200.Engine would: answers ``knn`` from the oracle, so the host-side mirror
    (argument handling, CSR assembly, dtype per method, where the result goes) is checked without a GPU."""

    def __init__(self):
        self.calls = []

    def knn(self, points, n_neighbors, q0=0, nq=None):
        n = points.shape[0]
        nq = n - q0 if nq is None else nq
        self.calls.append((points.shape, n_neighbors, q0, nq))
        g = oracle.knn.nearest_neighbour_graph(points, n_neighbors, rows=np.arange(q0, q0 + nq))
        K = max(0, min(n_neighbors, n - 1))
        return g.indices.reshape(nq, K).astype(np.int32), g.data.reshape(nq, K)


def test_host_mirror_equals_reference_wrapper_vectors_with_a_stub_engine():
    z = _golden()
    P = z["points"]
    eng = _OracleBackedEngine()
    ad = MiniAnnData(np.ones((400, 3)))
    ad.obsm["X_spectral"] = P
    ad.obsm["other"] = P[:, ::-1].copy()
    _same_graph(pp.knn(ad, n_neighbors=10, inplace=False, engine=eng), _csr(z, "k10"))
    _same_graph(pp.knn(ad, n_neighbors=10, use_dims=5, inplace=False, engine=eng), _csr(z, "k10_dims5"))
    _same_graph(pp.knn(ad, n_neighbors=10, use_dims=[0, 3, 7], inplace=False, engine=eng), _csr(z, "k10_dimslist"))
    _same_graph(pp.knn(ad, n_neighbors=7, use_rep="other", inplace=False, engine=eng), _csr(z, "k7_other"))
    assert pp.knn(ad, n_neighbors=25, engine=eng) is None                     # inplace by default: stored, nothing returned
    _same_graph(ad.obsp["distances"], _csr(z, "k25_inplace"))
    _same_graph(pp.knn(P, n_neighbors=10, inplace=True, engine=eng), _csr(z, "k10_ndarray"))   # ndarray: returned
    big = pp.knn(ad, n_neighbors=450, inplace=False, engine=eng)              # more neighbours than other points
    np.testing.assert_array_equal(np.diff(big.indptr), z["k450_row_lengths"])
    np.testing.assert_array_equal(big.data[: big.indptr[1]], z["k450_row0_data"])
    assert pp.knn(P, n_neighbors=10, method="hora", engine=eng).dtype == np.float32
    assert eng.calls[0] == ((400, 12), 10, 0, 400) and eng.calls[1][0] == (400, 5) and eng.calls[2][0] == (400, 3)
    one = pp.knn(P[:1], n_neighbors=5, engine=eng)                             # a single observation: an empty graph
    assert one.shape == (1, 1) and one.nnz == 0


# ------------------------------------------------------------------------------------------------ GPU

@pytest.mark.gpu
def test_cuda_graph_equals_reference_wrapper_vectors(engine):
    z = _golden()
    P = z["points"]
    ad = MiniAnnData(np.ones((400, 3)))
    ad.obsm["X_spectral"] = P
    ad.obsm["other"] = P[:, ::-1].copy()
    _same_graph(pp.knn(ad, n_neighbors=10, inplace=False, engine=engine), _csr(z, "k10"))
    _same_graph(pp.knn(ad, n_neighbors=10, use_dims=5, inplace=False, engine=engine), _csr(z, "k10_dims5"))
    _same_graph(pp.knn(ad, n_neighbors=10, use_dims=[0, 3, 7], inplace=False, engine=engine), _csr(z, "k10_dimslist"))
    _same_graph(pp.knn(ad, n_neighbors=7, use_rep="other", inplace=False, engine=engine), _csr(z, "k7_other"))
    assert pp.knn(ad, n_neighbors=25, engine=engine) is None
    _same_graph(ad.obsp["distances"], _csr(z, "k25_inplace"))
    _same_graph(pp.knn(P, n_neighbors=10, inplace=True, engine=engine), _csr(z, "k10_ndarray"))
    # hora: float32 distances of the same graph; pynndescent: the exact graph
    h = pp.knn(P, n_neighbors=10, method="hora", engine=engine)
    assert h.dtype == np.float32
    np.testing.assert_array_equal(h.indices, z["k10_indices"])
    _same_graph(pp.knn(P, n_neighbors=10, method="pynndescent", engine=engine), _csr(z, "k10"))


@pytest.mark.gpu
@pytest.mark.parametrize("n,d,k,seed", [
    (5000, 30, 50, 1),        # the default use: 30 components, 50 neighbours
    (3001, 30, 100, 2),       # the largest list, ragged last tile
    (1000, 64, 74, 3),        # widest points
    (777, 33, 20, 4),         # first dimension count of the 64-wide kernel
    (4000, 16, 15, 5), (2500, 8, 10, 6), (1500, 3, 30, 7), (900, 1, 12, 8),
    (129, 5, 128, 9),         # k clamps to n - 1 = 128 > 100: refused (see below)
    (60, 4, 80, 10),          # k > n - 1: every other point, 59 per row
    (2, 2, 5, 11), (1, 3, 5, 12),
])
def test_cuda_graph_equals_oracle(engine, n, d, k, seed):
    P = _blobs(n, d, seed)
    if n == 129:
        with pytest.raises(RuntimeError, match="at most 100"):
            engine.knn(P, k)
        return
    want = oracle.knn.nearest_neighbour_graph(P, k)
    got = pp.nearest_neighbour_graph(P, k, engine=engine)
    _same_graph(got, want)
    assert got.has_sorted_indices


@pytest.mark.gpu
def test_cuda_graph_on_hard_inputs(engine):
    rng = np.random.default_rng(5)
    # far from the origin and tightly packed (the float32 filter sees almost nothing: everything is decided in float64)
    P = 1e6 + rng.normal(size=(2000, 30)) * 1e-3
    _same_graph(pp.nearest_neighbour_graph(P, 30, engine=engine), oracle.knn.nearest_neighbour_graph(P, 30))
    # two tight clusters a long way apart: centring cannot help, every same-cluster pair passes the float32 filter
    P = np.where(rng.random((2000, 1)) < 0.5, 1e6, -1e6) + rng.normal(size=(2000, 30)) * 1e-3
    _same_graph(pp.nearest_neighbour_graph(P, 30, engine=engine), oracle.knn.nearest_neighbour_graph(P, 30))
    # exact duplicates, many of them, and a regular grid (ties everywhere at the k-th distance)
    P = np.repeat(rng.normal(size=(40, 6)), 25, axis=0)
    _same_graph(pp.nearest_neighbour_graph(P, 30, engine=engine), oracle.knn.nearest_neighbour_graph(P, 30))
    g = np.stack(np.meshgrid(np.arange(20.0), np.arange(20.0), np.arange(5.0), indexing="ij"), axis=-1).reshape(-1, 3)
    _same_graph(pp.nearest_neighbour_graph(g, 7, engine=engine), oracle.knn.nearest_neighbour_graph(g, 7))
    # wildly different scales per dimension, points sorted by distance from a corner (adversarial arrival order)
    P = rng.normal(size=(3000, 10)) * np.logspace(-4, 4, 10)
    P = P[np.argsort(-np.linalg.norm(P - P.min(axis=0), axis=1))]
    _same_graph(pp.nearest_neighbour_graph(P, 25, engine=engine), oracle.knn.nearest_neighbour_graph(P, 25))


@pytest.mark.gpu
def test_cuda_refuses_non_finite_points(engine):
    """The reference's kd-tree refuses NaN / inf coordinates (kdtree: NonFiniteCoordinate, unwrapped at knn.rs:20)."""
    P = _blobs(500, 8, 3)
    for bad in (np.nan, np.inf, -np.inf):
        Q = P.copy()
        Q[123, 4] = bad
        with pytest.raises(RuntimeError, match="non-finite"):
            engine.knn(Q, 10)
    idx, _ = engine.knn(P, 10)          # the context is usable afterwards
    assert idx.shape == (500, 10)


@pytest.mark.gpu
def test_cuda_query_ranges_and_reproducibility(engine):
    """A rank of a row-sharded run searches its own rows against all points; three runs are bit-identical
    (the reference's test_reproducibility, tests/test_tools.py:111-115)."""
    P = _blobs(6000, 30, 21)
    full = oracle.knn.nearest_neighbour_graph_kdtree(P, 50)
    idx, dst = engine.knn(P, 50)
    np.testing.assert_array_equal(idx.ravel(), full.indices)
    np.testing.assert_array_equal(dst.ravel(), full.data)
    for q0, nq in [(0, 1), (127, 130), (4000, 2000), (5999, 1), (300, 0)]:
        i2, d2 = engine.knn(P, 50, q0=q0, nq=nq)
        np.testing.assert_array_equal(i2, idx[q0:q0 + nq])
        np.testing.assert_array_equal(d2, dst[q0:q0 + nq])
    for _ in range(2):
        i3, d3 = engine.knn(P, 50)
        np.testing.assert_array_equal(i3, idx)
        np.testing.assert_array_equal(d3, dst)


@pytest.mark.gpu
def test_knn_of_a_spectral_embedding(engine):
    """End of the pipeline as a user runs it: tl.spectral then pp.knn on what it stored."""
    from snapatac2_b200 import synth, tl
    X = synth.generate_csr(synth.make_spec(4000, 20000, 300, n_clusters=6, seed=2))
    ad = MiniAnnData(X)
    tl.spectral(ad, n_comps=15, features=None, engine=engine)
    pp.knn(ad, n_neighbors=20, engine=engine)
    _same_graph(ad.obsp["distances"], oracle.knn.nearest_neighbour_graph(ad.obsm["X_spectral"], 20))

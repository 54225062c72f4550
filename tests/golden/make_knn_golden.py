"""Golden vectors for ``pp.knn``: the reference's OWN wrapper executed on a fixture.

Run from the repo root, in the container that has ``/root/reference``:

    python tests/golden/make_knn_golden.py

``snapatac2-python/python/snapatac2/preprocessing/_knn.py`` is loaded unmodified with ``importlib``;
its two imports of the compiled extension are stub modules whose ``nearest_neighbour_graph`` is answered
by the oracle's exhaustive search (the Rust kd-tree cannot be built here: no cargo).  What the file pins
is therefore everything the wrapper does around the native call (:53-87): the choice of ``obsm[use_rep]``,
``use_dims`` as an int and as a list, ndarray input forcing ``inplace=False``, the empty-matrix error,
where the result is stored -- plus the oracle's graph of the fixture itself, which
``tests/test_oracle.py`` checks against ``scipy.spatial.cKDTree``.

Output: ``tests/golden/knn_400x12.npz`` (points + graphs; CSR triplets per case).
"""

import importlib.util
import sys
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

REF = Path("/root/reference/snapatac2-python/python/snapatac2/preprocessing/_knn.py")
OUT = Path(__file__).resolve().parent


def load_reference_knn_module(oracle):
    pkg = types.ModuleType("snapatac2")
    pkg.__path__ = []
    utils = types.ModuleType("snapatac2._utils")
    utils.is_anndata = lambda x: hasattr(x, "obsm")
    internal = types.ModuleType("snapatac2._snapatac2")
    internal.AnnData = object
    internal.AnnDataSet = object
    internal.nearest_neighbour_graph = lambda data, k: oracle.knn.nearest_neighbour_graph(data, k)
    sys.modules["snapatac2"] = pkg
    sys.modules["snapatac2._utils"] = utils
    sys.modules["snapatac2._snapatac2"] = internal
    pkg._utils, pkg._snapatac2 = utils, internal
    spec = importlib.util.spec_from_file_location("snapatac2_reference_knn", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def fixture_points():
    """400 points in 12 dimensions: five blobs of different spread, a handful of exact duplicates."""
    rng = np.random.default_rng(11)
    centres = rng.normal(scale=3.0, size=(5, 12))
    lab = rng.integers(0, 5, size=400)
    P = centres[lab] + rng.normal(size=(400, 12)) * rng.uniform(0.05, 1.0, size=5)[lab][:, None]
    P[37] = P[5]
    P[111] = P[5]
    P[399] = P[398]
    return P


def main():
    import oracle
    from snapatac2_b200 import MiniAnnData
    mod = load_reference_knn_module(oracle)
    P = fixture_points()
    out = {"points": P}

    def put(tag, adj):
        adj = adj.tocsr()
        out[f"{tag}_indptr"] = adj.indptr.astype(np.int64)
        out[f"{tag}_indices"] = adj.indices.astype(np.int64)
        out[f"{tag}_data"] = adj.data.astype(np.float64)

    ad = MiniAnnData(np.ones((400, 3)))
    ad.obsm["X_spectral"] = P
    ad.obsm["other"] = P[:, ::-1].copy()
    put("k10", mod.knn(ad, n_neighbors=10, inplace=False))
    put("k10_dims5", mod.knn(ad, n_neighbors=10, use_dims=5, inplace=False))
    put("k10_dimslist", mod.knn(ad, n_neighbors=10, use_dims=[0, 3, 7], inplace=False))
    put("k7_other", mod.knn(ad, n_neighbors=7, use_rep="other", inplace=False))
    big = mod.knn(ad, n_neighbors=450, inplace=False).tocsr()          # more neighbours than other points: n - 1 per row
    out["k450_row_lengths"] = np.diff(big.indptr).astype(np.int64)
    out["k450_row0_data"] = big.data[: big.indptr[1]].astype(np.float64)
    assert mod.knn(ad, n_neighbors=25) is None                          # default inplace=True stores and returns nothing
    put("k25_inplace", ad.obsp["distances"])
    put("k10_ndarray", mod.knn(P, n_neighbors=10, inplace=True))        # ndarray: inplace is overridden, the graph is returned
    try:
        mod.knn(np.zeros((0, 4)))
        raise AssertionError("expected ValueError")
    except ValueError as e:
        out["empty_error"] = np.array(str(e))
    try:
        mod.knn(P, method="nope")
        raise AssertionError("expected ValueError")
    except ValueError as e:
        out["method_error"] = np.array(str(e))
    np.savez_compressed(OUT / "knn_400x12.npz", **out)
    print("knn_400x12.npz:", {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items() if k.endswith("_indptr")})


if __name__ == "__main__":
    main()

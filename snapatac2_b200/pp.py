"""Host-side mirror of ``snapatac2.preprocessing._knn.knn`` (preprocessing/_knn.py:10-87).

Same signature, same argument handling and errors; the search itself is
``snapb200_knn`` (csrc/knn.cu), the B200 replacement of
``internal.nearest_neighbour_graph`` (snapatac2-core/src/utils/knn.rs:9-33): an exact
Euclidean k-nearest-neighbour graph, the point itself excluded, as a CSR matrix of
distances with sorted rows.

``method``: the reference offers an exact kd-tree and two approximate searches
('hora' = HNSW, 'pynndescent').  All three names are accepted and all three get the
exact graph -- on the GPU the exact search is the fast one; 'hora' returns float32
distances as the reference's ``CsrMatrix<f32>`` does (knn.rs:34-56).

Under ``torch.distributed`` every rank passes its own rows (the layout ``tl.spectral``
leaves in ``obsm``); the points are all-gathered on the host (n x d doubles: small) and
each rank searches its own rows: the result holds the rank's rows, columns are global.
"""

from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from . import dist as _dist
from . import tl as _tl

_METHODS = ("kdtree", "hora", "pynndescent")


def _gather_points(data: np.ndarray):
    """(all points, first global row of this rank's block)."""
    rank, ws = _dist.world()
    if ws == 1:
        return data, 0
    blocks = _dist.allgather_objects(np.ascontiguousarray(data))
    q0 = sum(b.shape[0] for b in blocks[:rank])
    return np.concatenate(blocks, axis=0), q0


def nearest_neighbour_graph(data: np.ndarray, k: int, engine=None) -> sp.csr_matrix:
    """``internal.nearest_neighbour_graph(data, k)`` (snapatac2-python/src/knn.rs:8-16)."""
    eng = engine if engine is not None else _tl.default_engine()
    data = np.ascontiguousarray(data, dtype=np.float64)
    n_local = data.shape[0]
    points, q0 = _gather_points(data)
    n = points.shape[0]
    idx, dst = eng.knn(points, k, q0=q0, nq=n_local)
    K = idx.shape[1]
    indptr = np.arange(0, (n_local + 1) * K, max(K, 1), dtype=np.int64)[: n_local + 1] if K > 0 else np.zeros(n_local + 1, np.int64)
    return sp.csr_matrix((dst.ravel(), idx.ravel(), indptr), shape=(n_local, n))


def knn(
    adata,
    n_neighbors: int = 50,
    use_dims: int | list[int] | None = None,
    use_rep: str = "X_spectral",
    method: str = "kdtree",
    inplace: bool = True,
    random_state: int = 0,
    *,
    engine=None,
):
    """Compute a neighborhood graph of observations (Euclidean).

    Mirrors ``snap.pp.knn`` (preprocessing/_knn.py:10-87): ``adata`` is an AnnData-like object
    (the matrix is ``adata.obsm[use_rep]``) or a numpy array (then ``inplace`` is ignored and the
    graph is returned); ``use_dims`` keeps the first ``use_dims`` columns (int) or the listed
    ones; the result is the ``n x n`` CSR matrix of distances to the ``n_neighbors`` nearest other
    observations, stored in ``adata.obsp['distances']`` when ``inplace``.  ``random_state`` is
    accepted for signature compatibility (the search is exact and deterministic).
    """
    if hasattr(adata, "obsm"):
        data = adata.obsm[use_rep]
    else:
        inplace = False
        data = adata
    data = np.asarray(data)
    if data.size == 0:
        raise ValueError("matrix is empty")

    if use_dims is not None:
        if isinstance(use_dims, int):
            data = data[:, :use_dims]
        else:
            data = data[:, use_dims]

    if method not in _METHODS:
        raise ValueError("method must be one of 'hora', 'pynndescent', 'kdtree'")
    adj = nearest_neighbour_graph(data, n_neighbors, engine=engine)
    if method == "hora":
        adj = adj.astype(np.float32)

    if inplace:
        adata.obsp["distances"] = adj
    else:
        return adj

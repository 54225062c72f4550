"""CPU restatement of the reference's exact neighbour graph (``snap.pp.knn(method='kdtree')``).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Reference path: ``preprocessing/_knn.py:10-87`` (wrapper, pure Python) ->
``internal.nearest_neighbour_graph(data, k)`` (snapatac2-python/src/knn.rs:8-16) ->
``snapatac2-core/src/utils/knn.rs:9-33``: every point goes into a kd-tree (crate ``kdtree = "0.7"``,
snapatac2-core/Cargo.toml:23, not vendored under /root/reference), and per point
``iter_nearest(point, &squared_euclidean)`` is walked in ascending distance, the point's own index is
dropped (:27), the first ``k`` survivors are kept with ``distance.sqrt()``; ``to_csr_matrix`` (:58-79)
sorts every row by column index.

What is restated: the *result* of that search, which does not depend on the tree -- the k nearest other
points under ``squared_euclidean`` -- and the crate's published distance function
(kdtree 0.7 ``distance.rs``: ``a.iter().zip(b).map(|(x, y)| (x - y) * (x - y)).fold(0, +)``, a left-to-right
float64 sum without fused multiply-add), so distances are comparable bit for bit.  Ties at the k-th
distance: the crate's heap order is unspecified; the restatement (and the CUDA path) keep the smaller index.

PINNING: the Rust search cannot be executed here (no cargo).  The wrapper IS executed:
``tests/golden/make_knn_golden.py`` runs the reference's ``_knn.py`` unmodified with
``internal.nearest_neighbour_graph`` answered by this module and commits what it returns / stores
(``tests/golden/knn_*_ref.npz``); the search itself is cross-checked against ``scipy.spatial.cKDTree``
(an independent exact kd-tree) in ``tests/test_oracle.py``.
"""

from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def squared_euclidean(a: np.ndarray, B: np.ndarray) -> np.ndarray:
    """kdtree 0.7 ``squared_euclidean(a, b)`` for one point ``a`` (d,) or a block (r, 1, d) against rows
    of ``B`` (.., d): differences, squares, left-to-right sum -- one rounding per operation."""
    a = np.asarray(a, dtype=np.float64)
    B = np.asarray(B, dtype=np.float64)
    d = B.shape[-1]
    acc = np.zeros(np.broadcast_shapes(a.shape[:-1], B.shape[:-1]), dtype=np.float64)
    for k in range(d):
        diff = a[..., k] - B[..., k]
        acc = acc + diff * diff
    return acc


def nearest_neighbour_graph(points, k: int, rows=None) -> sp.csr_matrix:
    """knn.rs:9-33 by exhaustive search (any n the memory allows; used for the small fixtures).
    ``rows``: restrict the queries to these point indices (the result then has len(rows) rows)."""
    P = np.ascontiguousarray(points, dtype=np.float64)
    n, d = P.shape
    q = np.arange(n) if rows is None else np.asarray(rows, dtype=np.int64)
    K = max(0, min(int(k), n - 1))
    idx = np.empty((q.size, K), dtype=np.int64)
    dst = np.empty((q.size, K), dtype=np.float64)
    step = max(1, int(4e6 // max(n, 1)))
    cols = np.arange(n)
    for a in range(0, q.size, step):
        rows_a = q[a:a + step]
        d2 = squared_euclidean(P[rows_a][:, None, :], P[None, :, :])          # (r, n)
        d2[np.arange(rows_a.size), rows_a] = np.inf                            # the point itself (:27)
        # ascending (distance, index): stable sort on the distance over index-ordered columns
        order = np.argsort(d2, axis=1, kind="stable")[:, :K]
        order.sort(axis=1)                                                     # to_csr_matrix: by column (:65)
        idx[a:a + step] = cols[order]
        dst[a:a + step] = np.sqrt(np.take_along_axis(d2, order, axis=1))
    indptr = np.arange(q.size + 1, dtype=np.int64) * K
    return sp.csr_matrix((dst.ravel(), idx.ravel(), indptr), shape=(q.size, n))


def nearest_neighbour_graph_kdtree(points, k: int, rows=None, workers: int = 1) -> sp.csr_matrix:
    """The same graph through ``scipy.spatial.cKDTree`` -- an exact kd-tree like the reference's, usable at
    sizes the exhaustive search is not (the CPU baseline of ``scripts/bench_knn.py``).  Distances are
    re-evaluated with ``squared_euclidean`` so they carry the reference's rounding."""
    from scipy.spatial import cKDTree
    P = np.ascontiguousarray(points, dtype=np.float64)
    n, d = P.shape
    q = np.arange(n) if rows is None else np.asarray(rows, dtype=np.int64)
    K = max(0, min(int(k), n - 1))
    tree = cKDTree(P)
    kk = min(n, K + 1 + 8)                      # self + slack for equal distances around the k-th
    _, cand = tree.query(P[q], k=kk, workers=workers)
    cand = np.asarray(cand).reshape(q.size, kk)
    d2 = squared_euclidean(P[q][:, None, :], P[cand])
    d2[cand == q[:, None]] = np.inf
    # ascending (distance, index)
    order = np.lexsort((cand, d2), axis=1)[:, :K]
    sel = np.take_along_axis(cand, order, axis=1)
    sd = np.take_along_axis(d2, order, axis=1)
    by_col = np.argsort(sel, axis=1, kind="stable")
    sel = np.take_along_axis(sel, by_col, axis=1)
    sd = np.take_along_axis(sd, by_col, axis=1)
    indptr = np.arange(q.size + 1, dtype=np.int64) * K
    return sp.csr_matrix((np.sqrt(sd).ravel(), sel.ravel(), indptr), shape=(q.size, n))


def knn(adata, n_neighbors=50, use_dims=None, use_rep="X_spectral", method="kdtree", inplace=True, random_state=0):
    """``preprocessing/_knn.py:53-87`` with the native call answered by ``nearest_neighbour_graph``."""
    if hasattr(adata, "obsm"):
        data = adata.obsm[use_rep]
    else:
        inplace = False
        data = adata
    if data.size == 0:
        raise ValueError("matrix is empty")
    if use_dims is not None:
        data = data[:, :use_dims] if isinstance(use_dims, int) else data[:, use_dims]
    if method != "kdtree":
        raise ValueError("the oracle restates method='kdtree' only")
    adj = nearest_neighbour_graph(data, n_neighbors)
    if inplace:
        adata.obsp["distances"] = adj
    else:
        return adj

"""Synthetic binarised tile matrices (SURVEY.md section 8d recipe).

Planted-cluster cell x bin pattern matrices, deterministic in ``(seed, row)``
so that any row shard regenerates exactly the same rows.  All sampling
arithmetic is integer (a splitmix64 counter hash, 32.32 fixed-point CDFs,
integer binary search, sort, dedup), so the numpy generator here and the CUDA
generator in ``csrc/synth.cu`` produce *bit-identical* CSR arrays from the same
``SynthSpec`` tables.

Recipe: ``K`` clusters with geometric sizes (ratio 0.93); feature popularity
``base_j ~ Gamma(0.5, 1)``; cluster ``z`` boosts its private block of ``m/K``
features by a factor ``U(3, 9)``; every row draws ``nnz_row`` columns from its
cluster's distribution, then dedups and sorts them (achieved nnz/row is a few
percent below nominal).
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import scipy.sparse as sp

_M1 = np.uint64(0x9E3779B97F4A7C15)
_M2 = np.uint64(0xBF58476D1CE4E5B9)
_M3 = np.uint64(0x94D049BB133111EB)
_ONE32 = 1 << 32


def mix64(seed, row, t):
    """splitmix64 finaliser over ``seed*M1 + row*M2 + t*M3`` (mod 2^64).

    Mirrors ``mix64`` in csrc/common.cuh; vectorised over numpy uint64 arrays.
    """
    with np.errstate(over="ignore"):
        x = (np.uint64(seed) * _M1 + np.asarray(row, dtype=np.uint64) * _M2
             + np.asarray(t, dtype=np.uint64) * _M3)
        x = x ^ (x >> np.uint64(30))
        x = x * _M2
        x = x ^ (x >> np.uint64(27))
        x = x * _M3
        x = x ^ (x >> np.uint64(31))
    return x


@dataclass
class SynthSpec:
    """Sampling tables shared by the host and device generators."""
    n: int
    m: int
    nnz_row: int
    n_clusters: int
    seed: int
    feat_cdf: np.ndarray      # uint64[m+1], values in [0, 2^32], feat_cdf[m] == 2^32
    cluster_cdf: np.ndarray   # uint64[K+1], values in [0, 2^32]
    block_start: np.ndarray   # int64[K+1], private feature block of cluster z
    alpha: np.ndarray         # uint64[K], P(draw from global base) in 0.32 fixed point


def make_spec(n, m, nnz_row, n_clusters=48, seed=0, size_ratio=0.93,
              boost=(3.0, 9.0)) -> SynthSpec:
    rng = np.random.default_rng(seed)
    base = rng.gamma(0.5, 1.0, size=m)
    base = np.maximum(base, 1e-300)
    cdf = np.concatenate([[0.0], np.cumsum(base)])
    cdf /= cdf[-1]
    feat_cdf = np.floor(cdf * float(_ONE32)).astype(np.uint64)
    feat_cdf[0] = 0
    feat_cdf[-1] = _ONE32
    feat_cdf = np.maximum.accumulate(feat_cdf)

    sizes = size_ratio ** np.arange(n_clusters)
    ccdf = np.concatenate([[0.0], np.cumsum(sizes)])
    ccdf /= ccdf[-1]
    cluster_cdf = np.floor(ccdf * float(_ONE32)).astype(np.uint64)
    cluster_cdf[0] = 0
    cluster_cdf[-1] = _ONE32

    block_start = (np.arange(n_clusters + 1, dtype=np.int64) * m) // n_clusters
    boosts = rng.uniform(boost[0], boost[1], size=n_clusters)
    alpha = np.empty(n_clusters, dtype=np.uint64)
    for z in range(n_clusters):
        mass = int(feat_cdf[block_start[z + 1]]) - int(feat_cdf[block_start[z]])
        total = float(_ONE32) + (boosts[z] - 1.0) * mass
        alpha[z] = min(_ONE32 - 1, int(float(_ONE32) * float(_ONE32) / total))
    return SynthSpec(n, m, nnz_row, n_clusters, seed, feat_cdf, cluster_cdf, block_start, alpha)


def cluster_of_rows(spec: SynthSpec, rows) -> np.ndarray:
    """Planted cluster label of each global row (stream ``t = 2^40``)."""
    u = mix64(spec.seed, rows, np.uint64(1 << 40)) >> np.uint64(32)
    return (np.searchsorted(spec.cluster_cdf, u, side="right") - 1).astype(np.int64)


def _draw_columns(spec: SynthSpec, rows: np.ndarray) -> np.ndarray:
    """``len(rows) x nnz_row`` raw column draws (unsorted, with duplicates)."""
    rows = np.asarray(rows, dtype=np.uint64)
    z = cluster_of_rows(spec, rows)
    t = np.arange(spec.nnz_row, dtype=np.uint64)
    h = mix64(spec.seed, rows[:, None], t[None, :])
    sel = h >> np.uint64(32)             # mixture selector, 32 bits
    pos = h & np.uint64(0xFFFFFFFF)      # position, 32 bits
    lo = spec.feat_cdf[spec.block_start[z]][:, None]
    hi = spec.feat_cdf[spec.block_start[z + 1]][:, None]
    private = sel >= spec.alpha[z][:, None]
    target = np.where(private, lo + ((pos * (hi - lo)) >> np.uint64(32)), pos)
    col = np.searchsorted(spec.feat_cdf, target, side="right") - 1
    # a private draw must stay inside its block even if CDF entries tie
    blo = spec.block_start[z][:, None]
    bhi = spec.block_start[z + 1][:, None] - 1
    col = np.where(private, np.clip(col, blo, bhi), np.clip(col, 0, spec.m - 1))
    return col.astype(np.int64)


def generate_rows(spec: SynthSpec, row0: int, row1: int, batch: int = 2048):
    """Rows ``[row0, row1)`` as ``(indptr int64, indices int32)`` (sorted, unique)."""
    lens, chunks = [], []
    for s in range(row0, row1, batch):
        rows = np.arange(s, min(row1, s + batch), dtype=np.uint64)
        col = np.sort(_draw_columns(spec, rows), axis=1)
        keep = np.ones_like(col, dtype=bool)
        keep[:, 1:] = col[:, 1:] != col[:, :-1]
        lens.append(keep.sum(axis=1))
        chunks.append(col[keep].astype(np.int32))
    lens = np.concatenate(lens) if lens else np.zeros(0, dtype=np.int64)
    indptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    indices = np.concatenate(chunks) if chunks else np.zeros(0, dtype=np.int32)
    return indptr, indices


def generate_csr(spec: SynthSpec, row0: int = 0, row1: int | None = None,
                 dtype=np.float32) -> sp.csr_matrix:
    """Host generator: binarised scipy CSR of rows ``[row0, row1)``."""
    row1 = spec.n if row1 is None else row1
    indptr, indices = generate_rows(spec, row0, row1)
    data = np.ones(indices.shape[0], dtype=dtype)
    idx_dtype = np.int32 if indices.shape[0] < 2**31 - 1 else np.int64
    return sp.csr_matrix((data, indices.astype(idx_dtype), indptr.astype(idx_dtype)),
                         shape=(row1 - row0, spec.m))

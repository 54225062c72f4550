"""``snap.tl.spectral`` / ``snap.tl.multi_spectral`` on a B200.

Host-side mirror of the reference's operator interface for this path
(``snapatac2-python/python/snapatac2/tools/_embedding.py:129-295`` and
``:483-540``): same names, same argument meaning, same side effects
(``adata.obsm['X_spectral']``, ``adata.uns['spectral_eigenvalue']``), same
exceptions.  Where the reference calls ``internal.spectral_embedding``
(``_embedding.py:249`` -> ``snapatac2-python/src/embedding.rs:24-59``) this
module drives ``libsnapb200.so`` through ctypes instead.

``sample_size`` selects the Nystrom path exactly as in the reference
(``internal.spectral_embedding_nystrom``, embedding.rs:61-129, followed by
``orthogonalize``, _embedding.py:397-413).  There is no CPU fallback: the jaccard
metric (a dense similarity matrix, outside this build) raises NotImplementedError.
"""

from __future__ import annotations

import logging
from typing import Literal

import numpy as np
import scipy.sparse as sp

from . import dist
from .engine import Engine

_engine: Engine | None = None


def _check_engine(eng: Engine) -> Engine:
    """Under torch.distributed an engine must span the world: a context that never joined the
    communicator would silently treat its shard as the whole matrix (all-reduces are no-ops)."""
    rank, ws = dist.world()
    if ws > 1 and (eng.nranks != ws or eng.rank != rank):
        if eng.nranks == 1:
            dist.attach_engine_comm(eng)
        else:
            raise RuntimeError(f"engine belongs to a communicator of {eng.nranks} ranks (rank {eng.rank}); "
                               f"torch.distributed world is {ws} (rank {rank})")
    return eng


def default_engine() -> Engine:
    """Process-wide engine on ``cuda:LOCAL_RANK`` (created on first use and, under
    torch.distributed, joined to a communicator spanning the world)."""
    global _engine
    if _engine is None:
        _engine = Engine()
        dist.attach_engine_comm(_engine)
    return _engine


def _resolve_features(adata, features):
    """_embedding.py:225-229: a string names a boolean column of ``adata.var``."""
    if isinstance(features, str):
        if features in adata.var:
            col = adata.var[features]
            return col.to_numpy() if hasattr(col, "to_numpy") else np.asarray(col)
        raise NameError("Please call `select_features` first or explicitly set `features = None`")
    return features


def _feature_mask(features, n_vars, feature_weights):
    """Boolean keep-mask (+ weights re-ordered to ascending column order).

    The reference slices columns in the order given (``to_select_elem``,
    embedding.rs:36-39) and indexes ``feature_weights`` by the *sliced*
    position (:321).  Column order does not change the embedding, so integer
    index arrays are turned into a mask and the weights permuted to match.
    """
    if features is None:
        return None, feature_weights
    f = np.asarray(features)
    if f.dtype == bool:
        if f.shape[0] != n_vars:
            raise ValueError("boolean feature mask must have length n_vars")
        return f, feature_weights
    idx = f.astype(np.int64)
    idx = np.where(idx < 0, idx + n_vars, idx)
    if idx.size and (idx.min() < 0 or idx.max() >= n_vars):
        raise IndexError("feature index out of range")
    if np.unique(idx).size != idx.size:
        raise ValueError("duplicate feature indices are not supported")
    mask = np.zeros(n_vars, dtype=bool)
    mask[idx] = True
    if feature_weights is not None:
        order = np.argsort(idx, kind="stable")
        feature_weights = np.asarray(feature_weights, dtype=np.float64)[order]
    return mask, feature_weights


def _get_csr(adata):
    """``adata.X`` as a scipy CSR matrix -- or, when it is not one (a backed AnnData element, a lazy
    array, a generator of row blocks), a :class:`RowBlocks` source the engine assembles on the device
    block by block (``with_anndata!`` dispatch, utils/anndata.rs:174-236: in-memory and backed data
    take the same path)."""
    X = adata.X
    if isinstance(X, np.ndarray):
        X = sp.csr_matrix(X)
    if sp.issparse(X):
        if X.format != "csr":
            raise ValueError("adata.X must be a CSR matrix")
        return X
    return RowBlocks(X, adata.n_vars)


class RowBlocks:
    """Row-block view of an ``adata.X`` that is not an in-memory scipy matrix.

    Accepted sources, in this order: an object with ``chunked(chunk_size)`` (SnapATAC2's backed
    elements; yields ``(block, start, end)`` or blocks), an object with ``shape`` and row slicing
    (``anndata``'s backed sparse dataset, h5py-style lazy arrays), or any iterable of CSR blocks
    (a one-shot generator works: the data is only walked once)."""

    def __init__(self, source, n_vars):
        self.source = source
        self.n_vars = int(n_vars)

    @classmethod
    def from_csr_group(cls, group, n_vars=None):
        """The on-disk CSR layout of an ``.h5ad`` / anndata-rs file (``src/utils/anndata.rs:213-229`` reads it
        through ``chunked_X``): a group with the datasets ``indptr``, ``indices``, ``data`` and the attribute
        ``shape``.  ``group`` is an ``h5py.Group`` / ``zarr`` group -- anything whose members slice like arrays;
        one block of rows is read per step, nothing else is held in memory."""
        attrs = getattr(group, "attrs", {})
        enc = attrs.get("encoding-type", "csr_matrix")
        if isinstance(enc, bytes):
            enc = enc.decode()
        if enc not in ("csr_matrix", "csr_array"):
            raise ValueError(f"X is stored as {enc!r}: the cosine path reads rows, a CSR layout is required")
        indptr, indices, data = group["indptr"], group["indices"], group["data"]
        n_obs = int(indptr.shape[0]) - 1
        shape = attrs.get("shape", None)
        if n_vars is None:
            if shape is None:
                raise ValueError("n_vars is neither given nor stored in the group's 'shape' attribute")
            n_vars = int(shape[1])

        def blocks(chunk_size):
            for i in range(0, n_obs, chunk_size):
                j = min(i + chunk_size, n_obs)
                ptr = np.asarray(indptr[i:j + 1], dtype=np.int64)
                lo, hi = int(ptr[0]), int(ptr[-1])
                yield sp.csr_matrix((np.asarray(data[lo:hi]), np.asarray(indices[lo:hi]), ptr - lo), shape=(j - i, n_vars))

        src = type("CsrGroupRows", (), {"chunked": staticmethod(blocks)})()
        return cls(src, n_vars)

    @classmethod
    def from_h5ad(cls, path, key="X"):
        """Rows of ``path[key]`` (an ``.h5ad`` written by anndata / SnapATAC2 with a CSR ``X``), read block by
        block with h5py.  h5py is not part of this image: the import happens here, on use."""
        import h5py      # noqa: deliberately late
        f = h5py.File(path, "r")
        rb = cls.from_csr_group(f[key])
        rb._file = f     # keep the file open as long as the view lives
        return rb

    def blocks(self, chunk_size):
        src = self.source
        if hasattr(src, "chunked"):
            for item in src.chunked(chunk_size):
                yield item[0] if isinstance(item, tuple) else item
        elif hasattr(src, "shape") and hasattr(src, "__getitem__"):
            for i in range(0, int(src.shape[0]), chunk_size):
                yield src[i:i + chunk_size]
        else:
            yield from src


def spectral_embedding(engine: Engine, X, selected_features, n_components, random_state,
                       feature_weights=None, *, n_global=None, row0=0, binarized=None,
                       tol=0.0, block=0, max_basis=0, max_ops=0, return_parts=False, scale_by_sqrt_eval=False,
                       chunk_size=20000):
    """Counterpart of the PyO3 entry ``internal.spectral_embedding``
    (embedding.rs:24-59): load -> [select] -> [weights] -> prepare -> eigsh.
    ``scale_by_sqrt_eval``: see :func:`_weight_by_sd`."""
    blocks = isinstance(X, RowBlocks)
    mask, fw = _feature_mask(selected_features, X.n_vars if blocks else X.shape[1], feature_weights)
    if blocks:                      # backed / lazy X: assembled on the device block by block
        engine.load_blocks(X.blocks(chunk_size), X.n_vars)
        if n_global is not None:
            engine.set_geometry(n_global, row0)
    else:
        # the all-ones scan of X.data runs on background threads while the GPU prepares and solves; the verdict
        # is collected before the result is returned (values other than 1: ship them and compute again)
        engine.load_csr(X, n_global=n_global, row0=row0, binarized=binarized, defer_value_scan=not binarized)

    def compute():
        if mask is not None:
            engine.select_features(mask)
        engine.set_feature_weights(fw)
        idf_, degree_ = engine.prepare(want_outputs=return_parts)
        ev_, evec_ = engine.eigsh(n_components, seed=random_state, tol=tol, block=block,
                                  max_basis=max_basis, max_ops=max_ops, scale_by_sqrt_eval=scale_by_sqrt_eval)
        return ev_, evec_, idf_, degree_

    def all_ones():                     # the verdict must be the same on every rank: the recomputation is collective
        ok = engine.values_all_ones()
        if dist.world()[1] > 1:
            ok = bool(dist.allreduce_array(np.array([1.0 if ok else 0.0]), "min")[0] > 0.5)
        return ok

    try:
        evals, evecs, idf, degree = compute()
    except Exception:
        if blocks or binarized or engine.values_all_ones():
            raise
        evals = None                    # the pattern-only attempt may fail where the valued matrix does not
    if not blocks and not binarized and not all_ones():
        if mask is not None:            # the resident pattern is already column-selected: select the values the same way
            engine.load_csr(X, n_global=n_global, row0=row0, binarized=False)
        else:
            engine.load_values()
        evals, evecs, idf, degree = compute()
    if return_parts:
        return evals, evecs, idf, degree
    return evals, evecs


def _weight_by_sd(evals, evecs, already_scaled):
    """``weighted_by_sd`` (tools/_embedding.py:286-289): keep the components with a positive eigenvalue,
    scaled by ``sqrt(eigenvalue)``.  On the full-matrix path the scaling is folded into the final basis
    rotation on the device (``already_scaled``), so a 1M x 30 result is not copied twice more on the host."""
    keep = evals > 0
    if not already_scaled:
        evecs = evecs[:, keep] * np.sqrt(evals[keep])
    elif not keep.all():
        evecs = np.ascontiguousarray(evecs[:, keep])
    return evals[keep], evecs


def orthogonalize(evals, evecs):
    """``orthogonalize`` of the reference wrapper (tools/_embedding.py:397-413): turns the Nystrom
    extension into an orthogonal eigenbasis (k x k algebra after one thin SVD; host side).

    Row-sharded (``torchrun``): ``evecs`` is this rank's block of rows.  The singular values and right
    singular vectors of the whole matrix come from its k x k Gram matrix, summed over the shards
    (``sigma^2, V = eigh(U^T U)``); everything after that is k x k or row-local."""
    if dist.world()[1] > 1:
        gram = dist.allreduce_array(evecs.T @ evecs, "sum")
        s2, v = np.linalg.eigh(gram)
        order = np.argsort(s2)[::-1]
        sigma, v = np.sqrt(np.maximum(s2[order], 0.0)), v[:, order]
    else:
        _, sigma, vt = np.linalg.svd(evecs, full_matrices=False)
        v = vt.T
    b = np.multiply(v.T, evals.reshape((1, -1))) @ v
    b = b * sigma.reshape((-1, 1)) * sigma.reshape((1, -1))
    evals_new, evecs_new = np.linalg.eig(b)
    ix = evals_new.argsort()[::-1]
    evals_new = evals_new[ix]
    evecs_new = evecs_new[:, ix] / sigma.reshape((-1, 1))
    return evals_new, evecs @ v @ evecs_new


def _nystrom_normalise(q, v, chunk_size, n_global, row0):
    """The per-chunk degree normalisation of ``nystrom`` (embedding.rs:224-227), over ``chunk_size`` row
    blocks of the GLOBAL row order exactly as the reference streams them; a block that spans two
    shards gets its column sums and its smallest positive degree from both (two tiny all-reduces)."""
    n_chunks = -(-n_global // chunk_size)
    k = q.shape[1]
    sums = np.zeros((n_chunks, k), dtype=np.float64)
    spans = list(dist.chunk_overlaps(row0, q.shape[0], chunk_size))
    for c, lo, hi in spans:
        sums[c] = q[lo:hi].sum(axis=0)
    sums = dist.allreduce_array(sums, "sum")
    mins = np.full(n_chunks, np.inf, dtype=np.float64)
    dds = {}
    for c, lo, hi in spans:
        t = sums[c] * v                                      # :224
        dd = q[lo:hi] @ t                                    # :225
        dds[c] = dd
        pos = dd[dd > 0]
        if pos.size:
            mins[c] = pos.min()
    mins = dist.allreduce_array(mins, "min")
    for c, lo, hi in spans:
        dd = dds[c]
        dd[dd <= 0] = mins[c]                                # :226
        q[lo:hi] /= np.sqrt(dd)[:, None]                     # :227
    return q


def _streamed_degrees(device, chunk_csr, n_local, chunk_size, w):
    """``compute_degrees`` (embedding.rs:328-360) with one ``chunk_size`` block of cells on the device at a
    time: pass 1 accumulates the column sums of the normalised rows (all-reduced over the shards), pass 2
    forms ``d = Xhat colsum - 1`` block by block.  The blocks go through a context of their own that is
    not part of the communicator (shards have different numbers of blocks)."""
    csum = np.zeros(w.shape[0], dtype=np.float64)
    solo = Engine(device)
    try:
        for i in range(0, n_local, chunk_size):
            solo.load_csr(chunk_csr(i))
            solo.set_feature_weights(w)
            solo.prepare_projection(want_outputs=False)
            csum += solo.project_t_ones()
        csum = dist.allreduce_array(csum, "sum")
        deg_local = np.empty(n_local, dtype=np.float64)
        for i in range(0, n_local, chunk_size):
            solo.load_csr(chunk_csr(i))
            solo.set_feature_weights(w)
            deg_local[i:i + chunk_size] = solo.project(csum[:, None].astype(np.float32))[:, 0].astype(np.float64) - 1.0
    finally:
        solo.close()
    return deg_local


def _draw_landmarks(n, sample_size, degree_all):
    """``rand::seq::index::sample`` / ``sample_weighted`` with ``StdRng::seed_from_u64(2023)`` (embedding.rs:87-94)
    cannot be reproduced outside Rust: numpy's ``default_rng(2023)`` stands in (identical on every rank)."""
    rng = np.random.default_rng(2023)
    if degree_all is not None:                               # compute_probs (:362-365)
        p = 1.0 / degree_all
        return rng.choice(n, size=sample_size, replace=False, p=p / p.sum())
    return rng.choice(n, size=sample_size, replace=False)


def spectral_embedding_nystrom(engine: Engine, X, selected_features, n_components, sample_size,
                               weighted_by_degree, chunk_size, feature_weights=None, *, landmarks=None,
                               seed_engine: Engine | None = None, tol=0.0, block=0, return_parts=False,
                               stream: bool | None = None):
    """Counterpart of ``internal.spectral_embedding_nystrom`` (embedding.rs:61-129): spectral
    embedding of ``sample_size`` landmark cells, extended to every cell (``nystrom``, :194-267).

    Device work: IDF weights and row norms of all cells, the embedding of the landmark rows (a
    second context that shares the stream and communicator: device-side row gather, prepare and
    eigsh with the global weights), ``seed.T @ evecs`` (``snapb200_project`` transposed, summed
    over the row shards) and ``sample @ (...)`` for every cell (``snapb200_project``).  The per-chunk
    degree normalisation (:224-227) is n x k algebra on the host over the same global ``chunk_size``
    row blocks the reference streams.  Landmarks: see :func:`_draw_landmarks`; pass ``landmarks``
    (global row ids) to fix them.

    Row-sharded under ``torchrun``: every rank passes its block of cells (or ``X=None`` if the engine
    already holds it, e.g. generated on the device); the landmark matrix is itself row-sharded --
    every rank contributes the landmarks that fall into its block -- and embedded by the ordinary
    distributed solve, so no rows ever travel between GPUs.

    ``stream=True`` keeps only one ``chunk_size`` block of cells on the device at a time, the way
    the reference iterates a backed AnnData (:76-84, :109-118): document frequencies are
    accumulated chunk by chunk, degree-weighted sampling takes two more streaming passes
    (``compute_degrees``, :328-360), then every chunk is loaded, normalised and projected on its
    own.  Default: stream when the shard would not fit next to its tiled copy (> 6e9 stored entries).
    """
    rank, world = dist.world()
    resident = X is None
    if resident:
        n_local, m_all, _ = engine.shape()
        mask, fw, cols = None, feature_weights, None
        stream = False
    else:
        mask, fw = _feature_mask(selected_features, X.shape[1], feature_weights)
        n_local = X.shape[0]
        cols = None if mask is None else np.flatnonzero(mask)
        if stream is None:
            stream = X.nnz > 6_000_000_000
    if world > 1:
        n_locals = dist.allgather_ints(n_local)
        n, row0 = sum(n_locals), dist.shard_offsets(n_locals)[rank]
    else:
        n, row0 = n_local, 0

    def chunk_csr(i):
        Xc = X[i:i + chunk_size]
        if cols is not None:
            Xc = Xc[:, cols]
        return sp.csr_matrix(Xc)

    degree_all = None
    if stream:
        if fw is not None:
            w = np.asarray(fw, dtype=np.float64)
        else:                                               # idf_from_chunks (:288-312)
            m_sel = X.shape[1] if cols is None else cols.size
            df = np.zeros(m_sel, dtype=np.int64)
            for i in range(0, n_local, chunk_size):
                df += np.bincount(chunk_csr(i).indices, minlength=m_sel)
            df = dist.allreduce_array(df, "sum")
            if np.all(df == df[0]):
                w = np.ones(m_sel, dtype=np.float64)
            else:
                d = df.astype(np.float64)
                d[d == 0] = 1.0
                d[d == n] = n - 1.0
                w = np.log(n / d)
        if landmarks is None and weighted_by_degree:        # compute_degrees (:328-360): two streaming passes
            deg_local = _streamed_degrees(engine.device, chunk_csr, n_local, chunk_size, w)
            degree_all = np.concatenate(dist.allgather_objects(deg_local)) if world > 1 else deg_local
    else:
        if not resident:
            engine.load_csr(X, n_global=n, row0=row0)
            if mask is not None:
                engine.select_features(mask)
        engine.set_feature_weights(fw)
        if landmarks is None and weighted_by_degree:
            _, deg_local = engine.prepare()                 # compute_degrees (:328-360)
            degree_all = np.concatenate(dist.allgather_objects(deg_local)) if world > 1 else deg_local
        w, _rho = engine.prepare_projection()               # idf over all cells (:76-84) + row norms
    if landmarks is None:
        landmarks = _draw_landmarks(n, sample_size, degree_all)
    lm_given = np.asarray(landmarks, dtype=np.int64)
    lm_order = np.argsort(lm_given, kind="stable")
    landmarks = lm_given[lm_order]                           # row order of the seed matrix is immaterial: ascending
    mine = landmarks[(landmarks >= row0) & (landmarks < row0 + n_local)] - row0
    if world > 1:
        counts = dist.allgather_ints(mine.size)
        s_total, s_row0 = sum(counts), dist.shard_offsets(counts)[rank]
    else:
        s_total, s_row0 = int(mine.size), 0

    # landmark rows (:95-99), embedded on a second context with the global weights
    own_seed = seed_engine is None
    if own_seed:
        seed = Engine(engine.device)
        engine.attach_view(seed)                             # same stream, same communicator
    else:
        seed = seed_engine
    try:
        if stream:
            Xs = X[mine]
            if cols is not None:
                Xs = Xs[:, cols]
            seed.load_csr(sp.csr_matrix(Xs), n_global=s_total, row0=s_row0)
        else:
            engine.gather_rows_into(mine, seed, n_global=s_total, row0=s_row0)
        seed.set_feature_weights(w)
        _, d = seed.prepare()
        v, u = seed.eigsh(n_components, seed=0, tol=tol, block=block)   # spectral_mf(seed, k, 0) (:104)
        u = u / np.sqrt(d)[:, None]                          # :206-210
        u = u / v[None, :]                                   # :211-215
        proj = seed.project_t(u)                             # seed.T @ evecs, summed over the shards
    finally:
        if own_seed:
            seed.close()
    if stream:                                               # sample @ (seed.T @ evecs), chunk by chunk
        q = np.empty((n_local, proj.shape[1]), dtype=np.float64)
        solo = Engine(engine.device)
        try:
            for i in range(0, n_local, chunk_size):
                solo.load_csr(chunk_csr(i))
                solo.set_feature_weights(w)
                q[i:i + chunk_size] = solo.project(proj)
        finally:
            solo.close()
    else:
        q = engine.project(proj).astype(np.float64)          # ... or for every cell at once
    q = _nystrom_normalise(q, v, chunk_size, n, row0)
    if return_parts:
        if world == 1:                                       # landmark degrees back in the order the caller gave
            d_sorted, d = d, np.empty_like(d)
            d[lm_order] = d_sorted
        return v, q, w, d, lm_given
    return v, q


def spectral(
    adata,
    n_comps: int = 30,
    features: str | np.ndarray | None = "selected",
    random_state: int = 0,
    sample_size: int | float | None = None,
    sample_method: Literal["random", "degree"] = "random",
    chunk_size: int = 20000,
    distance_metric: Literal["jaccard", "cosine"] = "cosine",
    weighted_by_sd: bool = True,
    feature_weights: list[float] | None = None,
    inplace: bool = True,
    *,
    engine: Engine | None = None,
    tol: float = 0.0,
    block: int = 0,
    on_degenerate: Literal["raise", "nan"] = "raise",
) -> tuple[np.ndarray, np.ndarray] | None:
    """Laplacian-eigenmaps embedding, matrix-free, on the GPU.

    Same call as ``snap.tl.spectral`` (tools/_embedding.py:129-141).  Under
    ``torchrun`` every rank passes the AnnData holding *its* contiguous block
    of cells (rank order = row order); eigenvalues are replicated and every
    rank receives the embedding of its own rows.

    Extra keyword-only knobs: ``engine`` (reuse a context), ``tol`` (relative
    residual, default 1e-5; the reference asks ARPACK for machine precision),
    ``block`` (Lanczos block width 4/8/16, default 4), ``on_degenerate``: a cell
    whose row is empty after feature selection, or whose degree is not positive,
    makes the reference divide by zero -- its row norm / ``1/d`` become inf or NaN
    (embedding.rs:146,152,323), the NaN spreads through the operator and ARPACK
    returns NaN everywhere.  ``"raise"`` (default) reports the cells instead;
    ``"nan"`` reproduces the reference's outcome (all-NaN eigenvalues and embedding)
    for callers that rely on it.
    """
    np.random.seed(random_state)                                    # :223

    features = _resolve_features(adata, features)                   # :225-229

    rank, ws = dist.world()
    n_local = adata.n_obs
    if ws > 1:
        n_locals = dist.allgather_ints(n_local)
        n_global, row0 = sum(n_locals), dist.shard_offsets(n_locals)[rank]
    else:
        n_global, row0 = n_local, 0

    n_comps = min(adata.n_vars - 1, n_global - 1, n_comps)          # :231

    n_sample = n_global                                             # :233-245
    if sample_size is None:
        sample_size = n_sample
    elif isinstance(sample_size, int):
        if sample_size <= 1:
            raise ValueError("when sample_size is an integer, it should be > 1")
        if sample_size > n_sample:
            sample_size = n_sample
    else:
        if sample_size <= 0.0 or sample_size > 1.0:
            raise ValueError("when sample_size is a float, it should be > 0 and <= 1")
        sample_size = int(sample_size * n_sample)

    if distance_metric != "cosine":
        raise NotImplementedError("only distance_metric='cosine' (the matrix-free path) runs on the GPU")

    eng = _check_engine(engine) if engine is not None else default_engine()
    X = _get_csr(adata)
    if sample_size < n_sample:                                      # :257-265
        logging.getLogger(__name__).info("Perform spectral embedding using the Nystrom algorithm...")
        # like the reference (:263), the Nystrom call does not receive feature_weights: IDF from all cells
        if isinstance(X, RowBlocks):     # backed / lazy X: assemble the shard on the device first, then as resident data
            mask, _ = _feature_mask(features, X.n_vars, None)
            eng.load_blocks(X.blocks(chunk_size), X.n_vars)
            eng.set_geometry(n_global, row0)
            if mask is not None:
                eng.select_features(mask)
            X, features = None, None
        v, u = spectral_embedding_nystrom(eng, X, features, n_comps, sample_size, sample_method != "random",
                                          chunk_size, tol=tol, block=block)
        evals, evecs = orthogonalize(v, u)
        scaled = False
    else:
        try:
            evals, evecs = spectral_embedding(eng, X, features, n_comps, random_state, feature_weights,
                                              n_global=n_global, row0=row0, tol=tol, block=block,
                                              scale_by_sqrt_eval=weighted_by_sd, chunk_size=chunk_size)   # :249
        except RuntimeError as e:
            if on_degenerate != "nan" or "non-positive degree" not in str(e):
                raise
            evals = np.full(n_comps, np.nan)
            evecs = np.full((n_local, n_comps), np.nan)
            scaled = False                  # (weighted_by_sd then keeps nothing: `evals[i] > 0` is false for NaN, as in the reference)
        else:
            scaled = weighted_by_sd
    logging.getLogger(__name__).info("spectral: %s", eng.stats())

    if weighted_by_sd:                                              # :286-289
        evals, evecs = _weight_by_sd(evals, evecs, scaled)

    if inplace:                                                     # :291-293
        adata.uns["spectral_eigenvalue"] = evals
        adata.obsm["X_spectral"] = evecs
        return None
    return evals, evecs


def _view_engines(engine: Engine, n_views: int) -> list[Engine]:
    """``engine`` plus ``n_views - 1`` further contexts on the same GPU that share its stream and
    communicator (kept with the engine and reused across calls)."""
    extra = getattr(engine, "_view_engines", None)
    if extra is None:
        extra = engine._view_engines = []
    while len(extra) < n_views - 1:
        v = Engine(engine.device)
        engine.attach_view(v)
        extra.append(v)
    return [engine] + extra[:n_views - 1]


def multi_spectral_embedding(engine: Engine, xs, selected_features, weights, n_components, random_state,
                             *, sample_rows=None, tol=0.0, block=0, return_parts=False, container="csr_matrix",
                             scale_by_sqrt_eval=False, n_local=None, chunk_size=20000):
    """Counterpart of ``internal.multi_spectral_embedding`` (embedding.rs:388-452), entirely on the
    device and without the column concatenation ever being formed.

    Every view gets its own context on the GPU: load, column selection, IDF weights, row norms,
    both tiled copies and its single-view degrees -- the ordinary ``prepare`` (:404-416).  The view
    normaliser ``frobenius_norm`` (:454-471) is evaluated on all rows when ``n <= 2000`` and on
    2000 sampled rows otherwise (:417-421); the reference samples with Rust's ``StdRng(2023)``,
    which cannot be reproduced outside Rust, so the sample is ``sample_rows`` if given, else
    ``numpy.random.RandomState(2023)`` -- a documented deviation.  ``container`` names the scipy
    container the embedded Python snippet receives: ``"csr_matrix"`` (what pyanndata builds, the
    default; ``np.power`` is then a matrix power and the snippet equals ``||(X X^T) 1||^2``, two
    SpMVs on the device) or ``"csr_array"`` (element-wise square, the Frobenius norm the name
    promises; the <= 2000 x 2000 Gram of the sampled rows is formed on the host).  The scaled views
    (:428-442) are then chained behind one operator, ``A = sum_v X~_v X~_v^T - D^-1`` (the hstack of
    :443 and ``spectral_mf`` of :447 in factored form; ``snapb200_combine_views``).  Under
    ``torchrun`` every rank passes its block of cells of every view.
    """
    rank, world = dist.world()
    if n_local is None:
        n_local = next(X.shape[0] for X in xs if not isinstance(X, RowBlocks))
    if world > 1:      # every rank passes its own contiguous block of cells, the same block of every view
        n_locals = dist.allgather_ints(n_local)
        n_global, row0 = sum(n_locals), dist.shard_offsets(n_locals)[rank]
    else:
        n_global, row0 = n_local, 0
    if n_global <= 2000:
        rows = np.arange(n_global)
    elif sample_rows is not None:
        rows = np.sort(np.asarray(sample_rows))
    else:
        rows = np.sort(np.random.RandomState(2023).choice(n_global, 2000, replace=False))
    mine = rows[(rows >= row0) & (rows < row0 + n_local)] - row0
    views = _view_engines(engine, len(xs))
    norms, idfs = [], []
    for eng, X, sel in zip(views, xs, selected_features):
        if isinstance(X, RowBlocks):                # backed / lazy view: assembled on the device block by block
            if container != "csr_matrix":
                raise NotImplementedError("the csr_array reading needs the sampled rows on the host (in-memory views only)")
            mask, _ = _feature_mask(sel, X.n_vars, None)
            eng.load_blocks(X.blocks(chunk_size), X.n_vars)
            if eng.n_local != n_local:
                raise ValueError("all views must hold the same cells")
            eng.set_geometry(n_global, row0)
        else:
            if not sp.issparse(X) or X.format != "csr":
                X = sp.csr_matrix(X)
            if X.shape[0] != n_local:
                raise ValueError("all views must hold the same cells")
            mask, _ = _feature_mask(sel, X.shape[1], None)
            eng.load_csr(X, n_global=n_global, row0=row0)
        if mask is not None:
            eng.select_features(mask)
        eng.set_feature_weights(None)           # the views always use their own IDF (:413)
        idf, _ = eng.prepare(want_outputs=return_parts)
        if container == "csr_matrix":
            total = eng.view_frobenius(mine)
        elif container == "csr_array":
            w, rho = eng.get_vector("weights"), eng.get_vector("rho")
            Xm = X[mine] if mask is None else X[mine][:, np.flatnonzero(mask)]
            xhat_s = sp.csr_matrix(sp.diags(1.0 / rho[mine]) @ (sp.csr_matrix(Xm, dtype=np.float64) @ sp.diags(w)))
            if world > 1:   # the sampled unit rows of all shards on every rank
                xhat_s = sp.csr_matrix(sp.vstack(dist.allgather_objects(xhat_s), format="csr"))
            g = xhat_s @ xhat_s.T
            total = float(g.multiply(g).sum())
        else:
            raise ValueError("container must be 'csr_matrix' or 'csr_array'")
        norms.append(float(np.sqrt(total - len(rows))))
        idfs.append(idf)
    ws = [w / nrm for w, nrm in zip(weights, norms)]
    w_sum = float(sum(ws))
    scales = [float(np.sqrt(w / w_sum)) for w in ws]
    degree = engine.combine_views(views, scales, want_degree=return_parts)
    evals, evecs = engine.eigsh(n_components, seed=random_state, tol=tol, block=block, scale_by_sqrt_eval=scale_by_sqrt_eval)
    if return_parts:
        return evals, evecs, np.concatenate(idfs), degree, norms
    return evals, evecs


def multi_spectral(adatas, n_comps: int = 30, features="selected", weights=None,
                   random_state: int = 0, weighted_by_sd: bool = True, *,
                   engine: Engine | None = None, sample_rows=None, container: str = "csr_matrix"):
    """Laplacian eigenmaps on several modalities at once -- same call as
    ``snap.tl.multi_spectral`` (tools/_embedding.py:483-540); returns
    ``(evals, evecs)`` and does not write into the AnnData objects."""
    np.random.seed(random_state)                                            # :523
    if features is None or isinstance(features, str):                       # :525-528
        features = [features] * len(adatas)
    if all(isinstance(f, str) for f in features):
        features = [_resolve_features(a, f) for a, f in zip(adatas, features)]
    if weights is None:                                                     # :530-531
        weights = [1.0 for _ in adatas]
    # (no n_comps clamp here: the reference's multi_spectral has none, :523-533)
    eng = _check_engine(engine) if engine is not None else default_engine()
    evals, evecs = multi_spectral_embedding(eng, [_get_csr(a) for a in adatas], features, weights,
                                            n_comps, random_state, sample_rows=sample_rows, container=container,
                                            scale_by_sqrt_eval=weighted_by_sd, n_local=adatas[0].n_obs)   # :533
    if weighted_by_sd:                                                      # :535-538
        evals, evecs = _weight_by_sd(evals, evecs, True)
    return evals, evecs

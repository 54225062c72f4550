"""Numpy/scipy restatement of SnapATAC2's matrix-free spectral embedding.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  PINNED against outputs of the
reference's own code executed in the build container
(``tests/golden/make_ref_golden.py`` -> ``tests/golden/*_ref.npz``): the pure-Python
``SpectralMatrixFree.fit``/``_eigen``/``orthogonalize``/``spectral`` of
``tools/_embedding.py`` run unmodified, and the Python snippet embedded in
``frobenius_norm`` (embedding.rs:456-460) executed on both scipy containers.  What
stays restatement-only: the Rust IDF (embedding.rs:269-286, a closed form checked
separately) and the Rust RNG draws (landmarks, multi-view row sample).

Every function names the reference lines it follows.  Paths are relative to
``/root/reference``; ``embedding.rs`` = ``snapatac2-python/src/embedding.rs`` and
``_embedding.py`` = ``snapatac2-python/python/snapatac2/tools/_embedding.py``.

Third-party arithmetic on the path that is *not* vendored in the reference:
  * scipy ``sparse.linalg.eigsh`` (ARPACK) and sparse ``@`` -- constraint
    ``scipy>=1.4,<2`` (snapatac2-python/pyproject.toml:45); the oracle calls
    the installed scipy directly, exactly as the reference does.
  * nalgebra-sparse 0.10 CSR x dense vector (Cargo.toml:33) -- row-wise dot
    products, restated with scipy ``@``.
  * numpy legacy global RNG for ARPACK's start vector (embedding.rs:161,167).
"""

from __future__ import annotations

import numpy as np
import scipy.sparse as sp
from scipy.sparse.linalg import LinearOperator, eigsh


# --------------------------------------------------------------------------
# a3: IDF feature weights                                 embedding.rs:269-286
# --------------------------------------------------------------------------
def idf(mat: sp.csr_matrix) -> np.ndarray:
    """Inverse document frequency from the *stored-entry pattern*.

    embedding.rs:270-271 counts every stored column index (explicit zeros
    included); :273-274 returns all-ones if every column has the same count;
    :277-281 clamp ``0 -> 1`` and ``n -> n-1``; :282 ``ln(n / df)``.
    """
    n, m = mat.shape
    df = np.bincount(mat.indices, minlength=m).astype(np.float64)
    if m == 0 or np.all(df == df[0]):
        return np.ones(m, dtype=np.float64)
    df = df.copy()
    df[df == 0.0] = 1.0
    df[df == float(n)] = float(n) - 1.0
    return np.log(float(n) / df)


# --------------------------------------------------------------------------
# a4: feature weighting + row L2 normalisation            embedding.rs:315-326
# --------------------------------------------------------------------------
def normalize(mat: sp.csr_matrix, feature_weights: np.ndarray) -> sp.csr_matrix:
    """``x_ij *= w_j`` (:321) then divide each row by its L2 norm (:323-324).

    No guard for empty rows: norm 0 gives NaN exactly as the reference does.
    Returns a new float64 CSR (the reference mutates in place).
    """
    w = np.asarray(feature_weights, dtype=np.float64)
    out = sp.csr_matrix(
        (mat.data.astype(np.float64) * w[mat.indices], mat.indices.copy(), mat.indptr.copy()),
        shape=mat.shape,
    )
    sq = out.data * out.data
    indptr = out.indptr
    # per-row sums of squares; reduceat only over non-empty rows (an empty row
    # keeps norm 0 and turns into NaN below, like the reference)
    norms = np.zeros(out.shape[0], dtype=np.float64)
    nz = np.flatnonzero(np.diff(indptr) > 0)
    if nz.size:
        norms[nz] = np.sqrt(np.add.reduceat(sq, indptr[nz]))
    with np.errstate(divide="ignore", invalid="ignore"):
        out.data /= np.repeat(norms, np.diff(indptr))
    return out


# --------------------------------------------------------------------------
# a5 + a6 + a7: degrees, operator, ARPACK                 embedding.rs:133-191
# --------------------------------------------------------------------------
def operator_pieces(xhat: sp.csr_matrix):
    """Degree block of ``spectral_mf`` (embedding.rs:139-152).

    Returns ``(X_tilde, dinv, col_sum, degree)`` where ``col_sum = X^T 1``
    (:139-144), ``degree = X col_sum - 1`` and ``dinv = 1/degree`` (:145-146),
    and ``X_tilde`` has row ``i`` scaled by ``sqrt(dinv_i)`` (:149-152).
    """
    col_sum = np.zeros(xhat.shape[1], dtype=np.float64)
    np.add.at(col_sum, xhat.indices, xhat.data)
    degree = xhat @ col_sum - 1.0
    with np.errstate(divide="ignore", invalid="ignore"):
        dinv = 1.0 / degree
        xt = sp.diags(np.sqrt(dinv)) @ xhat
    return sp.csr_matrix(xt), dinv, col_sum, degree


def _arpack(xt: sp.csr_matrix, dinv: np.ndarray, k: int, seed: int, counter=None):
    """The embedded Python snippet of embedding.rs:158-171, restated.

    ``numpy.random.seed(seed)``; operator ``v -> X (X^T v) - dinv*v`` written
    as ``X @ (v.T @ X).T - D * v`` (:162-163); ``eigsh(A, k=k, v0=rand(n))``
    with scipy defaults (which='LM', ncv=max(2k+1,20), tol=0, maxiter=10n)
    (:166-167); eigenpairs sorted by descending eigenvalue (:168-170).
    """
    np.random.seed(seed)

    def f(v):
        if counter is not None:
            counter[0] += 1
        return xt @ (v.T @ xt).T - dinv * v

    n = xt.shape[0]
    op = LinearOperator((n, n), matvec=f, dtype=np.float64)
    evals, evecs = eigsh(op, k=k, v0=np.random.rand(n))
    order = evals.argsort()[::-1]
    return evals[order], evecs[:, order]


def spectral_mf(xhat: sp.csr_matrix, n_components: int, random_state: int, counter=None):
    """``spectral_mf`` (embedding.rs:133-191): returns ``(evals, evecs, degree)``."""
    xt, dinv, _, degree = operator_pieces(xhat)
    evals, evecs = _arpack(xt, dinv, n_components, random_state, counter)
    return evals, evecs, degree


# --------------------------------------------------------------------------
# a2: native entry point                                    embedding.rs:24-59
# --------------------------------------------------------------------------
def _select_columns(x, selected_features):
    """``to_select_elem`` + ``slice_axis(1, ..)`` (embedding.rs:36-41).

    ``None`` keeps all columns; a boolean mask or integer index array selects.
    The result is a float64 CSR (``try_convert::<CsrMatrix<f64>>``).
    """
    x = sp.csr_matrix(x)
    if selected_features is not None:
        sel = np.asarray(selected_features)
        if sel.dtype == bool:
            sel = np.flatnonzero(sel)
        x = x[:, sel]
    x = sp.csr_matrix(x, dtype=np.float64)
    x.sort_indices()
    return x


def spectral_embedding(x, selected_features, n_components, random_state,
                       feature_weights=None, counter=None, return_parts=False):
    """``spectral_embedding`` (embedding.rs:24-59).

    ``feature_weights`` (if given) is indexed by the *sliced* column index
    (:36-43, :321).  Returns ``(evals[k], evecs[n,k])``; with
    ``return_parts`` also the IDF weights and the degree vector (the two
    intermediate quantities north_star gates at 1e-5).
    """
    mat = _select_columns(x, selected_features)
    w = idf(mat) if feature_weights is None else np.asarray(feature_weights, dtype=np.float64)
    xhat = normalize(mat, w)
    evals, evecs, degree = spectral_mf(xhat, n_components, random_state, counter)
    if return_parts:
        return evals, evecs, w, degree
    return evals, evecs


# --------------------------------------------------------------------------
# Nystrom path                         embedding.rs:61-129, 194-267, 328-365
# --------------------------------------------------------------------------
def compute_degrees(x, selected_features, feature_weights):
    """``compute_degrees`` (embedding.rs:328-360): degrees of every cell under the
    full-data normalisation, ``Xhat (Xhat^T 1) - 1``."""
    xhat = normalize(_select_columns(x, selected_features), feature_weights)
    col_sum = np.asarray(xhat.sum(axis=0)).ravel()
    return xhat @ col_sum - 1.0


def nystrom(seed: sp.csr_matrix, evals, evecs, degrees, chunks):
    """``nystrom`` (embedding.rs:194-267).  ``chunks`` yields normalised row blocks."""
    evecs = evecs / np.sqrt(degrees)[:, None]          # :206-210
    evecs = evecs / evals[None, :]                     # :211-215
    proj = seed.T @ evecs                              # seed.T @ evecs                (:223)
    out = []
    for sample in chunks:
        q = sample @ proj                              # :223
        t = q.sum(axis=0) * evals                      # :224
        d = q @ t.reshape((-1, 1))                     # :225
        d[d <= 0] = np.min(d[d > 0])                   # :226
        q = q / np.sqrt(d)                             # :227
        out.append(q)
    return evals.copy(), np.vstack(out)


def orthogonalize(evals, evecs):
    """``orthogonalize`` (tools/_embedding.py:397-413)."""
    _, sigma, vt = np.linalg.svd(evecs, full_matrices=False)
    v = vt.T
    b = np.multiply(v.T, evals.reshape((1, -1))) @ v
    b = b * sigma.reshape((-1, 1)) * sigma.reshape((1, -1))
    evals_new, evecs_new = np.linalg.eig(b)
    ix = evals_new.argsort()[::-1]
    evals_new = evals_new[ix]
    evecs_new = evecs_new[:, ix]
    evecs_new = evecs_new / sigma.reshape((-1, 1))
    return evals_new, evecs @ v @ evecs_new


def spectral_embedding_nystrom(x, selected_features, n_components, landmarks, chunk_size,
                               feature_weights=None, return_parts=False):
    """``spectral_embedding_nystrom`` (embedding.rs:61-129) for a GIVEN landmark index list.

    The reference draws the landmarks with Rust's ``rand::StdRng::seed_from_u64(2023)``
    (:87-94), a stream that cannot be reproduced here; everything downstream of the draw
    is restated.  IDF weights come from all cells (:76-84), the seed rows are normalised
    with them (:95-103), ``spectral_mf(seed, k, 0)`` (:104), then ``nystrom`` over
    ``chunk_size`` row blocks of the normalised data (:106-119).
    """
    mat = _select_columns(x, selected_features)
    w = idf(mat) if feature_weights is None else np.asarray(feature_weights, dtype=np.float64)
    seed = normalize(sp.csr_matrix(mat[np.asarray(landmarks)]), w)
    v, u, d = spectral_mf(seed, n_components, 0)
    chunks = (normalize(sp.csr_matrix(mat[i:i + chunk_size]), w) for i in range(0, mat.shape[0], chunk_size))
    evals, q = nystrom(seed, v, u, d, chunks)
    if return_parts:
        return evals, q, w, d
    return evals, q


# --------------------------------------------------------------------------
# a1: Python wrapper semantics                         _embedding.py:129-295
# --------------------------------------------------------------------------
def spectral(adata, n_comps=30, features="selected", random_state=0, sample_size=None,
             distance_metric="cosine", weighted_by_sd=True, feature_weights=None,
             inplace=True):
    """Wrapper semantics of ``snap.tl.spectral`` for the full-matrix cosine path.

    Follows _embedding.py:223 (seed), :225-229 (features lookup / NameError),
    :231 (n_comps clamp with the *unselected* n_vars), :247-249 (dispatch),
    :286-289 (weighted_by_sd), :291-295 (inplace / return).
    """
    np.random.seed(random_state)
    if isinstance(features, str):
        if features in adata.var:
            features = np.asarray(adata.var[features])
        else:
            raise NameError("Please call `select_features` first or explicitly set `features = None`")
    n_comps = min(adata.n_vars - 1, adata.n_obs - 1, n_comps)
    if sample_size is not None or distance_metric != "cosine":
        raise NotImplementedError("oracle covers the full-matrix cosine path only")
    evals, evecs = spectral_embedding(adata.X, features, n_comps, random_state, feature_weights)
    if weighted_by_sd:
        keep = [i for i in range(evals.shape[0]) if evals[i] > 0]
        evals = evals[keep]
        evecs = evecs[:, keep] * np.sqrt(evals)
    if inplace:
        adata.uns["spectral_eigenvalue"] = evals
        adata.obsm["X_spectral"] = evecs
        return None
    return evals, evecs


# --------------------------------------------------------------------------
# a8: multi-view                                         embedding.rs:367-477
# --------------------------------------------------------------------------
def _frobenius_snippet(X):
    """The Python snippet embedded in ``frobenius_norm`` (embedding.rs:456-460), verbatim."""
    import numpy as np
    return np.power(X @ X.T, 2).sum()


def _frobenius_norm(xhat: sp.csr_matrix, container: str = "csr_matrix") -> float:
    """``frobenius_norm`` (embedding.rs:454-471): ``sqrt(snippet(X) - n)``.

    What the snippet computes depends on the scipy container it is handed, because
    ``np.power`` on a scipy sparse object falls through to the object's ``__pow__``:

    * ``csr_matrix`` (the ``spmatrix`` interface): ``__pow__`` is the MATRIX power, so the
      snippet returns ``sum((X X^T) @ (X X^T)) = || (X X^T) 1 ||^2``;
    * ``csr_array`` (the ``sparray`` interface): ``__pow__`` is element-wise, so the snippet
      returns ``|| X X^T ||_F^2`` -- what the function name promises.

    The reference passes ``PyArrayData::from(ArrayData::from(x))`` (embedding.rs:467);
    pyanndata is pinned to kaizhang/anndata-rs rev 0d27ac4 (snapatac2-python/Cargo.toml:22)
    and is not vendored under /root/reference.  Its CSR -> Python conversion builds
    ``scipy.sparse.csr_matrix`` (recalled from the published source of that crate; it cannot
    be checked offline), so ``container="csr_matrix"`` is the default here and in the CUDA
    path; both readings are pinned against the executed snippet
    (tests/golden/frobenius_snippet_ref.npz).
    """
    if container == "csr_matrix":
        X = sp.csr_matrix(xhat)
    elif container == "csr_array":
        X = sp.csr_array(xhat)
    else:
        raise ValueError("container must be 'csr_matrix' or 'csr_array'")
    total = float(_frobenius_snippet(X))
    return float(np.sqrt(total - xhat.shape[0]))


def multi_spectral_embedding(xs, selected_features, weights, n_components, random_state,
                             sample_rows=None, return_parts=False, container="csr_matrix"):
    """``multi_spectral_embedding`` (embedding.rs:388-452).

    Per view: slice, IDF, normalise (:404-416); Frobenius norm of the
    off-diagonal similarity on all rows if ``n <= 2000`` else on 2000 sampled
    rows (:417-421).  The reference samples with Rust ``rand 0.8``
    ``StdRng::seed_from_u64(2023)`` (:473-476), which cannot be reproduced
    outside Rust; ``sample_rows`` supplies the explicit index list instead
    (oracle and CUDA path share it -- "parity vs restatement, RNG sample not
    reference-identical").  Views are scaled by ``sqrt((w_i/norm_i)/sum)``
    (:428-442), concatenated column-wise (:443, :367-385) and handed to
    ``spectral_mf`` (:447).
    """
    mats, norms = [], []
    for x, sel in zip(xs, selected_features):
        mat = _select_columns(x, sel)
        xhat = normalize(mat, idf(mat))
        if xhat.shape[0] <= 2000:
            nrm = _frobenius_norm(xhat, container)
        else:
            if sample_rows is None:
                raise ValueError("n > 2000: pass sample_rows (the reference's Rust RNG is not reproducible here)")
            nrm = _frobenius_norm(xhat[np.asarray(sample_rows)], container)
        mats.append(xhat)
        norms.append(nrm)
    ws = [w / nrm for w, nrm in zip(weights, norms)]
    w_sum = float(sum(ws))
    scaled = [m * np.sqrt(w / w_sum) for m, w in zip(mats, ws)]
    stacked = sp.csr_matrix(sp.hstack(scaled, format="csr"))
    evals, evecs, degree = spectral_mf(stacked, n_components, random_state)
    if return_parts:
        return evals, evecs, norms, degree
    return evals, evecs


def multi_spectral(adatas, n_comps=30, features="selected", weights=None, random_state=0,
                   weighted_by_sd=True, sample_rows=None, container="csr_matrix"):
    """Wrapper semantics of ``snap.tl.multi_spectral`` (_embedding.py:483-540)."""
    np.random.seed(random_state)
    if features is None or isinstance(features, str):
        features = [features] * len(adatas)
    if all(isinstance(f, str) for f in features):
        features = [np.asarray(a.var[f]) for a, f in zip(adatas, features)]
    if weights is None:
        weights = [1.0 for _ in adatas]
    evals, evecs = multi_spectral_embedding([a.X for a in adatas], features, weights,
                                            n_comps, random_state, sample_rows=sample_rows, container=container)
    if weighted_by_sd:
        keep = [i for i in range(evals.shape[0]) if evals[i] > 0]
        evals = evals[keep]
        evecs = evecs[:, keep] * np.sqrt(evals)
    return evals, evecs


# --------------------------------------------------------------------------
# Pinning aids
# --------------------------------------------------------------------------
def matrix_free_twin(x, k, feature_weights=None):
    """The reference's own second statement of the algorithm:
    ``SpectralMatrixFree.fit`` + ``_eigen`` (_embedding.py:447-481).

    ``mat @ diags(w)``; ``s = 1/sqrt(rowsum(mat^2))``; ``X = diags(s) @ mat``;
    ``D = X @ X.sum(0).T - 1``; ``X = diags(1/sqrt(D)) @ X``;
    ``eigsh(v -> X @ (v.T @ X).T - (1/D) v, k)``; sort descending.
    """
    mat = sp.csr_matrix(x, dtype=np.float64)
    if feature_weights is not None:
        mat = mat @ sp.diags(np.asarray(feature_weights, dtype=np.float64))
    s = 1.0 / np.sqrt(np.ravel(mat.power(2).sum(axis=1)))
    xn = sp.diags(s) @ mat
    deg = np.ravel(xn @ np.asarray(xn.sum(axis=0)).T) - 1.0
    xn = sp.csr_matrix(sp.diags(1.0 / np.sqrt(deg)) @ xn)
    dinv = 1.0 / deg

    def f(v):
        return xn @ (v.T @ xn).T - dinv * v

    n = xn.shape[0]
    evals, evecs = eigsh(LinearOperator((n, n), matvec=f, dtype=np.float64), k=k)
    order = evals.argsort()[::-1]
    return evals[order], evecs[:, order], deg


def dense_check(x, k, selected_features=None, feature_weights=None):
    """Independent dense statement: ``A = D^-1/2 (S - I) D^-1/2`` with
    ``S = Xhat Xhat^T``, ``D = diag(S 1 - 1)``; ``numpy.linalg.eigh``; pick the
    ``k`` largest-|lambda| (ARPACK ``which='LM'``) and sort descending.
    Returns ``(evals, evecs, idf, degree)``.  O(n^2) memory: small n only.
    """
    mat = _select_columns(x, selected_features)
    w = idf(mat) if feature_weights is None else np.asarray(feature_weights, dtype=np.float64)
    xhat = normalize(mat, w).toarray()
    s = xhat @ xhat.T
    np.fill_diagonal(s, 0.0)
    degree = s.sum(axis=1)
    dm = 1.0 / np.sqrt(degree)
    a = dm[:, None] * s * dm[None, :]
    a = 0.5 * (a + a.T)
    ev, evec = np.linalg.eigh(a)
    pick = np.argsort(-np.abs(ev))[:k]
    pick = pick[np.argsort(-ev[pick])]
    return ev[pick], evec[:, pick], w, degree

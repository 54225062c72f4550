"""Thin object wrapper over one ``snapb200_ctx`` (one GPU).

Host-side plumbing only: numpy buffers in, numpy buffers out.  All compute
happens in the hand-written CUDA kernels behind the C ABI.
"""

from __future__ import annotations

import ctypes as C
import os

import numpy as np
import scipy.sparse as sp

from . import _lib


class Engine:
    def __init__(self, device: int | None = None):
        lib = _lib.load()
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        self._lib = lib
        self._ctx = C.c_void_p()
        _lib.check(lib.snapb200_create(int(device), C.byref(self._ctx)))
        self.device = int(device)
        self.rank, self.nranks = 0, 1
        self.n_local = self.n_global = self.row0 = self.m = 0

    # ------------------------------------------------------------ lifetime
    def close(self):
        for v in getattr(self, "_view_engines", []):
            v.close()
        self._view_engines = []
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self._lib.snapb200_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ------------------------------------------------------------ multi-GPU
    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        _lib.check(_lib.load().snapb200_comm_unique_id(buf))
        return buf.raw

    def init_comm(self, rank: int, nranks: int, unique_id: bytes):
        assert len(unique_id) == 128
        _lib.check(self._lib.snapb200_comm_init(self._ctx, int(rank), int(nranks), unique_id))
        self.rank, self.nranks = int(rank), int(nranks)

    # ------------------------------------------------------------ data in
    def load_arrays(self, indptr, indices, values, n_local, m, n_global=None, row0=0, on_device=False,
                    defer_value_scan=False):
        """Load a CSR row shard from raw arrays (numpy, or torch CUDA tensors
        with ``on_device=True``).  ``values=None`` means a binarised pattern."""
        n_global = n_local if n_global is None else n_global
        bits = lambda a: 8 * (a.itemsize if isinstance(a, np.ndarray) else a.element_size())
        vk = 0
        if values is not None:
            dt = values.dtype if isinstance(values, np.ndarray) else np.dtype(str(values.dtype).replace("torch.", ""))
            vk = _lib.value_kind(dt)
            if vk is None:
                raise TypeError(f"unsupported value dtype {dt}")
        # (the switch is per load: a context never keeps it on for a later caller)
        _lib.check(self._lib.snapb200_set_defer_value_scan(self._ctx, 1 if (defer_value_scan and not on_device) else 0))
        _lib.check(self._lib.snapb200_load_csr(
            self._ctx, int(n_local), int(n_global), int(row0), int(m),
            _lib.ptr(indptr), bits(indptr), _lib.ptr(indices), bits(indices),
            _lib.ptr(values), int(vk), 1 if on_device else 0))
        self.n_local, self.n_global, self.row0, self.m = int(n_local), int(n_global), int(row0), int(m)

    def load_csr(self, X, n_global=None, row0=0, binarized: bool | None = None, defer_value_scan: bool = False):
        """Load a scipy CSR row shard as it is: int32 or int64 index arrays and any supported value
        dtype go to the library untouched (pageable memory is fine -- the library stages them with
        a thread team, narrows 64-bit indices on the way and recognises an all-ones value array
        without shipping it).  ``binarized=True`` skips even the scan of the values (caller
        guarantees every stored entry is 1).  Rows must be sorted and duplicate-free; the device
        checks that, and a matrix that fails the check is canonicalised on the host and retried.
        ``defer_value_scan=True``: the pattern is loaded at once and the all-ones scan of the values
        runs on background threads while the caller goes on (prepare, eigsh); the caller must call
        :meth:`values_all_ones` before trusting a result and :meth:`load_values` + recompute if it
        returns False.  ``X`` must stay alive until then."""
        if not sp.issparse(X):
            X = sp.csr_matrix(np.asarray(X))
        if X.format != "csr":
            raise ValueError("X must be a CSR matrix (the reference rejects CSC input as well)")

        self._pending_values = None

        def attempt(M):
            indptr = np.ascontiguousarray(M.indptr)
            indices = np.ascontiguousarray(M.indices)
            values = None
            if not binarized:
                values = np.ascontiguousarray(M.data)
                if _lib.value_kind(values.dtype) is None:
                    values = values.astype(np.float64)
            self.load_arrays(indptr, indices, values, M.shape[0], M.shape[1], n_global, row0,
                             defer_value_scan=defer_value_scan)
            self._pending_values = values if defer_value_scan else None

        try:
            attempt(X)
        except RuntimeError as e:
            if "strictly increasing" not in str(e):
                raise
            X = X.copy()
            X.sum_duplicates()      # sorts the rows and merges duplicates (scipy's canonical format)
            attempt(X)

    def values_all_ones(self) -> bool:
        """Verdict of a deferred value scan (joins it).  True: the loaded pattern is the matrix."""
        ok = C.c_int()
        _lib.check(self._lib.snapb200_values_verdict(self._ctx, C.byref(ok)))
        return bool(ok.value)

    def load_values(self, values=None):
        """Ship the value array of the loaded pattern after all (deferred scan found values other than 1)."""
        values = self._pending_values if values is None else np.ascontiguousarray(values)
        if _lib.value_kind(values.dtype) is None:
            values = values.astype(np.float64)
        _lib.check(self._lib.snapb200_load_values(self._ctx, _lib.ptr(values), int(_lib.value_kind(values.dtype))))
        self._pending_values = None

    def load_blocks(self, blocks, m, n_global=None, row0=0, rows_hint=0, nnz_hint=0):
        """Assemble this rank's shard on the device from an iterable of scipy CSR row blocks (the
        chunks of a backed AnnData): the host never holds more than one block.  ``n_global=None``:
        the blocks are the whole matrix."""
        _lib.check(self._lib.snapb200_load_begin(self._ctx, int(m), int(rows_hint), int(nnz_hint)))
        for blk in blocks:
            if not sp.issparse(blk):
                blk = sp.csr_matrix(np.asarray(blk))
            blk = blk.tocsr()
            if blk.shape[1] != m:
                raise ValueError("every row block must have the matrix's number of columns")
            if not blk.has_sorted_indices:      # per block: cheap next to reading it from disk
                blk = blk.copy()
                blk.sum_duplicates()
            indptr = np.ascontiguousarray(blk.indptr)
            indices = np.ascontiguousarray(blk.indices)
            values = np.ascontiguousarray(blk.data)
            if _lib.value_kind(values.dtype) is None:
                values = values.astype(np.float64)
            bits = lambda a: 8 * a.itemsize
            _lib.check(self._lib.snapb200_load_append(
                self._ctx, int(blk.shape[0]), _lib.ptr(indptr), bits(indptr), _lib.ptr(indices), bits(indices),
                _lib.ptr(values), int(_lib.value_kind(values.dtype))))
        _lib.check(self._lib.snapb200_load_end(self._ctx, -1 if n_global is None else int(n_global), int(row0)))
        n, m2, _ = self.shape()
        self.n_local, self.m, self.row0 = int(n), int(m2), int(row0)
        self.n_global = self.n_local if n_global is None else int(n_global)

    def set_geometry(self, n_global, row0):
        """Place an already loaded shard inside the global matrix (block-wise loads learn their own size last)."""
        _lib.check(self._lib.snapb200_set_geometry(self._ctx, int(n_global), int(row0)))
        self.n_global, self.row0 = int(n_global), int(row0)

    def generate(self, spec, row0=0, n_local=None):
        """Synthetic planted-cluster rows generated on the device."""
        n_local = spec.n - row0 if n_local is None else n_local
        fc = np.ascontiguousarray(spec.feat_cdf, dtype=np.uint64)
        cc = np.ascontiguousarray(spec.cluster_cdf, dtype=np.uint64)
        bs = np.ascontiguousarray(spec.block_start, dtype=np.int64)
        al = np.ascontiguousarray(spec.alpha, dtype=np.uint64)
        _lib.check(self._lib.snapb200_generate(
            self._ctx, int(n_local), int(spec.n), int(row0), int(spec.m), int(spec.nnz_row),
            int(spec.n_clusters), C.c_uint64(spec.seed), _lib.ptr(fc), _lib.ptr(cc), _lib.ptr(bs), _lib.ptr(al)))
        self.n_local, self.n_global, self.row0, self.m = int(n_local), int(spec.n), int(row0), int(spec.m)

    def shape(self):
        n, m, nnz = C.c_int64(), C.c_int64(), C.c_int64()
        _lib.check(self._lib.snapb200_shape(self._ctx, C.byref(n), C.byref(m), C.byref(nnz)))
        return n.value, m.value, nnz.value

    def export_arrays(self, indptr=None, indices=None, values=None, with_values=False):
        """Copy the resident CSR back into (optionally caller-provided) host arrays."""
        n, m, nnz = self.shape()
        indptr = np.empty(n + 1, dtype=np.int64) if indptr is None else indptr
        indices = np.empty(nnz, dtype=np.int32) if indices is None else indices
        if with_values and values is None:
            values = np.empty(nnz, dtype=np.float32)
        _lib.check(self._lib.snapb200_export_csr(self._ctx, _lib.ptr(indptr), _lib.ptr(indices), _lib.ptr(values)))
        return indptr, indices, values

    def export_csr(self) -> sp.csr_matrix:
        n, m, nnz = self.shape()
        indptr, indices, values = self.export_arrays(with_values=True)
        if nnz < 2**31 - 1:
            indptr = indptr.astype(np.int32)
        return sp.csr_matrix((values, indices.astype(indptr.dtype, copy=False), indptr), shape=(n, m))

    # ------------------------------------------------------------ path
    def select_features(self, keep_mask):
        keep = np.ascontiguousarray(np.asarray(keep_mask, dtype=bool).astype(np.uint8))
        _lib.check(self._lib.snapb200_select_features(self._ctx, _lib.ptr(keep), int(keep.shape[0])))
        self.m = int(keep.sum())

    def set_feature_weights(self, w):
        if w is None:
            _lib.check(self._lib.snapb200_set_feature_weights(self._ctx, None, 0))
        else:
            w = np.ascontiguousarray(w, dtype=np.float64)
            _lib.check(self._lib.snapb200_set_feature_weights(self._ctx, _lib.ptr(w), int(w.shape[0])))

    def prepare(self, want_outputs=True):
        """a3-a5.  Returns ``(idf[m], degree[n_local])`` (or None, None)."""
        idf = np.empty(self.m, dtype=np.float64) if want_outputs else None
        deg = np.empty(self.n_local, dtype=np.float64) if want_outputs else None
        _lib.check(self._lib.snapb200_prepare(self._ctx, _lib.ptr(idf), _lib.ptr(deg)))
        return idf, deg

    def view_norms(self):
        """IDF weights and IDF-weighted row norms of the loaded view (multi_spectral)."""
        idf = np.empty(self.m, dtype=np.float64)
        rho = np.empty(self.n_local, dtype=np.float64)
        _lib.check(self._lib.snapb200_view_norms(self._ctx, _lib.ptr(idf), _lib.ptr(rho)))
        return idf, rho

    # ------------------------------------------------------------ multi-view (virtual hstack)
    def attach_view(self, view: "Engine"):
        """``view`` adopts this engine's stream and communicator (multi_spectral); this engine must
        outlive it."""
        _lib.check(self._lib.snapb200_attach_view(self._ctx, view._ctx))
        view.rank, view.nranks = self.rank, self.nranks

    def view_frobenius(self, sample_rows_local) -> float:
        """Value of the reference's ``frobenius_norm`` snippet on the sampled unit rows (csr_matrix
        reading: ``sum((X X^T) @ (X X^T))``); collective over the row shards."""
        rows = np.ascontiguousarray(sample_rows_local, dtype=np.int64)
        out = C.c_double()
        _lib.check(self._lib.snapb200_view_frobenius(self._ctx, _lib.ptr(rows) if rows.size else None, int(rows.size), C.byref(out)))
        return out.value

    def combine_views(self, views: "list[Engine]", scales, want_degree=False):
        """Chain the prepared ``views`` (``views[0]`` is this engine) with the view scales ``c_v``."""
        assert views and views[0] is self
        arr = (C.c_void_p * len(views))(*[v._ctx for v in views])
        cv = np.ascontiguousarray(scales, dtype=np.float64)
        deg = np.empty(self.n_local, dtype=np.float64) if want_degree else None
        _lib.check(self._lib.snapb200_combine_views(self._ctx, arr, _lib.ptr(cv), len(views), _lib.ptr(deg)))
        return deg

    def get_vector(self, which: str) -> np.ndarray:
        code = {"weights": 0, "rho": 1, "degree": 2, "colsum": 3}[which]
        out = np.empty(self.m if code in (0, 3) else self.n_local, dtype=np.float64)
        _lib.check(self._lib.snapb200_get_vector(self._ctx, code, _lib.ptr(out)))
        return out

    def gather_rows_into(self, rows_local, dst: "Engine", n_global=None, row0=0):
        """``dst`` <- the rows ``rows_local`` of this engine's resident matrix (device to device)."""
        rows = np.ascontiguousarray(rows_local, dtype=np.int64)
        n_global = rows.size if n_global is None else n_global
        _lib.check(self._lib.snapb200_gather_rows(self._ctx, _lib.ptr(rows) if rows.size else None, int(rows.size), dst._ctx,
                                                  int(n_global), int(row0)))
        dst.n_local, dst.n_global, dst.row0, dst.m = int(rows.size), int(n_global), int(row0), self.m

    def prepare_projection(self, want_outputs=True):
        """What :meth:`project` needs (weights, row norms, cell-major tiled copy), without the
        transpose.  Returns ``(w[m], rho[n_local])`` (or None, None)."""
        w = np.empty(self.m, dtype=np.float64) if want_outputs else None
        rho = np.empty(self.n_local, dtype=np.float64) if want_outputs else None
        _lib.check(self._lib.snapb200_prepare_projection(self._ctx, _lib.ptr(w), _lib.ptr(rho)))
        return w, rho

    def project(self, M):
        """``Xhat @ M`` for the feature-weighted, row-normalised matrix (M: m x k) -> n_local x k."""
        M = np.ascontiguousarray(M, dtype=np.float32)
        assert M.ndim == 2 and M.shape[0] == self.m
        out = np.empty((self.n_local, M.shape[1]), dtype=np.float32)
        _lib.check(self._lib.snapb200_project(self._ctx, 0, _lib.ptr(M), int(M.shape[1]), _lib.ptr(out)))
        return out

    def project_t(self, U):
        """``Xhat.T @ U`` (U: n_local x k) -> m x k, summed over the row shards."""
        U = np.ascontiguousarray(U, dtype=np.float32)
        assert U.ndim == 2 and U.shape[0] == self.n_local
        out = np.empty((self.m, U.shape[1]), dtype=np.float32)
        _lib.check(self._lib.snapb200_project(self._ctx, 1, _lib.ptr(U), int(U.shape[1]), _lib.ptr(out)))
        return out

    def project_t_ones(self):
        """``Xhat.T @ 1`` of the loaded block (streamed Nystrom degree pass, on a context outside the
        communicator): the column sums of the normalised rows, fp32 through the SpMM path."""
        ones = np.ones((self.n_local, 1), dtype=np.float32)
        out = np.empty((self.m, 1), dtype=np.float32)
        _lib.check(self._lib.snapb200_project(self._ctx, 1, _lib.ptr(ones), 1, _lib.ptr(out)))
        return out[:, 0].astype(np.float64)

    def operator_apply(self, V):
        V = np.ascontiguousarray(V, dtype=np.float32)
        assert V.ndim == 2 and V.shape[0] == self.n_local
        Y = np.empty_like(V)
        _lib.check(self._lib.snapb200_operator_apply(self._ctx, _lib.ptr(V), _lib.ptr(Y), int(V.shape[1])))
        return Y

    def operator_time(self, b=8, iters=5, flush_l2=True):
        p1, cm, p2 = C.c_double(), C.c_double(), C.c_double()
        _lib.check(self._lib.snapb200_operator_time(self._ctx, int(b), int(iters), 1 if flush_l2 else 0,
                                                    C.byref(p1), C.byref(cm), C.byref(p2)))
        return p1.value, cm.value, p2.value

    def eigsh(self, k, seed=0, tol=0.0, block=0, max_basis=0, max_ops=0, out_evecs=None, scale_by_sqrt_eval=False):
        evals = np.empty(k, dtype=np.float64)
        evecs = np.empty((self.n_local, k), dtype=np.float64) if out_evecs is None else out_evecs
        _lib.check(self._lib.snapb200_eigsh(self._ctx, int(k), int(seed), float(tol), int(block),
                                            int(max_basis), int(max_ops), _lib.ptr(evals), _lib.ptr(evecs),
                                            1 if scale_by_sqrt_eval else 0))
        st = self.stats()
        if not st["converged"]:
            # scipy's eigsh, which the reference calls (embedding.rs:166-167), raises here as well
            from scipy.sparse.linalg import ArpackNoConvergence
            raise ArpackNoConvergence(
                f"block Lanczos: no convergence after {st['n_ops']} block operator applications "
                f"(worst residual {st['max_residual']:.3e})", evals, evecs)
        return evals, evecs

    def stats(self) -> dict:
        s = _lib.Stats()
        _lib.check(self._lib.snapb200_get_stats(self._ctx, C.byref(s)))
        return s.as_dict()

    def set_spmm_mode(self, mode: str | int):
        """'auto' | 'csr' (gather out of L2) | 'tiled' (shared-memory sliced-ELL)."""
        code = {"auto": 0, "csr": 1, "tiled": 2}.get(mode, mode)
        _lib.check(self._lib.snapb200_set_spmm_mode(self._ctx, int(code)))

    def set_block(self, block: int):
        """Default Lanczos block width (4, 8 or 16)."""
        _lib.check(self._lib.snapb200_set_block(self._ctx, int(block)))

    def stream_handle(self) -> int:
        """Raw ``cudaStream_t`` of the context (for torch.cuda.ExternalStream)."""
        h = C.c_void_p()
        _lib.check(self._lib.snapb200_get_stream(self._ctx, C.byref(h)))
        return int(h.value or 0)

    def knn(self, points, n_neighbors, q0=0, nq=None):
        """Exact neighbour graph of ``points`` (n x d float64): for the queries ``points[q0:q0+nq]`` the
        ``K = min(n_neighbors, n - 1)`` nearest other points.  Returns ``(indices int32, distances float64)``,
        both ``nq x K`` with every row sorted by index (``snapb200_knn``; knn.rs:9-33)."""
        P = np.ascontiguousarray(points, dtype=np.float64)
        if P.ndim != 2:
            raise ValueError("points must be a 2-d array")
        n, d = P.shape
        nq = n - q0 if nq is None else int(nq)
        K = max(0, min(int(n_neighbors), n - 1))
        idx = np.empty((nq, K), dtype=np.int32)
        dist = np.empty((nq, K), dtype=np.float64)
        _lib.check(self._lib.snapb200_knn(self._ctx, n, d, _lib.ptr(P), 0, int(q0), nq, int(n_neighbors),
                                          _lib.ptr(idx), _lib.ptr(dist)))
        return idx, dist

    def ortho_selftest(self, n=5000, ncols=64, block=4) -> float:
        err = C.c_double()
        _lib.check(self._lib.snapb200_ortho_selftest(self._ctx, int(n), int(ncols), int(block), C.byref(err)))
        return err.value

    def dense_selftest(self, n=4099, ncq=136, p=30) -> float:
        err = C.c_double()
        _lib.check(self._lib.snapb200_dense_selftest(self._ctx, int(n), int(ncq), int(p), C.byref(err)))
        return err.value


def sym_eig(a: np.ndarray):
    """Host Rayleigh-Ritz eigensolver of the library (no GPU needed)."""
    a = np.array(a, dtype=np.float64, order="C")
    n = a.shape[0]
    w = np.empty(n, dtype=np.float64)
    _lib.check(_lib.load().snapb200_sym_eig(n, _lib.ptr(a), _lib.ptr(w)))
    return w, a

/*
 * snapb200.h -- C ABI of the B200-native matrix-free spectral embedding.
 *
 * This is the drop-in boundary for the native entry points the reference binds
 * through PyO3 (all paths relative to the SnapATAC2 source tree):
 *
 *   spectral_embedding(anndata, selected_features, n_components, random_state,
 *                      feature_weights=None) -> (evals f64[k], evecs f64[n,k])
 *       snapatac2-python/src/embedding.rs:24-59   (registered src/lib.rs:79)
 *   multi_spectral_embedding(anndata[], selected_features[], weights[],
 *                      n_components, random_state)
 *       snapatac2-python/src/embedding.rs:388-452 (registered src/lib.rs:80)
 *
 * The reference does everything behind one call; here the same work is split
 * into load -> prepare -> eigsh so that the intermediate quantities north_star
 * gates (IDF weights, degree vector) and the operator alone can be checked
 * and profiled.  The Python mirror `snapatac2_b200.tl.spectral` strings them
 * together behind the reference's signature (tools/_embedding.py:129-141).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; the message
 *     is available from snapb200_last_error() (thread-local);
 *   - plain pointers and sizes only; the caller owns every host pointer, the
 *     library owns all device memory inside the context;
 *   - one context = one GPU = one host thread; multi-GPU is one process (or
 *     thread) per GPU, joined by snapb200_comm_init (NCCL over NVLink);
 *   - rows (cells) are sharded: a context holds global rows
 *     [row0, row0 + n_local) of an n_global x m matrix;
 *   - there is no CPU fallback: every entry point needs a CUDA device.
 */
#ifndef SNAPB200_H
#define SNAPB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct snapb200_ctx snapb200_ctx;

/* Phase timings and solver diagnostics of the last prepare/eigsh calls. */
typedef struct snapb200_stats {
    double ms_load;        /* host->device copy + format conversion           */
    double ms_transpose;   /* building the feature-major copy                  */
    double ms_prepare;     /* IDF, row norms, column sums, degrees (a3-a5)     */
    double ms_eigsh;       /* whole eigensolve (a6+a7)                         */
    double ms_spmm;        /* time inside the two SpMM passes                  */
    double ms_ortho;       /* Gram / projection / CholQR kernels               */
    double ms_comm;        /* NCCL all-reduces                                 */
    double ms_host;        /* host-side Rayleigh-Ritz                          */
    double max_residual;   /* max ||A v - theta v|| estimate over the k pairs  */
    int64_t n_ops;         /* block operator applications                      */
    int64_t n_restarts;    /* thick restarts                                   */
    int64_t basis_cols;    /* Krylov basis width at exit                       */
    int64_t block;         /* block width b                                    */
    int64_t nnz_local;     /* stored entries on this shard                     */
    int64_t kernel_launches; /* CUDA kernels launched by the library so far    */
    double ms_format;      /* building the shared-memory tiled (sliced-ELL) copies */
    int64_t spmm_tiled;    /* 1 if the last operator ran the tiled kernels     */
    double ms_prepare_wall; /* host wall clock of the whole prepare call        */
    double ms_pool;        /* host time spent in cudaMalloc/cudaFree by the caching allocator (cumulative) */
    int64_t pool_mallocs;  /* driver allocations so far (cumulative; a steady state adds none)             */
    int64_t converged;     /* last eigsh: 1 = every pair met the tolerance (or the Krylov space was exhausted),
                              0 = max_ops reached first (scipy's eigsh raises ArpackNoConvergence there)   */
    int64_t n_spec_ops;    /* operator applications enqueued ahead of the convergence check and discarded  */
    double ms_d2h;         /* last eigsh: final rotation + copy of the eigenvectors to the host            */
    int64_t bytes_h2d;     /* last load_csr: bytes that crossed PCIe (indptr + int32 indices [+ f32 values]) */
    int64_t host_threads;  /* last load_csr: size of the host staging team                                 */
    int64_t fused_allreduce; /* last eigsh: 1 = the small fp64 all-reduces of the block step ran inside the Gram
                              kernels over NVLink peer memory (csrc/peer.cuh), 0 = NCCL (or a single rank)   */
    int64_t bytes_h2d_indices; /* last index transfer: bytes that crossed PCIe for the column indices (2 per entry when
                              delta-encoded, see csrc/ingest.cu; 4 otherwise)                               */
    double ms_knn;         /* last knn: device time of centring + filter/exact scan (CUDA events)          */
    double ms_knn_wall;    /* last knn: wall clock of the call, uploads and the copy of the result included */
    int64_t knn_mma;       /* last knn: 1 = the filter ran on the tensor cores (mma.sync tf32 x 3, SNAPB200_KNN_MMA) */
} snapb200_stats;

/* Library / error plumbing. */
const char* snapb200_last_error(void);
int  snapb200_version(void);

/* Context: binds `device` (cudaSetDevice ordinal) and a private stream. */
int  snapb200_create(int device, snapb200_ctx** out);
int  snapb200_destroy(snapb200_ctx* ctx);

/* Multi-GPU (one rank per context).  `id` is an opaque 128-byte NCCL unique
 * id produced on rank 0 and distributed by the caller (torch.distributed,
 * MPI, a file ...).  Without comm_init the context is a single-rank job.
 * comm_init also sets up CUDA-IPC mailboxes in every rank's memory for the small all-reduces that are
 * fused into the eigensolver's Gram kernels (plain stores over NVLink); if that is not possible
 * (ranks in one process, no peer access, SNAPB200_NO_PEER set) those go through NCCL. */
int  snapb200_comm_unique_id(char id[128]);
int  snapb200_comm_init(snapb200_ctx* ctx, int rank, int nranks, const char id[128]);

/* Load this rank's row shard of X as CSR (replaces the slice + try_convert of
 * embedding.rs:36-41).  `indptr` has n_local+1 entries (relative to the
 * shard), `indices` are column ids < m, sorted within a row.  *_bits is 32 or
 * 64.  `values` may be NULL (binarised pattern: every stored entry is 1);
 * otherwise value_kind selects 1=f32, 2=f64, 3=u32, 4=i32, 5=i64, 6=u64,
 * 7=u8 (also numpy bool), 8=i8, 9=u16, 10=i16 and the values are converted to f32; an
 * all-ones value array is recognised (host arrays: by a threaded scan before anything is
 * shipped) and dropped.  Host arrays may be pageable: a team of host threads
 * (SNAPB200_THREADS, default min(16, cores / local ranks)) narrows 64-bit indices to 32 bits
 * into a ring of pinned buffers and issues the DMAs chunk by chunk.  `on_device` != 0 means the pointers are
 * device pointers on this context's GPU (e.g. torch CUDA tensors). */
int  snapb200_load_csr(snapb200_ctx* ctx, int64_t n_local, int64_t n_global, int64_t row0,
                       int64_t m, const void* indptr, int indptr_bits,
                       const void* indices, int indices_bits,
                       const void* values, int value_kind, int on_device);

/* Deferred value scan.  With set_defer_value_scan(1), snapb200_load_csr on host arrays loads the pattern only
 * and starts the scan of the value array on background threads, so that the check "is every stored
 * value 1?" overlaps the GPU's work on the matrix instead of preceding it (C3: 19.6 GB of values, ~140 ms).
 * The caller must keep the value array alive and ask for the verdict before it trusts any result:
 * values_verdict joins the scan (*all_ones = 1: the pattern-only result stands); otherwise it ships the
 * values with load_values (f32 on the device; invalidates prepare) and repeats prepare / eigsh. */
int  snapb200_set_defer_value_scan(snapb200_ctx* ctx, int on);
int  snapb200_values_verdict(snapb200_ctx* ctx, int* all_ones);
int  snapb200_load_values(snapb200_ctx* ctx, const void* values, int value_kind);

/* The same load, block by block: the matrix arrives as a sequence of CSR row blocks (the chunks a
 * backed AnnData yields -- the reference iterates `chunked_X`, embedding.rs:76-84 -- or any iterator
 * of scipy CSR blocks) and is assembled on the device, so the host never holds more than one block.
 * load_begin(m, hints) -> load_append(block)... -> load_end(n_global, row0); n_global < 0 means the
 * blocks are the whole matrix.  A block's indptr may start anywhere (differences are used); index and
 * value conventions as for snapb200_load_csr (host arrays only). */
int  snapb200_load_begin(snapb200_ctx* ctx, int64_t m, int64_t rows_hint, int64_t nnz_hint);
int  snapb200_load_append(snapb200_ctx* ctx, int64_t n_rows, const void* indptr, int indptr_bits,
                          const void* indices, int indices_bits, const void* values, int value_kind);
int  snapb200_load_end(snapb200_ctx* ctx, int64_t n_global, int64_t row0);
/* Place an already loaded shard inside the global matrix (a block-wise load learns its own row count
 * last; under multi-GPU the global count follows from an exchange of the shard sizes). */
int  snapb200_set_geometry(snapb200_ctx* ctx, int64_t n_global, int64_t row0);

/* Column selection on the device (to_select_elem + slice_axis(1, ..),
 * embedding.rs:36-39): keep[j] != 0 keeps column j; surviving columns are
 * renumbered densely.  Must be called between load and prepare. */
int  snapb200_select_features(snapb200_ctx* ctx, const uint8_t* keep, int64_t m);

/* Synthetic planted-cluster pattern rows generated on the device
 * (SURVEY.md 8d); bit-identical to snapatac2_b200.synth.generate_rows for the
 * same tables.  feat_cdf has m+1 entries, cluster_cdf/block_start K+1,
 * alpha K (32.32 fixed point, see synth.py). */
int  snapb200_generate(snapb200_ctx* ctx, int64_t n_local, int64_t n_global, int64_t row0,
                       int64_t m, int nnz_row, int n_clusters, uint64_t seed,
                       const uint64_t* feat_cdf, const uint64_t* cluster_cdf,
                       const int64_t* block_start, const uint64_t* alpha);

/* Shard geometry and export of the loaded/generated CSR (tests, e2e bench). */
int  snapb200_shape(snapb200_ctx* ctx, int64_t* n_local, int64_t* m, int64_t* nnz_local);
int  snapb200_export_csr(snapb200_ctx* ctx, int64_t* indptr, int32_t* indices, float* values_or_null);

/* User feature weights (embedding.rs:42-43); NULL restores IDF.  Length m
 * (the *selected* feature count). */
int  snapb200_set_feature_weights(snapb200_ctx* ctx, const double* w, int64_t m);

/* a3-a5: IDF weights (embedding.rs:269-286), row L2 norms (:315-326), column
 * sums and degrees with self-similarity removed (:139-152).  Also builds the
 * feature-major copy used by pass 1.  Outputs may be NULL; idf_out has m
 * entries, degree_out n_local (degree = X c - 1, i.e. 1/dinv). */
int  snapb200_prepare(snapb200_ctx* ctx, double* idf_out, double* degree_out);

/* Multi-view support (multi_spectral_embedding, embedding.rs:398-416): the
 * per-view statistics of the loaded (and column-selected) view -- its IDF
 * weights (idf_out, m entries) and the L2 norms of its IDF-weighted rows
 * (rho_out, n_local entries).  The caller scales and concatenates the views
 * (embedding.rs:428-443) and runs the ordinary load/prepare/eigsh on the result. */
int  snapb200_view_norms(snapb200_ctx* ctx, double* idf_out, double* rho_out);

/* Multi-view embedding as a VIRTUAL column concatenation (multi_spectral_embedding,
 * embedding.rs:388-452; hstack :367-385): one context per view on the same GPU.
 *   attach_view    the view context adopts the main context's stream and communicator (call before
 *                  loading the view; the main context must outlive its views);
 *   [load_csr / select_features / prepare on every view context -- the ordinary single-view calls]
 *   view_frobenius the value of the Python snippet inside frobenius_norm (embedding.rs:456-460) on the
 *                  unit-norm rows `sample_rows` (this rank's local ids) of a prepared view, as the
 *                  snippet evaluates on a scipy.sparse.csr_matrix: sum((X X^T) @ (X X^T)); the
 *                  caller forms norm = sqrt(value - n_sample) and the view scales
 *                  c_v = sqrt((weight_v / norm_v) / sum) (embedding.rs:423-442);
 *   combine_views  views[0] = main; combined degrees d = sum_v c_v^2 (d_v + 1) - 1, D^-1, the
 *                  trivial eigenvector and every view's operator row scale; afterwards
 *                  snapb200_eigsh / operator_apply on the main context use
 *                  A = sum_v X~_v X~_v^T - D^-1 without the concatenated matrix ever existing.
 * get_vector copies a prepared context's fp64 vectors to the host: which = 0 feature weights (m),
 * 1 row norms (n_local), 2 degrees (n_local), 3 column sums (m). */
int  snapb200_attach_view(snapb200_ctx* main_ctx, snapb200_ctx* view);
int  snapb200_view_frobenius(snapb200_ctx* ctx, const int64_t* sample_rows, int64_t n_sample_local, double* snippet_sum);
int  snapb200_combine_views(snapb200_ctx* main_ctx, snapb200_ctx** views, const double* view_scale, int n_views,
                            double* degree_out);
int  snapb200_get_vector(snapb200_ctx* ctx, int which, double* out);

/* dst <- the rows `rows` (local row ids of src's shard, any order) of src's resident matrix; dst
 * becomes rows [row0_dst, row0_dst + n_rows) of an n_global_dst-row matrix with the same columns.
 * Both contexts on the same GPU.  The Nystrom path takes its landmark rows this way
 * (select_axis(0, ..), embedding.rs:95-99). */
int  snapb200_gather_rows(snapb200_ctx* src, const int64_t* rows, int64_t n_rows, snapb200_ctx* dst,
                          int64_t n_global_dst, int64_t row0_dst);

/* Nystrom extension (spectral_embedding_nystrom / nystrom, embedding.rs:61-129, 194-267): products
 * with the feature-weighted, row-normalised matrix  Xhat = diag(1/rho) P diag(w)  on k dense
 * columns (row-major f32 host buffers):
 *   transposed == 0:  out[n_local x k] = Xhat   in[m x k]       -- "sample @ (...)" for every cell
 *   transposed != 0:  out[m x k]       = Xhat^T in[n_local x k] -- "seed.T @ evecs" (summed over the
 *                                                                  row shards)
 * prepare_projection computes what the non-transposed product needs (IDF or user weights, row
 * norms, the cell-major tiled copy) without the transpose; it is implied by the first project call.
 * Its outputs may be NULL (idf_out: m doubles, rho_out: n_local doubles). */
int  snapb200_prepare_projection(snapb200_ctx* ctx, double* idf_out, double* rho_out);
int  snapb200_project(snapb200_ctx* ctx, int transposed, const float* in, int k, float* out);

/* a6 alone: Y = X~ (X~^T V) - dinv .* V on b vectors (embedding.rs:162-163).
 * V and Y are host, row-major n_local x b, b in {4, 8, 16}. */
int  snapb200_operator_apply(snapb200_ctx* ctx, const float* V, float* Y, int b);

/* Device-only timing of `iters` operator applications on resident random
 * vectors (for the roofline line of bench.py); returns average milliseconds
 * of pass 1 (X^T V), the all-reduce, and pass 2 (X W). */
int  snapb200_operator_time(snapb200_ctx* ctx, int b, int iters, int flush_l2,
                            double* ms_pass1, double* ms_comm, double* ms_pass2);

/* a7: k largest-magnitude eigenpairs of the normalised similarity operator,
 * sorted by descending eigenvalue, trivial pair (lambda = 1) included -- what
 * scipy eigsh(which='LM') + argsort()[::-1] returns at embedding.rs:164-171.
 * Block Lanczos with full re-orthogonalisation and thick restart.
 * evals: k doubles.  evecs: n_local x k doubles, row-major (this rank's rows).
 * tol <= 0 selects the default (1e-5 relative residual); block 0 -> default.
 * scale_by_sqrt_eval != 0: eigenvector columns whose eigenvalue is positive are returned
 * multiplied by sqrt(eigenvalue) -- the wrapper's weighted_by_sd step
 * (tools/_embedding.py:286-289) folded into the final basis rotation on the device.
 * The stats field `converged` tells whether the tolerance was met (scipy raises
 * ArpackNoConvergence otherwise; the Python mirror does the same). */
int  snapb200_eigsh(snapb200_ctx* ctx, int k, int64_t seed, double tol,
                    int block, int max_basis, int max_ops,
                    double* evals, double* evecs, int scale_by_sqrt_eval);

int  snapb200_get_stats(snapb200_ctx* ctx, snapb200_stats* out);

/* SpMM kernel selection: 0 = automatic (tiled for >= 2^25 stored entries and
 * b = 4 or 8), 1 = CSR gather out of L2, 2 = shared-memory tiled sliced-ELL
 * (16-bit tile-local entries ordered by a padded class rotation so that the
 * lanes of a quarter warp read distinct shared-memory bank groups).
 * Takes effect at the next prepare. */
int  snapb200_set_spmm_mode(snapb200_ctx* ctx, int mode);

/* Default Lanczos block width b (4, 8 or 16; initially 4).  prepare() builds
 * the tiled copies for it and eigsh(block = 0) uses it. */
int  snapb200_set_block(snapb200_ctx* ctx, int block);

/* The context's CUDA stream (a cudaStream_t), so a caller can record its own
 * events around library calls (bench.py wraps it in torch.cuda.ExternalStream). */
int  snapb200_get_stream(snapb200_ctx* ctx, void** stream);

/* Test hooks (no reference counterpart).  dense_selftest runs the FP64
 * tensor-core Gram / projection / rotation kernels against plain fp64 loops on
 * random data and returns the worst relative error; sym_eig is the host-side
 * Rayleigh-Ritz eigensolver (a: n x n row-major, destroyed; eigenvalues
 * ascending in w, eigenvectors in the columns of a) and needs no GPU. */
int  snapb200_dense_selftest(snapb200_ctx* ctx, int64_t n, int ncq, int p, double* max_rel_err);
/* Exact k-nearest-neighbour graph of an embedding: replaces `nearest_neighbour_graph`
 * (snapatac2-core/src/utils/knn.rs:9-33, bound as `internal.nearest_neighbour_graph(data, k)` at
 * snapatac2-python/src/knn.rs:8-16 and called from preprocessing/_knn.py:80).  `points` is the n x d float64
 * row-major matrix (host memory, or device memory when on_device != 0; d <= 64).  The queries are the points
 * [q0, q0 + nq) -- a rank of a row-sharded run passes all points and its own row range.  Per query the
 * K = min(k, n - 1) nearest other points (Euclidean, the query's own index excluded, ties at the K-th distance
 * broken by the smaller index): out_indices / out_distances are nq x K host arrays, each row sorted by index,
 * i.e. the `indices` / `data` of the reference's CSR result with indptr = K * arange(nq + 1).  Distances are
 * bit-identical to sqrt(squared_euclidean) of the reference's kd-tree crate (left-to-right float64 sum).
 * K <= 100 (74 when d > 32). */
int  snapb200_knn(snapb200_ctx* ctx, int64_t n, int d, const double* points, int on_device, int64_t q0, int64_t nq,
                  int k, int32_t* out_indices, double* out_distances);
int  snapb200_knn_limits(int* max_neighbors, int* max_dim);

/* Host-only replay of the delta encoding the index transfer uses (csrc/ingest.cu: 2 bytes per stored
 * entry over PCIe, decoded by a kernel): encodes `count` indices chunk by chunk and decodes them with a
 * scalar loop.  Returns 0 = identical, 1 = a chunk would fall back to plain int32, -1 = mismatch, -2 = bad
 * arguments.  No device work; test infrastructure for the CPU suite. */
int  snapb200_delta_selftest_host(const void* indices, int index_bits, int64_t count, int64_t* n_side);

/* ortho_selftest builds an orthonormal basis of `ncols` columns from random blocks of width `block`
 * with the eigensolver's fused orthogonalisation kernels and returns max |Q^T Q - I|. */
int  snapb200_ortho_selftest(snapb200_ctx* ctx, int64_t n, int ncols, int block, double* max_err);
int  snapb200_sym_eig(int n, double* a, double* w);

#ifdef __cplusplus
}
#endif
#endif /* SNAPB200_H */

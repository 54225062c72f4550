// Context and device-side data structures of libsnapb200.
#pragma once

#include "common.cuh"
#include "../../include/snapb200.h"

#include <atomic>
#include <thread>
#include <vector>

namespace snapb {

// Caching device allocator (pool.cu): stream-ordered reuse of freed blocks.
void* pool_alloc(size_t bytes);
void pool_free(void* p);
void pool_trim();
void pool_set_stream(cudaStream_t s);
void pool_forget_stream(cudaStream_t s);   // before cudaStreamDestroy
void pool_counters(long long* mallocs, long long* reuses, long long* trims, double* ms);

// Owning device buffer (released to the pool with the context or on reassignment).
template <typename T>
struct DevBuf {
    T* p = nullptr;
    int64_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) pool_free(p);
        p = nullptr;
        n = 0;
    }
    void alloc(int64_t count) {
        release();
        if (count > 0) p = static_cast<T*>(pool_alloc(sizeof(T) * static_cast<size_t>(count)));
        n = count;
    }
    void ensure(int64_t count) {
        if (count > n) alloc(count);
    }
    void swap(DevBuf& o) {
        T* tp = p; p = o.p; o.p = tp;
        int64_t tn = n; n = o.n; o.n = tn;
    }
};

template <typename T>
struct PinBuf {
    T* p = nullptr;
    int64_t n = 0;
    PinBuf() = default;
    PinBuf(const PinBuf&) = delete;
    PinBuf& operator=(const PinBuf&) = delete;
    ~PinBuf() { if (p) cudaFreeHost(p); }
    void ensure(int64_t count) {
        if (count <= n) return;
        if (p) cudaFreeHost(p);
        p = nullptr;
        SB_CUDA(cudaMallocHost(reinterpret_cast<void**>(&p), sizeof(T) * static_cast<size_t>(count)));
        n = count;
    }
};

// CSR pattern (+ optional f32 values) resident on the device.
struct Csr {
    int64_t nrows = 0, ncols = 0, nnz = 0;
    DevBuf<int64_t> ptr;   // nrows + 1
    DevBuf<int32_t> idx;   // nnz, sorted within a row
    DevBuf<float> val;     // nnz or empty (binarised)
    bool has_values() const { return val.p != nullptr; }
    void clear() {
        nrows = ncols = nnz = 0;
        ptr.release();
        idx.release();
        val.release();
    }
};

// Tile-major transpose of the pattern (prep.cu: transpose_tiled), the input of the feature-major
// tiled copy.  The local cells are cut into tiles of `tile_rows` (= the column tile of that copy);
// tile t stores, feature by feature, the tile-local ids (16 bit, ascending) of its cells that have
// the feature: segment (t, j) = ids[tile_base[t] + segoff[t*m + j] .. + cnt[t*m + j]).
struct TileT {
    int tile_rows = 0, n_tiles = 0;
    int64_t m = 0, nnz = 0;
    DevBuf<uint16_t> cnt;        // n_tiles * m   segment lengths
    DevBuf<uint32_t> segoff;     // n_tiles * m   segment starts relative to the tile
    DevBuf<int64_t> tile_base;   // n_tiles + 1
    DevBuf<uint16_t> ids;        // nnz
    DevBuf<float> vals;          // nnz or empty
    bool built = false;
    void clear() {
        cnt.release(); segoff.release(); tile_base.release(); ids.release(); vals.release();
        built = false; nnz = 0; n_tiles = 0;
    }
};

// Column-tiled sliced-ELL copy of a CSR matrix for the shared-memory SpMM
// (sell_build.cu / spmm_tiled.cu).  Columns are cut into tiles of `tile_cols`
// columns so that a tile of the dense operand (tile_cols x b floats) fits in
// shared memory; rows into `n_windows` windows of at most kSellWindowRows rows.
// The chunk list is TILE-MAJOR: tile 0 of every window, then tile 1, ...  Inside
// one (tile, window) the row segments are sorted by length and cut into chunks
// of 32 (one lane per row segment; every row appears in every tile, possibly
// with an empty segment, so each (tile, row) partial result is written exactly
// once).  An entry is the 16-bit tile-local column (0xFFFF = empty slot): 2
// bytes per stored entry.  A chunk stores its entries interleaved, eight per
// lane at a time (one "group" = 32 lanes x 16 bytes), so a warp reads 512
// contiguous bytes per step.  Within a lane the entries are ordered so that the
// eight lanes of a quarter warp hit eight different 16-byte bank groups (see
// sell_build.cu).
struct Sell {
    int b = 0;                                // dense block width the tiles are sized for (4 or 8)
    int n_windows = 0, n_tiles = 0, tile_cols = 0;
    int64_t chunks_per_tile = 0;              // the same for every tile
    int64_t nrows = 0, ncols = 0;
    int64_t n_chunks = 0, n_entries = 0;      // n_entries counts slots (entries + empty)
    DevBuf<int64_t> window_start;             // n_windows + 1 row boundaries
    DevBuf<int64_t> window_chunk0;            // n_windows + 1 first chunk of a window inside a tile
    DevBuf<int32_t> chunk_rows;               // n_chunks * 32 (global row id, -1 = none)
    DevBuf<int32_t> chunk_span;               // n_chunks * 32: a lane of a split row covers pieces [lo, lo + cnt) of its
                                              //   segment: lo | cnt << 16; 0xFFFF in the high half = the whole segment
    DevBuf<int32_t> chunk_groups;             // n_chunks   (groups of 8 slots per lane)
    DevBuf<int64_t> chunk_off;                // n_chunks + 1, in groups (256 slots)
    DevBuf<uint16_t> data;                    // n_entries
    DevBuf<float> vals;                       // n_entries or empty
    bool built = false;
    void clear() {
        window_start.release(); window_chunk0.release(); chunk_rows.release(); chunk_span.release(); chunk_groups.release();
        chunk_off.release(); data.release(); vals.release();
        built = false; n_chunks = n_entries = 0;
    }
};
constexpr int kSellTileBytes = 196608;     // 192 KB of shared memory: 6144 rows (b=8) / 12288 rows (b=4)
constexpr int kSellWindowRows = 8192;      // rows sorted together (one CTA-wide sort)

struct Comm;     // NCCL wrapper (comm.cu)
struct PeerBox;  // mailboxes of the fused small all-reduce over peer memory (peer.cuh)

}  // namespace snapb

// The public opaque type.
struct snapb200_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;    // the context's stream: every call is ordered on it
    int num_sms = snapb::kNumSMsB200;

    // multi-GPU
    snapb::Comm* comm = nullptr;
    int rank = 0, nranks = 1;

    // shard geometry
    int64_t n_local = 0, n_global = 0, row0 = 0, m = 0;
    bool loaded = false, prepared = false;

    snapb::Csr X;    // cells x features (this rank's rows)
    snapb::Csr Xt;   // features x local cells, CSR: only built for the CSR-gather operator (ensure_xt)
    bool xt_built = false;
    snapb::TileT XtT;  // tile-major 16-bit transpose: the input of S1 (tiled path), released once S1 is built
    snapb::Sell S2;  // tiled copy of X  (pass 2: gathers W rows by feature)
    snapb::Sell S1;  // tiled copy of Xt (pass 1: gathers r.*V rows by cell)
    int spmm_mode = 0;   // 0 = auto, 1 = CSR gather from L2, 2 = shared-memory tiled SELL
    // stored entries per rank averaged over the communicator (set by prepare / prepare_projection,
    // -1 = unknown): the automatic kernel choice must be the same on every rank, or the ranks would
    // disagree on block widths and all-reduce counts
    int64_t nnz_mode = -1;
    // default Lanczos block width (prepare builds the tiled copies for it).  4: a dense row is one
    // 16-byte bank group, half the shared-memory traffic per entry of b = 8; the solver needs ~1.6x the
    // operator applications but each costs less than half.
    int block = 4;

    // deferred scan of the host value array (load_csr with defer_value_scan): the matrix is loaded as a
    // pattern, a background team checks that every stored value is 1 while the GPU already works;
    // values_verdict() joins it
    bool defer_value_scan = false;
    bool scan_pending = false;
    std::thread scan_thread;
    std::atomic<int> scan_not_one{0};

    // block-wise load in progress (load_begin / load_append / load_end): rows and entries appended so far
    bool appending = false;
    int64_t app_rows = 0, app_nnz = 0;

    // user feature weights (host copy, optional)
    std::vector<double> user_weights;

    // prepare() products
    snapb::DevBuf<double> w;        // m   feature weights (IDF or user)
    snapb::DevBuf<double> rho;      // n   row L2 norms of the weighted rows
    snapb::DevBuf<double> csum;     // m   column sums of Xhat
    snapb::DevBuf<double> degree;   // n   d_i = Xhat c - 1
    snapb::DevBuf<float> r;         // n   sqrt(dinv)/rho  (row scale of X~)
    snapb::DevBuf<float> dinv;      // n   1/d
    snapb::DevBuf<float> w2;        // m   w^2 (f32)
    snapb::DevBuf<float> u1;        // n   sqrt(d)/||sqrt(d)||, trivial eigenvector
    double view_scale = 1.0;        // multi-view scale folded into w
    snapb::DevBuf<float> w_f;       // m   w as f32      (projection products)
    snapb::DevBuf<float> rhoinv_f;  // n   1/rho as f32
    bool proj_ready = false;

    // operator workspaces
    snapb::DevBuf<float> Vr;        // n x b  r .* V
    snapb::DevBuf<float> W;         // m x b  X^T (r V), then w^2-scaled
    snapb::DevBuf<float> opV, opY;  // operator_apply / operator_time staging
    snapb::DevBuf<float> partial;   // n_tiles x nrows x 8 per-tile partial sums of the tiled SpMM

    // scratch
    snapb::DevBuf<unsigned char> scratch;
    snapb::PinBuf<unsigned char> pinned;
    snapb::PinBuf<unsigned char> ring;          // pinned ring of the host staging team (ingest.cu)
    std::vector<cudaEvent_t> ring_events;
    snapb::DevBuf<unsigned char> ring_dev;      // device side of the ring (delta-encoded index chunks before decoding)

    // multi-view (multi_spectral): further views chained behind this context by combine_views();
    // they share this context's stream and communicator (attach_view) and are owned by the caller
    std::vector<snapb200_ctx*> views;
    snapb::DevBuf<float> neg_one;   // n  (accumulating pass 2 of the views after the first)
    bool borrowed = false;          // stream / communicator belong to another context

    snapb200_stats stats{};

    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

namespace snapb {

// ---- comm.cu
void comm_unique_id(char id[128]);
void comm_init(snapb200_ctx* c, int rank, int nranks, const char id[128]);
void comm_destroy(snapb200_ctx* c);
void peer_setup(snapb200_ctx* c);
// descriptor of the next fused exchange (advances the sequence number); false: use NCCL
bool peer_box(snapb200_ctx* c, PeerBox* out);
bool peer_error(snapb200_ctx* c);
void allreduce_f32(snapb200_ctx* c, float* buf, int64_t count);
void allreduce_f64(snapb200_ctx* c, double* buf, int64_t count);
void allreduce_i64(snapb200_ctx* c, int64_t* buf, int64_t count);

// ---- ingest.cu: threaded staging between pageable host arrays and the device
int host_threads(const snapb200_ctx* c);
bool stage_indices(snapb200_ctx* c, const void* src, int bits, int64_t count, int32_t* dst_dev);
int delta_selftest_host(const void* src, int bits, int64_t count, int64_t* n_side);
bool host_values_all_ones(snapb200_ctx* c, const void* values, int kind, int64_t count);
void stage_values(snapb200_ctx* c, const void* src, int kind, int64_t count, float* dst_dev);
void copy_to_host(snapb200_ctx* c, void* dst, const void* src_dev, size_t bytes);

// ---- knn.cu: exact k-nearest-neighbour graph of an n x d float64 point set (the consumer of X_spectral)
void knn(snapb200_ctx* c, int64_t n, int d, const double* points, int on_device, int64_t q0, int64_t nq, int k,
         int32_t* out_indices, double* out_distances);
int knn_max_neighbors();
int knn_max_dim();

// ---- synth.cu
void generate_rows(snapb200_ctx* c, int64_t n_local, int64_t n_global, int64_t row0, int64_t m,
                   int nnz_row, int n_clusters, uint64_t seed, const uint64_t* feat_cdf,
                   const uint64_t* cluster_cdf, const int64_t* block_start, const uint64_t* alpha);

// ---- util.cu
void exclusive_scan_i64(snapb200_ctx* c, const int64_t* in, int64_t* out, int64_t n);  // out has n+1
void exclusive_scan_i32_to_i64(snapb200_ctx* c, const int32_t* in, int64_t* out, int64_t n);
void fill_f32(snapb200_ctx* c, float* p, float v, int64_t n);
void flush_l2(snapb200_ctx* c);

// ---- prep.cu
void select_features(snapb200_ctx* c, const uint8_t* keep_host, int64_t m);
void ensure_xt(snapb200_ctx* c);   // builds the CSR feature-major copy c->Xt if it is missing
// tile-major transpose c->XtT for cell tiles of `tile_rows`; also leaves the local document
// frequencies in df_local (m int64) when it is non-null
void transpose_tiled(snapb200_ctx* c, int tile_rows, int64_t* df_local);
void prepare(snapb200_ctx* c, double* idf_out, double* degree_out);
void view_norms(snapb200_ctx* c, double* idf_out, double* rho_out);
double view_frobenius(snapb200_ctx* c, const int64_t* sample_rows, int64_t ns_local);
void combine_views(snapb200_ctx* main, snapb200_ctx** views, const double* cv, int n_views, double* degree_out);
void gather_rows(snapb200_ctx* src, const int64_t* rows_host, int64_t nr, snapb200_ctx* dst, int64_t n_global_dst,
                 int64_t row0_dst);
// IDF (or user) weights and weighted row norms of the loaded matrix into device buffers (no transpose)
void weights_and_norms(snapb200_ctx* c, double* w_dev, double* rho_dev);

// ---- spmm.cu
// Y[n x b] (leading dim ldy) = X~ X~^T V - dinv .* V, V with leading dim ldv.
// evs (optional, 4 events): recorded before pass 1, after pass 1, after the
// all-reduce, after pass 2.
// vr_ready: c->Vr already holds r .* V (the eigensolver's append kernel writes it), skip the scaling kernel.
void operator_apply_dev(snapb200_ctx* c, const float* V, int64_t ldv, float* Y, int64_t ldy, int b,
                        cudaEvent_t* evs = nullptr, bool vr_ready = false);

// products with Xhat = diag(1/rho) P diag(w) on k dense columns (Nystrom extension)
void prepare_projection(snapb200_ctx* c);
void project(snapb200_ctx* c, bool transposed, const float* in_host, int k, float* out_host);

// ---- sell_build.cu / spmm_tiled.cu
void sell_build(snapb200_ctx* c, const Csr& M, Sell& S, int b);
// the same from the tile-major transpose (rows = features, columns = local cells; T.tile_rows must
// equal the format's column tile for b)
void sell_build_transposed(snapb200_ctx* c, const TileT& T, int64_t n_cells, Sell& S, int b);
// out[row, 0:b] = scale[row] * (M in)[row, 0:b] - (sub ? subscale[row] * sub[row*lds + 0:b] : 0)
// through the tiled copy (b = S.b); `in` and `out` are packed (leading dimension b).
void sell_spmm(snapb200_ctx* c, const Sell& S, const float* in, float* out, const float* scale,
               const float* subscale, const float* sub, int64_t lds);
// fp64 SpMV through the tiled copy (prepare): mode 0: out[row] = sqrt(sum_j v_ij^2 x[j]);
// mode 1: out[row] = scale[row] * sum_j v_ij x[j] + shift  (v = 1 without values, scale may be null)
void sell_spmv64(snapb200_ctx* c, const Sell& S, const double* x, int mode, const double* scale, double shift,
                 double* out);
bool use_tiled(const snapb200_ctx* c, int b);
// agree on the stored-entry count the automatic kernel choice looks at (collective over the ranks)
void decide_spmm_mode(snapb200_ctx* c);
// (re)build S1/S2 for block width b if they are missing or sized for another width
void ensure_tiled(snapb200_ctx* c, int b);

// ---- lanczos.cu
void eigsh(snapb200_ctx* c, int k, int64_t seed, double tol, int block, int max_basis, int max_ops,
           double* evals, double* evecs, bool scale_by_sqrt_eval);

inline void count_launch(snapb200_ctx* c, int64_t n = 1) { c->stats.kernel_launches += n; }

}  // namespace snapb

// prepare(): feature selection, deterministic transpose, IDF weights, row
// norms, column sums and degrees (reference: snapatac2-python/src/embedding.rs
// :269-286 idf, :315-326 normalize, :139-152 degree block of spectral_mf).
//
// All of these are one-off O(nnz) passes.  The m- and n-length vectors they
// gather from are kept in fp64 so the IDF / degree parity gates (1e-5) are met
// with a wide margin; the index stream is the only large traffic.
#include "ctx.cuh"

#include <math.h>
#include <chrono>
#include <vector>

namespace snapb {

namespace {

// ------------------------------------------------------------------------
// Feature selection (embedding.rs:36-39)
// ------------------------------------------------------------------------
__global__ void count_kept_kernel(const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                                  const int64_t* __restrict__ remap, const uint8_t* __restrict__ keep,
                                  int64_t nrows, int32_t* __restrict__ len) {
    int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    for (int64_t r = warp; r < nrows; r += nwarps) {
        int cnt = 0;
        for (int64_t p = ptr[r] + lane; p < ptr[r + 1]; p += 32) cnt += keep[idx[p]] ? 1 : 0;
        cnt = static_cast<int>(warp_sum(static_cast<float>(cnt)) + 0.5f);
        if (lane == 0) len[r] = cnt;
    }
}

__global__ void compact_kept_kernel(const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                                    const float* __restrict__ val, const int64_t* __restrict__ remap,
                                    const uint8_t* __restrict__ keep, int64_t nrows,
                                    const int64_t* __restrict__ nptr, int32_t* __restrict__ nidx,
                                    float* __restrict__ nval) {
    int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    for (int64_t r = warp; r < nrows; r += nwarps) {
        int64_t out = nptr[r];
        int64_t s = ptr[r], e = ptr[r + 1];
        for (int64_t base = s; base < e; base += 32) {
            int64_t p = base + lane;
            int j = (p < e) ? idx[p] : 0;
            bool k = (p < e) && keep[j];
            unsigned mask = __ballot_sync(0xffffffffu, k);
            if (k) {
                int off = __popc(mask & ((1u << lane) - 1u));
                nidx[out + off] = static_cast<int32_t>(remap[j]);
                if (val) nval[out + off] = val[p];
            }
            out += __popc(mask);
        }
    }
}

// ------------------------------------------------------------------------
// Deterministic transpose: column histogram, scan, then per row slab a
// feature x slab-row bitmap (atomicOr is order independent) that one thread
// per feature walks in row order.  Output rows (features) therefore list
// their cells in ascending order regardless of scheduling.
// ------------------------------------------------------------------------
__global__ void col_count_kernel(const int32_t* __restrict__ idx, int64_t nnz, int32_t* __restrict__ cnt) {
    int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (; i < nnz; i += stride) atomicAdd(&cnt[ld_stream_int(idx + i)], 1);
}

__global__ void bitmap_set_kernel(const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                                  int64_t r0, int64_t r1, int64_t m, uint32_t* __restrict__ bm) {
    // kSetSplit warps share a row so that a 1024-row slab keeps every SM busy
    // (one warp per row left ~7 warps per SM and the kernel latency bound)
    constexpr int kSetSplit = 8;
    int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    const int64_t n_units = (r1 - r0) * kSetSplit;
    for (int64_t u = warp; u < n_units; u += nwarps) {
        const int64_t r = r0 + u / kSetSplit;
        const int part = static_cast<int>(u % kSetSplit);
        const int64_t il = r - r0;
        uint32_t* plane = bm + (il >> 5) * m;
        const uint32_t bit = 1u << (il & 31);
        const int64_t s = ptr[r], len = ptr[r + 1] - s;
        const int64_t lo = s + (len * part) / kSetSplit, hi = s + (len * (part + 1)) / kSetSplit;
        int64_t p = lo + lane;
        for (; p + 96 < hi; p += 128) {   // four index loads in flight per lane
            const int j0 = ld_stream_int(idx + p), j1 = ld_stream_int(idx + p + 32);
            const int j2 = ld_stream_int(idx + p + 64), j3 = ld_stream_int(idx + p + 96);
            atomicOr(plane + j0, bit);
            atomicOr(plane + j1, bit);
            atomicOr(plane + j2, bit);
            atomicOr(plane + j3, bit);
        }
        for (; p < hi; p += 32) atomicOr(plane + ld_stream_int(idx + p), bit);
    }
}

// position of column j inside row r (indices sorted) -- values path only
__device__ __forceinline__ int64_t find_in_row(const int32_t* __restrict__ idx, int64_t s, int64_t e, int32_t j) {
    while (s < e) {
        int64_t mid = (s + e) >> 1;
        if (idx[mid] < j) s = mid + 1; else e = mid;
    }
    return s;
}

// One thread per feature walks the slab's planes in row order (eight coalesced word loads in
// flight) and appends the set bits to the feature's list; consumed words are cleared in place so
// the slab needs no separate memset.  (A 32 x 32 feature x plane CTA tiling with a shared-memory
// prefix was tried and was 3x slower: profiles/README.md.)
__global__ void bitmap_emit_kernel(uint32_t* __restrict__ bm, int64_t m, int planes, int64_t r0,
                                   int64_t* __restrict__ cursor, int32_t* __restrict__ tidx,
                                   const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                                   const float* __restrict__ val, float* __restrict__ tval) {
    int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (j >= m) return;
    int64_t cur = cursor[j];
    for (int k0 = 0; k0 < planes; k0 += 8) {
        uint32_t words[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)   // eight independent (coalesced) loads in flight
            words[u] = (k0 + u < planes) ? bm[static_cast<int64_t>(k0 + u) * m + j] : 0u;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            uint32_t word = words[u];
            if (word) bm[static_cast<int64_t>(k0 + u) * m + j] = 0u;
            while (word) {
                int bit = __ffs(word) - 1;
                word &= word - 1;
                int64_t r = r0 + (k0 + u) * 32 + bit;
                tidx[cur] = static_cast<int32_t>(r);
                if (val) tval[cur] = val[find_in_row(idx, ptr[r], ptr[r + 1], static_cast<int32_t>(j))];
                ++cur;
            }
        }
    }
    cursor[j] = cur;
}

// ------------------------------------------------------------------------
// IDF (embedding.rs:269-286)
// ------------------------------------------------------------------------
__global__ void df_minmax_kernel(const int64_t* __restrict__ df, int64_t m, int64_t* __restrict__ mm) {
    // single block
    __shared__ int64_t smin[256], smax[256];
    int64_t lo = INT64_MAX, hi = INT64_MIN;
    for (int64_t j = threadIdx.x; j < m; j += blockDim.x) {
        int64_t v = df[j];
        lo = v < lo ? v : lo;
        hi = v > hi ? v : hi;
    }
    smin[threadIdx.x] = lo;
    smax[threadIdx.x] = hi;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            smin[threadIdx.x] = min(smin[threadIdx.x], smin[threadIdx.x + s]);
            smax[threadIdx.x] = max(smax[threadIdx.x], smax[threadIdx.x + s]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { mm[0] = smin[0]; mm[1] = smax[0]; }
}

__global__ void idf_kernel(const int64_t* __restrict__ df, const int64_t* __restrict__ mm, int64_t m,
                           double n_total, double* __restrict__ w) {
    int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (j >= m) return;
    if (mm[0] == mm[1]) { w[j] = 1.0; return; }      // all_equal -> ones (:273-274)
    double x = static_cast<double>(df[j]);
    if (x == 0.0) x = 1.0;                            // :277-278
    else if (x == n_total) x = n_total - 1.0;         // :279-280
    w[j] = log(n_total / x);                          // :282
}

__global__ void i32_to_i64_kernel(const int32_t* __restrict__ in, int64_t* __restrict__ out, int64_t n) {
    int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}

// ------------------------------------------------------------------------
// fp64 gather SpMVs, one warp per row
// ------------------------------------------------------------------------
// MODE 0: out[r] = sqrt(sum (val*vec[idx])^2)            row norms   (:321-323)
// MODE 1: out[r] = scale[r] * sum val*vec[idx]           column sums (:139-144) / degrees (:145)
template <int MODE>
__global__ void __launch_bounds__(256)
spmv_f64_kernel(const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx, const float* __restrict__ val,
                const double* __restrict__ vec, const double* __restrict__ scale, double shift, int64_t nrows,
                double* __restrict__ out) {
    int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    for (int64_t r = warp; r < nrows; r += nwarps) {
        int64_t s = ptr[r], e = ptr[r + 1];
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int64_t p = s + lane;
        for (; p + 96 < e; p += 128) {
            int j0 = ld_stream_int(idx + p), j1 = ld_stream_int(idx + p + 32);
            int j2 = ld_stream_int(idx + p + 64), j3 = ld_stream_int(idx + p + 96);
            double x0 = vec[j0], x1 = vec[j1], x2 = vec[j2], x3 = vec[j3];
            if (val) {
                x0 *= static_cast<double>(val[p]);      x1 *= static_cast<double>(val[p + 32]);
                x2 *= static_cast<double>(val[p + 64]); x3 *= static_cast<double>(val[p + 96]);
            }
            if (MODE == 0) { a0 += x0 * x0; a1 += x1 * x1; a2 += x2 * x2; a3 += x3 * x3; }
            else           { a0 += x0;      a1 += x1;      a2 += x2;      a3 += x3; }
        }
        for (; p < e; p += 32) {
            double x0 = vec[ld_stream_int(idx + p)];
            if (val) x0 *= static_cast<double>(val[p]);
            a0 += (MODE == 0) ? x0 * x0 : x0;
        }
        double acc = warp_sum((a0 + a1) + (a2 + a3));
        if (lane == 0) {
            if (MODE == 0) out[r] = sqrt(acc);
            else out[r] = (scale ? scale[r] : 1.0) * acc + shift;
        }
    }
}

__global__ void recip_kernel(const double* __restrict__ in, double* __restrict__ out, int64_t n) {
    int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = 1.0 / in[i];
}
__global__ void sqr_kernel(const double* __restrict__ in, double* __restrict__ out, int64_t n) {
    int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] * in[i];
}
__global__ void mul_kernel(const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out, int64_t n) {
    int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] * b[i];
}

// sum of degrees + count of degenerate rows (block partials, fixed order)
__global__ void degree_stats_kernel(const double* __restrict__ d, int64_t n, double* __restrict__ part_sum,
                                    int64_t* __restrict__ part_bad) {
    __shared__ double ssum[256];
    __shared__ int64_t sbad[256];
    double s = 0.0;
    int64_t bad = 0;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        double v = d[i];
        if (!(v > 0.0) || isinf(v)) ++bad; else s += v;
    }
    ssum[threadIdx.x] = s;
    sbad[threadIdx.x] = bad;
    __syncthreads();
    for (int k = blockDim.x / 2; k > 0; k >>= 1) {
        if (threadIdx.x < k) { ssum[threadIdx.x] += ssum[threadIdx.x + k]; sbad[threadIdx.x] += sbad[threadIdx.x + k]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { part_sum[blockIdx.x] = ssum[0]; part_bad[blockIdx.x] = sbad[0]; }
}

__global__ void derive_rows_kernel(const double* __restrict__ d, const double* __restrict__ rho, double inv_norm,
                                   int64_t n, float* __restrict__ r, float* __restrict__ dinv, float* __restrict__ u1) {
    int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double di = d[i];
    double inv = 1.0 / di;
    dinv[i] = static_cast<float>(inv);
    r[i] = static_cast<float>(sqrt(inv) / rho[i]);
    u1[i] = static_cast<float>(sqrt(di) * inv_norm);
}
__global__ void derive_cols_kernel(const double* __restrict__ w, int64_t m, float* __restrict__ w2) {
    int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (j < m) w2[j] = static_cast<float>(w[j] * w[j]);
}

inline int grid_for_rows(snapb200_ctx* c, int64_t nrows) {
    return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(ceil_div(nrows, 8), static_cast<int64_t>(c->num_sms) * 16)));
}
inline unsigned grid1d(int64_t n, int threads = 256) { return static_cast<unsigned>(std::max<int64_t>(1, ceil_div(n, threads))); }

}  // namespace

// --------------------------------------------------------------------------
void select_features(snapb200_ctx* c, const uint8_t* keep_host, int64_t m) {
    SB_CHECK(c->loaded, "select_features: no matrix loaded");
    SB_CHECK(m == c->m, "select_features: mask length must equal the number of columns");
    Csr& X = c->X;
    DevBuf<uint8_t> keep;
    DevBuf<int32_t> keep32, len;
    DevBuf<int64_t> remap, nptr;
    keep.alloc(m);
    SB_CUDA(cudaMemcpyAsync(keep.p, keep_host, static_cast<size_t>(m), cudaMemcpyHostToDevice, c->stream));
    std::vector<int32_t> k32(static_cast<size_t>(m));
    int64_t m_new = 0;
    for (int64_t j = 0; j < m; ++j) { k32[j] = keep_host[j] ? 1 : 0; m_new += k32[j]; }
    SB_CHECK(m_new > 0, "select_features: no feature selected");
    keep32.alloc(m);
    SB_CUDA(cudaMemcpyAsync(keep32.p, k32.data(), sizeof(int32_t) * m, cudaMemcpyHostToDevice, c->stream));
    remap.alloc(m + 1);
    exclusive_scan_i32_to_i64(c, keep32.p, remap.p, m);

    len.alloc(std::max<int64_t>(1, X.nrows));
    nptr.alloc(X.nrows + 1);
    int g = grid_for_rows(c, X.nrows);
    count_kept_kernel<<<g, 256, 0, c->stream>>>(X.ptr.p, X.idx.p, remap.p, keep.p, X.nrows, len.p);
    SB_LAUNCH_CHECK();
    exclusive_scan_i32_to_i64(c, len.p, nptr.p, X.nrows);
    int64_t nnz = 0;
    SB_CUDA(cudaMemcpyAsync(&nnz, nptr.p + X.nrows, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    SB_CUDA(cudaStreamSynchronize(c->stream));
    DevBuf<int32_t> nidx;
    DevBuf<float> nval;
    nidx.alloc(std::max<int64_t>(1, nnz));
    if (X.has_values()) nval.alloc(std::max<int64_t>(1, nnz));
    compact_kept_kernel<<<g, 256, 0, c->stream>>>(X.ptr.p, X.idx.p, X.val.p, remap.p, keep.p, X.nrows, nptr.p,
                                                 nidx.p, nval.p);
    SB_LAUNCH_CHECK();
    count_launch(c, 2);
    SB_CUDA(cudaStreamSynchronize(c->stream));
    X.ptr.swap(nptr);
    X.idx.swap(nidx);
    if (X.has_values()) X.val.swap(nval);
    X.nnz = nnz;
    X.ncols = m_new;
    c->m = m_new;
    c->prepared = false;
    c->proj_ready = false;
    c->Xt.clear();
    c->xt_built = false;
    c->XtT.clear();
    c->S1.clear();
    c->S2.clear();
    c->nnz_mode = -1;
    c->stats.nnz_local = nnz;
}

// --------------------------------------------------------------------------
// Builds c->Xt and leaves the local column counts in `cnt_out` (m int32).
// Temporaries of the transpose that must outlive its (asynchronous) slab loop.
struct TransposeState {
    DevBuf<int64_t> cursor;
    DevBuf<uint32_t> bm;
    int64_t S = 0;
};

// Part 1: local column counts (left in `cnt`), row pointers of Xt, output allocation.
static void transpose_begin(snapb200_ctx* c, DevBuf<int32_t>& cnt, TransposeState& ts) {
    Csr& X = c->X;
    Csr& T = c->Xt;
    const int64_t m = c->m, n = X.nrows, nnz = X.nnz;
    T.nrows = m;
    T.ncols = n;
    T.nnz = nnz;
    cnt.alloc(m);
    SB_CUDA(cudaMemsetAsync(cnt.p, 0, sizeof(int32_t) * m, c->stream));
    if (nnz > 0) {
        int blocks = static_cast<int>(std::min<int64_t>(ceil_div(nnz, 256), static_cast<int64_t>(c->num_sms) * 32));
        col_count_kernel<<<blocks, 256, 0, c->stream>>>(X.idx.p, nnz, cnt.p);
        SB_LAUNCH_CHECK();
        count_launch(c);
    }
    T.ptr.alloc(m + 1);
    exclusive_scan_i32_to_i64(c, cnt.p, T.ptr.p, m);
    T.idx.alloc(std::max<int64_t>(1, nnz));
    if (X.has_values()) T.val.alloc(std::max<int64_t>(1, nnz)); else T.val.release();
    ts.S = 0;
    if (nnz == 0 || n == 0) return;
    ts.cursor.alloc(m);
    SB_CUDA(cudaMemcpyAsync(ts.cursor.p, T.ptr.p, sizeof(int64_t) * m, cudaMemcpyDeviceToDevice, c->stream));
    // slab height: bitmap of m x S bits kept around 64 MB (L2 resident)
    int64_t S = ((512ll << 20) / std::max<int64_t>(m, 1)) / 32 * 32;
    S = std::max<int64_t>(32, std::min<int64_t>(S, 8192));
    S = std::min<int64_t>(S, ceil_div(n, 32) * 32);
    ts.S = S;
    ts.bm.alloc((S / 32) * m);
    // cleared once; the emit kernel zeroes every word it consumes
    SB_CUDA(cudaMemsetAsync(ts.bm.p, 0, sizeof(uint32_t) * static_cast<size_t>(S / 32) * m, c->stream));
}

// Part 2: the slab loop, enqueued on the context's stream without waiting for it.
static void transpose_enqueue(snapb200_ctx* c, TransposeState& ts) {
    if (ts.S == 0) return;
    Csr& X = c->X;
    Csr& T = c->Xt;
    const int64_t m = c->m, n = X.nrows, S = ts.S;
    for (int64_t r0 = 0; r0 < n; r0 += S) {
        int64_t r1 = std::min<int64_t>(n, r0 + S);
        int g = grid_for_rows(c, (r1 - r0) * 8);   // 8 warps per row (kSetSplit)
        bitmap_set_kernel<<<g, 256, 0, c->stream>>>(X.ptr.p, X.idx.p, r0, r1, m, ts.bm.p);
        SB_LAUNCH_CHECK();
        int pl = static_cast<int>(ceil_div(r1 - r0, 32));
        bitmap_emit_kernel<<<grid1d(m), 256, 0, c->stream>>>(ts.bm.p, m, pl, r0, ts.cursor.p, T.idx.p, X.ptr.p,
                                                            X.idx.p, X.val.p, T.val.p);
        SB_LAUNCH_CHECK();
        count_launch(c, 2);
    }
}

void ensure_xt(snapb200_ctx* c) {
    if (c->xt_built) return;
    DevBuf<int32_t> cnt;
    TransposeState ts;
    transpose_begin(c, cnt, ts);
    transpose_enqueue(c, ts);
    SB_CUDA(cudaStreamSynchronize(c->stream));
    c->xt_built = true;
}

// --------------------------------------------------------------------------
// prepare(): document frequencies -> IDF, both layouts of the pattern the operator gathers over,
// row norms, column sums, degrees, and the derived fp32 vectors.
//   tiled path (large problems): tile-major transpose -> S1, S2 from the rows; the three fp64 SpMVs
//   run through the tiled copies.
//   CSR path (small problems): CSR transpose, fp64 gather SpMVs.
void prepare(snapb200_ctx* c, double* idf_out, double* degree_out) {
    SB_CHECK(c->loaded, "prepare: no matrix loaded");
    Csr& X = c->X;
    const int64_t m = c->m, n = c->n_local;
    SB_CHECK(m >= 1, "prepare: matrix has no columns");
    cudaStream_t st = c->stream;
    const auto wall0 = std::chrono::steady_clock::now();
    auto since = [](std::chrono::steady_clock::time_point t0) {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    };
    c->views.clear();   // a new prepare invalidates a multi-view combination
    decide_spmm_mode(c);
    const bool tiled = use_tiled(c, c->block);
    const bool user_w = !c->user_weights.empty();
    if (user_w)
        SB_CHECK(static_cast<int64_t>(c->user_weights.size()) == m,
                 "feature_weights length must equal the number of selected features");

    c->w.alloc(m);
    c->rho.alloc(std::max<int64_t>(1, n));
    c->degree.alloc(std::max<int64_t>(1, n));
    c->csum.alloc(m);
    DevBuf<double> rinv, wc;
    DevBuf<int64_t> df, mm;
    rinv.alloc(n + 2);   // (+2: the tiled fp64 SpMV stages whole 16-byte units)
    wc.alloc(m + 2);
    df.alloc(m);
    mm.alloc(2);
    c->S1.clear();
    c->S2.clear();
    c->XtT.clear();
    double ms_format = 0.0;

    // ---- transpose (also yields the local document frequencies)
    SB_CUDA(cudaEventRecord(c->ev0, st));
    if (tiled) {
        c->Xt.clear();
        c->xt_built = false;
        transpose_tiled(c, kSellTileBytes / (4 * c->block), df.p);
    } else {
        c->xt_built = false;
        DevBuf<int32_t> cnt;
        TransposeState ts;
        transpose_begin(c, cnt, ts);
        i32_to_i64_kernel<<<grid1d(m), 256, 0, st>>>(cnt.p, df.p, m);
        SB_LAUNCH_CHECK();
        count_launch(c);
        transpose_enqueue(c, ts);
        SB_CUDA(cudaStreamSynchronize(st));
        c->xt_built = true;
    }
    SB_CUDA(cudaEventRecord(c->ev1, st));
    if (user_w) {
        SB_CUDA(cudaMemcpyAsync(c->w.p, c->user_weights.data(), sizeof(double) * m, cudaMemcpyHostToDevice, st));
    } else {
        allreduce_i64(c, df.p, m);
    }

    // ---- IDF weights (:269-286)
    if (!user_w) {
        df_minmax_kernel<<<1, 256, 0, st>>>(df.p, m, mm.p);
        SB_LAUNCH_CHECK();
        idf_kernel<<<grid1d(m), 256, 0, st>>>(df.p, mm.p, m, static_cast<double>(c->n_global), c->w.p);
        SB_LAUNCH_CHECK();
        count_launch(c, 2);
    }
    // ---- both tiled copies, then row norms rho_i = || w .* x_i ||   (:315-326)
    if (tiled) {
        const auto w0 = std::chrono::steady_clock::now();
        sell_build_transposed(c, c->XtT, n, c->S1, c->block);
        c->XtT.clear();   // only the input of S1: released (rebuilt on demand if the block width changes)
        sell_build(c, c->X, c->S2, c->block);
        ms_format = since(w0);
        if (n > 0) {
            sqr_kernel<<<grid1d(m), 256, 0, st>>>(c->w.p, wc.p, m);   // wc is free until the degrees
            SB_LAUNCH_CHECK();
            sell_spmv64(c, c->S2, wc.p, 0, nullptr, 0.0, c->rho.p);
            count_launch(c);
        }
    } else if (n > 0) {
        spmv_f64_kernel<0><<<grid_for_rows(c, n), 256, 0, st>>>(X.ptr.p, X.idx.p, X.val.p, c->w.p, nullptr, 0.0, n,
                                                               c->rho.p);
        SB_LAUNCH_CHECK();
        count_launch(c);
    }
    if (n > 0) {
        recip_kernel<<<grid1d(n), 256, 0, st>>>(c->rho.p, rinv.p, n);
        SB_LAUNCH_CHECK();
        count_launch(c);
    }

    // ---- column sums c_j = w_j sum_i x_ij / rho_i               (:139-144)
    if (tiled) {
        sell_spmv64(c, c->S1, rinv.p, 1, c->w.p, 0.0, c->csum.p);
    } else {
        spmv_f64_kernel<1><<<grid_for_rows(c, m), 256, 0, st>>>(c->Xt.ptr.p, c->Xt.idx.p, c->Xt.val.p, rinv.p, c->w.p,
                                                               0.0, m, c->csum.p);
        SB_LAUNCH_CHECK();
        count_launch(c);
    }
    allreduce_f64(c, c->csum.p, m);
    mul_kernel<<<grid1d(m), 256, 0, st>>>(c->w.p, c->csum.p, wc.p, m);
    SB_LAUNCH_CHECK();
    count_launch(c);
    // ---- degrees d_i = (1/rho_i) sum_j x_ij w_j c_j - 1         (:145-146)
    if (n > 0) {
        if (tiled) {
            sell_spmv64(c, c->S2, wc.p, 1, rinv.p, -1.0, c->degree.p);
        } else {
            spmv_f64_kernel<1><<<grid_for_rows(c, n), 256, 0, st>>>(X.ptr.p, X.idx.p, X.val.p, wc.p, rinv.p, -1.0, n,
                                                                   c->degree.p);
            SB_LAUNCH_CHECK();
            count_launch(c);
        }
    }
    // ---- global sum of degrees, degenerate-row check
    const int nb = 256;
    DevBuf<double> psum;
    DevBuf<int64_t> pbad;
    psum.alloc(nb);
    pbad.alloc(nb);
    degree_stats_kernel<<<nb, 256, 0, st>>>(c->degree.p, n, psum.p, pbad.p);
    SB_LAUNCH_CHECK();
    count_launch(c);
    std::vector<double> hsum(nb);
    std::vector<int64_t> hbad(nb);
    SB_CUDA(cudaMemcpyAsync(hsum.data(), psum.p, sizeof(double) * nb, cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaMemcpyAsync(hbad.data(), pbad.p, sizeof(int64_t) * nb, cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    double tot[2] = {0.0, 0.0};
    for (int i = 0; i < nb; ++i) { tot[0] += hsum[i]; tot[1] += static_cast<double>(hbad[i]); }
    if (c->nranks > 1) {
        DevBuf<double> t2;
        t2.alloc(2);
        SB_CUDA(cudaMemcpyAsync(t2.p, tot, sizeof(double) * 2, cudaMemcpyHostToDevice, st));
        allreduce_f64(c, t2.p, 2);
        SB_CUDA(cudaMemcpyAsync(tot, t2.p, sizeof(double) * 2, cudaMemcpyDeviceToHost, st));
        SB_CUDA(cudaStreamSynchronize(st));
    }
    if (degree_out && n > 0)
        SB_CUDA(cudaMemcpyAsync(degree_out, c->degree.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    if (idf_out) SB_CUDA(cudaMemcpyAsync(idf_out, c->w.p, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    if (tot[1] > 0.0) {
        char buf[256];
        snprintf(buf, sizeof(buf),
                 "%lld cell(s) have an empty row or a non-positive degree after feature weighting; the reference "
                 "produces NaN here (embedding.rs:146,152,323)", static_cast<long long>(tot[1]));
        throw Error(buf);
    }

    // ---- derived f32 vectors used by the operator
    c->r.alloc(std::max<int64_t>(1, n));
    c->dinv.alloc(std::max<int64_t>(1, n));
    c->u1.alloc(std::max<int64_t>(1, n));
    c->w2.alloc(m);
    if (n > 0) {
        derive_rows_kernel<<<grid1d(n), 256, 0, st>>>(c->degree.p, c->rho.p, 1.0 / sqrt(tot[0]), n, c->r.p, c->dinv.p,
                                                     c->u1.p);
        SB_LAUNCH_CHECK();
    }
    derive_cols_kernel<<<grid1d(m), 256, 0, st>>>(c->w.p, m, c->w2.p);
    SB_LAUNCH_CHECK();
    count_launch(c, 2);
    SB_CUDA(cudaStreamSynchronize(st));
    {
        float ms_t = 0.f;
        cudaEventElapsedTime(&ms_t, c->ev0, c->ev1);
        c->stats.ms_transpose = ms_t;
    }
    c->stats.ms_format = ms_format;
    c->stats.ms_prepare_wall = since(wall0);
    c->stats.ms_prepare = std::max(0.0, c->stats.ms_prepare_wall - c->stats.ms_transpose - ms_format);
    c->prepared = true;
    c->proj_ready = false;
}

void weights_and_norms(snapb200_ctx* c, double* w_dev, double* rho_dev) {
    Csr& X = c->X;
    const int64_t m = c->m, n = c->n_local;
    cudaStream_t st = c->stream;
    if (!c->user_weights.empty()) {
        SB_CHECK(static_cast<int64_t>(c->user_weights.size()) == m,
                 "feature_weights length must equal the number of selected features");
        SB_CUDA(cudaMemcpyAsync(w_dev, c->user_weights.data(), sizeof(double) * m, cudaMemcpyHostToDevice, st));
    } else {
        DevBuf<int32_t> cnt;
        DevBuf<int64_t> df, mm;
        cnt.alloc(m);
        df.alloc(m);
        mm.alloc(2);
        SB_CUDA(cudaMemsetAsync(cnt.p, 0, sizeof(int32_t) * m, st));
        if (X.nnz > 0) {
            int blocks = static_cast<int>(std::min<int64_t>(ceil_div(X.nnz, 256), static_cast<int64_t>(c->num_sms) * 32));
            col_count_kernel<<<blocks, 256, 0, st>>>(X.idx.p, X.nnz, cnt.p);
            SB_LAUNCH_CHECK();
        }
        i32_to_i64_kernel<<<grid1d(m), 256, 0, st>>>(cnt.p, df.p, m);
        SB_LAUNCH_CHECK();
        allreduce_i64(c, df.p, m);
        df_minmax_kernel<<<1, 256, 0, st>>>(df.p, m, mm.p);
        SB_LAUNCH_CHECK();
        idf_kernel<<<grid1d(m), 256, 0, st>>>(df.p, mm.p, m, static_cast<double>(c->n_global), w_dev);
        SB_LAUNCH_CHECK();
        count_launch(c, 4);
        SB_CUDA(cudaStreamSynchronize(st));   // temporaries
    }
    if (n > 0) {
        spmv_f64_kernel<0><<<grid_for_rows(c, n), 256, 0, st>>>(X.ptr.p, X.idx.p, X.val.p, w_dev, nullptr, 0.0, n, rho_dev);
        SB_LAUNCH_CHECK();
        count_launch(c);
    }
}

// Per-view statistics for multi_spectral (embedding.rs:413-416): IDF weights of the loaded
// (and column-selected) view and the L2 norms of its IDF-weighted rows.  No transposition.
void view_norms(snapb200_ctx* c, double* idf_out, double* rho_out) {
    SB_CHECK(c->loaded, "view_norms: no matrix loaded");
    const int64_t m = c->m, n = c->n_local;
    cudaStream_t st = c->stream;
    DevBuf<double> w, rho;
    w.alloc(m);
    rho.alloc(std::max<int64_t>(1, n));
    // the per-view statistics always use the view's own IDF (embedding.rs:413), never user weights
    std::vector<double> saved;
    saved.swap(c->user_weights);
    try {
        weights_and_norms(c, w.p, rho.p);
    } catch (...) {
        saved.swap(c->user_weights);
        throw;
    }
    saved.swap(c->user_weights);
    if (idf_out) SB_CUDA(cudaMemcpyAsync(idf_out, w.p, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
    if (rho_out && n > 0) SB_CUDA(cudaMemcpyAsync(rho_out, rho.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
}

// ==========================================================================
// Multi-view embedding on the device (multi_spectral_embedding, embedding.rs:388-452) as a
// VIRTUAL column concatenation: every view keeps its own context (pattern, tiled copies, IDF
// weights, row norms) and the views only share the dense block V and the degree vector,
//     A V = sum_v X~_v (X~_v^T V) - D^-1 V,   X~_v = diag(r_v) P_v diag(w_v),
//     r_v,i = sqrt(1/d_i) c_v / rho_v,i,      c_v = sqrt((weight_v / norm_v) / sum)   (:428-442)
// -- exactly the operator of the hstack-ed matrix (:443, :367-385) without ever forming it, and a
// binarised view keeps its 2-byte pattern-only entries next to a valued one.  The stacked rows
// have unit norm (sum_v c_v^2 = 1), so  d_i = sum_v c_v^2 (d_v,i + 1) - 1  with d_v the degrees
// the ordinary single-view prepare() computes for view v.
// ==========================================================================
namespace {

__global__ void sample_mask_kernel(const int64_t* __restrict__ rows, int64_t ns, const double* __restrict__ rho,
                                   double* __restrict__ msk) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < ns) msk[rows[i]] = 1.0 / rho[rows[i]];
}
__global__ void sumsq_kernel(const double* __restrict__ v, int64_t n, double* __restrict__ part) {
    __shared__ double ss[256];
    double s = 0.0;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) s += v[i] * v[i];
    ss[threadIdx.x] = s;
    __syncthreads();
    for (int k = blockDim.x / 2; k > 0; k >>= 1) {
        if (threadIdx.x < k) ss[threadIdx.x] += ss[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.x] = ss[0];
}
// d (in: first view's degrees) <- c0^2 (d + 1) - 1   or   d += cv^2 (dv + 1)
__global__ void combine_degree_kernel(double* __restrict__ d, const double* __restrict__ dv, double c2, int first, int64_t n) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (first) d[i] = c2 * (d[i] + 1.0) - 1.0;
    else d[i] += c2 * (dv[i] + 1.0);
}
__global__ void view_rowscale_kernel(const double* __restrict__ d, const double* __restrict__ rho, double cv, int64_t n,
                                     float* __restrict__ r) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) r[i] = static_cast<float>(sqrt(1.0 / d[i]) * cv / rho[i]);
}

// out_cols[m] = scale_cols .* (P^T x_rows)   through whichever feature-major copy the context has
void col_product(snapb200_ctx* c, const double* x_rows, const double* scale_cols, double* out_cols) {
    if (c->S1.built) {
        sell_spmv64(c, c->S1, x_rows, 1, scale_cols, 0.0, out_cols);
    } else {
        ensure_xt(c);
        spmv_f64_kernel<1><<<grid_for_rows(c, c->m), 256, 0, c->stream>>>(c->Xt.ptr.p, c->Xt.idx.p, c->Xt.val.p, x_rows,
                                                                           scale_cols, 0.0, c->m, out_cols);
        SB_LAUNCH_CHECK();
        count_launch(c);
    }
}
// out_rows[n] = scale_rows .* (P x_cols) + shift
void row_product(snapb200_ctx* c, const double* x_cols, const double* scale_rows, double shift, double* out_rows) {
    if (c->n_local == 0) return;
    if (c->S2.built) {
        sell_spmv64(c, c->S2, x_cols, 1, scale_rows, shift, out_rows);
    } else {
        spmv_f64_kernel<1><<<grid_for_rows(c, c->n_local), 256, 0, c->stream>>>(c->X.ptr.p, c->X.idx.p, c->X.val.p, x_cols,
                                                                                 scale_rows, shift, c->n_local, out_rows);
        SB_LAUNCH_CHECK();
        count_launch(c);
    }
}

double sum_of_squares(snapb200_ctx* c, const double* v, int64_t n) {
    const int nb = 256;
    DevBuf<double> part;
    part.alloc(nb);
    sumsq_kernel<<<nb, 256, 0, c->stream>>>(v, n, part.p);
    SB_LAUNCH_CHECK();
    count_launch(c);
    std::vector<double> h(nb);
    SB_CUDA(cudaMemcpyAsync(h.data(), part.p, sizeof(double) * nb, cudaMemcpyDeviceToHost, c->stream));
    SB_CUDA(cudaStreamSynchronize(c->stream));
    double t = 0.0;
    for (double x : h) t += x;
    if (c->nranks > 1) {
        DevBuf<double> d;
        d.alloc(1);
        SB_CUDA(cudaMemcpyAsync(d.p, &t, sizeof(double), cudaMemcpyHostToDevice, c->stream));
        allreduce_f64(c, d.p, 1);
        SB_CUDA(cudaMemcpyAsync(&t, d.p, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        SB_CUDA(cudaStreamSynchronize(c->stream));
    }
    return t;
}

}  // namespace

// The value the Python snippet inside frobenius_norm (embedding.rs:456-460) returns for the unit-norm
// rows `sample_rows` (local ids, this rank's part of the sample) when it is handed a
// scipy.sparse.csr_matrix:  np.power(S, 2) is then the MATRIX square of S = Xs Xs^T, whose sum is
// || S 1 ||^2 = sum_i ( xhat_i . sum_{i' in sample} xhat_i' )^2  -- two fp64 SpMVs over the
// prepared view, no Gram matrix.  Collective over the row shards.
double view_frobenius(snapb200_ctx* c, const int64_t* sample_rows, int64_t ns_local) {
    SB_CHECK(c->prepared, "view_frobenius: call prepare first");
    const int64_t n = c->n_local, m = c->m;
    cudaStream_t st = c->stream;
    DevBuf<double> msk, cs, wc, y;
    DevBuf<int64_t> rows;
    msk.alloc(n + 2);
    cs.alloc(m);
    wc.alloc(m + 2);
    y.alloc(std::max<int64_t>(1, n));
    SB_CUDA(cudaMemsetAsync(msk.p, 0, sizeof(double) * (n + 2), st));
    if (ns_local > 0) {
        for (int64_t i = 0; i < ns_local; ++i) SB_CHECK(sample_rows[i] >= 0 && sample_rows[i] < n, "view_frobenius: sample row out of range");
        rows.alloc(ns_local);
        SB_CUDA(cudaMemcpyAsync(rows.p, sample_rows, sizeof(int64_t) * ns_local, cudaMemcpyHostToDevice, st));
        sample_mask_kernel<<<grid1d(ns_local), 256, 0, st>>>(rows.p, ns_local, c->rho.p, msk.p);
        SB_LAUNCH_CHECK();
        count_launch(c);
    }
    col_product(c, msk.p, c->w.p, cs.p);                       // sum of the sampled unit rows
    allreduce_f64(c, cs.p, m);
    mul_kernel<<<grid1d(m), 256, 0, st>>>(c->w.p, cs.p, wc.p, m);
    SB_LAUNCH_CHECK();
    count_launch(c);
    SB_CUDA(cudaMemsetAsync(y.p, 0, sizeof(double) * std::max<int64_t>(1, n), st));
    row_product(c, wc.p, msk.p, 0.0, y.p);                     // (S 1)_i on the sampled rows, 0 elsewhere
    const double total = sum_of_squares(c, y.p, n);
    SB_CUDA(cudaStreamSynchronize(st));                        // `rows` came from a host temporary
    return total;
}

// Chain `n_views` prepared view contexts behind `main` (views[0] must be main itself) with the view
// scales c_v: combined degrees, D^-1, the trivial eigenvector and every view's operator row scale.
void combine_views(snapb200_ctx* main, snapb200_ctx** views, const double* cv, int n_views, double* degree_out) {
    SB_CHECK(n_views >= 1 && views[0] == main, "combine_views: views[0] must be the main context");
    const int64_t n = main->n_local;
    cudaStream_t st = main->stream;
    for (int v = 0; v < n_views; ++v) {
        snapb200_ctx* x = views[v];
        SB_CHECK(x->prepared, "combine_views: every view must be prepared");
        SB_CHECK(x->n_local == n && x->n_global == main->n_global && x->row0 == main->row0, "combine_views: views must hold the same cells");
        SB_CHECK(x->stream == st && x->device == main->device, "combine_views: attach the view contexts first");
    }
    for (int v = 0; v < n_views; ++v) {
        if (n > 0) {
            combine_degree_kernel<<<grid1d(n), 256, 0, st>>>(main->degree.p, views[v]->degree.p, cv[v] * cv[v], v == 0 ? 1 : 0, n);
            SB_LAUNCH_CHECK();
        }
    }
    const int nb = 256;
    DevBuf<double> psum;
    DevBuf<int64_t> pbad;
    psum.alloc(nb);
    pbad.alloc(nb);
    degree_stats_kernel<<<nb, 256, 0, st>>>(main->degree.p, n, psum.p, pbad.p);
    SB_LAUNCH_CHECK();
    std::vector<double> hsum(nb);
    std::vector<int64_t> hbad(nb);
    SB_CUDA(cudaMemcpyAsync(hsum.data(), psum.p, sizeof(double) * nb, cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaMemcpyAsync(hbad.data(), pbad.p, sizeof(int64_t) * nb, cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    double tot[2] = {0.0, 0.0};
    for (int i = 0; i < nb; ++i) { tot[0] += hsum[i]; tot[1] += static_cast<double>(hbad[i]); }
    if (main->nranks > 1) {
        DevBuf<double> t2;
        t2.alloc(2);
        SB_CUDA(cudaMemcpyAsync(t2.p, tot, sizeof(double) * 2, cudaMemcpyHostToDevice, st));
        allreduce_f64(main, t2.p, 2);
        SB_CUDA(cudaMemcpyAsync(tot, t2.p, sizeof(double) * 2, cudaMemcpyDeviceToHost, st));
        SB_CUDA(cudaStreamSynchronize(st));
    }
    SB_CHECK(!(tot[1] > 0.0), "combine_views: a cell has a non-positive combined degree");
    if (n > 0) {
        // D^-1 and the trivial eigenvector from the combined degrees (rho is not used for them)
        derive_rows_kernel<<<grid1d(n), 256, 0, st>>>(main->degree.p, main->rho.p, 1.0 / sqrt(tot[0]), n, main->r.p, main->dinv.p,
                                                     main->u1.p);
        SB_LAUNCH_CHECK();
        for (int v = 0; v < n_views; ++v) {
            view_rowscale_kernel<<<grid1d(n), 256, 0, st>>>(main->degree.p, views[v]->rho.p, cv[v], n, views[v]->r.p);
            SB_LAUNCH_CHECK();
        }
        main->neg_one.alloc(n);
        fill_f32(main, main->neg_one.p, -1.f, n);
    }
    count_launch(main, n_views * 2 + 3);
    if (degree_out && n > 0) SB_CUDA(cudaMemcpyAsync(degree_out, main->degree.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    main->views.assign(views + 1, views + n_views);
}

// --------------------------------------------------------------------------
// Row gather between two contexts on the same device: dst <- the rows `rows` (local ids, any order)
// of src's resident CSR.  The Nystrom path takes its landmark rows this way (select_axis(0, ..),
// embedding.rs:95-99) without a round trip through the host.
// --------------------------------------------------------------------------
namespace {
__global__ void gather_len_kernel(const int64_t* __restrict__ ptr, const int64_t* __restrict__ rows, int64_t nr,
                                  int32_t* __restrict__ len) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < nr) len[i] = static_cast<int32_t>(ptr[rows[i] + 1] - ptr[rows[i]]);
}
__global__ void gather_copy_kernel(const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx, const float* __restrict__ val,
                                   const int64_t* __restrict__ rows, int64_t nr, const int64_t* __restrict__ optr,
                                   int32_t* __restrict__ oidx, float* __restrict__ oval) {
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    for (int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; i < nr; i += nwarps) {
        const int64_t s = ptr[rows[i]], e = ptr[rows[i] + 1], o = optr[i];
        for (int64_t p = s + lane; p < e; p += 32) {
            oidx[o + (p - s)] = idx[p];
            if (val) oval[o + (p - s)] = val[p];
        }
    }
}
}  // namespace

void gather_rows(snapb200_ctx* src, const int64_t* rows_host, int64_t nr, snapb200_ctx* dst, int64_t n_global_dst,
                 int64_t row0_dst) {
    SB_CHECK(src->loaded, "gather_rows: no matrix loaded in the source context");
    SB_CHECK(src != dst && src->device == dst->device, "gather_rows: need two contexts on the same device");
    SB_CHECK(nr >= 0 && row0_dst >= 0 && row0_dst + nr <= n_global_dst, "gather_rows: bad shard geometry");
    for (int64_t i = 0; i < nr; ++i) SB_CHECK(rows_host[i] >= 0 && rows_host[i] < src->n_local, "gather_rows: row out of range");
    if (src->stream != dst->stream) SB_CUDA(cudaStreamSynchronize(src->stream));
    cudaStream_t st = dst->stream;
    const Csr& X = src->X;
    Csr& Y = dst->X;
    dst->loaded = false;
    dst->prepared = false;
    dst->proj_ready = false;
    dst->views.clear();
    dst->nnz_mode = -1;
    dst->S1.clear(); dst->S2.clear(); dst->Xt.clear(); dst->xt_built = false; dst->XtT.clear();
    DevBuf<int64_t> rows;
    DevBuf<int32_t> len;
    rows.alloc(std::max<int64_t>(1, nr));
    len.alloc(std::max<int64_t>(1, nr));
    Y.nrows = nr;
    Y.ncols = X.ncols;
    Y.ptr.alloc(nr + 1);
    if (nr > 0) {
        SB_CUDA(cudaMemcpyAsync(rows.p, rows_host, sizeof(int64_t) * nr, cudaMemcpyHostToDevice, st));
        gather_len_kernel<<<grid1d(nr), 256, 0, st>>>(X.ptr.p, rows.p, nr, len.p);
        SB_LAUNCH_CHECK();
    }
    exclusive_scan_i32_to_i64(dst, len.p, Y.ptr.p, nr);
    int64_t nnz = 0;
    SB_CUDA(cudaMemcpyAsync(&nnz, Y.ptr.p + nr, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    Y.nnz = nnz;
    Y.idx.alloc(std::max<int64_t>(1, nnz));
    if (X.has_values()) Y.val.alloc(std::max<int64_t>(1, nnz)); else Y.val.release();
    if (nr > 0 && nnz > 0) {
        gather_copy_kernel<<<grid_for_rows(dst, nr), 256, 0, st>>>(X.ptr.p, X.idx.p, X.val.p, rows.p, nr, Y.ptr.p, Y.idx.p,
                                                                  Y.val.p);
        SB_LAUNCH_CHECK();
    }
    count_launch(dst, 2);
    SB_CUDA(cudaStreamSynchronize(st));
    dst->n_local = nr;
    dst->n_global = n_global_dst;
    dst->row0 = row0_dst;
    dst->m = src->m;
    dst->loaded = true;
    dst->stats.nnz_local = nnz;
}

}  // namespace snapb

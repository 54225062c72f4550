// Device-side synthetic planted-cluster pattern generator.
// Integer-only sampling so the output is bit-identical to the numpy generator
// in snapatac2_b200/synth.py for the same tables.
#include "ctx.cuh"

#include <algorithm>

namespace snapb {

namespace {

constexpr int kGenThreads = 256;
constexpr int kMaxDraws = 8192;

struct GenTables {
    const uint64_t* feat_cdf;     // m + 1
    const uint64_t* cluster_cdf;  // K + 1
    const int64_t* block_start;   // K + 1
    const uint64_t* alpha;        // K
};

// largest j with cdf[j] <= target (cdf non-decreasing, cdf[0] = 0)
__device__ __forceinline__ int64_t cdf_find(const uint64_t* __restrict__ cdf, int64_t len, uint64_t target) {
    int64_t lo = 0, hi = len;  // first index with cdf > target lies in (lo, hi]
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (__ldg(cdf + mid) <= target) lo = mid + 1; else hi = mid;
    }
    return lo - 1;
}

__global__ void __launch_bounds__(kGenThreads)
gen_rows_kernel(GenTables tb, int64_t n_local, int64_t row0, int64_t m, int nnz_row, int pow2, int n_clusters,
                uint64_t seed, int32_t* __restrict__ tmp, int32_t* __restrict__ row_len) {
    extern __shared__ uint32_t keys[];  // pow2 entries
    __shared__ int s_cluster;
    __shared__ int s_warp_tot[kGenThreads / 32];

    for (int64_t lr = blockIdx.x; lr < n_local; lr += gridDim.x) {
        const uint64_t row = static_cast<uint64_t>(row0 + lr);
        if (threadIdx.x == 0) {
            uint64_t u = mix64(seed, row, 1ull << 40) >> 32;
            s_cluster = static_cast<int>(cdf_find(tb.cluster_cdf, n_clusters + 1, u));
        }
        __syncthreads();
        const int z = s_cluster;
        const int64_t b_lo = tb.block_start[z], b_hi = tb.block_start[z + 1];
        const uint64_t f_lo = tb.feat_cdf[b_lo], f_hi = tb.feat_cdf[b_hi];
        const uint64_t a = tb.alpha[z];

        for (int t = threadIdx.x; t < pow2; t += kGenThreads) {
            uint32_t key = 0xFFFFFFFFu;
            if (t < nnz_row) {
                uint64_t h = mix64(seed, row, static_cast<uint64_t>(t));
                uint64_t sel = h >> 32, pos = h & 0xFFFFFFFFull;
                bool priv = sel >= a;
                uint64_t target = priv ? f_lo + ((pos * (f_hi - f_lo)) >> 32) : pos;
                int64_t col = cdf_find(tb.feat_cdf, m + 1, target);
                if (priv) col = max(b_lo, min(col, b_hi - 1));
                else col = max(static_cast<int64_t>(0), min(col, m - 1));
                key = static_cast<uint32_t>(col);
            }
            keys[t] = key;
        }
        __syncthreads();

        // bitonic sort (ascending) of pow2 keys
        for (int k = 2; k <= pow2; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = threadIdx.x; t < (pow2 >> 1); t += kGenThreads) {
                    int i = 2 * t - (t & (j - 1));  // index with bit j clear
                    int p = i + j;
                    bool up = ((i & k) == 0);
                    uint32_t x = keys[i], y = keys[p];
                    if ((x > y) == up) { keys[i] = y; keys[p] = x; }
                }
                __syncthreads();
            }
        }

        // dedup + compaction: thread t owns items [t*per, (t+1)*per)
        const int per = pow2 / kGenThreads > 0 ? pow2 / kGenThreads : 1;
        const int beg = threadIdx.x * per;
        int cnt = 0;
        for (int q = 0; q < per; ++q) {
            int i = beg + q;
            if (i < nnz_row && (i == 0 || keys[i] != keys[i - 1])) ++cnt;
        }
        // block exclusive scan of cnt
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int v = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += v;
        }
        if ((threadIdx.x & 31) == 31) s_warp_tot[threadIdx.x >> 5] = incl;
        __syncthreads();
        int warp_off = 0, total = 0;
        for (int w = 0; w < kGenThreads / 32; ++w) {
            int v = s_warp_tot[w];
            if (w < (threadIdx.x >> 5)) warp_off += v;
            total += v;
        }
        int pos = warp_off + incl - cnt;
        int32_t* out = tmp + lr * static_cast<int64_t>(nnz_row);
        for (int q = 0; q < per; ++q) {
            int i = beg + q;
            if (i < nnz_row && (i == 0 || keys[i] != keys[i - 1])) out[pos++] = static_cast<int32_t>(keys[i]);
        }
        if (threadIdx.x == 0) row_len[lr] = total;
        __syncthreads();
    }
}

__global__ void compact_rows_kernel(const int32_t* __restrict__ tmp, const int64_t* __restrict__ ptr,
                                    int64_t n_local, int nnz_row, int32_t* __restrict__ idx) {
    // one warp per row
    int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    for (int64_t r = warp; r < n_local; r += nwarps) {
        int64_t s = ptr[r];
        int len = static_cast<int>(ptr[r + 1] - s);
        const int32_t* src = tmp + r * static_cast<int64_t>(nnz_row);
        for (int q = lane; q < len; q += 32) idx[s + q] = src[q];
    }
}

}  // namespace

void generate_rows(snapb200_ctx* c, int64_t n_local, int64_t n_global, int64_t row0, int64_t m, int nnz_row,
                   int n_clusters, uint64_t seed, const uint64_t* feat_cdf, const uint64_t* cluster_cdf,
                   const int64_t* block_start, const uint64_t* alpha) {
    SB_CHECK(nnz_row >= 1 && nnz_row <= kMaxDraws, "nnz_row must be in [1, 8192]");
    SB_CHECK(m >= 1 && m < (1ll << 31), "m must be < 2^31");
    SB_CHECK(n_clusters >= 1 && n_clusters <= 4096, "bad n_clusters");
    SB_CHECK(n_local >= 0 && row0 >= 0 && row0 + n_local <= n_global, "bad shard geometry");

    DevBuf<uint64_t> d_feat, d_ccdf, d_alpha;
    DevBuf<int64_t> d_block;
    d_feat.alloc(m + 1);
    d_ccdf.alloc(n_clusters + 1);
    d_alpha.alloc(n_clusters);
    d_block.alloc(n_clusters + 1);
    SB_CUDA(cudaMemcpyAsync(d_feat.p, feat_cdf, sizeof(uint64_t) * (m + 1), cudaMemcpyHostToDevice, c->stream));
    SB_CUDA(cudaMemcpyAsync(d_ccdf.p, cluster_cdf, sizeof(uint64_t) * (n_clusters + 1), cudaMemcpyHostToDevice, c->stream));
    SB_CUDA(cudaMemcpyAsync(d_alpha.p, alpha, sizeof(uint64_t) * n_clusters, cudaMemcpyHostToDevice, c->stream));
    SB_CUDA(cudaMemcpyAsync(d_block.p, block_start, sizeof(int64_t) * (n_clusters + 1), cudaMemcpyHostToDevice, c->stream));

    int pow2 = 32;
    while (pow2 < nnz_row) pow2 <<= 1;
    if (pow2 < kGenThreads) pow2 = kGenThreads;

    DevBuf<int32_t> tmp, len;
    tmp.alloc(n_local * static_cast<int64_t>(nnz_row));
    len.alloc(n_local > 0 ? n_local : 1);

    GenTables tb{d_feat.p, d_ccdf.p, d_block.p, d_alpha.p};
    if (n_local > 0) {
        size_t smem = sizeof(uint32_t) * pow2;
        int blocks = static_cast<int>(std::min<int64_t>(n_local, static_cast<int64_t>(c->num_sms) * 6));
        gen_rows_kernel<<<blocks, kGenThreads, smem, c->stream>>>(tb, n_local, row0, m, nnz_row, pow2, n_clusters,
                                                                   seed, tmp.p, len.p);
        SB_LAUNCH_CHECK();
        count_launch(c);
    }

    Csr& X = c->X;
    X.nrows = n_local;
    X.ncols = m;
    X.val.release();
    X.ptr.alloc(n_local + 1);
    exclusive_scan_i32_to_i64(c, len.p, X.ptr.p, n_local);
    int64_t nnz = 0;
    SB_CUDA(cudaMemcpyAsync(&nnz, X.ptr.p + n_local, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    SB_CUDA(cudaStreamSynchronize(c->stream));
    X.nnz = nnz;
    X.idx.alloc(nnz > 0 ? nnz : 1);
    if (n_local > 0) {
        int blocks = static_cast<int>(std::min<int64_t>(ceil_div(n_local, 8), static_cast<int64_t>(c->num_sms) * 16));
        compact_rows_kernel<<<blocks, 256, 0, c->stream>>>(tmp.p, X.ptr.p, n_local, nnz_row, X.idx.p);
        SB_LAUNCH_CHECK();
        count_launch(c);
    }
    SB_CUDA(cudaStreamSynchronize(c->stream));

    c->n_local = n_local;
    c->n_global = n_global;
    c->row0 = row0;
    c->m = m;
    c->loaded = true;
    c->prepared = false;
    c->proj_ready = false;
    c->nnz_mode = -1;
    c->stats.nnz_local = nnz;
}

}  // namespace snapb

// Builds the column-tiled sliced-ELL copy (struct Sell, ctx.cuh) of a CSR
// matrix for the shared-memory SpMM in spmm_tiled.cu.
//
// Why this layout: at 4 B of index per stored entry the SpMM would be HBM
// bound, but every entry also gathers a dense row of b fp32 (32 B for b = 8).
// Gathered from L2 that is ~5.7 TB/s of sector traffic and caps the kernel at
// ~10% of the HBM roofline (profiles/r01_bench_c3_v1_csr_gather.json).  Staging
// a column tile of the dense operand in shared memory moves the gathers on
// chip; what then limits the kernel is the shared-memory pipe, so the build
// also orders every lane's entries such that the eight lanes of a quarter warp
// read eight different 16-byte bank groups:
//   b = 8: dense row j sits at byte j*32 -> bank groups 2*(j%4), 2*(j%4)+1;
//          lane l reads half (l&1) first, so even lanes need distinct j%4 and
//          odd lanes need distinct j%4 within a quarter warp: 4 classes,
//          lane offset o = (l>>1) & 3;
//   b = 4: dense row j sits at byte j*16 -> bank group j%8: 8 classes,
//          lane offset o = l & 7;
//   lane l consumes class (k + o) % NC at step k: the q-th entry of class c is
//   placed at step NC*q + ((c - o) mod NC) while every class still has
//   entries; the remainder follows in column order.
// The chunk list is tile-major so that a persistent CTA streams one long
// contiguous range of it and touches only one or two tiles.
// Everything is deterministic (sorts with total order, no atomics on data).
#include "ctx.cuh"

#include <algorithm>
#include <vector>

namespace snapb {

namespace {

constexpr int kPlanThreads = 256;
constexpr uint32_t kLenBias = 16383;   // tile_cols <= 12288 < 2^14

// segptr[row*(T+1) + t] = number of entries of `row` with column < t*tile_cols
__global__ void seg_bounds_kernel(const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx, int64_t nrows,
                                  int n_tiles, int tile_cols, int32_t* __restrict__ segptr) {
    const int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t total = nrows * (n_tiles + 1);
    if (g >= total) return;
    const int64_t row = g / (n_tiles + 1);
    const int t = static_cast<int>(g - row * (n_tiles + 1));
    const int64_t s = ptr[row], e = ptr[row + 1];
    const int64_t bound = static_cast<int64_t>(t) * tile_cols;
    int64_t lo = s, hi = e;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (static_cast<int64_t>(idx[mid]) < bound) lo = mid + 1; else hi = mid;
    }
    segptr[g] = static_cast<int32_t>(lo - s);
}

// One CTA per (tile, window): sort the window's rows by segment length
// (descending, ties by row) and cut them into chunks of 32 lanes.
__global__ void __launch_bounds__(kPlanThreads)
sell_plan_kernel(const int64_t* __restrict__ window_start, const int64_t* __restrict__ window_chunk0,
                 const int32_t* __restrict__ segptr, int n_tiles, int n_windows, int64_t chunks_per_tile,
                 int32_t* __restrict__ chunk_rows, int32_t* __restrict__ chunk_len4) {
    __shared__ uint32_t keys[kSellWindowRows];
    const int t = blockIdx.x / n_windows, w = blockIdx.x % n_windows;
    const int64_t r0 = window_start[w];
    const int nr = static_cast<int>(window_start[w + 1] - r0);
    if (nr == 0) return;
    int pow2 = 2;
    while (pow2 < nr) pow2 <<= 1;
    for (int i = threadIdx.x; i < pow2; i += kPlanThreads) {
        uint32_t key = 0xFFFFFFFFu;
        if (i < nr) {
            const int32_t* sp = segptr + (r0 + i) * (n_tiles + 1) + t;
            const int len = sp[1] - sp[0];
            key = ((kLenBias - static_cast<uint32_t>(len)) << 13) | static_cast<uint32_t>(i);
        }
        keys[i] = key;
    }
    __syncthreads();
    for (int k = 2; k <= pow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int q = threadIdx.x; q < (pow2 >> 1); q += kPlanThreads) {
                const int i = 2 * q - (q & (j - 1));
                const int p = i + j;
                const bool up = ((i & k) == 0);
                const uint32_t x = keys[i], y = keys[p];
                if ((x > y) == up) { keys[i] = y; keys[p] = x; }
            }
            __syncthreads();
        }
    }
    const int nch = (nr + 31) / 32;
    const int64_t base = static_cast<int64_t>(t) * chunks_per_tile + window_chunk0[w];
    for (int q = threadIdx.x; q < nch * 32; q += kPlanThreads) {
        const bool valid = q < nr;
        const uint32_t key = valid ? keys[q] : 0u;
        chunk_rows[base * 32 + q] = valid ? static_cast<int32_t>(r0 + (key & 8191u)) : -1;
        if ((q & 31) == 0) {
            const int len = static_cast<int>(kLenBias - (key >> 13));   // longest segment of the chunk
            chunk_len4[base + (q >> 5)] = (len + 3) >> 2;
        }
    }
}

// Per-lane class counters packed 16 bits each into 64-bit words (register only).
template <int NC>
struct ClassCounters {
    uint64_t w[NC / 4];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int i = 0; i < NC / 4; ++i) w[i] = 0;
    }
    // returns the count before the increment
    __device__ __forceinline__ int bump(int cl) {
        const int sh = 16 * (cl & 3);
        int old = 0;
#pragma unroll
        for (int i = 0; i < NC / 4; ++i)
            if ((cl >> 2) == i) {
                old = static_cast<int>((w[i] >> sh) & 0xFFFFull);
                w[i] += 1ull << sh;
            }
        return old;
    }
    __device__ __forceinline__ int min_count() const {
        int m = 0x7fffffff;
#pragma unroll
        for (int i = 0; i < NC / 4; ++i)
#pragma unroll
            for (int s = 0; s < 4; ++s) m = min(m, static_cast<int>((w[i] >> (16 * s)) & 0xFFFFull));
        return m;
    }
};

// One warp per chunk: every lane lays its row segment out in the bank-conflict
// free order described at the top of the file.  NC = classes (4 for b=8, 8 for b=4).
template <bool HAS_VAL, int NC>
__global__ void __launch_bounds__(256)
sell_fill_kernel(const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx, const float* __restrict__ val,
                 const int32_t* __restrict__ segptr, int n_tiles, int tile_cols, int row_bytes, int64_t n_chunks,
                 int64_t chunks_per_tile, const int32_t* __restrict__ chunk_rows,
                 const int32_t* __restrict__ chunk_len4, const int64_t* __restrict__ chunk_off,
                 int32_t* __restrict__ data, float* __restrict__ vals) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    const int o = (NC == 4) ? ((lane >> 1) & 3) : (lane & 7);
    for (int64_t c = warp; c < n_chunks; c += nwarps) {
        const int steps = chunk_len4[c] * 4;
        if (steps == 0) continue;
        const int row = chunk_rows[c * 32 + lane];
        const int t = static_cast<int>(c / chunks_per_tile);
        int32_t* d = data + chunk_off[c] * 128 + lane * 4;
        float* dv = HAS_VAL ? vals + chunk_off[c] * 128 + lane * 4 : nullptr;
        int64_t s = 0, e = 0;
        if (row >= 0) {
            const int32_t* sp = segptr + static_cast<int64_t>(row) * (n_tiles + 1) + t;
            s = ptr[row] + sp[0];
            e = ptr[row] + sp[1];
        }
        ClassCounters<NC> cnt;
        cnt.clear();
        for (int64_t p = s; p < e; ++p) cnt.bump(idx[p] & (NC - 1));
        const int mmin = cnt.min_count();
        cnt.clear();
        int left = 0;
        const int col0 = t * tile_cols;
        for (int64_t p = s; p < e; ++p) {
            const int j = idx[p];
            const int cl = j & (NC - 1);
            const int q = cnt.bump(cl);
            const int k = (q < mmin) ? NC * q + ((cl - o) & (NC - 1)) : NC * mmin + left++;
            const int64_t pos = static_cast<int64_t>(k >> 2) * 128 + (k & 3);
            d[pos] = (j - col0) * row_bytes;
            if (HAS_VAL) dv[pos] = val[p];
        }
        for (int k = static_cast<int>(e - s); k < steps; ++k) {
            const int64_t pos = static_cast<int64_t>(k >> 2) * 128 + (k & 3);
            d[pos] = -1;
            if (HAS_VAL) dv[pos] = 0.f;
        }
    }
}

}  // namespace

void sell_build(snapb200_ctx* c, const Csr& M, Sell& S, int b) {
    SB_CHECK(b == 4 || b == 8, "tiled format: block width must be 4 or 8");
    cudaStream_t st = c->stream;
    S.clear();
    const int64_t R = M.nrows;
    S.b = b;
    S.nrows = R;
    S.ncols = M.ncols;
    S.tile_cols = kSellTileBytes / (4 * b);
    S.n_tiles = static_cast<int>(std::max<int64_t>(1, ceil_div(M.ncols, S.tile_cols)));
    const int T = S.n_tiles;

    // ---- windows of <= 8192 consecutive rows (equal row counts)
    const int nw = static_cast<int>(std::max<int64_t>(1, ceil_div(R, kSellWindowRows)));
    S.n_windows = nw;
    std::vector<int64_t> ws(nw + 1), wc(nw + 1);
    wc[0] = 0;
    for (int w = 0; w <= nw; ++w) ws[w] = (R * w) / nw;
    for (int w = 0; w < nw; ++w) wc[w + 1] = wc[w] + ceil_div(ws[w + 1] - ws[w], 32);
    S.chunks_per_tile = wc[nw];
    S.n_chunks = S.chunks_per_tile * T;
    S.window_start.alloc(nw + 1);
    S.window_chunk0.alloc(nw + 1);
    SB_CUDA(cudaMemcpyAsync(S.window_start.p, ws.data(), sizeof(int64_t) * (nw + 1), cudaMemcpyHostToDevice, st));
    SB_CUDA(cudaMemcpyAsync(S.window_chunk0.p, wc.data(), sizeof(int64_t) * (nw + 1), cudaMemcpyHostToDevice, st));
    const int64_t n_chunks = S.n_chunks;

    // ---- per-row tile boundaries
    DevBuf<int32_t> segptr;
    const int64_t nseg = R * (T + 1);
    segptr.alloc(std::max<int64_t>(1, nseg));
    if (nseg > 0) {
        seg_bounds_kernel<<<static_cast<unsigned>(ceil_div(nseg, 256)), 256, 0, st>>>(M.ptr.p, M.idx.p, R, T,
                                                                                   S.tile_cols, segptr.p);
        SB_LAUNCH_CHECK();
    }
    // ---- plan: sorted chunk membership and chunk lengths, then offsets
    S.chunk_rows.alloc(std::max<int64_t>(1, n_chunks * 32));
    S.chunk_len4.alloc(std::max<int64_t>(1, n_chunks));
    S.chunk_off.alloc(n_chunks + 1);
    if (n_chunks > 0) {
        sell_plan_kernel<<<static_cast<unsigned>(static_cast<int64_t>(T) * nw), kPlanThreads, 0, st>>>(
            S.window_start.p, S.window_chunk0.p, segptr.p, T, nw, S.chunks_per_tile, S.chunk_rows.p, S.chunk_len4.p);
        SB_LAUNCH_CHECK();
    }
    exclusive_scan_i32_to_i64(c, S.chunk_len4.p, S.chunk_off.p, n_chunks);
    int64_t n_groups = 0;
    SB_CUDA(cudaMemcpyAsync(&n_groups, S.chunk_off.p + n_chunks, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));   // also keeps ws / wc alive until the copies are done
    S.n_entries = n_groups * 128;

    // ---- fill
    S.data.alloc(std::max<int64_t>(4, S.n_entries));
    if (M.has_values()) S.vals.alloc(std::max<int64_t>(4, S.n_entries));
    if (n_chunks > 0 && S.n_entries > 0) {
        const int blocks = static_cast<int>(std::min<int64_t>(ceil_div(n_chunks, 8), static_cast<int64_t>(c->num_sms) * 32));
        const int rb = 4 * b;
#define SB_FILL(HV, NC)                                                                                            \
    sell_fill_kernel<HV, NC><<<blocks, 256, 0, st>>>(M.ptr.p, M.idx.p, M.val.p, segptr.p, T, S.tile_cols, rb, n_chunks, \
                                                     S.chunks_per_tile, S.chunk_rows.p, S.chunk_len4.p, S.chunk_off.p, \
                                                     S.data.p, S.vals.p)
        if (M.has_values()) { if (b == 8) SB_FILL(true, 4); else SB_FILL(true, 8); }
        else                { if (b == 8) SB_FILL(false, 4); else SB_FILL(false, 8); }
#undef SB_FILL
        SB_LAUNCH_CHECK();
    }
    count_launch(c, 3);
    S.built = true;
}

}  // namespace snapb

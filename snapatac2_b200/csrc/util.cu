// Small device utilities: exclusive scans, fills, L2 flush.
#include "ctx.cuh"

#include <algorithm>

namespace snapb {

namespace {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;                        // per thread
constexpr int kScanTile = kScanThreads * kScanItems;  // per block

template <typename In>
__global__ void scan_tile_sums(const In* __restrict__ in, int64_t n, int64_t* __restrict__ tile_sums) {
    __shared__ int64_t warp_part[kScanThreads / 32];
    int64_t base = static_cast<int64_t>(blockIdx.x) * kScanTile;
    int64_t s = 0;
    for (int k = 0; k < kScanItems; ++k) {
        int64_t i = base + k * kScanThreads + threadIdx.x;
        if (i < n) s += static_cast<int64_t>(in[i]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t t = 0;
        for (int w = 0; w < kScanThreads / 32; ++w) t += warp_part[w];
        tile_sums[blockIdx.x] = t;
    }
}

// One block: exclusive scan of the tile sums in place (sequential chunks per
// thread, then a block scan of the per-thread totals).
__global__ void scan_tile_offsets(int64_t* __restrict__ tile_sums, int64_t n_tiles) {
    __shared__ int64_t part[1024];
    int t = threadIdx.x;
    int64_t per = (n_tiles + blockDim.x - 1) / blockDim.x;
    int64_t lo = min(n_tiles, static_cast<int64_t>(t) * per);
    int64_t hi = min(n_tiles, lo + per);
    int64_t s = 0;
    for (int64_t i = lo; i < hi; ++i) s += tile_sums[i];
    part[t] = s;
    __syncthreads();
    if (t == 0) {
        int64_t run = 0;
        for (int i = 0; i < blockDim.x; ++i) {
            int64_t v = part[i];
            part[i] = run;
            run += v;
        }
    }
    __syncthreads();
    int64_t run = part[t];
    for (int64_t i = lo; i < hi; ++i) {
        int64_t v = tile_sums[i];
        tile_sums[i] = run;
        run += v;
    }
}

template <typename In>
__global__ void scan_apply(const In* __restrict__ in, int64_t n, const int64_t* __restrict__ tile_offsets,
                           int64_t* __restrict__ out) {
    // thread t owns items [t*kScanItems, (t+1)*kScanItems) of the tile
    __shared__ int64_t part[kScanThreads];
    int64_t base = static_cast<int64_t>(blockIdx.x) * kScanTile + static_cast<int64_t>(threadIdx.x) * kScanItems;
    int64_t local[kScanItems];
    int64_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        int64_t i = base + k;
        int64_t v = (i < n) ? static_cast<int64_t>(in[i]) : 0;
        local[k] = s;
        s += v;
    }
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t run = tile_offsets[blockIdx.x];
        for (int i = 0; i < kScanThreads; ++i) {
            int64_t v = part[i];
            part[i] = run;
            run += v;
        }
        if (static_cast<int64_t>(blockIdx.x + 1) * kScanTile >= n) out[n] = run;
    }
    __syncthreads();
    int64_t off = part[threadIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        int64_t i = base + k;
        if (i < n) out[i] = off + local[k];
    }
}

template <typename In>
void exclusive_scan_impl(snapb200_ctx* c, const In* in, int64_t* out, int64_t n) {
    if (n == 0) {
        SB_CUDA(cudaMemsetAsync(out, 0, sizeof(int64_t), c->stream));
        return;
    }
    int64_t n_tiles = ceil_div(n, kScanTile);
    DevBuf<int64_t> tiles;
    tiles.alloc(n_tiles);
    scan_tile_sums<In><<<static_cast<unsigned>(n_tiles), kScanThreads, 0, c->stream>>>(in, n, tiles.p);
    SB_LAUNCH_CHECK();
    scan_tile_offsets<<<1, 1024, 0, c->stream>>>(tiles.p, n_tiles);
    SB_LAUNCH_CHECK();
    scan_apply<In><<<static_cast<unsigned>(n_tiles), kScanThreads, 0, c->stream>>>(in, n, tiles.p, out);
    SB_LAUNCH_CHECK();
    count_launch(c, 3);
    SB_CUDA(cudaStreamSynchronize(c->stream));  // `tiles` is freed on return
}

__global__ void fill_kernel(float* p, float v, int64_t n) {
    int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}

}  // namespace

void exclusive_scan_i64(snapb200_ctx* c, const int64_t* in, int64_t* out, int64_t n) {
    exclusive_scan_impl<int64_t>(c, in, out, n);
}
void exclusive_scan_i32_to_i64(snapb200_ctx* c, const int32_t* in, int64_t* out, int64_t n) {
    exclusive_scan_impl<int32_t>(c, in, out, n);
}

void fill_f32(snapb200_ctx* c, float* p, float v, int64_t n) {
    if (n == 0) return;
    int blocks = static_cast<int>(std::min<int64_t>(ceil_div(n, 256), c->num_sms * 8));
    fill_kernel<<<blocks, 256, 0, c->stream>>>(p, v, n);
    SB_LAUNCH_CHECK();
    count_launch(c);
}

// Evict everything from L2 by overwriting a buffer larger than the cache.
void flush_l2(snapb200_ctx* c) {
    const int64_t bytes = 256ll << 20;
    c->scratch.ensure(bytes);
    SB_CUDA(cudaMemsetAsync(c->scratch.p, 0xA5, static_cast<size_t>(bytes), c->stream));
}

}  // namespace snapb

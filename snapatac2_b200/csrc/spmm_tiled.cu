// Shared-memory tiled SpMM over the sliced-ELL copy built by sell_build.cu
// (b = 4 or 8 dense columns):
//     partial[t][row, 0:b] = sum over the row's entries in column tile t of in[col, 0:b]
//     out[row, 0:b]        = scale[row] * sum_t partial[t][row, 0:b]  (- subscale[row] * sub[row, 0:b])
//
// sell_spmm_kernel: one persistent CTA per SM.  The tile-major chunk list is
// cut into gridDim.x ranges of equal stored entries; a CTA walks its range,
// which touches one or two column tiles.  For each of them it stages the dense
// operand tile (192 KB: 6144 rows of 32 B or 12288 rows of 16 B) in shared
// memory with one TMA bulk copy (cp.async.bulk + mbarrier), then its 24 warps
// pull chunks from a shared counter: one lane per row segment, eight 16-bit
// entries per 128-bit load (512 contiguous bytes per warp-load, double buffered
// in registers), b/4 LDS.128 per entry.  The build orders entries so the
// LDS.128 of a quarter warp are (mostly) bank-conflict free.  Every
// (tile, row) partial is written exactly once with a plain store: no atomics,
// bitwise reproducible.
//
// Bytes per stored entry: 2 from HBM (the 16-bit tile-local column; CSR's int32
// index is the 4-byte algorithmic figure the roofline line of bench.py uses) and
// 4b bytes of shared-memory bandwidth, which is what bounds the kernel.
#include "ctx.cuh"

#include <stdlib.h>
#include <algorithm>

namespace snapb {

namespace {

constexpr int kTiledThreads = 768;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    const uint32_t addr = smem_u32(bar);
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(addr), "r"(phase) : "memory");
    }
}
// 1-D TMA bulk copy global -> shared, completion signalled on the mbarrier
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float4 ld_stream_float4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_float4(float4* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ float4 shfl_down4(const float4& v, int d) {
    return make_float4(__shfl_down_sync(0xffffffffu, v.x, d), __shfl_down_sync(0xffffffffu, v.y, d),
                       __shfl_down_sync(0xffffffffu, v.z, d), __shfl_down_sync(0xffffffffu, v.w, d));
}
__device__ __forceinline__ void acc4(float4& a, const float4& x, float v, bool has_val) {
    if (has_val) {
        a.x = fmaf(v, x.x, a.x); a.y = fmaf(v, x.y, a.y); a.z = fmaf(v, x.z, a.z); a.w = fmaf(v, x.w, a.w);
    } else {
        a.x += x.x; a.y += x.y; a.z += x.z; a.w += x.w;
    }
}

// one stored entry (16-bit tile-local column, 0xFFFF = empty slot): B/4 LDS.128 from the staged tile
template <int B, bool HAS_VAL>
__device__ __forceinline__ void gather_one(uint32_t col, float v, uint32_t tile1, uint32_t tile2, float4& a, float4& b) {
    if (col != 0xFFFFu) {
        const uint32_t off = col * (4u * B);
        const float4 x1 = lds128(tile1 + off);
        acc4(a, x1, v, HAS_VAL);
        if (B == 8) {
            const float4 x2 = lds128(tile2 + off);
            acc4(b, x2, v, HAS_VAL);
        }
    }
}
template <int B, bool HAS_VAL>
__device__ __forceinline__ void gather_eight(const int4& e, const float4& v0, const float4& v1, uint32_t tile1,
                                             uint32_t tile2, float4& a, float4& b) {
    const uint32_t x = static_cast<uint32_t>(e.x), y = static_cast<uint32_t>(e.y);
    const uint32_t z = static_cast<uint32_t>(e.z), w = static_cast<uint32_t>(e.w);
    gather_one<B, HAS_VAL>(x & 0xFFFFu, v0.x, tile1, tile2, a, b);
    gather_one<B, HAS_VAL>(x >> 16, v0.y, tile1, tile2, a, b);
    gather_one<B, HAS_VAL>(y & 0xFFFFu, v0.z, tile1, tile2, a, b);
    gather_one<B, HAS_VAL>(y >> 16, v0.w, tile1, tile2, a, b);
    gather_one<B, HAS_VAL>(z & 0xFFFFu, v1.x, tile1, tile2, a, b);
    gather_one<B, HAS_VAL>(z >> 16, v1.y, tile1, tile2, a, b);
    gather_one<B, HAS_VAL>(w & 0xFFFFu, v1.z, tile1, tile2, a, b);
    gather_one<B, HAS_VAL>(w >> 16, v1.w, tile1, tile2, a, b);
}

// End of a chunk: one partial sum per (tile, row).  The lanes of a split row (sell_build.cu) are
// adjacent: they are added up in a fixed order and the first one stores.  Kept out of line so that
// it does not disturb the code generation of the gather loop (predicated LDS.128 in flight).
template <int B>
__device__ __noinline__ void store_partial(float4 a, float4 b, int row, int lane, float* __restrict__ part_t) {
    // canonical halves (b = 8: odd lanes gathered the upper half first)
    float4 lo = (B == 8 && (lane & 1)) ? b : a;
    float4 hi = (B == 8 && (lane & 1)) ? a : b;
    const int nrow = __shfl_down_sync(0xffffffffu, row, 1);
    if (__any_sync(0xffffffffu, row >= 0 && lane < 31 && nrow == row)) {
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int orow = __shfl_down_sync(0xffffffffu, row, d);
            const float4 ol = shfl_down4(lo, d);
            const float4 oh = (B == 8) ? shfl_down4(hi, d) : hi;
            if (row >= 0 && lane + d < 32 && orow == row) {
                lo.x += ol.x; lo.y += ol.y; lo.z += ol.z; lo.w += ol.w;
                if (B == 8) { hi.x += oh.x; hi.y += oh.y; hi.z += oh.z; hi.w += oh.w; }
            }
        }
    }
    const int prow = __shfl_up_sync(0xffffffffu, row, 1);
    if (row >= 0 && (lane == 0 || prow != row)) {
        float4* o = reinterpret_cast<float4*>(part_t + static_cast<int64_t>(row) * B);
        st_stream_float4(o, lo);
        if (B == 8) st_stream_float4(o + 1, hi);
    }
}

// first chunk index c in [0, n] with chunk_off[c] >= target
// Cost of the chunk list up to chunk c: its groups of 8 slots plus a fixed per-chunk charge (fetching
// the chunk's metadata, the epilogue and its store cost about as much as kChunkCost groups); ranges
// of many short chunks would otherwise take longer than ranges of few long ones (ncu on a 1/8
// shard of C3: slowest SM 1.45x the average in pass 1).
constexpr int64_t kChunkCostDefault = 3;
__device__ int64_t chunk_lower_bound(const int64_t* __restrict__ chunk_off, int64_t n, int64_t target, int64_t kChunkCost) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (chunk_off[mid] + kChunkCost * mid < target) lo = mid + 1; else hi = mid;
    }
    return lo;
}
// (SNAPB200_CHUNK_COST overrides the per-chunk charge: tuning aid)
static int64_t chunk_cost() {
    static const int64_t v = [] {
        const char* e = getenv("SNAPB200_CHUNK_COST");
        return e ? static_cast<int64_t>(atoi(e)) : kChunkCostDefault;
    }();
    return v;
}

template <int B, bool HAS_VAL, int U>
__global__ void __launch_bounds__(kTiledThreads, 1)
sell_spmm_kernel(const int32_t* __restrict__ chunk_rows, const int32_t* __restrict__ chunk_groups,
                 const int64_t* __restrict__ chunk_off, const uint16_t* __restrict__ data, const float* __restrict__ vals,
                 const float* __restrict__ in, float* __restrict__ partial, int64_t n_chunks, int64_t chunks_per_tile,
                 int tile_cols, int64_t ncols, int64_t nrows, int64_t kChunkCost) {
    constexpr int RB = 4 * B;   // bytes per dense row
    extern __shared__ __align__(128) unsigned char smem[];
    float* tile = reinterpret_cast<float*>(smem);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(tile_cols) * RB);
    unsigned long long* next_chunk = reinterpret_cast<unsigned long long*>(bar + 1);
    int64_t* range = reinterpret_cast<int64_t*>(bar + 2);   // [2]
    const int lane = threadIdx.x & 31;
    // b = 8: lane reads half (lane & 1) of the dense row first, the other half second
    const uint32_t tile1 = smem_u32(tile) + ((B == 8) ? (lane & 1) * 16 : 0);
    const uint32_t tile2 = smem_u32(tile) + 16 - (lane & 1) * 16;

    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        // this CTA's share of the tile-major chunk list: equal cost
        const int64_t groups = chunk_off[n_chunks] + kChunkCost * n_chunks;
        const int64_t g_lo = static_cast<int64_t>((static_cast<__int128>(groups) * blockIdx.x) / gridDim.x);
        const int64_t g_hi = static_cast<int64_t>((static_cast<__int128>(groups) * (blockIdx.x + 1)) / gridDim.x);
        range[0] = (blockIdx.x == 0) ? 0 : chunk_lower_bound(chunk_off, n_chunks, g_lo, kChunkCost);
        range[1] = (blockIdx.x == gridDim.x - 1) ? n_chunks : chunk_lower_bound(chunk_off, n_chunks, g_hi, kChunkCost);
    }
    __syncthreads();
    const int64_t c_begin = range[0], c_end = range[1];
    if (c_begin >= c_end) return;
    uint32_t phase = 0;

    for (int64_t t = c_begin / chunks_per_tile; t * chunks_per_tile < c_end; ++t) {
        const int64_t lo = max(c_begin, t * chunks_per_tile);
        const int64_t hi = min(c_end, (t + 1) * chunks_per_tile);
        __syncthreads();   // every warp is done with the previous tile and counter
        if (threadIdx.x == 0) {
            *next_chunk = static_cast<unsigned long long>(lo);
            const int64_t c0 = t * tile_cols;
            const uint32_t bytes = static_cast<uint32_t>(min(static_cast<int64_t>(tile_cols), ncols - c0)) * RB;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(bar, bytes);
            tma_load_1d(tile, in + c0 * B, bytes, bar);
        }
        __syncthreads();   // counter visible
        mbar_wait(bar, phase);
        phase ^= 1;
        float* part_t = partial + t * nrows * B;

        while (true) {
            unsigned long long cu = 0;
            if (lane == 0) cu = atomicAdd(next_chunk, 1ull);
            const int64_t c = static_cast<int64_t>(__shfl_sync(0xffffffffu, cu, 0));
            if (c >= hi) break;
            const int row = chunk_rows[c * 32 + lane];
            const int ng = chunk_groups[c];
            const int4* d = reinterpret_cast<const int4*>(data) + chunk_off[c] * 32 + lane;   // a group: 32 lanes x 16 B
            // values: 8 floats per lane and group = two float4 (a group: 64 float4)
            const float4* dv = HAS_VAL ? reinterpret_cast<const float4*>(vals) + chunk_off[c] * 64 + lane * 2 : nullptr;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
            const float4 ones = make_float4(1.f, 1.f, 1.f, 1.f);
            const int4 none = make_int4(-1, -1, -1, -1);
            int4 cur[U], nxt[U];
            float4 vcur[2 * U], vnxt[2 * U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                cur[u] = none;
                vcur[2 * u] = vcur[2 * u + 1] = ones;
                if (u < ng) {
                    cur[u] = ld_stream_int4(d + u * 32);
                    if (HAS_VAL) {
                        vcur[2 * u] = ld_stream_float4(dv + u * 64);
                        vcur[2 * u + 1] = ld_stream_float4(dv + u * 64 + 1);
                    }
                }
            }
            for (int g = 0; g < ng; g += U) {
#pragma unroll
                for (int u = 0; u < U; ++u) {   // prefetch the next U groups while this batch is consumed
                    nxt[u] = none;
                    vnxt[2 * u] = vnxt[2 * u + 1] = ones;
                    if (g + U + u < ng) {
                        nxt[u] = ld_stream_int4(d + (g + U + u) * 32);
                        if (HAS_VAL) {
                            vnxt[2 * u] = ld_stream_float4(dv + (g + U + u) * 64);
                            vnxt[2 * u + 1] = ld_stream_float4(dv + (g + U + u) * 64 + 1);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u)
                    gather_eight<B, HAS_VAL>(cur[u], vcur[2 * u], vcur[2 * u + 1], tile1, tile2, a, b);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    cur[u] = nxt[u];
                    vcur[2 * u] = vnxt[2 * u];
                    vcur[2 * u + 1] = vnxt[2 * u + 1];
                }
            }
            store_partial<B>(a, b, row, lane, part_t);
        }
    }
}

// out float4 g (row = g / (B/4)) = scale[row] * sum_t partial[t] - subscale[row] * sub[row]
template <int B, bool HAS_SUB>
__global__ void __launch_bounds__(256)
reduce_tiles_kernel(const float* __restrict__ partial, int n_tiles, int64_t nrows, const float* __restrict__ scale,
                    const float* __restrict__ subscale, const float* __restrict__ sub, int64_t lds,
                    float* __restrict__ out) {
    constexpr int Q = B / 4;
    const int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;   // float4 index
    if (g >= nrows * Q) return;
    const float4* p = reinterpret_cast<const float4*>(partial) + g;
    const int64_t stride = nrows * Q;
    float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
    int t = 0;
    for (; t + 1 < n_tiles; t += 2) {      // two fixed-order chains for memory-level parallelism
        const float4 x0 = ld_stream_float4(p + static_cast<int64_t>(t) * stride);
        const float4 x1 = ld_stream_float4(p + static_cast<int64_t>(t + 1) * stride);
        s0.x += x0.x; s0.y += x0.y; s0.z += x0.z; s0.w += x0.w;
        s1.x += x1.x; s1.y += x1.y; s1.z += x1.z; s1.w += x1.w;
    }
    if (t < n_tiles) {
        const float4 x0 = ld_stream_float4(p + static_cast<int64_t>(t) * stride);
        s0.x += x0.x; s0.y += x0.y; s0.z += x0.z; s0.w += x0.w;
    }
    const int64_t row = g / Q;
    const float sc = scale[row];
    float4 y = make_float4(sc * (s0.x + s1.x), sc * (s0.y + s1.y), sc * (s0.z + s1.z), sc * (s0.w + s1.w));
    if (HAS_SUB) {
        const float ss = subscale[row];
        const float* sp = sub + row * lds + (g % Q) * 4;
        y.x -= ss * sp[0]; y.y -= ss * sp[1]; y.z -= ss * sp[2]; y.w -= ss * sp[3];
    }
    reinterpret_cast<float4*>(out)[g] = y;
}

// ---- fp64 SpMV through the same format (prepare: row norms, column sums, degrees) ----
// The dense operand is one fp64 vector, staged as 8-byte rows (tile_cols * 8 bytes per tile).
// partial64[t][row] is written once.
//   SQ: the stored value is squared (row norms of the weighted rows).
__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
template <bool HAS_VAL, bool SQ>
__device__ __forceinline__ void gather64(uint32_t col, float v, uint32_t tile, double& a) {
    if (col != 0xFFFFu) {
        const double x = lds_f64(tile + col * 8u);
        if (HAS_VAL) {
            double dv = static_cast<double>(v);
            if (SQ) dv *= dv;
            a = fma(dv, x, a);
        } else {
            a += x;
        }
    }
}

template <bool HAS_VAL, bool SQ>
__global__ void __launch_bounds__(kTiledThreads, 1)
sell_spmv64_kernel(const int32_t* __restrict__ chunk_rows, const int32_t* __restrict__ chunk_groups,
                   const int64_t* __restrict__ chunk_off, const uint16_t* __restrict__ data, const float* __restrict__ vals,
                   const double* __restrict__ in, double* __restrict__ partial, int64_t n_chunks, int64_t chunks_per_tile,
                   int tile_cols, int64_t ncols, int64_t nrows, int64_t kChunkCost) {
    constexpr int U = 2;
    extern __shared__ __align__(128) unsigned char smem[];
    double* tile = reinterpret_cast<double*>(smem);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(tile_cols) * 8);
    unsigned long long* next_chunk = reinterpret_cast<unsigned long long*>(bar + 1);
    int64_t* range = reinterpret_cast<int64_t*>(bar + 2);
    const int lane = threadIdx.x & 31;
    const uint32_t tile_a = smem_u32(tile);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        const int64_t groups = chunk_off[n_chunks] + kChunkCost * n_chunks;
        const int64_t g_lo = static_cast<int64_t>((static_cast<__int128>(groups) * blockIdx.x) / gridDim.x);
        const int64_t g_hi = static_cast<int64_t>((static_cast<__int128>(groups) * (blockIdx.x + 1)) / gridDim.x);
        range[0] = (blockIdx.x == 0) ? 0 : chunk_lower_bound(chunk_off, n_chunks, g_lo, kChunkCost);
        range[1] = (blockIdx.x == gridDim.x - 1) ? n_chunks : chunk_lower_bound(chunk_off, n_chunks, g_hi, kChunkCost);
    }
    __syncthreads();
    const int64_t c_begin = range[0], c_end = range[1];
    if (c_begin >= c_end) return;
    uint32_t phase = 0;
    for (int64_t t = c_begin / chunks_per_tile; t * chunks_per_tile < c_end; ++t) {
        const int64_t lo = max(c_begin, t * chunks_per_tile);
        const int64_t hi = min(c_end, (t + 1) * chunks_per_tile);
        __syncthreads();
        if (threadIdx.x == 0) {
            *next_chunk = static_cast<unsigned long long>(lo);
            const int64_t c0 = t * tile_cols;
            // bulk copies move multiples of 16 bytes: an odd tail reads one double of slack (callers pad x)
            const uint32_t bytes = (static_cast<uint32_t>(min(static_cast<int64_t>(tile_cols), ncols - c0)) * 8 + 15u) & ~15u;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(bar, bytes);
            tma_load_1d(tile, in + c0, bytes, bar);
        }
        __syncthreads();
        mbar_wait(bar, phase);
        phase ^= 1;
        double* part_t = partial + t * nrows;
        while (true) {
            unsigned long long cu = 0;
            if (lane == 0) cu = atomicAdd(next_chunk, 1ull);
            const int64_t c = static_cast<int64_t>(__shfl_sync(0xffffffffu, cu, 0));
            if (c >= hi) break;
            const int row = chunk_rows[c * 32 + lane];
            const int ng = chunk_groups[c];
            const int4* d = reinterpret_cast<const int4*>(data) + chunk_off[c] * 32 + lane;
            const float4* dv = HAS_VAL ? reinterpret_cast<const float4*>(vals) + chunk_off[c] * 64 + lane * 2 : nullptr;
            double a0 = 0.0, a1 = 0.0;   // two fixed chains
            for (int g = 0; g < ng; g += U) {
                int4 e[U];
                float4 v[2 * U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    e[u] = make_int4(-1, -1, -1, -1);
                    v[2 * u] = v[2 * u + 1] = make_float4(1.f, 1.f, 1.f, 1.f);
                    if (g + u < ng) {
                        e[u] = ld_stream_int4(d + (g + u) * 32);
                        if (HAS_VAL) {
                            v[2 * u] = ld_stream_float4(dv + (g + u) * 64);
                            v[2 * u + 1] = ld_stream_float4(dv + (g + u) * 64 + 1);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const uint32_t x = static_cast<uint32_t>(e[u].x), y = static_cast<uint32_t>(e[u].y);
                    const uint32_t z = static_cast<uint32_t>(e[u].z), w = static_cast<uint32_t>(e[u].w);
                    gather64<HAS_VAL, SQ>(x & 0xFFFFu, v[2 * u].x, tile_a, a0);
                    gather64<HAS_VAL, SQ>(x >> 16, v[2 * u].y, tile_a, a1);
                    gather64<HAS_VAL, SQ>(y & 0xFFFFu, v[2 * u].z, tile_a, a0);
                    gather64<HAS_VAL, SQ>(y >> 16, v[2 * u].w, tile_a, a1);
                    gather64<HAS_VAL, SQ>(z & 0xFFFFu, v[2 * u + 1].x, tile_a, a0);
                    gather64<HAS_VAL, SQ>(z >> 16, v[2 * u + 1].y, tile_a, a1);
                    gather64<HAS_VAL, SQ>(w & 0xFFFFu, v[2 * u + 1].z, tile_a, a0);
                    gather64<HAS_VAL, SQ>(w >> 16, v[2 * u + 1].w, tile_a, a1);
                }
            }
            double acc = a0 + a1;
            const int nrow = __shfl_down_sync(0xffffffffu, row, 1);
            if (__any_sync(0xffffffffu, row >= 0 && lane < 31 && nrow == row)) {   // split rows (see sell_spmm_kernel)
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int orow = __shfl_down_sync(0xffffffffu, row, d);
                    const double oa = __shfl_down_sync(0xffffffffu, acc, d);
                    if (row >= 0 && lane + d < 32 && orow == row) acc += oa;
                }
            }
            const int prow = __shfl_up_sync(0xffffffffu, row, 1);
            if (row >= 0 && (lane == 0 || prow != row)) part_t[row] = acc;
        }
    }
}

// mode 0: out = sqrt(sum_t partial)    mode 1: out = scale * sum_t partial + shift
__global__ void __launch_bounds__(256)
reduce_tiles64_kernel(const double* __restrict__ partial, int n_tiles, int64_t nrows, int mode,
                      const double* __restrict__ scale, double shift, double* __restrict__ out) {
    const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    double s = 0.0;
    for (int t = 0; t < n_tiles; ++t) s += partial[static_cast<int64_t>(t) * nrows + r];
    out[r] = (mode == 0) ? sqrt(s) : (scale ? scale[r] : 1.0) * s + shift;
}

template <int B>
void spmm_impl(snapb200_ctx* c, const Sell& S, const float* in, float* out, const float* scale, const float* subscale,
               const float* sub, int64_t lds) {
    cudaStream_t st = c->stream;
    c->partial.ensure(static_cast<int64_t>(S.n_tiles) * S.nrows * B);
    const size_t smem = static_cast<size_t>(S.tile_cols) * 4 * B + 64;
    const int grid = c->num_sms;
    if (S.vals.p) {
        auto k = sell_spmm_kernel<B, true, 1>;
        SB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        k<<<grid, kTiledThreads, smem, st>>>(S.chunk_rows.p, S.chunk_groups.p, S.chunk_off.p, S.data.p, S.vals.p, in,
                                            c->partial.p, S.n_chunks, S.chunks_per_tile, S.tile_cols, S.ncols, S.nrows, chunk_cost());
    } else {
        auto k = sell_spmm_kernel<B, false, 2>;
        SB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        k<<<grid, kTiledThreads, smem, st>>>(S.chunk_rows.p, S.chunk_groups.p, S.chunk_off.p, S.data.p, nullptr, in,
                                            c->partial.p, S.n_chunks, S.chunks_per_tile, S.tile_cols, S.ncols, S.nrows, chunk_cost());
    }
    SB_LAUNCH_CHECK();
    const unsigned rb = static_cast<unsigned>(ceil_div(S.nrows * (B / 4), 256));
    if (sub)
        reduce_tiles_kernel<B, true><<<rb, 256, 0, st>>>(c->partial.p, S.n_tiles, S.nrows, scale, subscale, sub, lds, out);
    else
        reduce_tiles_kernel<B, false><<<rb, 256, 0, st>>>(c->partial.p, S.n_tiles, S.nrows, scale, nullptr, nullptr, 0, out);
    SB_LAUNCH_CHECK();
    count_launch(c, 2);
}

}  // namespace

bool use_tiled(const snapb200_ctx* c, int b) {
    if (b != 8 && b != 4) return false;
    if (c->spmm_mode == 1) return false;
    if (c->spmm_mode == 2) return true;
    const int64_t nnz = c->nnz_mode >= 0 ? c->nnz_mode : c->X.nnz;
    return nnz >= (1ll << 25);   // small problems: the CSR kernel avoids staging tiles at all
}

void decide_spmm_mode(snapb200_ctx* c) {
    int64_t nnz = c->X.nnz;
    if (c->nranks > 1) {
        DevBuf<int64_t> t;
        t.alloc(1);
        SB_CUDA(cudaMemcpyAsync(t.p, &nnz, sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
        allreduce_i64(c, t.p, 1);
        SB_CUDA(cudaMemcpyAsync(&nnz, t.p, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
        SB_CUDA(cudaStreamSynchronize(c->stream));
        nnz /= c->nranks;
    }
    c->nnz_mode = nnz;
}

void ensure_tiled(snapb200_ctx* c, int b) {
    if (c->S1.built && c->S2.built && c->S1.b == b && c->S2.b == b) return;
    // block width differs from the one prepare() built for: rebuild both copies
    cudaStream_t st = c->stream;
    SB_CUDA(cudaEventRecord(c->ev0, st));
    c->S1.clear();
    c->S2.clear();
    sell_build(c, c->X, c->S2, b);
    transpose_tiled(c, kSellTileBytes / (4 * b), nullptr);
    sell_build_transposed(c, c->XtT, c->n_local, c->S1, b);
    c->XtT.clear();
    SB_CUDA(cudaEventRecord(c->ev1, st));
    SB_CUDA(cudaEventSynchronize(c->ev1));
    float ms = 0.f;
    SB_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->stats.ms_format = ms;
}

void sell_spmv64(snapb200_ctx* c, const Sell& S, const double* x, int mode, const double* scale, double shift,
                 double* out) {
    SB_CHECK(S.built, "tiled SpMV: format not built");
    if (S.nrows == 0) return;
    cudaStream_t st = c->stream;
    // the fp32 partial buffer doubles as the fp64 one (n_tiles x nrows doubles)
    c->partial.ensure(static_cast<int64_t>(S.n_tiles) * S.nrows * 2);
    double* part = reinterpret_cast<double*>(c->partial.p);
    const size_t smem = static_cast<size_t>(S.tile_cols) * 8 + 64;
    const bool sq = mode == 0;
#define SB_SPMV64(HV, SQ)                                                                                              \
    do {                                                                                                               \
        auto k = sell_spmv64_kernel<HV, SQ>;                                                                           \
        SB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));         \
        k<<<c->num_sms, kTiledThreads, smem, st>>>(S.chunk_rows.p, S.chunk_groups.p, S.chunk_off.p, S.data.p,          \
                                                   S.vals.p, x, part, S.n_chunks, S.chunks_per_tile, S.tile_cols,      \
                                                   S.ncols, S.nrows, chunk_cost());                                    \
    } while (0)
    if (S.vals.p) { if (sq) SB_SPMV64(true, true); else SB_SPMV64(true, false); }
    else          SB_SPMV64(false, false);
#undef SB_SPMV64
    SB_LAUNCH_CHECK();
    reduce_tiles64_kernel<<<static_cast<unsigned>(ceil_div(S.nrows, 256)), 256, 0, st>>>(part, S.n_tiles, S.nrows, mode,
                                                                                        scale, shift, out);
    SB_LAUNCH_CHECK();
    count_launch(c, 2);
}

void sell_spmm(snapb200_ctx* c, const Sell& S, const float* in, float* out, const float* scale,
               const float* subscale, const float* sub, int64_t lds) {
    SB_CHECK(S.built, "tiled SpMM: format not built");
    if (S.nrows == 0) return;
    if (S.b == 8) spmm_impl<8>(c, S, in, out, scale, subscale, sub, lds);
    else spmm_impl<4>(c, S, in, out, scale, subscale, sub, lds);
}

}  // namespace snapb

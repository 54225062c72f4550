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
constexpr uint32_t kLenBias = 32767;   // padded segment length <= 2 * tile_cols = 24576 < 2^15

// Padded class rotation.  A lane with `len` entries consumes its classes in rotation for R rounds
// (slot NC*q + ((class - o) mod NC) holds the q-th entry of a class, -1 if the class has fewer than
// q+1 entries); entries beyond R per class go to a tail after slot NC*R in column order, and if the
// tail is full into leftover holes of the rotation region.  R slightly below len/NC keeps ~90% of
// the entries in the conflict-free rotation; the lane owns len + a(len) slots (~11% padding for
// len ~ 100), a function of len alone so that the chunk plan needs no per-class statistics.
__host__ __device__ __forceinline__ int rotation_depth(int len, int nc) {
    const int r = (len + nc - 1) / nc - 1;
    return r > 0 ? r : 0;
}
__host__ __device__ __forceinline__ int padded_slots(int len, int nc) {
    if (rotation_depth(len, nc) == 0) return len;
    return len + static_cast<int>(0.45f * sqrtf(static_cast<float>(nc * len))) + 1;
}
// A lane lays its segment out in pieces of kPiece entries, each with its own rotation, so that the
// fill kernel can stage one piece window of every lane of a chunk in shared memory whatever the
// segment length.  Every full piece owns the same number of slots (kPieceSlots), which keeps the
// piece windows of the 32 lanes of a chunk aligned; only a lane's last, partial piece is shorter.
constexpr int kPiece = 256;
template <int NC> struct PieceSlots { static constexpr int value = (NC == 8) ? 280 : 272; };   // round4(padded_slots(kPiece, NC))
__host__ __device__ __forceinline__ int piece_slots(int nc) { return nc == 8 ? PieceSlots<8>::value : PieceSlots<4>::value; }
__host__ __device__ __forceinline__ int total_slots(int len, int nc) {
    const int full = len / kPiece;
    return full * piece_slots(nc) + padded_slots(len - full * kPiece, nc);
}

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
                 int pad_nc, int32_t* __restrict__ chunk_rows, int32_t* __restrict__ chunk_len4) {
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
            int len = sp[1] - sp[0];
            if (pad_nc) len = total_slots(len, pad_nc);   // slots, not entries
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

// Per-lane counters packed 16 bits each into 64-bit words (register only).
template <int NC>
struct Packed {
    uint64_t w[NC / 4];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int i = 0; i < NC / 4; ++i) w[i] = 0;
    }
    __device__ __forceinline__ int get(int c) const {
        uint64_t x = w[0];
#pragma unroll
        for (int i = 1; i < NC / 4; ++i)
            if ((c >> 2) == i) x = w[i];
        return static_cast<int>((x >> (16 * (c & 3))) & 0xFFFFull);
    }
    __device__ __forceinline__ void add(int c, int v) {
        const uint64_t inc = static_cast<uint64_t>(v) << (16 * (c & 3));
#pragma unroll
        for (int i = 0; i < NC / 4; ++i)
            if ((c >> 2) == i) w[i] += inc;
    }
    __device__ __forceinline__ int min_all() const {
        int m = 0x7fffffff;
#pragma unroll
        for (int c = 0; c < NC; ++c) m = min(m, get(c));
        return m;
    }
};

// Lanes that must read distinct bank-group classes in one LDS.128 wavefront
// form a "group": b = 4 (NC = 8): the 8 lanes of a quarter warp;
// b = 8 (NC = 4): the 4 even or the 4 odd lanes of a quarter warp.
template <int NC>
__device__ __forceinline__ int group_member(int lane) { return NC == 8 ? (lane & 7) : ((lane >> 1) & 3); }
template <int NC>
__device__ __forceinline__ int group_lane(int lane, int member) {
    return NC == 8 ? ((lane & ~7) | member) : ((lane & ~7) | (member << 1) | (lane & 1));
}
template <int NC>
__device__ __forceinline__ unsigned group_bits(unsigned ballot, int lane) {
    if (NC == 8) return (ballot >> (lane & ~7)) & 0xFFu;
    const unsigned x = (ballot >> ((lane & ~7) + (lane & 1))) & 0x55u;
    return (x & 1u) | ((x >> 1) & 2u) | ((x >> 2) & 4u) | ((x >> 3) & 8u);
}
template <int NC>
__device__ __forceinline__ unsigned rotl_nc(unsigned x, int r) {
    return ((x << r) | (x >> (NC - r))) & ((1u << NC) - 1u);
}

// Plain per-lane layout straight from global memory: class rotation while every
// class has entries (depth = smallest class count, no holes), the rest in column order.
template <bool HAS_VAL, int NC>
__device__ void fill_lane_simple(const int32_t* __restrict__ idx, const float* __restrict__ val, int64_t s, int64_t e,
                                 int o, int col0, int row_bytes, int steps, int32_t* __restrict__ d,
                                 float* __restrict__ dv) {
    // eight consecutive entries (one 32-byte sector) are loaded per iteration so that a lane exposes
    // one memory latency per sector instead of one per entry
    Packed<NC> cnt;
    cnt.clear();
    for (int64_t p = s; p < e; p += 8) {
        int jj[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) jj[u] = (p + u < e) ? idx[p + u] : -1;
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (jj[u] >= 0) cnt.add(jj[u] & (NC - 1), 1);
    }
    const int len = static_cast<int>(e - s);
    const int R = cnt.min_all();
    cnt.clear();
    int left = 0;
    for (int64_t p = s; p < e; p += 8) {
        int jj[8];
        float vv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            jj[u] = (p + u < e) ? idx[p + u] : -1;
            vv[u] = (HAS_VAL && p + u < e) ? val[p + u] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (jj[u] < 0) continue;
            const int cl = jj[u] & (NC - 1);
            const int q = cnt.get(cl);
            cnt.add(cl, 1);
            const int k = (q < R) ? NC * q + ((cl - o) & (NC - 1)) : NC * R + left++;
            const int64_t pos = static_cast<int64_t>(k >> 2) * 128 + (k & 3);
            d[pos] = (jj[u] - col0) * row_bytes;
            if (HAS_VAL) dv[pos] = vv[u];
        }
    }
    for (int k = len; k < steps; ++k) {
        const int64_t pos = static_cast<int64_t>(k >> 2) * 128 + (k & 3);
        d[pos] = -1;
        if (HAS_VAL) dv[pos] = 0.f;
    }
}

// ---- plain order: one warp per chunk, every lane lays out its own segment ----
template <bool HAS_VAL, int NC>
__global__ void __launch_bounds__(256)
sell_fill_simple_kernel(const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx, const float* __restrict__ val,
                        const int32_t* __restrict__ segptr, int n_tiles, int tile_cols, int row_bytes, int64_t n_chunks,
                        int64_t chunks_per_tile, const int32_t* __restrict__ chunk_rows,
                        const int32_t* __restrict__ chunk_len4, const int64_t* __restrict__ chunk_off,
                        int32_t* __restrict__ data, float* __restrict__ vals) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    const int o = group_member<NC>(lane);
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
        fill_lane_simple<HAS_VAL, NC>(idx, val, s, e, o, t * tile_cols, row_bytes, steps, d, dv);
    }
}

// ---- padded class rotation (default): staged in shared memory -------------
// One warp per chunk, one lane per row segment.  A lane walks its segment once, in pieces of kPiece
// entries read with aligned 16-byte loads, and drops every entry at its slot of a per-lane column
// in shared memory (16-bit tile-local columns, 0xFFFF = empty); the column is then written out as
// 16-byte groups, so that the warp stores 512 contiguous bytes per instruction instead of 32
// scattered words.  Slot of the q-th entry of class cl in a piece of plen entries:
//   q < R (= rotation_depth(plen)):  NC*q + ((cl - o) mod NC)        (conflict-free rotation)
//   else, while the tail has room:   NC*R + (running tail count)      (column order)
//   else (tail full, ~1 lane in 5):  the leftover rotation holes in class order -- those are only
//        known once the class totals are, so such entries wait in a small per-lane list (walked
//        again from global memory if even that overflows, e.g. all columns in one class).
// With values, the same walk runs a second time staging the entry's position in the piece, and the
// write-out gathers val[] through it.
constexpr int kStageWarps = 12;
constexpr int kOvfCap = 16;
template <int NC> struct StagePitch { static constexpr int value = 4 * ((PieceSlots<NC>::value / 4) | 1); };   // u16 units; odd in 8-byte words: conflict-free LDS.64

__device__ __forceinline__ int4 load_idx4(const int32_t* __restrict__ idx, int64_t q, int64_t nnz) {
    if (q + 3 < nnz) return *reinterpret_cast<const int4*>(idx + q);
    int4 v = make_int4(0, 0, 0, 0);
    if (q < nnz) v.x = idx[q];
    if (q + 1 < nnz) v.y = idx[q + 1];
    if (q + 2 < nnz) v.z = idx[q + 2];
    return v;
}

template <bool HAS_VAL, int NC>
__global__ void __launch_bounds__(kStageWarps * 32)
sell_fill_staged_kernel(const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx, const float* __restrict__ val,
                        int64_t nnz, const int32_t* __restrict__ segptr, int n_tiles, int tile_cols, int row_bytes,
                        int64_t n_chunks, int64_t chunks_per_tile, const int32_t* __restrict__ chunk_rows,
                        const int32_t* __restrict__ chunk_len4, const int64_t* __restrict__ chunk_off,
                        int32_t* __restrict__ data, float* __restrict__ vals) {
    constexpr int SP = PieceSlots<NC>::value, PITCH = StagePitch<NC>::value;
    extern __shared__ __align__(16) unsigned char stage_smem[];
    const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
    uint16_t* col = reinterpret_cast<uint16_t*>(stage_smem) + static_cast<size_t>(wic) * (32 * PITCH + kOvfCap * 32) + lane * PITCH;
    uint16_t* ovf = reinterpret_cast<uint16_t*>(stage_smem) + static_cast<size_t>(wic) * (32 * PITCH + kOvfCap * 32) + 32 * PITCH + lane;
    uint64_t* col64 = reinterpret_cast<uint64_t*>(col);
    const int64_t warp = static_cast<int64_t>(blockIdx.x) * kStageWarps + wic;
    const int64_t nwarps = static_cast<int64_t>(gridDim.x) * kStageWarps;
    const int o = group_member<NC>(lane);
    for (int64_t c = warp; c < n_chunks; c += nwarps) {
        const int steps = chunk_len4[c] * 4;
        if (steps == 0) continue;
        const int row = chunk_rows[c * 32 + lane];
        const int t = static_cast<int>(c / chunks_per_tile);
        const int col0 = t * tile_cols;
        const int64_t obase = chunk_off[c] * 128 + lane * 4;
        int64_t s = 0, e = 0;
        if (row >= 0) {
            const int32_t* sp = segptr + static_cast<int64_t>(row) * (n_tiles + 1) + t;
            s = ptr[row] + sp[0];
            e = ptr[row] + sp[1];
        }
        for (int k0 = 0; k0 < steps; k0 += SP, s += kPiece) {
            const int nst4 = min(SP, steps - k0) >> 2;
            const int64_t pe = (e - s > kPiece) ? s + kPiece : e;
            const int plen = pe > s ? static_cast<int>(pe - s) : 0;
            const int R = rotation_depth(plen, NC);
            const int tail_cap = padded_slots(plen, NC) - NC * R;
#pragma unroll 1
            for (int pass = 0; pass < (HAS_VAL ? 2 : 1); ++pass) {
                for (int g = 0; g < nst4; ++g) col64[g] = ~0ull;
                if (plen > 0) {
                    Packed<NC> cnt;
                    cnt.clear();
                    int left = 0, novf = 0;
                    for (int64_t a = s & ~static_cast<int64_t>(3); a < pe; a += 16) {
                        int4 v[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            v[u] = (a + 4 * u < pe) ? load_idx4(idx, a + 4 * u, nnz) : make_int4(0, 0, 0, 0);
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int jj[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                            for (int w = 0; w < 4; ++w) {
                                const int64_t p = a + 4 * u + w;
                                if (p < s || p >= pe) continue;
                                const int cl = jj[w] & (NC - 1);
                                const int q = cnt.get(cl);
                                cnt.add(cl, 1);
                                const uint16_t item = static_cast<uint16_t>(pass == 0 ? jj[w] - col0 : static_cast<int>(p - s));
                                if (q < R) {
                                    col[NC * q + ((cl - o) & (NC - 1))] = item;
                                } else if (left < tail_cap) {
                                    col[NC * R + left++] = item;
                                } else {
                                    if (novf < kOvfCap) ovf[novf * 32] = item;
                                    ++novf;
                                }
                            }
                        }
                    }
                    if (novf > 0) {   // tail overflow: fill the leftover holes of the rotation region in class order
                        int hole_c = 0, hole_q = cnt.get(0);
                        auto put = [&](uint16_t item) {
                            while (hole_q >= R) { ++hole_c; hole_q = cnt.get(hole_c); }
                            col[NC * hole_q + ((hole_c - o) & (NC - 1))] = item;
                            ++hole_q;
                        };
                        if (novf <= kOvfCap) {
                            for (int i = 0; i < novf; ++i) put(ovf[i * 32]);
                        } else {   // the list was too small: walk the piece again
                            Packed<NC> c2;
                            c2.clear();
                            int left2 = 0;
                            for (int64_t p = s; p < pe; ++p) {
                                const int j = idx[p];
                                const int cl = j & (NC - 1);
                                const int q = c2.get(cl);
                                c2.add(cl, 1);
                                if (q < R) continue;
                                if (left2 < tail_cap) { ++left2; continue; }
                                put(static_cast<uint16_t>(pass == 0 ? j - col0 : static_cast<int>(p - s)));
                            }
                        }
                    }
                }
                // write-out: 4 slots (16 bytes) per lane and instruction, 512 contiguous bytes per warp
                const int64_t ob = obase + static_cast<int64_t>(k0) * 32;
                for (int g = 0; g < nst4; ++g) {
                    const uint64_t w4 = col64[g];
                    const int c0 = static_cast<int>(w4 & 0xFFFFu), c1 = static_cast<int>((w4 >> 16) & 0xFFFFu);
                    const int c2 = static_cast<int>((w4 >> 32) & 0xFFFFu), c3 = static_cast<int>(w4 >> 48);
                    if (pass == 0) {
                        int4 out;
                        out.x = c0 == 0xFFFF ? -1 : c0 * row_bytes;
                        out.y = c1 == 0xFFFF ? -1 : c1 * row_bytes;
                        out.z = c2 == 0xFFFF ? -1 : c2 * row_bytes;
                        out.w = c3 == 0xFFFF ? -1 : c3 * row_bytes;
                        __stcs(reinterpret_cast<int4*>(data + ob + static_cast<int64_t>(g) * 128), out);
                    } else {
                        float4 out;
                        out.x = c0 == 0xFFFF ? 0.f : val[s + c0];
                        out.y = c1 == 0xFFFF ? 0.f : val[s + c1];
                        out.z = c2 == 0xFFFF ? 0.f : val[s + c2];
                        out.w = c3 == 0xFFFF ? 0.f : val[s + c3];
                        __stcs(reinterpret_cast<float4*>(vals + ob + static_cast<int64_t>(g) * 128), out);
                    }
                }
            }
        }
    }
}

// ---- matched order --------------------------------------------------------
// One warp per chunk.
//  1. The 32 row segments of the chunk are loaded cooperatively (coalesced,
//     four segments in flight) into shared memory as 16-bit tile-local columns.
//  2. Every lane buckets its entries by bank-group class (walks over shared
//     memory).  Per-lane state is kept in ROTATED class order (slot x holds
//     class (x + o) mod NC), so that the class a lane wants at step k, slot
//     k mod NC, is a compile-time register index once the step loop is
//     unrolled by NC.
//  3. Regular rounds: while every lane of the warp still has an entry of
//     every class, a round of NC steps is a plain rotation -- conflict free by
//     construction, no cross-lane traffic.
//  4. Tail: step by step, a lane takes an entry of its rotation class when it
//     has one; the others are matched, inside their group, to the classes the
//     group's primaries leave free (parallel proposals, lowest lane wins, a
//     few rounds); a lane with slack idles rather than take a conflicting
//     class, a lane without slack accepts the conflict.
constexpr int kFillWarps = 8;
constexpr int kColPitch = 33;   // uint16 columns staged as [position][lane], padded against bank conflicts

template <int NC>
__device__ __forceinline__ int slot_get(const int (&a)[NC], int xi) {
    int v = a[0];
#pragma unroll
    for (int x = 1; x < NC; ++x)
        if (x == xi) v = a[x];
    return v;
}
template <int NC>
__device__ __forceinline__ void slot_inc(int (&a)[NC], int xi) {
#pragma unroll
    for (int x = 0; x < NC; ++x)
        if (x == xi) ++a[x];
}

template <bool HAS_VAL, int NC, int CAP>
__global__ void __launch_bounds__(kFillWarps * 32)
sell_fill_matched_kernel(const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx, const float* __restrict__ val,
                         const int32_t* __restrict__ segptr, int n_tiles, int tile_cols, int row_bytes,
                         int64_t n_chunks, int64_t chunks_per_tile, const int32_t* __restrict__ chunk_rows,
                         const int32_t* __restrict__ chunk_len4, const int64_t* __restrict__ chunk_off,
                         int32_t* __restrict__ data, float* __restrict__ vals) {
    extern __shared__ __align__(16) unsigned char fill_smem[];
    const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
    const unsigned full = 0xffffffffu;
    uint16_t* cols = reinterpret_cast<uint16_t*>(fill_smem) + static_cast<size_t>(wic) * CAP * kColPitch;
    uint8_t* bkt = fill_smem + static_cast<size_t>(kFillWarps) * CAP * kColPitch * 2 + static_cast<size_t>(wic) * CAP * 32;
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    const int o = group_member<NC>(lane);
    const unsigned gmask = (NC == 8) ? (0xFFu << (lane & ~7)) : (0x55u << ((lane & ~7) + (lane & 1)));
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
        const int col0 = t * tile_cols;
        const int len = static_cast<int>(e - s);
        if (steps > CAP) {   // warp-uniform; rare (very long segments): simple order straight from global memory
            fill_lane_simple<HAS_VAL, NC>(idx, val, s, e, o, col0, row_bytes, steps, d, dv);
            continue;
        }
        // ---- 1. cooperative staging of the chunk's columns
        __syncwarp();
        for (int seg0 = 0; seg0 < 32; seg0 += 4) {
            int64_t ss[4];
            int ll[4], maxl = 0;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                ss[u] = __shfl_sync(full, s, seg0 + u);
                ll[u] = __shfl_sync(full, len, seg0 + u);
                maxl = max(maxl, ll[u]);
            }
            for (int q = lane; q < maxl; q += 32) {
                int v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = (q < ll[u]) ? ld_stream_int(idx + ss[u] + q) : 0;
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (q < ll[u]) cols[q * kColPitch + seg0 + u] = static_cast<uint16_t>(v[u] - col0);
            }
        }
        __syncwarp();
        // ---- 2. bucket by class; state in rotated slot order (slot x <-> class (x + o) mod NC)
        int cnt[NC], head[NC];
#pragma unroll
        for (int x = 0; x < NC; ++x) cnt[x] = 0;
        for (int q = 0; q < len; ++q) slot_inc<NC>(cnt, (static_cast<int>(cols[q * kColPitch + lane]) - o) & (NC - 1));
        int mmin = cnt[0];
        {
            int run = 0;
#pragma unroll
            for (int x = 0; x < NC; ++x) {
                head[x] = run;          // bucket start of slot x
                run += cnt[x];
                mmin = min(mmin, cnt[x]);
            }
        }
        {
            int fillp[NC];
#pragma unroll
            for (int x = 0; x < NC; ++x) fillp[x] = head[x];
            for (int q = 0; q < len; ++q) {
                const int xi = (static_cast<int>(cols[q * kColPitch + lane]) - o) & (NC - 1);
                bkt[slot_get<NC>(fillp, xi) * 32 + lane] = static_cast<uint8_t>(q);
                slot_inc<NC>(fillp, xi);
            }
        }
        int endp[NC];
#pragma unroll
        for (int x = 0; x < NC; ++x) endp[x] = head[x] + cnt[x];
        // ---- 3. regular rounds (every lane that has entries still has one of every class)
        const int rreg = __reduce_min_sync(full, len > 0 ? mmin : 0x7fffffff);
        const int n_reg = (rreg == 0x7fffffff) ? 0 : rreg;
        for (int r = 0; r < n_reg; ++r) {
#pragma unroll
            for (int x = 0; x < NC; ++x) {
                const int k = NC * r + x;
                int out = -1;
                float outv = 0.f;
                if (len > 0) {
                    const int rel = bkt[(head[x] + r) * 32 + lane];
                    out = static_cast<int>(cols[rel * kColPitch + lane]) * row_bytes;
                    if (HAS_VAL) outv = val[s + rel];
                }
                const int64_t pos = static_cast<int64_t>(k >> 2) * 128 + (k & 3);
                d[pos] = out;
                if (HAS_VAL) dv[pos] = outv;
            }
        }
        int total = len;
        unsigned avail = 0;   // bit x: slot x still has entries
        if (len > 0) {
            total = len - NC * n_reg;
#pragma unroll
            for (int x = 0; x < NC; ++x) {
                head[x] += n_reg;
                if (head[x] < endp[x]) avail |= 1u << x;
            }
        }
        // ---- 4. matched tail; k0 is a multiple of NC so slot (k mod NC) is the compile-time x
        for (int k0 = NC * n_reg; k0 < steps; k0 += NC) {
#pragma unroll
            for (int x = 0; x < NC; ++x) {
                const int k = k0 + x;
                if (k >= steps) break;   // warp-uniform
                const bool has = total > 0;
                const bool primary = has && ((avail >> x) & 1u);
                int slot = primary ? x : -1;
                const unsigned np = __ballot_sync(full, !primary);
                const unsigned holes = __ballot_sync(full, has && !primary);
                if (holes) {
                    // classes of the group not used by a primary this step.  Member i wants class
                    // (k + i) mod NC; in MY slot numbering class c is slot (c - o) mod NC, so the free
                    // slots are the group's non-primary member bits rotated by (member - o) = 0: bit i of
                    // group_bits(np) frees class (k + i), i.e. my slot (x + i - o) mod NC.
                    const unsigned gb = group_bits<NC>(np, lane);
                    unsigned freeslots = rotl_nc<NC>(gb, (x - o) & (NC - 1));
                    bool pending = has && !primary;
                    while (__any_sync(full, pending)) {
                        int want = -1;
                        if (pending) {
                            const unsigned usable = freeslots & avail;
                            if (usable) {
                                want = __ffs(usable) - 1;
                            } else {
                                if (total >= steps - k) slot = __ffs(avail) - 1;   // no slack: accept a conflict
                                pending = false;                                     // else idle this step
                            }
                        }
                        // the class behind my slot `want`, as a group-wide key: lowest lane per class wins
                        const int want_cls = (want + o) & (NC - 1);
                        const unsigned key = (want >= 0) ? static_cast<unsigned>((lane >> 3) * 64 + (NC == 4 ? (lane & 1) * 16 : 0) + want_cls)
                                                         : (0x1000u + static_cast<unsigned>(lane));
                        const unsigned same = __match_any_sync(full, key);
                        const bool win = pending && want >= 0 && (lane == __ffs(same) - 1);
                        if (win) { slot = want; pending = false; }
                        const unsigned wcls = win ? (1u << want_cls) : 0u;
                        const unsigned taken_cls = __reduce_or_sync(gmask, wcls);
                        // back to my slot numbering: class c -> slot (c - o) mod NC  (rotate right by o)
                        freeslots &= ~rotl_nc<NC>(taken_cls, (NC - o) & (NC - 1));
                    }
                }
                int out = -1;
                float outv = 0.f;
                if (slot >= 0) {
                    const int h = (slot == x) ? head[x] : slot_get<NC>(head, slot);
                    const int rel = bkt[h * 32 + lane];
                    if (slot == x) ++head[x]; else slot_inc<NC>(head, slot);
                    const int en = (slot == x) ? endp[x] : slot_get<NC>(endp, slot);
                    if (h + 1 == en) avail &= ~(1u << slot);
                    --total;
                    out = static_cast<int>(cols[rel * kColPitch + lane]) * row_bytes;
                    if (HAS_VAL) outv = val[s + rel];
                }
                const int64_t pos = static_cast<int64_t>(k >> 2) * 128 + (k & 3);
                d[pos] = out;
                if (HAS_VAL) dv[pos] = outv;
            }
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
    const int pad_mode = (c->fill_mode == 0) ? 1 : 0;   // 0: padded rotation (default), 1: matched, 2: plain
    const int nc = (b == 8) ? 4 : 8;
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
            S.window_start.p, S.window_chunk0.p, segptr.p, T, nw, S.chunks_per_tile, pad_mode ? nc : 0, S.chunk_rows.p,
            S.chunk_len4.p);
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
        const bool matched = c->fill_mode == 1;
        const int rb = 4 * b;
#define SB_FILL_ARGS                                                                                                   \
    segptr.p, T, S.tile_cols, rb, n_chunks, S.chunks_per_tile, S.chunk_rows.p, S.chunk_len4.p, S.chunk_off.p,          \
        S.data.p, S.vals.p
#define SB_FILL(HV, NC, CAP)                                                                                           \
    do {                                                                                                               \
        if (pad_mode) {                                                                                                \
            SB_CHECK(PieceSlots<NC>::value == ((padded_slots(kPiece, NC) + 3) & ~3), "piece slot table out of date");  \
            const size_t fsm = static_cast<size_t>(kStageWarps) * (32 * StagePitch<NC>::value + kOvfCap * 32) * 2;     \
            auto kern = sell_fill_staged_kernel<HV, NC>;                                                               \
            SB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(fsm)));   \
            const int blocks = static_cast<int>(std::min<int64_t>(ceil_div(n_chunks, kStageWarps), c->num_sms));       \
            kern<<<blocks, kStageWarps * 32, fsm, st>>>(M.ptr.p, M.idx.p, M.val.p, M.nnz, SB_FILL_ARGS);               \
        } else if (matched) {                                                                                          \
            const size_t fsm = static_cast<size_t>(kFillWarps) * CAP * (kColPitch * 2 + 32);                           \
            auto kern = sell_fill_matched_kernel<HV, NC, CAP>;                                                         \
            SB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(fsm)));   \
            const int per_sm = std::max<int>(1, static_cast<int>((220u << 10) / fsm));                                 \
            const int blocks = static_cast<int>(std::min<int64_t>(ceil_div(n_chunks, kFillWarps),                      \
                                                                  static_cast<int64_t>(c->num_sms) * per_sm));         \
            kern<<<blocks, kFillWarps * 32, fsm, st>>>(M.ptr.p, M.idx.p, M.val.p, SB_FILL_ARGS);                       \
        } else {                                                                                                       \
            const int blocks = static_cast<int>(std::min<int64_t>(ceil_div(n_chunks, 8),                               \
                                                                  static_cast<int64_t>(c->num_sms) * 8));              \
            sell_fill_simple_kernel<HV, NC><<<blocks, 256, 0, st>>>(M.ptr.p, M.idx.p, M.val.p, SB_FILL_ARGS);          \
        }                                                                                                              \
    } while (0)
        const bool hv = M.has_values();
        if (b == 8) { if (hv) SB_FILL(true, 4, 128); else SB_FILL(false, 4, 128); }
        else        { if (hv) SB_FILL(true, 8, 256); else SB_FILL(false, 8, 256); }
#undef SB_FILL_ARGS
#undef SB_FILL
        SB_LAUNCH_CHECK();
    }
    count_launch(c, 3);
    S.built = true;
}

}  // namespace snapb

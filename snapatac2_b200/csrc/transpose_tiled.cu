// Tile-major 16-bit transpose of the pattern (struct TileT, ctx.cuh): the input
// of the feature-major tiled copy S1 (pass 1 of the operator gathers over
// features; reference: the X^T product inside f(v), embedding.rs:162-163).
//
// The local cells are cut into tiles of H rows (H = the column tile of S1, at
// most 12288, so a tile-local cell id fits 16 bits).  For every (tile, feature)
// the transpose lists the tile's cells that have the feature, in ascending
// order.  Two streaming passes over the CSR rows, all the scattered work in
// shared memory, no global atomics on data and a result that does not depend
// on scheduling:
//   count:   one unit = (tile, 65536-feature range): a shared-memory histogram
//            (16-bit counters, integer atomics: order independent) -> cnt[t][j];
//   offsets: exclusive scan of cnt[t][.] per tile, tile bases, local document
//            frequencies df[j] = sum_t cnt[t][j];
//   emit:    one unit = (tile, group of 1024-feature ranges).  A per-row cursor
//            in shared memory walks the (sorted) rows range by range.  For each
//            range and each sub-block of 1024 rows the CTA sets a 1024 x 1024
//            feature x row bitmap in shared memory (one thread per row), then
//            one thread per feature walks its 1024 bits in row order and appends
//            the set rows to the feature's segment.  Segments start on 16-byte
//            boundaries (lengths padded to multiples of eight, so the format
//            build reads them with aligned vector loads); the segments of a range
//            are contiguous in the output (~250 KB at 1% density), so the appends
//            merge in L2 before they reach HBM.
// Units are handed out through an atomic counter (only the assignment of units
// to CTAs depends on timing, never the output).
#include "ctx.cuh"

#include <stdlib.h>
#include <algorithm>
#include <vector>

namespace snapb {

namespace {

constexpr int kTtThreads = 1024;
constexpr int kTtHist = 65536;     // features per counting unit (two 16-bit counters per word: 128 KB)
constexpr int kTtMaxRows = 12288;  // cursor capacity = largest tile

__device__ __forceinline__ int64_t lower_bound_idx(const int32_t* __restrict__ idx, int64_t lo, int64_t hi, int64_t key) {
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (static_cast<int64_t>(idx[mid]) < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// two 16-bit counters per word; an index outside the unit (unsorted input) is dropped and shows up
// as a count mismatch on the host
__device__ __forceinline__ void hist_add(uint32_t* hist, int k) {
    if (static_cast<unsigned>(k) < static_cast<unsigned>(kTtHist)) atomicAdd(&hist[k >> 1], (k & 1) ? 0x10000u : 1u);
}

__global__ void __launch_bounds__(kTtThreads, 1)
tile_count_kernel(const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx, int64_t n, int64_t m, int H,
                  int n_hr, int64_t n_units, uint16_t* __restrict__ cnt, unsigned long long* __restrict__ counter) {
    extern __shared__ __align__(16) uint32_t hist[];   // kTtHist / 2 words
    __shared__ long long s_unit;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    while (true) {
        if (tid == 0) s_unit = static_cast<long long>(atomicAdd(counter, 1ull));
        __syncthreads();
        const int64_t u = s_unit;
        if (u >= n_units) break;
        const int64_t t = u / n_hr;
        const int hr = static_cast<int>(u - t * n_hr);
        const int64_t f0 = static_cast<int64_t>(hr) * kTtHist, f1 = min(m, f0 + kTtHist);
        for (int i = tid; i < kTtHist / 2; i += kTtThreads) hist[i] = 0u;
        __syncthreads();
        const int64_t r_lo = t * H;
        const int nr = static_cast<int>(min(static_cast<int64_t>(H), n - r_lo));
        for (int rb = warp * 32; rb < nr; rb += kTtThreads) {
            int64_t s0 = 0, s1 = 0;
            if (rb + lane < nr) {   // one row per lane: its entries inside [f0, f1)
                const int64_t rs = ptr[r_lo + rb + lane], re = ptr[r_lo + rb + lane + 1];
                s0 = (f0 == 0) ? rs : lower_bound_idx(idx, rs, re, f0);
                s1 = (f1 >= m) ? re : lower_bound_idx(idx, s0, re, f1);
            }
            const int rows_here = min(32, nr - rb);
            for (int i = 0; i < rows_here; ++i) {   // the warp streams each of the 32 segments
                const int64_t a = __shfl_sync(0xffffffffu, s0, i), b = __shfl_sync(0xffffffffu, s1, i);
                int64_t p = a + lane;
                for (; p + 96 < b; p += 128) {
                    const int j0 = ld_stream_int(idx + p), j1 = ld_stream_int(idx + p + 32);
                    const int j2 = ld_stream_int(idx + p + 64), j3 = ld_stream_int(idx + p + 96);
                    const int k0 = j0 - static_cast<int>(f0), k1 = j1 - static_cast<int>(f0);
                    const int k2 = j2 - static_cast<int>(f0), k3 = j3 - static_cast<int>(f0);
                    hist_add(hist, k0);
                    hist_add(hist, k1);
                    hist_add(hist, k2);
                    hist_add(hist, k3);
                }
                for (; p < b; p += 32) {
                    const int k0 = ld_stream_int(idx + p) - static_cast<int>(f0);
                    hist_add(hist, k0);
                }
            }
        }
        __syncthreads();
        uint16_t* out = cnt + t * m + f0;
        for (int i = tid; i < static_cast<int>(f1 - f0); i += kTtThreads) {
            const uint32_t w = hist[i >> 1];
            out[i] = static_cast<uint16_t>((i & 1) ? (w >> 16) : (w & 0xFFFFu));
        }
        __syncthreads();
    }
}

// One CTA per tile: segoff[t][j] = exclusive scan of the segment lengths of tile t, each rounded up
// to a multiple of 8 entries (16 bytes); tile_total[t] = padded size, tile_total[n_tiles + t] = entries.
__global__ void __launch_bounds__(1024)
tile_scan_kernel(const uint16_t* __restrict__ cnt, int64_t m, uint32_t* __restrict__ segoff,
                 int64_t* __restrict__ tile_total) {
    __shared__ uint32_t wsum[32];
    __shared__ unsigned long long s_carry;
    __shared__ unsigned long long s_raw[32];
    const int64_t t = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0ull;
    unsigned long long raw = 0ull;
    __syncthreads();
    for (int64_t j0 = 0; j0 < m; j0 += 1024) {
        const int64_t j = j0 + tid;
        const uint32_t c0 = (j < m) ? cnt[t * m + j] : 0u;
        raw += c0;
        const uint32_t v = (c0 + 7u) & ~7u;
        uint32_t x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += y;
        }
        if (lane == 31) wsum[warp] = x;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = wsum[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, w, d);
                if (lane >= d) w += y;
            }
            wsum[lane] = w;   // inclusive
        }
        __syncthreads();
        const unsigned long long carry = s_carry;
        const unsigned long long excl = carry + (warp > 0 ? wsum[warp - 1] : 0u) + (x - v);
        // offsets beyond 2^32 - 1 are reported through tile_total (checked on the host)
        if (j < m) segoff[t * m + j] = static_cast<uint32_t>(excl);
        __syncthreads();
        if (tid == 0) s_carry = carry + wsum[31];
        __syncthreads();
    }
    // entries of the tile (fixed-order sum)
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) raw += __shfl_down_sync(0xffffffffu, raw, d);
    if (lane == 0) s_raw[warp] = raw;
    __syncthreads();
    if (tid == 0) {
        unsigned long long r = 0ull;
        for (int w = 0; w < 32; ++w) r += s_raw[w];
        tile_total[t] = static_cast<int64_t>(s_carry);
        tile_total[gridDim.x + t] = static_cast<int64_t>(r);
    }
}

__global__ void tile_df_kernel(const uint16_t* __restrict__ cnt, int64_t m, int n_tiles, int64_t* __restrict__ df) {
    const int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (j >= m) return;
    int64_t s = 0;
    for (int t = 0; t < n_tiles; ++t) s += cnt[static_cast<int64_t>(t) * m + j];
    df[j] = s;
}

// ---- emit ------------------------------------------------------------------
// One unit = (cell tile, group of 1024-feature ranges).  A per-row cursor in shared memory walks
// the (sorted) rows range by range.  For each range and each sub-block of 1024 rows the CTA sets
// a 1024 x 1024 feature x row bitmap in shared memory (one thread per row), then one thread per
// feature walks its 1024 bits in row order and appends the set rows to the feature's segment.
// The segments of a range are contiguous in the output (~250 KB at 1% density), so the small
// appends merge in L2 before they reach HBM.
// (Tried and slower on the C3 workload, whose feature popularity is heavily skewed -- see
// profiles/README.md: a warp-uniform bit loop with a per-feature plane summary, an entry-parallel
// variant with per-feature row masks and two barriers per 32 rows, and a barrier-free variant
// with one unit per warp.)
constexpr int kTtRange = 1024;     // features per emit range (one thread each)
constexpr int kTtSub = 1024;       // rows per emit sub-block (one thread each)

__global__ void __launch_bounds__(kTtThreads, 1)
tile_emit_kernel(const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx, int64_t nnz, int64_t n, int64_t m,
                 int H, int n_groups, int ranges_per_group, int n_ranges, int64_t n_units,
                 const uint16_t* __restrict__ cnt, const uint32_t* __restrict__ segoff,
                 const int64_t* __restrict__ tile_base, uint16_t* __restrict__ ids,
                 unsigned long long* __restrict__ counter) {
    // bitmap word of (plane w, feature f): bm[((w >> 2) * kTtRange + f) * 4 + (w & 3)] -- plane w holds
    // rows 32 w .. 32 w + 31 of the sub-block (= warp w of the set phase); the emit thread of
    // feature f reads four planes with one 16-byte load.
    extern __shared__ __align__(16) uint32_t tt_smem[];
    uint32_t* bm = tt_smem;                                    // 32 * kTtRange words (128 KB)
    uint32_t* cur = tt_smem + 32 * kTtRange;                   // kTtMaxRows cursors (48 KB)
    uint32_t* opos = cur + kTtMaxRows;                         // kTtRange: next free slot of a segment (tile relative)
    uint16_t* plist = reinterpret_cast<uint16_t*>(opos + kTtRange);   // kTtRange: the range's popular features
    __shared__ long long s_unit;
    __shared__ int s_npop;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 32 * kTtRange; i += kTtThreads) bm[i] = 0u;
    if (tid == 0) s_npop = 0;
    while (true) {
        __syncthreads();
        if (tid == 0) s_unit = static_cast<long long>(atomicAdd(counter, 1ull));
        __syncthreads();
        const int64_t u = s_unit;
        if (u >= n_units) break;
        const int64_t t = u / n_groups;
        const int g = static_cast<int>(u - t * n_groups);
        const int rg0 = g * ranges_per_group, rg1 = min(n_ranges, rg0 + ranges_per_group);
        const int64_t r_lo = t * H;
        const int nr = static_cast<int>(min(static_cast<int64_t>(H), n - r_lo));
        const int64_t fg0 = static_cast<int64_t>(rg0) * kTtRange;
        for (int lr = tid; lr < nr; lr += kTtThreads) {   // cursor = first entry of the row at or after the group
            const int64_t rs = ptr[r_lo + lr], re = ptr[r_lo + lr + 1];
            cur[lr] = (fg0 == 0) ? 0u : static_cast<uint32_t>(lower_bound_idx(idx, rs, re, fg0) - rs);
        }
        __syncthreads();
        const int64_t tbase = tile_base[t];
        for (int rg = rg0; rg < rg1; ++rg) {
            const int fb = rg * kTtRange;
            const int fe = static_cast<int>(min(m, static_cast<int64_t>(fb) + kTtRange));
            const bool has_feature = fb + tid < fe;
            // Feature popularity is heavily skewed: a thread that walked a popular feature's bits alone
            // would hold its whole warp up.  Features with more than ~24 cells per sub-block in this
            // tile are listed and emitted by a whole warp each (one lane per bitmap plane).
            const int n_sub = (nr + kTtSub - 1) / kTtSub;
            bool popular = false;
            if (tid == 0) s_npop = 0;
            __syncthreads();
            if (has_feature) {
                opos[tid] = segoff[t * m + fb + tid];
                popular = static_cast<int>(cnt[t * m + fb + tid]) > 24 * n_sub;
                if (popular) plist[atomicAdd(&s_npop, 1)] = static_cast<uint16_t>(tid);   // order irrelevant
            }
            __syncthreads();
            const int npop = s_npop;
            for (int sb = 0; sb < nr; sb += kTtSub) {
                // ---- set: one thread per row of the sub-block.  Rows are sorted, so the entries of a
                //      batch that fall into the range form a prefix: no dependency between them.
                const int lr = sb + tid;
                if (lr < nr) {
                    const int64_t rs = ptr[r_lo + lr];
                    uint32_t c0 = cur[lr];
                    int left = static_cast<int>(ptr[r_lo + lr + 1] - rs) - static_cast<int>(c0);   // entries not yet consumed
                    const int64_t pos = rs + c0;
                    int64_t a = pos & ~static_cast<int64_t>(3);   // two aligned 16-byte loads per batch
                    int skip = static_cast<int>(pos - a);
                    uint32_t* plane = bm + (static_cast<size_t>(warp >> 2) * kTtRange) * 4 + (warp & 3);
                    const uint32_t bit = 1u << lane;
                    while (left > 0) {
                        int j[8];
                        if (a + 7 < nnz) {
                            const int4 v0 = *reinterpret_cast<const int4*>(idx + a);
                            const int4 v1 = *reinterpret_cast<const int4*>(idx + a + 4);
                            j[0] = v0.x; j[1] = v0.y; j[2] = v0.z; j[3] = v0.w;
                            j[4] = v1.x; j[5] = v1.y; j[6] = v1.z; j[7] = v1.w;
                        } else {
#pragma unroll
                            for (int q = 0; q < 8; ++q) j[q] = (a + q < nnz) ? idx[a + q] : 0x7fffffff;
                        }
                        int took = 0;
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            if (q >= skip && q - skip < left && j[q] < fe) {
                                const int f = j[q] - fb;
                                if (f >= 0) atomicOr(plane + f * 4, bit);
                                ++took;
                            }
                        }
                        c0 += took;
                        if (took < 8 - skip) break;
                        left -= took;
                        a += 8;
                        skip = 0;
                    }
                    cur[lr] = c0;
                }
                __syncthreads();
                // ---- emit, ordinary features: one thread per feature, rows in ascending order
                if (has_feature && !popular) {
                    int64_t outpos = tbase + opos[tid];
                    uint4* mine = reinterpret_cast<uint4*>(bm) + tid;
#pragma unroll 1
                    for (int g8 = 0; g8 < 8; ++g8) {
                        const uint4 w4 = mine[g8 * kTtRange];
                        if ((w4.x | w4.y | w4.z | w4.w) == 0u) continue;
                        mine[g8 * kTtRange] = make_uint4(0u, 0u, 0u, 0u);
                        const uint32_t ww[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            uint32_t word = ww[k];
                            const int row0 = sb + (g8 * 4 + k) * 32;
                            while (word) {
                                const int b = __ffs(word) - 1;
                                word &= word - 1;
                                ids[outpos++] = static_cast<uint16_t>(row0 + b);
                            }
                        }
                    }
                    opos[tid] = static_cast<uint32_t>(outpos - tbase);
                }
                // ---- emit, popular features: one warp per feature, lane p owns plane p (rows 32p..32p+31)
                for (int i = warp; i < npop; i += kTtThreads / 32) {
                    const int f = plist[i];
                    uint32_t* wp = bm + (static_cast<size_t>(lane >> 2) * kTtRange + f) * 4 + (lane & 3);
                    uint32_t word = *wp;
                    *wp = 0u;
                    const int c = __popc(word);
                    int incl = c;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const int y = __shfl_up_sync(0xffffffffu, incl, d);
                        if (lane >= d) incl += y;
                    }
                    const uint32_t base = opos[f];
                    int64_t outpos = tbase + base + (incl - c);
                    const int row0 = sb + lane * 32;
                    while (word) {
                        const int b = __ffs(word) - 1;
                        word &= word - 1;
                        ids[outpos++] = static_cast<uint16_t>(row0 + b);
                    }
                    __syncwarp();
                    if (lane == 31) opos[f] = base + static_cast<uint32_t>(incl);
                }
                __syncthreads();
            }
        }
    }
}

// values of the transposed entries (values path only): one warp per (tile, feature) segment
__global__ void __launch_bounds__(256)
tile_vals_kernel(const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx, const float* __restrict__ val,
                 int64_t m, int H, int n_tiles, const uint16_t* __restrict__ cnt, const uint32_t* __restrict__ segoff,
                 const int64_t* __restrict__ tile_base, const uint16_t* __restrict__ ids, float* __restrict__ tvals) {
    const int lane = threadIdx.x & 31;
    const int64_t nseg = static_cast<int64_t>(n_tiles) * m;
    const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    for (int64_t sg = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; sg < nseg; sg += nwarps) {
        const int len = cnt[sg];
        if (len == 0) continue;
        const int64_t t = sg / m;
        const int64_t j = sg - t * m;
        const int64_t s = tile_base[t] + segoff[sg];
        for (int k = lane; k < len; k += 32) {
            const int64_t r = t * H + ids[s + k];
            tvals[s + k] = val[lower_bound_idx(idx, ptr[r], ptr[r + 1], j)];
        }
    }
}

// ==========================================================================
// Bucketed transpose (default; m <= kTrF * kTrMaxBuckets features).
//
// The emit kernel above walks the CSR rows of a cell tile once per 1024-feature range and finds
// ~10 entries per (row, range): one thread per row with a handful of entries each, then one thread
// per feature over a 1024 x 1024 bitmap that is 1 % full -- a few active lanes per warp on both
// sides (ncu: IPC 1.3, 95 GB of DRAM reads for a 19.6 GB index stream).  Here every pass over the
// entries is entry-parallel and coalesced, and every global write is a long contiguous piece:
//   A1  slabcnt[s][k]  entries of the 16-row slab s in feature bucket k (kTrF = 512 features): one
//                      warp streams the slab's rows, 128 contiguous bytes per load; rows are sorted,
//                      so a bucket is a run of consecutive entries;
//   A2  slaboff[s][k]  exclusive prefix of slabcnt over the slabs of s's cell tile, bucket totals;
//   A3  scatter        one CTA per slab: its entries are partitioned by bucket IN SHARED MEMORY
//                      (run lengths per (row, bucket), prefix over rows and buckets, 16-bit staging
//                      of up to ~93k entries) and every bucket's piece -- the slab's ~80 entries of
//                      that bucket, rows in order -- is written as one contiguous run to
//                      tmp[tile][bucket][slaboff ...] as (feature in bucket, tile-local cell).  (A warp
//                      writing each row's ~20-byte run by itself paid a DRAM fill for every partially
//                      written 32-byte sector: 30 ms for this pass alone on C3.)
//   B   per (tile, bucket): a stable counting sort of the bucket's ~62k entries by feature
//       (per-warp counters in shared memory over consecutive slices, ranks inside a 32-entry
//       batch from one ballot per feature bit) INTO SHARED MEMORY, then the bucket's whole
//       output -- the segments of its 512 features are adjacent in the TileT layout -- leaves as one
//       coalesced copy.  (Two-byte scattered global stores kept the first version on the L1 pipe.)
// All positions come from prefix sums: no atomics on data, the result does not depend on timing.
// A tile's buckets get regions padded for the worst case (7 slots per feature), so no pass has
// to know the exact padded segment lengths of earlier buckets.
// ==========================================================================
constexpr int kTrLog = 9;
constexpr int kTrF = 1 << kTrLog;          // features per bucket
constexpr int kTrMaxBuckets = 1024;        // run-length tables of A3: 16 rows x buckets x 2 B of shared memory
constexpr int kTrWarps = 16;               // warps per CTA in A1 / A3
constexpr int kTrSlab = 16;                // rows per slab (= warps of an A3 CTA)
constexpr int kTrSmemMax = 232448;         // 227 KB of dynamic shared memory per CTA

// Run lengths of one row per bucket, streamed by one warp: tab[k] += entries of the row in bucket k.
// The scatter pass hands the two halves of a row to two warps; the second half starts at entry
// tr_mid(len) (a multiple of 32, so it starts a batch) and needs to know how many entries of its
// first bucket's run lie in the first half: *mid_carry = bucket << 16 | count.
__device__ __forceinline__ int tr_mid(int len) { return ((len + 1) / 2 + 31) & ~31; }

__device__ __forceinline__ void tr_count_row(const int32_t* __restrict__ idx, int64_t rs, int64_t re, uint16_t* tab, int lane,
                                             uint32_t* mid_carry) {
    const int32_t* __restrict__ row = idx + rs;
    const int len = static_cast<int>(re - rs);          // a row has fewer than 2^31 entries (m < 2^31, no duplicates)
    const int mid = tr_mid(len);
    for (int o = lane; o < len + lane; o += 128) {      // (warp-uniform trip count: o - lane < len)
        int j[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) j[u] = (o + 32 * u < len) ? ld_stream_int(row + o + 32 * u) : -1;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int nvalid = min(32, len - (o - lane) - 32 * u);   // valid lanes are a prefix of the warp
            if (nvalid <= 0) break;
            const bool valid = lane < nvalid;
            const int k = valid ? (j[u] >> kTrLog) : -1;
            if (mid_carry != nullptr && (o - lane) + 32 * u == mid && lane == 0)
                *mid_carry = (static_cast<uint32_t>(k) << 16) | tab[k];          // tab holds this row's counts so far
            const int kprev = __shfl_up_sync(0xffffffffu, k, 1);
            const bool head = valid && (lane == 0 || k != kprev);
            const unsigned hmask = __ballot_sync(0xffffffffu, head);
            // the run this lane heads ends at the next head (or at the end of the batch)
            const unsigned above = hmask & ~((2u << lane) - 1u);
            const int end = above ? (__ffs(above) - 1) : nvalid;
            if (head) tab[k] = static_cast<uint16_t>(tab[k] + (end - lane));   // heads of a batch have distinct buckets
            __syncwarp();
        }
    }
}

// A1: one warp per slab (grid-stride in slab order): run lengths per (row, bucket) -- the scatter pass loads
// them instead of streaming its rows twice -- and their sums per (slab, bucket).
__global__ void __launch_bounds__(kTrWarps * 32)
tr_slabcnt_kernel(const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx, int64_t n, int H, int SPT,
                  int64_t n_slabs, int NBp, uint16_t* __restrict__ rowcnt, uint16_t* __restrict__ slabcnt,
                  uint32_t* __restrict__ midcarry) {
    extern __shared__ __align__(16) uint16_t tr_tab[];   // kTrWarps x 2 x NBp: current row, slab sums
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint16_t* tab = tr_tab + static_cast<size_t>(warp) * 2 * NBp;
    uint16_t* acc = tab + NBp;
    for (int i = lane; i < 2 * NBp; i += 32) tab[i] = 0;
    __syncwarp();
    const int64_t nw = static_cast<int64_t>(gridDim.x) * kTrWarps;
    for (int64_t sl = static_cast<int64_t>(blockIdx.x) * kTrWarps + warp; sl < n_slabs; sl += nw) {
        const int64_t t = sl / SPT;
        const int64_t r0 = t * H + (sl - t * SPT) * kTrSlab;
        const int64_t r1 = min(min(r0 + kTrSlab, (t + 1) * H), n);
        for (int64_t r = r0; r < r1; ++r) {
            tr_count_row(idx, ptr[r], ptr[r + 1], tab, lane, midcarry + r);
            __syncwarp();
            uint32_t* out = reinterpret_cast<uint32_t*>(rowcnt + r * NBp);      // NBp is a multiple of 32: rows are 64-byte aligned
            uint32_t* t32 = reinterpret_cast<uint32_t*>(tab);
            uint32_t* a32 = reinterpret_cast<uint32_t*>(acc);
            for (int i = lane; i < NBp / 2; i += 32) {
                const uint32_t v = t32[i];
                out[i] = v;
                a32[i] += v;          // two 16-bit sums per word; a slab holds at most 16 x 512 entries per bucket: no carry
                t32[i] = 0u;
            }
            __syncwarp();
        }
        uint16_t* out = slabcnt + sl * NBp;
        for (int i = lane; i < NBp; i += 32) {
            out[i] = acc[i];
            acc[i] = 0;
        }
        __syncwarp();
    }
}

// A2: one CTA per (cell tile, group of 32 buckets): exclusive prefix over the tile's slabs
__global__ void __launch_bounds__(1024)
tr_rowscan_kernel(const uint16_t* __restrict__ rowcnt, int64_t n, int H, int NBp, uint32_t* __restrict__ rowoff,
                  int64_t* __restrict__ btot) {
    __shared__ uint32_t wsum[32][33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int groups = NBp / 32;
    const int64_t t = blockIdx.x / groups;
    const int g = blockIdx.x % groups;
    const int64_t r_lo = t * H;
    const int nr = static_cast<int>(min(static_cast<int64_t>(H), n - r_lo));
    const int per = (nr + 31) / 32;
    const int a = min(nr, warp * per), b = min(nr, a + per);
    const uint16_t* src = rowcnt + r_lo * NBp + g * 32 + lane;
    uint32_t s = 0;
    for (int r = a; r < b; ++r) s += src[static_cast<int64_t>(r) * NBp];
    wsum[warp][lane] = s;
    __syncthreads();
    uint32_t run = 0;
    for (int w = 0; w < warp; ++w) run += wsum[w][lane];
    if (warp == 31) btot[t * NBp + g * 32 + lane] = static_cast<int64_t>(run + s);
    uint32_t* dst = rowoff + r_lo * NBp + g * 32 + lane;
    for (int r = a; r < b; ++r) {
        dst[static_cast<int64_t>(r) * NBp] = run;
        run += src[static_cast<int64_t>(r) * NBp];
    }
}

// A3: one CTA per slab (persistent, slabs in order), two warps per row of the slab (the kernel is bound by
// instruction issue: 32 resident warps instead of 16).  The run lengths per (row, bucket) come from A1's
// table, so the rows are streamed once (eight loads in flight per warp).
__global__ void __launch_bounds__(kTrWarps * 64, 1)
tr_slab_scatter_kernel(const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx, int64_t n, int H, int SPT,
                       int64_t n_slabs, int NBp, int cap, const uint16_t* __restrict__ rowcnt,
                       const uint32_t* __restrict__ slaboff, const uint32_t* __restrict__ trel,
                       const int64_t* __restrict__ tile_tmp0, const uint32_t* __restrict__ midcarry,
                       uint32_t* __restrict__ tmp_all, int* __restrict__ too_long) {
    extern __shared__ __align__(16) unsigned char tr_sm[];
    uint16_t* cnt = reinterpret_cast<uint16_t*>(tr_sm);                           // kTrSlab x NBp: run lengths, then row offsets
    uint32_t* tot = reinterpret_cast<uint32_t*>(cnt + static_cast<size_t>(kTrSlab) * NBp);   // NBp: entries of the group per bucket
    uint32_t* bstart = tot + NBp;                                                 // NBp: start of a bucket's piece in the staging
    uint32_t* gbase = bstart + NBp;                                               // NBp: where the bucket's next piece goes in tmp
    uint16_t* stage = reinterpret_cast<uint16_t*>(gbase + NBp);                   // cap entries: feature in bucket << 4 | row in slab
    __shared__ uint32_t wsum[2 * kTrWarps];
    __shared__ int s_g1, s_long;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wrow = warp >> 1, half = warp & 1;     // row of the slab and half of it this warp streams
    for (int64_t sl = blockIdx.x; sl < n_slabs; sl += gridDim.x) {
        const int64_t t = sl / SPT;
        const int64_t r0 = t * H + (sl - t * SPT) * kTrSlab;
        const int nrows = static_cast<int>(max(static_cast<int64_t>(0), min(min(r0 + kTrSlab, (t + 1) * H), n) - r0));
        if (nrows == 0) continue;
        uint32_t* tmp = tmp_all + tile_tmp0[t];
        const uint32_t rl0 = static_cast<uint32_t>(r0 - t * H);
        {
            const uint32_t* so = slaboff + sl * NBp;
            const uint32_t* tr = trel + t * NBp;
            for (int i = tid; i < NBp; i += blockDim.x) gbase[i] = tr[i] + so[i];
        }
        // rows [g0, g1) of the slab are staged together; almost always that is the whole slab
        for (int g0 = 0; g0 < nrows;) {
            __syncthreads();
            if (tid == 0) {
                int64_t acc = 0;
                int g1 = g0, any_long = 0;
                while (g1 < nrows) {
                    const int64_t len = ptr[r0 + g1 + 1] - ptr[r0 + g1];
                    if (g1 > g0 && acc + len > cap) break;
                    if (len > cap) any_long = 1;     // staged alone, skipped and flagged below
                    acc += len;
                    ++g1;
                }
                s_g1 = g1;
                s_long = any_long;
            }
            {   // run lengths of the slab's rows: one coalesced copy of A1's table (rows of the slab are adjacent)
                const uint4* src = reinterpret_cast<const uint4*>(rowcnt + r0 * NBp);
                uint4* dst = reinterpret_cast<uint4*>(cnt);
                const int n16 = nrows * NBp / 8;
                for (int i = tid; i < n16; i += blockDim.x) dst[i] = src[i];
            }
            __syncthreads();
            const int g1 = s_g1;
            bool mine = wrow >= g0 && wrow < g1;
            const int64_t rs = mine ? ptr[r0 + wrow] : 0, re = mine ? ptr[r0 + wrow + 1] : 0;
            // a single row longer than the staging cannot be handled here: flagged, the caller falls back to the
            // bitmap transpose
            if (mine && re - rs > cap) {
                mine = false;
                if (lane == 0) *too_long = 1;
                if (half == 0)
                    for (int i = lane; i < NBp; i += 32) cnt[static_cast<size_t>(wrow) * NBp + i] = 0;   // nothing of it is staged
            }
            if (s_long) __syncthreads();
            // prefix over the rows of every bucket, then over the buckets (one bucket per thread)
            uint32_t mine_tot = 0;
            if (tid < NBp) {
                uint32_t off = 0;
                for (int w = g0; w < g1; ++w) {
                    const uint16_t c0 = cnt[static_cast<size_t>(w) * NBp + tid];
                    cnt[static_cast<size_t>(w) * NBp + tid] = static_cast<uint16_t>(off);
                    off += c0;
                }
                tot[tid] = off;
                mine_tot = off;
            }
            uint32_t x = mine_tot;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
                if (lane >= d) x += y;
            }
            if (lane == 31) wsum[warp] = x;
            __syncthreads();
            uint32_t before = 0;
            for (int w = 0; w < warp; ++w) before += wsum[w];
            if (tid < NBp) bstart[tid] = before + x - mine_tot;
            __syncthreads();
            // scatter into the staging: a run that continues from the previous batch is the first of the batch
            if (mine) {
                const uint16_t* roff = cnt + static_cast<size_t>(wrow) * NBp;
                int carry_k = -1;
                uint32_t carry_cnt = 0;
                const int full = static_cast<int>(re - rs);
                const int mid = min(full, tr_mid(full));
                const int32_t* __restrict__ row = idx + rs + (half ? mid : 0);
                const int len = half ? full - mid : mid;
                if (half && len > 0) {       // the part of the first bucket's run that the other warp writes
                    const uint32_t cv = midcarry[r0 + wrow];
                    carry_k = static_cast<int>(cv >> 16);
                    carry_cnt = cv & 0xFFFFu;
                }
                const uint16_t tagw = static_cast<uint16_t>(wrow);
                for (int o = lane; o < len + lane; o += 256) {
                    int j[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) j[u] = (o + 32 * u < len) ? ld_stream_int(row + o + 32 * u) : -1;
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int nvalid = min(32, len - (o - lane) - 32 * u);   // valid lanes are a prefix of the warp
                        if (nvalid <= 0) break;
                        const bool valid = lane < nvalid;
                        const int k = valid ? (j[u] >> kTrLog) : -1;
                        const int kprev = __shfl_up_sync(0xffffffffu, k, 1);
                        const bool head = valid && (lane == 0 || k != kprev);
                        const unsigned hmask = __ballot_sync(0xffffffffu, head);
                        uint32_t off = 0;
                        if (head) off = bstart[k] + roff[k] + ((lane == 0 && k == carry_k) ? carry_cnt : 0u);
                        const int h = 31 - __clz(hmask & ((2u << lane) - 1u));   // my run's head lane
                        const uint32_t base = __shfl_sync(0xffffffffu, off, h);
                        if (valid) stage[base + (lane - h)] = static_cast<uint16_t>(((j[u] & (kTrF - 1)) << 4) | tagw);
                        const int hl = 31 - __clz(hmask);                        // head of the last run
                        const int kl = __shfl_sync(0xffffffffu, k, hl);
                        const uint32_t len_l = static_cast<uint32_t>(nvalid - hl);
                        carry_cnt = len_l + ((hl == 0 && kl == carry_k) ? carry_cnt : 0u);
                        carry_k = kl;
                    }
                }
            }
            __syncthreads();
            // every bucket's piece leaves as one contiguous run
            for (int k = warp; k < NBp; k += 2 * kTrWarps) {
                const uint32_t len = tot[k];
                if (len == 0) continue;
                const uint16_t* src = stage + bstart[k];
                const uint32_t gb = gbase[k];
                uint32_t* dst = tmp + gb;
                for (uint32_t i = lane; i < len; i += 32) {
                    const uint32_t e = src[i];
                    dst[i] = ((e >> 4) << 14) | (rl0 + (e & 15u));
                }
                __syncwarp();
                if (lane == 0) gbase[k] = gb + len;
            }
            g0 = g1;
        }
        __syncthreads();
    }
}

// B: one unit = (cell tile, bucket); persistent CTAs pull units from a counter.
// Counts per (warp slice, feature) by integer shared-memory atomics (order independent).  In the
// scatter phase the entries of a 32-entry batch that share a feature must take consecutive slots
// in lane (= row) order: the set of lanes with the same feature comes from one ballot per feature
// bit, the lowest of them advances the feature's cursor by the group size, and each lane's slot
// is the cursor plus its rank in the set.  (This is match.any spelled out: the instruction itself
// costs hundreds of cycles on 32 distinct keys and made the kernel 10x slower; a shared-memory
// mask table worked but added to the pipe that bounds the kernel.)
__global__ void __launch_bounds__(1024, 1)
tr_bucket_sort_kernel(const uint32_t* __restrict__ tmp, const int64_t* __restrict__ tmp_base, const int64_t* __restrict__ btot,
                      const uint32_t* __restrict__ ureg, const int64_t* __restrict__ tile_base, int64_t m, int NB, int NBp,
                      int64_t n_units, int cap, uint16_t* __restrict__ cnt, uint32_t* __restrict__ segoff,
                      uint16_t* __restrict__ ids, unsigned long long* __restrict__ counter) {
    extern __shared__ __align__(16) uint32_t bs_smem[];
    uint16_t* cw = reinterpret_cast<uint16_t*>(bs_smem);                 // 32 warps x kTrF counters, then running offsets
    uint32_t* segstart = bs_smem + 32 * kTrF / 2;                        // kTrF: segment starts relative to the unit's output
    uint16_t* stage = reinterpret_cast<uint16_t*>(segstart + kTrF);      // cap entries: the unit's output
    __shared__ uint32_t wsum[32];
    __shared__ long long s_unit;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint16_t* mine = cw + static_cast<size_t>(warp) * kTrF;
    uint32_t* mine32 = reinterpret_cast<uint32_t*>(mine);
    while (true) {
        __syncthreads();
        if (tid == 0) s_unit = static_cast<long long>(atomicAdd(counter, 1ull));
        __syncthreads();
        const int64_t u = s_unit;
        if (u >= n_units) break;
        const int64_t t = u / NB;
        const int k = static_cast<int>(u - t * NB);
        const int64_t E = btot[t * NBp + k];
        const uint32_t* in = tmp + tmp_base[t * NBp + k];
        const int64_t f0 = static_cast<int64_t>(k) << kTrLog;
        const int nf = static_cast<int>(min(static_cast<int64_t>(kTrF), m - f0));
        for (int i = tid; i < 32 * kTrF / 2; i += 1024) bs_smem[i] = 0u;
        __syncthreads();
        // ---- counts per (warp slice, feature); slices are consecutive, so slice order = row order
        const int En = static_cast<int>(E);                      // a unit has at most tile rows x kTrF entries
        const int per = ((En + 31) / 32 + 127) / 128 * 128;
        const int a = min(En, warp * per), b = min(En, a + per);
        auto load4 = [&](int p0, uint32_t* e) {
#pragma unroll
            for (int u2 = 0; u2 < 4; ++u2) {
                const int p = p0 + 32 * u2 + lane;
                e[u2] = (p < b) ? __ldg(in + p) : 0xFFFFFFFFu;   // a stored entry never has its top byte set
            }
        };
        {
            uint32_t cur[4], nxt[4];
            if (a < b) load4(a, cur);
            for (int p0 = a; p0 < b; p0 += 128) {
#pragma unroll
                for (int u2 = 0; u2 < 4; ++u2) nxt[u2] = 0xFFFFFFFFu;
                if (p0 + 128 < b) load4(p0 + 128, nxt);
#pragma unroll
                for (int u2 = 0; u2 < 4; ++u2) {
                    const uint32_t e = cur[u2];
                    if (e != 0xFFFFFFFFu) {
                        const uint32_t f = (e >> 14) & (kTrF - 1);
                        atomicAdd(&mine32[f >> 1], (f & 1u) ? 0x10000u : 1u);   // two 16-bit counters per word
                    }
                }
#pragma unroll
                for (int u2 = 0; u2 < 4; ++u2) cur[u2] = nxt[u2];
            }
        }
        __syncthreads();
        // ---- per feature: exclusive prefix over the warps, segment length, padded exclusive scan
        uint32_t out_len = 0;
        {
            uint32_t total = 0;
            if (tid < nf) {
                for (int w = 0; w < 32; ++w) {
                    const uint16_t c0 = cw[static_cast<size_t>(w) * kTrF + tid];
                    cw[static_cast<size_t>(w) * kTrF + tid] = static_cast<uint16_t>(total);
                    total += c0;
                }
                cnt[t * m + f0 + tid] = static_cast<uint16_t>(total);
            }
            const uint32_t v = (total + 7u) & ~7u;
            uint32_t x = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
                if (lane >= d) x += y;
            }
            if (lane == 31) wsum[warp] = x;
            __syncthreads();
            if (warp == 0) {
                uint32_t w = wsum[lane];
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t y = __shfl_up_sync(0xffffffffu, w, d);
                    if (lane >= d) w += y;
                }
                wsum[lane] = w;
            }
            __syncthreads();
            const uint32_t excl = (warp > 0 ? wsum[warp - 1] : 0u) + (x - v);
            if (tid < kTrF) segstart[tid] = excl;
            if (tid < nf) segoff[t * m + f0 + tid] = ureg[t * NBp + k] + excl;
            out_len = wsum[31];            // padded size of the unit's output (a multiple of 8)
        }
        __syncthreads();
        // ---- stable scatter: into the staging if the unit fits (it almost always does), else straight to global
        const bool staged = out_len <= static_cast<uint32_t>(cap);
        uint16_t* out_g = ids + tile_base[t] + ureg[t * NBp + k];
        {
            uint32_t cur[4], nxt[4];
            if (a < b) load4(a, cur);
            for (int p0 = a; p0 < b; p0 += 128) {
#pragma unroll
                for (int u2 = 0; u2 < 4; ++u2) nxt[u2] = 0xFFFFFFFFu;
                if (p0 + 128 < b) load4(p0 + 128, nxt);
#pragma unroll
                for (int u2 = 0; u2 < 4; ++u2) {
                    const uint32_t e = cur[u2];
                    const uint32_t f = (e >> 14) & (kTrF - 1);
                    const bool valid = e != 0xFFFFFFFFu;
                    unsigned same = __ballot_sync(0xffffffffu, valid);
#pragma unroll
                    for (int bit = 0; bit < kTrLog; ++bit) {
                        const bool one = (f >> bit) & 1u;
                        const unsigned mb = __ballot_sync(0xffffffffu, one);
                        same &= one ? mb : ~mb;
                    }
                    if (!valid) same = 1u << lane;
                    const int leader = __ffs(same) - 1;
                    uint32_t base = 0;
                    if (valid && lane == leader) {
                        base = mine[f];
                        mine[f] = static_cast<uint16_t>(base + __popc(same));
                    }
                    base = __shfl_sync(0xffffffffu, base, leader);
                    if (valid) {
                        const uint32_t slot = segstart[f] + base + __popc(same & ((1u << lane) - 1u));
                        const uint16_t row = static_cast<uint16_t>(e & 0x3FFFu);
                        if (staged) stage[slot] = row; else out_g[slot] = row;
                    }
                    __syncwarp();
                }
#pragma unroll
                for (int u2 = 0; u2 < 4; ++u2) cur[u2] = nxt[u2];
            }
        }
        if (staged) {
            __syncthreads();
            // the unit's whole output (segments of its features are adjacent) as one coalesced copy
            const uint4* src = reinterpret_cast<const uint4*>(stage);
            uint4* dst = reinterpret_cast<uint4*>(out_g);
            for (uint32_t i = tid; i < out_len / 8; i += 1024) dst[i] = src[i];
        }
    }
}

// returns false if a row did not fit the scatter pass's staging (the caller then uses the bitmap transpose)
static bool transpose_bucketed(snapb200_ctx* c, int tile_rows, int64_t* df_local) {
    const Csr& X = c->X;
    TileT& T = c->XtT;
    cudaStream_t st = c->stream;
    const int64_t n = X.nrows, m = c->m;
    const int nt = T.n_tiles;
    const bool debug = getenv("SNAPB200_DEBUG") != nullptr;
    cudaEvent_t dbg[6] = {};
    auto mark = [&](int i) {
        if (!debug) return;
        if (!dbg[i]) cudaEventCreate(&dbg[i]);
        cudaEventRecord(dbg[i], st);
    };
    mark(0);
    const int NB = static_cast<int>(ceil_div(m, kTrF));
    const int NBp = (NB + 31) / 32 * 32;
    const int SPT = static_cast<int>(ceil_div(tile_rows, kTrSlab));      // slabs per cell tile
    const int64_t n_slabs = static_cast<int64_t>(nt) * SPT;
    DevBuf<uint16_t> slabcnt, rowcnt;
    DevBuf<uint32_t> slaboff, ureg, tmp, trel, midcarry;
    DevBuf<int64_t> tile_tmp0, btot, tmp_base;
    DevBuf<unsigned long long> counter;
    DevBuf<int> too_long;
    too_long.alloc(1);
    SB_CUDA(cudaMemsetAsync(too_long.p, 0, sizeof(int), st));
    slabcnt.alloc(n_slabs * NBp);
    slaboff.alloc(n_slabs * NBp);
    rowcnt.alloc(std::max<int64_t>(1, n) * NBp + 8);
    midcarry.alloc(std::max<int64_t>(1, n));
    btot.alloc(static_cast<int64_t>(nt) * NBp);
    tmp_base.alloc(static_cast<int64_t>(nt) * NBp);
    ureg.alloc(static_cast<int64_t>(nt) * NBp);
    trel.alloc(static_cast<int64_t>(nt) * NBp);
    tile_tmp0.alloc(nt);
    counter.alloc(1);
    SB_CUDA(cudaMemsetAsync(counter.p, 0, sizeof(unsigned long long), st));
    const size_t smem_cnt = static_cast<size_t>(kTrWarps) * 2 * NBp * sizeof(uint16_t);
    const int walk_grid = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(ceil_div(n_slabs, kTrWarps), c->num_sms * 4)));
    SB_CUDA(cudaFuncSetAttribute(tr_slabcnt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_cnt)));
    tr_slabcnt_kernel<<<walk_grid, kTrWarps * 32, smem_cnt, st>>>(X.ptr.p, X.idx.p, n, tile_rows, SPT, n_slabs, NBp, rowcnt.p,
                                                                  slabcnt.p, midcarry.p);
    SB_LAUNCH_CHECK();
    mark(1);
    tr_rowscan_kernel<<<static_cast<unsigned>(static_cast<int64_t>(nt) * (NBp / 32)), 1024, 0, st>>>(slabcnt.p, n_slabs, SPT, NBp,
                                                                                                  slaboff.p, btot.p);
    SB_LAUNCH_CHECK();
    mark(2);
    // ---- bucket bases (host: nt x NB numbers): exact offsets into tmp, worst-case padded regions of the output
    std::vector<int64_t> hb(static_cast<size_t>(nt) * NBp), htb(static_cast<size_t>(nt) * NBp), tb(nt + 1);
    std::vector<uint32_t> hureg(static_cast<size_t>(nt) * NBp), hrel(static_cast<size_t>(nt) * NBp);
    std::vector<int64_t> ht0(nt);
    SB_CUDA(cudaMemcpyAsync(hb.data(), btot.p, sizeof(int64_t) * hb.size(), cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    int64_t run_tmp = 0, counted = 0;
    tb[0] = 0;
    for (int t = 0; t < nt; ++t) {
        int64_t reg = 0;
        ht0[t] = run_tmp;
        for (int k = 0; k < NBp; ++k) {
            const int64_t e = hb[static_cast<size_t>(t) * NBp + k];
            htb[static_cast<size_t>(t) * NBp + k] = run_tmp;
            hrel[static_cast<size_t>(t) * NBp + k] = static_cast<uint32_t>(run_tmp - ht0[t]);
            hureg[static_cast<size_t>(t) * NBp + k] = static_cast<uint32_t>(reg);
            run_tmp += e;
            counted += e;
            if (k < NB) {
                const int64_t nf = std::min<int64_t>(kTrF, m - static_cast<int64_t>(k) * kTrF);
                reg += (e + 7 * std::min<int64_t>(nf, e) + 7) / 8 * 8;   // every non-empty segment pads by at most 7
            }
        }
        SB_CHECK(reg < (1ll << 32) && run_tmp - ht0[t] < (1ll << 32), "transpose_tiled: more than 2^32 slots in one cell tile");
        tb[t + 1] = tb[t] + reg;
    }
    SB_CHECK(counted == X.nnz, "transpose_tiled: count mismatch (column index out of range or unsorted rows?)");
    SB_CUDA(cudaMemcpyAsync(tmp_base.p, htb.data(), sizeof(int64_t) * htb.size(), cudaMemcpyHostToDevice, st));
    SB_CUDA(cudaMemcpyAsync(ureg.p, hureg.data(), sizeof(uint32_t) * hureg.size(), cudaMemcpyHostToDevice, st));
    SB_CUDA(cudaMemcpyAsync(trel.p, hrel.data(), sizeof(uint32_t) * hrel.size(), cudaMemcpyHostToDevice, st));
    SB_CUDA(cudaMemcpyAsync(tile_tmp0.p, ht0.data(), sizeof(int64_t) * nt, cudaMemcpyHostToDevice, st));
    SB_CUDA(cudaMemcpyAsync(T.tile_base.p, tb.data(), sizeof(int64_t) * (nt + 1), cudaMemcpyHostToDevice, st));
    T.ids.alloc(tb[nt] + 8);
    tmp.alloc(std::max<int64_t>(1, X.nnz));
    if (n > 0 && X.nnz > 0) {
        const size_t tables = static_cast<size_t>(kTrSlab) * NBp * sizeof(uint16_t) + 3 * static_cast<size_t>(NBp) * sizeof(uint32_t);
        const int cap = static_cast<int>((kTrSmemMax - 256 - tables) / sizeof(uint16_t));
        const size_t smem = tables + static_cast<size_t>(cap) * sizeof(uint16_t);
        SB_CUDA(cudaFuncSetAttribute(tr_slab_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        const int grid = static_cast<int>(std::min<int64_t>(n_slabs, c->num_sms));
        tr_slab_scatter_kernel<<<grid, kTrWarps * 64, smem, st>>>(X.ptr.p, X.idx.p, n, tile_rows, SPT, n_slabs, NBp, cap, rowcnt.p,
                                                                  slaboff.p, trel.p, tile_tmp0.p, midcarry.p, tmp.p, too_long.p);
        SB_LAUNCH_CHECK();
        int h_long = 0;
        SB_CUDA(cudaMemcpyAsync(&h_long, too_long.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        SB_CUDA(cudaStreamSynchronize(st));
        if (h_long) return false;
    }
    mark(3);
    {
        const int64_t n_units = static_cast<int64_t>(nt) * NB;
        const size_t tables = static_cast<size_t>(32) * kTrF * sizeof(uint16_t) + kTrF * sizeof(uint32_t);
        const int cap = static_cast<int>((kTrSmemMax - 512 - tables) / sizeof(uint16_t)) / 8 * 8;
        const size_t smem = tables + static_cast<size_t>(cap) * sizeof(uint16_t);
        SB_CUDA(cudaFuncSetAttribute(tr_bucket_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        const int grid = static_cast<int>(std::min<int64_t>(n_units, c->num_sms));
        tr_bucket_sort_kernel<<<grid, 1024, smem, st>>>(tmp.p, tmp_base.p, btot.p, ureg.p, T.tile_base.p, m, NB, NBp, n_units, cap,
                                                        T.cnt.p, T.segoff.p, T.ids.p, counter.p);
        SB_LAUNCH_CHECK();
    }
    mark(4);
    if (df_local) {
        tile_df_kernel<<<static_cast<unsigned>(ceil_div(m, 256)), 256, 0, st>>>(T.cnt.p, m, nt, df_local);
        SB_LAUNCH_CHECK();
    }
    if (X.has_values() && X.nnz > 0) {
        T.vals.alloc(tb[nt] + 8);
        const int blocks = static_cast<int>(std::min<int64_t>(ceil_div(static_cast<int64_t>(nt) * m, 8),
                                                              static_cast<int64_t>(c->num_sms) * 16));
        tile_vals_kernel<<<blocks, 256, 0, st>>>(X.ptr.p, X.idx.p, X.val.p, m, tile_rows, nt, T.cnt.p, T.segoff.p,
                                                 T.tile_base.p, T.ids.p, T.vals.p);
        SB_LAUNCH_CHECK();
    }
    count_launch(c, 6);
    mark(5);
    SB_CUDA(cudaStreamSynchronize(st));   // host tables and temporaries stay alive until everything ran
    if (debug) {
        float ms[5] = {};
        for (int i = 0; i < 5; ++i)
            if (dbg[i] && dbg[i + 1]) cudaEventElapsedTime(&ms[i], dbg[i], dbg[i + 1]);
        fprintf(stderr, "[snapb200] bucketed transpose: slabcnt %.2f  scan %.2f  host tables %.2f  slab scatter %.2f  bucket sort %.2f ms\n",
                ms[0], ms[1], 0.0f, ms[2], ms[3]);
        for (auto& e : dbg) if (e) cudaEventDestroy(e);
    }
    T.built = true;
    return true;
}

}  // namespace

void transpose_tiled(snapb200_ctx* c, int tile_rows, int64_t* df_local) {
    SB_CHECK(tile_rows > 0 && tile_rows <= kTtMaxRows, "transpose_tiled: bad tile height");
    const Csr& X = c->X;
    TileT& T = c->XtT;
    cudaStream_t st = c->stream;
    const int64_t n = X.nrows, m = c->m;
    T.clear();
    T.tile_rows = tile_rows;
    T.n_tiles = static_cast<int>(std::max<int64_t>(1, ceil_div(n, tile_rows)));
    T.m = m;
    T.nnz = X.nnz;
    const int nt = T.n_tiles;
    T.cnt.alloc(static_cast<int64_t>(nt) * m);
    T.segoff.alloc(static_cast<int64_t>(nt) * m);
    T.tile_base.alloc(nt + 1);
    // Two transposes with the same output, bit for bit: the bitmap transpose above (round 1) and the bucketed
    // one below (round 2, default: 64 vs 80 ms on C3, 8.4 vs 11.7 ms on a 1/8 shard; profiles/README.md).
    // The bitmap transpose remains for matrices with more than kTrF * kTrMaxBuckets features or rows longer
    // than the scatter pass's staging, and under SNAPB200_TRANSPOSE=bitmap.
    const char* tr_mode = getenv("SNAPB200_TRANSPOSE");
    const bool bucketed = !(tr_mode != nullptr && tr_mode[0] == 'b' && tr_mode[1] == 'i');
    if (bucketed && ceil_div(m, kTrF) <= kTrMaxBuckets && tile_rows <= (1 << 14)) {
        if (transpose_bucketed(c, tile_rows, df_local)) return;
    }
    DevBuf<unsigned long long> counter;
    DevBuf<int64_t> totals;
    counter.alloc(2);
    totals.alloc(2 * nt);
    SB_CUDA(cudaMemsetAsync(counter.p, 0, 2 * sizeof(unsigned long long), st));

    // ---- count
    {
        const int n_hr = static_cast<int>(ceil_div(m, kTtHist));
        const int64_t n_units = static_cast<int64_t>(nt) * n_hr;
        const size_t smem = static_cast<size_t>(kTtHist / 2) * sizeof(uint32_t);
        SB_CUDA(cudaFuncSetAttribute(tile_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        const int grid = static_cast<int>(std::min<int64_t>(n_units, c->num_sms));
        tile_count_kernel<<<grid, kTtThreads, smem, st>>>(X.ptr.p, X.idx.p, n, m, tile_rows, n_hr, n_units, T.cnt.p,
                                                          counter.p);
        SB_LAUNCH_CHECK();
    }
    // ---- offsets, tile bases, document frequencies
    tile_scan_kernel<<<nt, 1024, 0, st>>>(T.cnt.p, m, T.segoff.p, totals.p);
    SB_LAUNCH_CHECK();
    if (df_local) {
        tile_df_kernel<<<static_cast<unsigned>(ceil_div(m, 256)), 256, 0, st>>>(T.cnt.p, m, nt, df_local);
        SB_LAUNCH_CHECK();
    }
    std::vector<int64_t> ht(2 * nt), hb(nt + 1);
    SB_CUDA(cudaMemcpyAsync(ht.data(), totals.p, sizeof(int64_t) * 2 * nt, cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    hb[0] = 0;
    int64_t counted = 0;
    for (int t = 0; t < nt; ++t) {
        SB_CHECK(ht[t] < (1ll << 32), "transpose_tiled: more than 2^32 stored entries in one cell tile");
        hb[t + 1] = hb[t] + ht[t];
        counted += ht[nt + t];
    }
    SB_CHECK(counted == X.nnz, "transpose_tiled: count mismatch (column index out of range or unsorted rows?)");
    SB_CUDA(cudaMemcpyAsync(T.tile_base.p, hb.data(), sizeof(int64_t) * (nt + 1), cudaMemcpyHostToDevice, st));
    T.ids.alloc(hb[nt] + 8);   // segments padded to 16-byte units

    // ---- emit
    if (X.nnz > 0) {
        const int n_ranges = static_cast<int>(ceil_div(m, kTtRange));
        int groups = static_cast<int>(std::min<int64_t>(n_ranges, std::max<int64_t>(1, ceil_div(4 * c->num_sms, nt))));
        const int rpg = static_cast<int>(ceil_div(n_ranges, groups));
        groups = static_cast<int>(ceil_div(n_ranges, rpg));
        const int64_t n_units = static_cast<int64_t>(nt) * groups;
        const size_t smem = (static_cast<size_t>(32) * kTtRange + kTtMaxRows + kTtRange) * sizeof(uint32_t) +
                            static_cast<size_t>(kTtRange) * sizeof(uint16_t);
        SB_CUDA(cudaFuncSetAttribute(tile_emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        const int grid = static_cast<int>(std::min<int64_t>(n_units, c->num_sms));
        tile_emit_kernel<<<grid, kTtThreads, smem, st>>>(X.ptr.p, X.idx.p, X.nnz, n, m, tile_rows, groups, rpg, n_ranges, n_units,
                                                         T.cnt.p, T.segoff.p, T.tile_base.p, T.ids.p, counter.p + 1);
        SB_LAUNCH_CHECK();
        if (X.has_values()) {
            T.vals.alloc(hb[nt] + 8);
            const int blocks = static_cast<int>(std::min<int64_t>(ceil_div(static_cast<int64_t>(nt) * m, 8),
                                                                  static_cast<int64_t>(c->num_sms) * 16));
            tile_vals_kernel<<<blocks, 256, 0, st>>>(X.ptr.p, X.idx.p, X.val.p, m, tile_rows, nt, T.cnt.p, T.segoff.p,
                                                     T.tile_base.p, T.ids.p, T.vals.p);
            SB_LAUNCH_CHECK();
        }
    }
    count_launch(c, 5);
    SB_CUDA(cudaStreamSynchronize(st));   // hb / counters stay alive until the copies and kernels are done
    T.built = true;
}

}  // namespace snapb

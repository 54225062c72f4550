// Builds the column-tiled sliced-ELL copy (struct Sell, ctx.cuh) of a CSR
// matrix for the shared-memory SpMM in spmm_tiled.cu.
//
// Why this layout: at 4 B of index per stored entry the SpMM would be HBM
// bound, but every entry also gathers a dense row of b fp32 (32 B for b = 8).
// Gathered from L2 that is ~5.7 TB/s of sector traffic and caps the kernel at
// ~10% of the HBM roofline (profiles/r01_bench_c3_v1_csr_gather.json).  Staging
// a column tile of the dense operand in shared memory moves the gathers on
// chip, and because a tile has at most 12288 columns an entry is stored as a
// 16-bit tile-local column: 2 bytes per stored entry instead of the 4 of CSR.
// What then limits the kernel is the shared-memory pipe, so the build also
// orders every lane's entries such that the eight lanes of a quarter warp read
// eight different 16-byte bank groups:
//   b = 8: dense row j sits at byte j*32 -> bank groups 2*(j%4), 2*(j%4)+1;
//          lane l reads half (l&1) first, so even lanes need distinct j%4 and
//          odd lanes need distinct j%4 within a quarter warp: 4 classes,
//          lane offset o = (l>>1) & 3;
//   b = 4: dense row j sits at byte j*16 -> bank group j%8: 8 classes,
//          lane offset o = l & 7;
//   lane l consumes class (k + o) % NC at step k: the q-th entry of class c is
//   placed at slot NC*q + ((c - o) mod NC) for the first R rounds (slots whose
//   class has run out stay empty: 0xFFFF), the remainder follows in column order.
// The chunk list is tile-major so that a persistent CTA streams one long
// contiguous range of it and touches only one or two tiles.
// Everything is deterministic (sorts with total order, no atomics on data).
#include "ctx.cuh"

#include <algorithm>
#include <vector>

namespace snapb {

namespace {

constexpr int kPlanThreads = 256;
constexpr uint32_t kLenBias = 32767;   // slots of a segment <= ~1.1 * tile_cols < 2^15

// Padded class rotation.  A lane piece with `len` entries consumes its classes in rotation for R
// rounds (slot NC*q + ((class - o) mod NC) holds the q-th entry of a class, empty if the class has
// fewer than q+1 entries); entries beyond R per class go to a tail after slot NC*R in column order,
// and if the tail is full into leftover holes of the rotation region.  R slightly below len/NC keeps
// ~90% of the entries in the conflict-free rotation; the piece owns len + a(len) slots (~11% padding
// for len ~ 100), a function of len alone so that the chunk plan needs no per-class statistics.
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
// segment length.  Every full piece owns the same number of slots (PieceSlots), which keeps the
// piece windows of the 32 lanes of a chunk aligned; only a lane's last, partial piece is shorter.
constexpr int kPiece = 256;
template <int NC> struct PieceSlots { static constexpr int value = (NC == 8) ? 280 : 272; };   // round8(padded_slots(kPiece, NC))
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

// Where the row segments of a (row, tile) pair come from.
//   CsrSegs:  a CSR matrix plus the per-row tile boundaries of seg_bounds_kernel (cell-major copy);
//   TileSegs: the tile-major 16-bit transpose (feature-major copy): entries are already tile-local.
struct CsrSegs {
    const int64_t* ptr;
    const int32_t* idx;
    const float* val;
    int64_t nnz;
    const int32_t* segptr;
    int n_tiles, tile_cols;
    static constexpr int kHeadMask = 3;   // entries per aligned 16-byte unit - 1
    __device__ __forceinline__ int len(int64_t row, int t) const {
        const int32_t* sp = segptr + row * (n_tiles + 1) + t;
        return sp[1] - sp[0];
    }
    __device__ __forceinline__ int64_t start(int64_t row, int t) const {
        return ptr[row] + segptr[row * (n_tiles + 1) + t];
    }
    __device__ __forceinline__ int col0(int t) const { return t * tile_cols; }
    __device__ __forceinline__ int at(int64_t p) const { return idx[p]; }
    // 16 consecutive entries starting at the 16-byte aligned position a (entries at or past nnz read as 0)
    __device__ __forceinline__ void load16(int64_t a, int n_needed, int (&out)[16]) const {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            int4 v = make_int4(0, 0, 0, 0);
            const int64_t q = a + 4 * u;
            if (4 * u < n_needed) {
                if (q + 3 < nnz) {
                    v = *reinterpret_cast<const int4*>(idx + q);
                } else {
                    if (q < nnz) v.x = idx[q];
                    if (q + 1 < nnz) v.y = idx[q + 1];
                    if (q + 2 < nnz) v.z = idx[q + 2];
                }
            }
            out[4 * u] = v.x; out[4 * u + 1] = v.y; out[4 * u + 2] = v.z; out[4 * u + 3] = v.w;
        }
    }
};
struct TileSegs {
    const uint16_t* cnt;
    const uint32_t* segoff;
    const int64_t* tile_base;
    const uint16_t* ids;     // allocated with 8 entries of slack
    const float* val;
    int64_t m;
    static constexpr int kHeadMask = 7;
    __device__ __forceinline__ int len(int64_t row, int t) const { return cnt[static_cast<int64_t>(t) * m + row]; }
    __device__ __forceinline__ int64_t start(int64_t row, int t) const {
        return tile_base[t] + segoff[static_cast<int64_t>(t) * m + row];
    }
    __device__ __forceinline__ int col0(int) const { return 0; }
    __device__ __forceinline__ int at(int64_t p) const { return ids[p]; }
    __device__ __forceinline__ void load16(int64_t a, int n_needed, int (&out)[16]) const {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (8 * u < n_needed) v = *reinterpret_cast<const uint4*>(ids + a + 8 * u);
            const uint32_t ww[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                out[8 * u + 2 * k] = static_cast<int>(ww[k] & 0xFFFFu);
                out[8 * u + 2 * k + 1] = static_cast<int>(ww[k] >> 16);
            }
        }
    }
};

// One CTA per (tile, window): sort the window's rows by segment length
// (descending, ties by row) and cut them into chunks of 32 lanes.
//
// Splitting.  A warp works through a chunk at 1/24 of an SM's shared-memory bandwidth, so one very
// long segment (a feature present in a large share of a tile's cells) can outlast the whole fair
// share of its CTA when the shard is small (C3 on 8 GPUs: pass 1 ran at 56% instead of 77% of the
// roofline).  Segments with more than `cap_slots` slots are therefore spread over p = 2..32
// adjacent lanes (whole 256-entry pieces each); the SpMM adds the lanes of a row up before its single
// store.  The longest rows come first in the sorted order and p is a non-increasing power of two,
// so a row's lanes never straddle a chunk.  Every (tile, window) owns kSplitChunks spare chunks
// for the extra lanes (unused ones have zero length).
constexpr int kMaxSplitRows = 32;   // rows split per (tile, window) at most (the longest ones)
__host__ __device__ __forceinline__ int pieces_of(int len) { return (len + kPiece - 1) / kPiece; }

template <typename SEGS>
__global__ void __launch_bounds__(kPlanThreads)
sell_plan_kernel(const int64_t* __restrict__ window_start, const int64_t* __restrict__ window_chunk0,
                 const SEGS segs, int n_tiles, int n_windows, int64_t chunks_per_tile, int nc, int cap_slots,
                 int32_t* __restrict__ chunk_rows, int32_t* __restrict__ chunk_span,
                 int32_t* __restrict__ chunk_groups) {
    __shared__ uint32_t keys[kSellWindowRows];
    __shared__ uint16_t split_slots[kMaxSplitRows * 32];   // slots of the lanes of split rows
    __shared__ int32_t split_row[kMaxSplitRows * 32];      // their row (window relative) and span
    __shared__ int32_t split_span[kMaxSplitRows * 32];
    __shared__ int s_nsplit, s_lsplit;
    const int t = blockIdx.x / n_windows, w = blockIdx.x % n_windows;
    const int64_t r0 = window_start[w];
    const int nr = static_cast<int>(window_start[w + 1] - r0);
    const int64_t base = static_cast<int64_t>(t) * chunks_per_tile + window_chunk0[w];
    const int nch = static_cast<int>(window_chunk0[w + 1] - window_chunk0[w]);   // chunks owned, spare ones included
    if (nch == 0) return;
    int pow2 = 2;
    while (pow2 < nr) pow2 <<= 1;
    for (int i = threadIdx.x; i < pow2; i += kPlanThreads) {
        uint32_t key = 0xFFFFFFFFu;
        if (i < nr) {
            const int len = total_slots(segs.len(r0 + i, t), nc);   // slots, not entries
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
    // ---- split rows: a prefix of the sorted order
    if (threadIdx.x == 0) {
        int ns = 0, lanes = 0;
        const int spare_lanes = (nch - (nr + 31) / 32) * 32;
        while (ns < kMaxSplitRows && ns < nr) {
            const uint32_t key = keys[ns];
            const int slots = static_cast<int>(kLenBias - (key >> 13));
            if (slots <= cap_slots) break;
            const int len = segs.len(r0 + (key & 8191u), t);
            const int np = pieces_of(len);
            int p = 2;
            while (p < 32 && p * cap_slots < slots) p <<= 1;
            while (p > np) p >>= 1;                      // at least one piece per lane
            if (p < 2 || lanes + p - 1 > spare_lanes) break;
            const int pp = (np + p - 1) / p;             // pieces per lane
            for (int j = 0; j < p; ++j) {
                const int lo = j * pp, cnt = max(0, min(np, lo + pp) - lo);
                int sl = 0;
                if (cnt > 0) sl = (lo + cnt == np) ? (cnt - 1) * piece_slots(nc) + padded_slots(len - (np - 1) * kPiece, nc)
                                                   : cnt * piece_slots(nc);
                split_slots[lanes + j] = static_cast<uint16_t>(sl);
                split_row[lanes + j] = static_cast<int32_t>(key & 8191u);
                split_span[lanes + j] = lo | (cnt << 16);
            }
            lanes += p;
            ++ns;
        }
        s_nsplit = ns;
        s_lsplit = lanes;
    }
    __syncthreads();
    const int ns = s_nsplit, ls = s_lsplit;
    const int n_lanes = ls + (nr - ns);
    for (int q = threadIdx.x; q < nch * 32; q += kPlanThreads) {
        int32_t row = -1, span = static_cast<int32_t>(0xFFFF0000u);   // whole segment
        if (q < ls) {
            row = static_cast<int32_t>(r0 + split_row[q]);
            span = split_span[q];
        } else if (q < n_lanes) {
            row = static_cast<int32_t>(r0 + (keys[ns + q - ls] & 8191u));
        }
        chunk_rows[base * 32 + q] = row;
        chunk_span[base * 32 + q] = span;
    }
    for (int ch = threadIdx.x; ch < nch; ch += kPlanThreads) {
        int mx = 0;
        for (int l = 0; l < 32; ++l) {
            const int q = ch * 32 + l;
            int sl = 0;
            if (q < ls) sl = split_slots[q];
            else if (q < n_lanes) sl = static_cast<int>(kLenBias - (keys[ns + q - ls] >> 13));
            mx = max(mx, sl);
        }
        chunk_groups[base + ch] = (mx + 7) >> 3;
    }
}

// Per-lane class counters, 8 bits each, in one or two 32-bit registers.  A counter is only ever
// incremented below the rotation depth R <= 63, so 8 bits are enough.
template <int NC>
struct ClassCount {
    uint32_t w0 = 0, w1 = 0;
    __device__ __forceinline__ int get(int c) const {
        const uint32_t x = (NC == 8 && (c & 4)) ? w1 : w0;
        return static_cast<int>(__byte_perm(x, 0u, 0x4440u | static_cast<uint32_t>(c & 3)));
    }
    __device__ __forceinline__ void inc(int c) {
        const uint32_t d = 1u << ((c & 3) * 8);
        if (NC == 8 && (c & 4)) w1 += d; else w0 += d;
    }
};

// Lanes that must read distinct bank-group classes in one LDS.128 wavefront: b = 4 (NC = 8): the 8
// lanes of a quarter warp; b = 8 (NC = 4): the 4 even or the 4 odd lanes of a quarter warp.
template <int NC>
__device__ __forceinline__ int group_member(int lane) { return NC == 8 ? (lane & 7) : ((lane >> 1) & 3); }

// ---- fill: staged in shared memory -----------------------------------------
// One warp per chunk, one lane per row segment.  A lane walks its segment once, in pieces of kPiece
// entries read with aligned 16-byte loads, and drops every entry at its slot of a per-lane column
// in shared memory (16-bit tile-local columns, 0xFFFF = empty); the column is then written out
// eight slots (16 bytes) at a time, so that the warp stores 512 contiguous bytes per instruction.
// Slot of the q-th entry of class cl in a piece of plen entries:
//   q < R (= rotation_depth(plen)):  NC*q + ((cl - o) mod NC)        (conflict-free rotation)
//   else, while the tail has room:   NC*R + (running tail count)      (column order)
//   else (tail full, ~1 lane in 5):  the leftover rotation holes in class order -- those are only
//        known once the class totals are, so such entries wait in a small per-lane list (walked
//        again from global memory if even that overflows, e.g. all columns in one class).
// With values, the same walk runs a second time staging the entry's position in the piece, and the
// write-out gathers val[] through it.
constexpr int kStageWarps = 12;
constexpr int kOvfCap = 16;
template <int NC> struct StagePitch { static constexpr int value = 8 * ((PieceSlots<NC>::value / 8) | 1); };   // u16 units; odd in 16-byte words: conflict-free LDS.128

template <bool HAS_VAL, int NC, typename SEGS>
__global__ void __launch_bounds__(kStageWarps * 32)
sell_fill_kernel(const SEGS segs, int64_t n_chunks, int64_t chunks_per_tile, const int32_t* __restrict__ chunk_rows,
                 const int32_t* __restrict__ chunk_span,
                 const int32_t* __restrict__ chunk_groups, const int64_t* __restrict__ chunk_off,
                 uint16_t* __restrict__ data, float* __restrict__ vals) {
    constexpr int SP = PieceSlots<NC>::value, PITCH = StagePitch<NC>::value;
    constexpr int kWarpStage = 32 * PITCH + kOvfCap * 32;   // u16 per warp
    extern __shared__ __align__(16) unsigned char stage_smem[];
    const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
    uint16_t* col = reinterpret_cast<uint16_t*>(stage_smem) + static_cast<size_t>(wic) * kWarpStage + lane * PITCH;
    uint16_t* ovf = reinterpret_cast<uint16_t*>(stage_smem) + static_cast<size_t>(wic) * kWarpStage + 32 * PITCH + lane;
    uint4* col128 = reinterpret_cast<uint4*>(col);
    const int64_t warp = static_cast<int64_t>(blockIdx.x) * kStageWarps + wic;
    const int64_t nwarps = static_cast<int64_t>(gridDim.x) * kStageWarps;
    const int o = group_member<NC>(lane);
    for (int64_t c = warp; c < n_chunks; c += nwarps) {
        const int steps = chunk_groups[c] * 8;
        if (steps == 0) continue;
        const int row = chunk_rows[c * 32 + lane];
        const int t = static_cast<int>(c / chunks_per_tile);
        const int col0 = segs.col0(t);
        const int64_t obase = chunk_off[c] * 256 + lane * 8;
        int64_t s = 0, e = 0;
        if (row >= 0) {
            s = segs.start(row, t);
            e = s + segs.len(row, t);
            const uint32_t span = static_cast<uint32_t>(chunk_span[c * 32 + lane]);
            if ((span >> 16) != 0xFFFFu) {   // one lane of a split row: pieces [lo, lo + cnt)
                s += static_cast<int64_t>(span & 0xFFFFu) * kPiece;
                e = min(e, s + static_cast<int64_t>(span >> 16) * kPiece);
            }
        }
        for (int k0 = 0; k0 < steps; k0 += SP, s += kPiece) {
            const int nst8 = min(SP, steps - k0) >> 3;
            const int plen = static_cast<int>(max(static_cast<int64_t>(0), min(e - s, static_cast<int64_t>(kPiece))));
            const int R = rotation_depth(plen, NC);
            const int tail_cap = padded_slots(plen, NC) - NC * R;
            // 16-byte aligned window over the piece: entry r (relative to s) sits at position s + r
            const int head = static_cast<int>(s & SEGS::kHeadMask);
#pragma unroll 1
            for (int pass = 0; pass < (HAS_VAL ? 2 : 1); ++pass) {
                for (int g = 0; g < nst8; ++g) col128[g] = make_uint4(~0u, ~0u, ~0u, ~0u);
                if (plen > 0) {
                    ClassCount<NC> cnt;
                    int left = 0, novf = 0;
                    for (int r0 = -head; r0 < plen; r0 += 16) {
                        int jj[16];
                        segs.load16(s + r0, plen - r0, jj);
#pragma unroll
                        for (int w = 0; w < 16; ++w) {
                            const int r = r0 + w;
                            if (static_cast<unsigned>(r) >= static_cast<unsigned>(plen)) continue;
                            const int cl = jj[w] & (NC - 1);
                            const int q = cnt.get(cl);
                            const uint16_t item = static_cast<uint16_t>(pass == 0 ? jj[w] - col0 : r);
                            if (q < R) {
                                cnt.inc(cl);
                                col[NC * q + ((cl - o) & (NC - 1))] = item;
                            } else if (left < tail_cap) {
                                col[NC * R + left++] = item;
                            } else {
                                if (novf < kOvfCap) ovf[novf * 32] = item;
                                ++novf;
                            }
                        }
                    }
                    if (novf > 0) {   // tail overflow: fill the leftover holes of the rotation region in class order
                        int hole_c = 0, hole_q = cnt.get(0), left2 = 0;
                        ClassCount<NC> c2;
                        const bool rewalk = novf > kOvfCap;   // the list was too small: walk the piece again
                        const int n_items = rewalk ? plen : novf;
                        for (int i = 0; i < n_items; ++i) {
                            uint16_t item;
                            if (rewalk) {
                                const int j = segs.at(s + i);
                                const int cl = j & (NC - 1);
                                if (c2.get(cl) < R) { c2.inc(cl); continue; }
                                if (left2 < tail_cap) { ++left2; continue; }
                                item = static_cast<uint16_t>(pass == 0 ? j - col0 : i);
                            } else {
                                item = ovf[i * 32];
                            }
                            while (hole_q >= R) { ++hole_c; hole_q = cnt.get(hole_c); }
                            col[NC * hole_q + ((hole_c - o) & (NC - 1))] = item;
                            ++hole_q;
                        }
                    }
                }
                // write-out: 8 slots per lane and instruction, 512 contiguous bytes per warp
                const int64_t ob = obase + static_cast<int64_t>(k0) * 32;
                for (int g = 0; g < nst8; ++g) {
                    const uint4 w8 = col128[g];
                    if (pass == 0) {
                        __stcs(reinterpret_cast<uint4*>(data + ob + static_cast<int64_t>(g) * 256), w8);
                    } else {
                        const uint32_t ww[4] = {w8.x, w8.y, w8.z, w8.w};
                        float f[8];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const uint32_t lo = ww[i] & 0xFFFFu, hi = ww[i] >> 16;
                            f[2 * i] = lo == 0xFFFFu ? 0.f : segs.val[s + lo];
                            f[2 * i + 1] = hi == 0xFFFFu ? 0.f : segs.val[s + hi];
                        }
                        float4* dst = reinterpret_cast<float4*>(vals + ob + static_cast<int64_t>(g) * 256);
                        __stcs(dst, make_float4(f[0], f[1], f[2], f[3]));
                        __stcs(dst + 1, make_float4(f[4], f[5], f[6], f[7]));
                    }
                }
            }
        }
    }
}

// Windows, plan, offsets and fill, common to both segment sources.  `make_segs(T)` is called once
// the number of column tiles is known (the CSR source needs it to size its boundary table).
template <typename SEGS, typename MAKE>
void build_impl(snapb200_ctx* c, int64_t nrows, int64_t ncols, int64_t nnz_estimate, bool has_values, Sell& S, int b,
                MAKE make_segs) {
    SB_CHECK(b == 4 || b == 8, "tiled format: block width must be 4 or 8");
    cudaStream_t st = c->stream;
    S.clear();
    const int64_t R = nrows;
    S.b = b;
    S.nrows = R;
    S.ncols = ncols;
    S.tile_cols = kSellTileBytes / (4 * b);
    S.n_tiles = static_cast<int>(std::max<int64_t>(1, ceil_div(ncols, S.tile_cols)));
    const int T = S.n_tiles;

    // ---- windows of <= 8192 consecutive rows (equal row counts)
    const int nw = static_cast<int>(std::max<int64_t>(1, ceil_div(R, kSellWindowRows)));
    S.n_windows = nw;
    std::vector<int64_t> ws(nw + 1), wc(nw + 1);
    wc[0] = 0;
    for (int w = 0; w <= nw; ++w) ws[w] = (R * w) / nw;
    // a chunk may last at most about half of a warp's fair share of the pass (24 warps x #SMs);
    // longer segments are split over lanes, for which every (tile, window) gets spare chunks
    const int nc = (b == 8) ? 4 : 8;
    const int64_t est_steps = (nnz_estimate + nnz_estimate / 6) / 32 + 1;   // per-lane slots in total (~1.16 x entries / 32)
    const int cap_slots = static_cast<int>(std::min<int64_t>(
        1 << 20, std::max<int64_t>(piece_slots(nc), est_steps / (static_cast<int64_t>(c->num_sms) * 48))));
    const bool may_split = cap_slots < total_slots(S.tile_cols, nc);
    const int spare = may_split ? kMaxSplitRows : 0;
    for (int w = 0; w < nw; ++w) wc[w + 1] = wc[w] + ceil_div(ws[w + 1] - ws[w], 32) + (ws[w + 1] > ws[w] ? spare : 0);
    S.chunks_per_tile = wc[nw];
    S.n_chunks = S.chunks_per_tile * T;
    S.window_start.alloc(nw + 1);
    S.window_chunk0.alloc(nw + 1);
    SB_CUDA(cudaMemcpyAsync(S.window_start.p, ws.data(), sizeof(int64_t) * (nw + 1), cudaMemcpyHostToDevice, st));
    SB_CUDA(cudaMemcpyAsync(S.window_chunk0.p, wc.data(), sizeof(int64_t) * (nw + 1), cudaMemcpyHostToDevice, st));
    const int64_t n_chunks = S.n_chunks;

    SB_CHECK(piece_slots(nc) == ((padded_slots(kPiece, nc) + 7) & ~7), "piece slot table out of date");
    const SEGS segs = make_segs(T, S.tile_cols);

    // ---- plan: sorted chunk membership and chunk lengths, then offsets
    S.chunk_rows.alloc(std::max<int64_t>(1, n_chunks * 32));
    S.chunk_span.alloc(std::max<int64_t>(1, n_chunks * 32));
    S.chunk_groups.alloc(std::max<int64_t>(1, n_chunks));
    S.chunk_off.alloc(n_chunks + 1);
    if (n_chunks > 0) {
        sell_plan_kernel<SEGS><<<static_cast<unsigned>(static_cast<int64_t>(T) * nw), kPlanThreads, 0, st>>>(
            S.window_start.p, S.window_chunk0.p, segs, T, nw, S.chunks_per_tile, nc, cap_slots, S.chunk_rows.p,
            S.chunk_span.p, S.chunk_groups.p);
        SB_LAUNCH_CHECK();
    }
    exclusive_scan_i32_to_i64(c, S.chunk_groups.p, S.chunk_off.p, n_chunks);
    int64_t n_groups = 0;
    SB_CUDA(cudaMemcpyAsync(&n_groups, S.chunk_off.p + n_chunks, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));   // also keeps ws / wc alive until the copies are done
    S.n_entries = n_groups * 256;

    // ---- fill
    S.data.alloc(std::max<int64_t>(8, S.n_entries));
    if (has_values) S.vals.alloc(std::max<int64_t>(8, S.n_entries));
    if (n_chunks > 0 && S.n_entries > 0) {
        const int blocks = static_cast<int>(std::min<int64_t>(ceil_div(n_chunks, kStageWarps), c->num_sms));
#define SB_FILL(HV, NC)                                                                                                \
    do {                                                                                                               \
        const size_t fsm = static_cast<size_t>(kStageWarps) * (32 * StagePitch<NC>::value + kOvfCap * 32) * 2;         \
        auto kern = sell_fill_kernel<HV, NC, SEGS>;                                                                    \
        SB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(fsm)));       \
        kern<<<blocks, kStageWarps * 32, fsm, st>>>(segs, n_chunks, S.chunks_per_tile, S.chunk_rows.p, S.chunk_span.p, \
                                                    S.chunk_groups.p, S.chunk_off.p, S.data.p, S.vals.p);              \
    } while (0)
        if (b == 8) { if (has_values) SB_FILL(true, 4); else SB_FILL(false, 4); }
        else        { if (has_values) SB_FILL(true, 8); else SB_FILL(false, 8); }
#undef SB_FILL
        SB_LAUNCH_CHECK();
    }
    count_launch(c, 3);
    SB_CUDA(cudaStreamSynchronize(st));   // temporaries of make_segs (boundary table) are released by the caller
    S.built = true;
}

}  // namespace

void sell_build(snapb200_ctx* c, const Csr& M, Sell& S, int b) {
    DevBuf<int32_t> segptr;
    build_impl<CsrSegs>(c, M.nrows, M.ncols, M.nnz, M.has_values(), S, b, [&](int T, int tile_cols) {
        // per-row tile boundaries
        const int64_t nseg = M.nrows * (T + 1);
        segptr.alloc(std::max<int64_t>(1, nseg));
        if (nseg > 0) {
            seg_bounds_kernel<<<static_cast<unsigned>(ceil_div(nseg, 256)), 256, 0, c->stream>>>(M.ptr.p, M.idx.p, M.nrows, T,
                                                                                              tile_cols, segptr.p);
            SB_LAUNCH_CHECK();
        }
        return CsrSegs{M.ptr.p, M.idx.p, M.val.p, M.nnz, segptr.p, T, tile_cols};
    });
}

void sell_build_transposed(snapb200_ctx* c, const TileT& Tt, int64_t n_cells, Sell& S, int b) {
    SB_CHECK(Tt.built, "tiled format: the tile-major transpose is missing");
    SB_CHECK(Tt.tile_rows == kSellTileBytes / (4 * b), "tiled format: transpose tile height does not match the block width");
    build_impl<TileSegs>(c, Tt.m, n_cells, Tt.nnz, Tt.vals.p != nullptr, S, b, [&](int T, int) {
        SB_CHECK(T == Tt.n_tiles, "tiled format: tile count mismatch");
        return TileSegs{Tt.cnt.p, Tt.segoff.p, Tt.tile_base.p, Tt.ids.p, Tt.vals.p, Tt.m};
    });
}

}  // namespace snapb

// Host <-> device staging for pageable host arrays (the scipy CSR of an in-memory AnnData,
// reference boundary: `adata.X` handed to internal.spectral_embedding, tools/_embedding.py:249,
// converted by `try_convert::<CsrMatrix<f64>>` at embedding.rs:36-41).
//
// A scipy CSR lives in pageable memory, its indices are int64 as soon as nnz > 2^31 and its
// values are whatever dtype the user stored.  The device wants int32 indices and -- unless the
// matrix is binarised -- f32 values.  A plain cudaMemcpy from pageable memory is staged by the
// driver on one thread (~10 GB/s) and would ship twice the index bytes; instead a team of host
// threads converts chunk by chunk straight into a ring of pinned buffers (narrowing the
// indices on the way, non-temporal stores) and each chunk's DMA is issued as soon as it is
// filled, so the conversion of chunk j + 1 overlaps the transfer of chunk j and the PCIe link
// only ever carries 4 bytes per stored entry.  Whether the values are all ones is decided by a
// threaded scan on the host: a binarised matrix never ships its values at all.
#include "ctx.cuh"
#include "delta_encode.h"

#include <immintrin.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <thread>
#include <vector>

namespace snapb {

int host_threads(const snapb200_ctx* c) {
    if (const char* e = getenv("SNAPB200_THREADS")) return std::max(1, atoi(e));
    int hw = static_cast<int>(std::thread::hardware_concurrency());
    if (hw <= 0) hw = 8;
    int local = 1;
    if (const char* e = getenv("LOCAL_WORLD_SIZE")) local = std::max(1, atoi(e));
    else if (c && c->nranks > 1) local = c->nranks;
    return std::max(1, std::min(16, hw / local));
}

namespace {

constexpr size_t kChunkBytes = 8u << 20;   // bytes of one pinned slot (2 Mi 4-byte elements)

struct Ring {
    snapb200_ctx* c;
    int slots;
    unsigned char* base;
    std::vector<cudaEvent_t>& ev;
    unsigned char* slot(int s) const { return base + static_cast<size_t>(s) * kChunkBytes; }
};

Ring ring_of(snapb200_ctx* c, int threads) {
    const int slots = 2 * threads;
    c->ring.ensure(static_cast<int64_t>(slots) * static_cast<int64_t>(kChunkBytes));
    while (static_cast<int>(c->ring_events.size()) < slots) {
        cudaEvent_t e;
        SB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->ring_events.push_back(e);
    }
    return Ring{c, slots, c->ring.p, c->ring_events};
}

// wall-clock accounting of the staging team (printed under SNAPB200_DEBUG): nanoseconds summed over threads
std::atomic<int64_t> g_ns_wait{0}, g_ns_work{0}, g_ns_api{0};
inline int64_t now_ns() {
    return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Run `work(j)` for j in [0, n_chunks) on a team of threads; chunk j uses ring slot j % slots and
// may only touch it once chunk j - slots has released it (release order is tracked per slot).
// `work(j, slot_ptr, event)` must leave the slot's last device operation recorded in `event`
// (H2D) or finish with the slot entirely (D2H).
template <typename Work>
void run_ring(Ring& r, int64_t n_chunks, int threads, Work work) {
    std::vector<std::atomic<int64_t>> released(r.slots);
    for (auto& a : released) a.store(-1);
    std::atomic<int64_t> next{0};
    std::atomic<int> failed{0};
    std::string err;
    auto body = [&]() {
        cudaSetDevice(r.c->device);
        try {
            while (!failed.load()) {
                const int64_t j = next.fetch_add(1);
                if (j >= n_chunks) break;
                const int s = static_cast<int>(j % r.slots);
                const int64_t want = j - r.slots;
                const int64_t t_w0 = now_ns();
                if (want >= 0) {
                    while (released[s].load(std::memory_order_acquire) != want) {
                        if (failed.load()) return;
                        std::this_thread::yield();
                    }
                    SB_CUDA(cudaEventSynchronize(r.ev[s]));
                }
                g_ns_wait.fetch_add(now_ns() - t_w0, std::memory_order_relaxed);
                work(j, r.slot(s), r.ev[s]);
                released[s].store(j, std::memory_order_release);
            }
        } catch (const std::exception& e) {
            if (!failed.exchange(1)) err = e.what();
        }
    };
    std::vector<std::thread> team;
    const int nt = static_cast<int>(std::min<int64_t>(threads, std::max<int64_t>(1, n_chunks)));
    for (int t = 1; t < nt; ++t) team.emplace_back(body);
    body();
    for (auto& t : team) t.join();
    if (failed.load()) throw Error(err.empty() ? "host staging failed" : err);
}

// ---- conversions into a pinned slot (4-byte outputs) ------------------------------------------
// int64 -> int32 with non-temporal stores; `orall` collects the OR of every input (bits >= 31 set
// means a negative or >= 2^31 index somewhere)
void narrow_i64(const int64_t* in, int32_t* out, int64_t n, uint64_t& orall) {
    int64_t i = 0;
    uint64_t acc = 0;
#if defined(__AVX2__)
    __m256i vor = _mm256_setzero_si256();
    const __m256i pick = _mm256_setr_epi32(0, 2, 4, 6, 0, 2, 4, 6);
    for (; i + 8 <= n; i += 8) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(in + i));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(in + i + 4));
        vor = _mm256_or_si256(vor, _mm256_or_si256(a, b));
        const __m256i lo = _mm256_permutevar8x32_epi32(a, pick);
        const __m256i hi = _mm256_permutevar8x32_epi32(b, pick);
        const __m256i v = _mm256_permute2x128_si256(lo, hi, 0x20);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(out + i), v);   // slots are 32-byte aligned, i % 8 == 0
    }
    alignas(32) uint64_t tmp[4];
    _mm256_store_si256(reinterpret_cast<__m256i*>(tmp), vor);
    acc = tmp[0] | tmp[1] | tmp[2] | tmp[3];
    _mm_sfence();
#endif
    for (; i < n; ++i) {
        acc |= static_cast<uint64_t>(in[i]);
        out[i] = static_cast<int32_t>(in[i]);
    }
    orall |= acc;
}

template <typename T>
void to_f32(const T* in, float* out, int64_t n) {
    for (int64_t i = 0; i < n; ++i) out[i] = static_cast<float>(in[i]);
}

template <typename T>
bool all_ones_t(const T* v, int64_t n) {
    // branch-free accumulation (vectorises); early exit per 64 Ki block
    for (int64_t b0 = 0; b0 < n; b0 += 65536) {
        const int64_t b1 = std::min(n, b0 + 65536);
        int bad = 0;
        for (int64_t i = b0; i < b1; ++i) bad |= (v[i] != static_cast<T>(1));
        if (bad) return false;
    }
    return true;
}

size_t kind_size(int kind) {
    switch (kind) {
        case 1: case 3: case 4: return 4;
        case 2: case 5: case 6: return 8;
        case 7: case 8: return 1;
        case 9: case 10: return 2;
        default: throw Error("unknown value_kind");
    }
}

bool all_ones_kind(const void* v, int kind, int64_t n) {
    switch (kind) {
        case 1: return all_ones_t(static_cast<const float*>(v), n);
        case 2: return all_ones_t(static_cast<const double*>(v), n);
        case 3: return all_ones_t(static_cast<const uint32_t*>(v), n);
        case 4: return all_ones_t(static_cast<const int32_t*>(v), n);
        case 5: return all_ones_t(static_cast<const int64_t*>(v), n);
        case 6: return all_ones_t(static_cast<const uint64_t*>(v), n);
        case 7: return all_ones_t(static_cast<const uint8_t*>(v), n);
        case 8: return all_ones_t(static_cast<const int8_t*>(v), n);
        case 9: return all_ones_t(static_cast<const uint16_t*>(v), n);
        case 10: return all_ones_t(static_cast<const int16_t*>(v), n);
        default: throw Error("unknown value_kind");
    }
}

void to_f32_kind(const void* v, int kind, int64_t off, int64_t n, float* out) {
    switch (kind) {
        case 1: memcpy(out, static_cast<const float*>(v) + off, sizeof(float) * n); break;
        case 2: to_f32(static_cast<const double*>(v) + off, out, n); break;
        case 3: to_f32(static_cast<const uint32_t*>(v) + off, out, n); break;
        case 4: to_f32(static_cast<const int32_t*>(v) + off, out, n); break;
        case 5: to_f32(static_cast<const int64_t*>(v) + off, out, n); break;
        case 6: to_f32(static_cast<const uint64_t*>(v) + off, out, n); break;
        case 7: to_f32(static_cast<const uint8_t*>(v) + off, out, n); break;
        case 8: to_f32(static_cast<const int8_t*>(v) + off, out, n); break;
        case 9: to_f32(static_cast<const uint16_t*>(v) + off, out, n); break;
        case 10: to_f32(static_cast<const int16_t*>(v) + off, out, n); break;
        default: throw Error("unknown value_kind");
    }
}

// ---- delta-encoded index transfer: encoder in delta_encode.h, decoder here ---------------------------
// one CTA per tile: segmented inclusive prefix sum (a marker restarts the sum at its side value)
__global__ void __launch_bounds__(256)
decode_deltas_kernel(const unsigned char* __restrict__ chunk, int32_t* __restrict__ out) {
    const DeltaHeader* hd = reinterpret_cast<const DeltaHeader*>(chunk);
    const int n = static_cast<int>(hd->n_entries), n_tiles = static_cast<int>(hd->n_tiles);
    const uint16_t* d16 = reinterpret_cast<const uint16_t*>(chunk + sizeof(DeltaHeader));
    const size_t d_bytes = (static_cast<size_t>(n) * 2 + 31) & ~static_cast<size_t>(31);
    const uint32_t* tile_side = reinterpret_cast<const uint32_t*>(chunk + sizeof(DeltaHeader) + d_bytes);
    const int32_t* side = reinterpret_cast<const int32_t*>(tile_side + n_tiles + 1);
    __shared__ int32_t s_out[kDeltaTile];
    __shared__ int s_wv[8], s_wf[8], s_wc[8];
    const int t = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int a = t * kDeltaTile;
    const int e0 = a + tid * 8;
    // thread-local fold of 8 consecutive entries: value since the last marker (or the plain sum), marker seen?, markers
    int dv[8];
    unsigned marks = 0;
    int v = 0, f = 0, cnt = 0;
    uint4 raw = make_uint4(0, 0, 0, 0);
    if (e0 < n) raw = *reinterpret_cast<const uint4*>(d16 + e0);        // 16-byte aligned; the delta area is padded to 32 B
    const unsigned rw[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int i = e0 + q;
        dv[q] = (i < n) ? static_cast<int>((rw[q >> 1] >> ((q & 1) * 16)) & 0xFFFFu) : 0;
        if (dv[q] == 0xFFFF) { f = 1; ++cnt; marks |= 1u << q; }        // the side value is placed once its rank is known
    }
    // my markers' side values: rank = markers before me in the tile
    int c_incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, c_incl, d);
        if (lane >= d) c_incl += y;
    }
    if (lane == 31) s_wc[warp] = c_incl;
    __syncthreads();
    int c_before = c_incl - cnt;
    for (int w = 0; w < warp; ++w) c_before += s_wc[w];
    const int32_t* my_side = side + tile_side[t] + c_before;
    // redo the local fold with the side values in place
    int k = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        if (marks >> q & 1u) v = my_side[k++];
        else v += dv[q];
        dv[q] = v;                                                      // running value since the thread's start / last marker
    }
    // warp scan of (f, v) with  (f1,v1) o (f2,v2) = (f1|f2, f2 ? v2 : v1 + v2)
    int sv = v, sf = f;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int yv = __shfl_up_sync(0xffffffffu, sv, d), yf = __shfl_up_sync(0xffffffffu, sf, d);
        if (lane >= d) { sv = sf ? sv : yv + sv; sf = sf | yf; }
    }
    if (lane == 31) { s_wv[warp] = sv; s_wf[warp] = sf; }
    __syncthreads();
    // carry into this thread = scan over the previous warps, then over the previous lanes of this warp
    int cv = 0, cf = 0;
    for (int w = 0; w < warp; ++w) { cv = s_wf[w] ? s_wv[w] : cv + s_wv[w]; cf |= s_wf[w]; }
    const int pv = __shfl_up_sync(0xffffffffu, sv, 1), pf = __shfl_up_sync(0xffffffffu, sf, 1);
    if (lane > 0) { cv = pf ? pv : cv + pv; }
    // apply the carry to the entries before the thread's first marker
    bool seen = false;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        seen = seen || (marks >> q & 1u);
        s_out[tid * 8 + q] = seen ? dv[q] : dv[q] + cv;
    }
    __syncthreads();
    const int cnt_tile = min(kDeltaTile, n - a);
    for (int i = tid; i < cnt_tile; i += 256) out[a + i] = s_out[i];
}

}  // namespace

// Host-only check of the encoder (no device work: usable by the CPU test suite): encode `count`
// indices chunk by chunk exactly as stage_indices does and replay the chunk format with a scalar
// loop.  0 = every index came back, 1 = a chunk overflowed its side list (stage_indices would ship
// plain int32), -1 = mismatch.  `n_side` receives the number of side-list entries.
int delta_selftest_host(const void* src, int bits, int64_t count, int64_t* n_side) {
    std::vector<unsigned char> slot_store(kChunkBytes + 64);
    unsigned char* slot = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(slot_store.data()) + 63) & ~static_cast<uintptr_t>(63));
    std::vector<int32_t> back(static_cast<size_t>(std::min<int64_t>(count, kDeltaPer)));
    int64_t total_side = 0;
    for (int64_t off = 0; off < count; off += kDeltaPer) {
        const int64_t len = std::min(kDeltaPer, count - off);
        uint64_t o = 0;
        size_t used = 0;
        const bool ok = (bits == 64)
            ? encode_deltas(static_cast<const int64_t*>(src) + off, len, slot, kChunkBytes, o, used)
            : encode_deltas(static_cast<const int32_t*>(src) + off, len, slot, kChunkBytes, o, used);
        if (!ok) return 1;
        if (used > kChunkBytes) return -1;
        if (!decode_deltas_host(slot, len, back.data())) return -1;
        uint64_t want_or = 0;
        for (int64_t i = 0; i < len; ++i) {
            const int64_t want = (bits == 64) ? static_cast<const int64_t*>(src)[off + i]
                                              : static_cast<int64_t>(static_cast<const int32_t*>(src)[off + i]);
            want_or |= static_cast<uint64_t>(want);
            if (back[i] != static_cast<int32_t>(want)) return -1;
        }
        if ((want_or >> 31 != 0) != (o >> 31 != 0)) return -1;      // the range verdict is part of the contract
        total_side += reinterpret_cast<const DeltaHeader*>(slot)->n_side;
    }
    if (n_side) *n_side = total_side;
    return 0;
}

// indices (int32 or int64, pageable or pinned host memory) -> int32 on the device.
// Returns false if some 64-bit index does not fit 31 bits (the range check against m runs on the device).
bool stage_indices(snapb200_ctx* c, const void* src, int bits, int64_t count, int32_t* dst_dev) {
    if (count == 0) return true;
    const int threads = host_threads(c);
    Ring r = ring_of(c, threads);
    std::atomic<uint64_t> orall{0};
    cudaStream_t st = c->stream;
    const bool no_delta = getenv("SNAPB200_NO_DELTA") != nullptr;
    const int64_t t_all0 = now_ns();
    g_ns_work.store(0); g_ns_api.store(0); g_ns_wait.store(0);
    if (!no_delta) {
        // 2 bytes per entry over PCIe: deltas + markers, decoded on the device (see above)
        c->ring_dev.ensure(static_cast<int64_t>(r.slots) * static_cast<int64_t>(kChunkBytes));
        const int64_t n_chunks = ceil_div(count, kDeltaPer);
        std::atomic<int> overflow{0};
        std::atomic<int64_t> shipped{0};
        run_ring(r, n_chunks, threads, [&](int64_t j, unsigned char* slot, cudaEvent_t ev) {
            const int64_t off = j * kDeltaPer, len = std::min(kDeltaPer, count - off);
            if (overflow.load()) {
                SB_CUDA(cudaEventRecord(ev, st));
                return;
            }
            uint64_t o = 0;
            size_t used = 0;
            const int64_t t_e0 = now_ns();
            const bool ok = (bits == 64)
                ? encode_deltas(static_cast<const int64_t*>(src) + off, len, slot, kChunkBytes, o, used)
                : encode_deltas(static_cast<const int32_t*>(src) + off, len, slot, kChunkBytes, o, used);
            const int64_t t_e1 = now_ns();
            g_ns_work.fetch_add(t_e1 - t_e0, std::memory_order_relaxed);
            if (o >> 31) orall.fetch_or(o);
            if (!ok) {
                overflow.store(1);
                SB_CUDA(cudaEventRecord(ev, st));
                return;
            }
            unsigned char* dslot = c->ring_dev.p + static_cast<size_t>(j % r.slots) * kChunkBytes;
            shipped.fetch_add(static_cast<int64_t>(used));
            SB_CUDA(cudaMemcpyAsync(dslot, slot, used, cudaMemcpyHostToDevice, st));
            decode_deltas_kernel<<<static_cast<unsigned>(ceil_div(len, kDeltaTile)), 256, 0, st>>>(dslot, dst_dev + off);
            SB_CUDA(cudaGetLastError());
            SB_CUDA(cudaEventRecord(ev, st));      // the slot (host and device side) is free once the decode has run
            g_ns_api.fetch_add(now_ns() - t_e1, std::memory_order_relaxed);
        });
        const int64_t t_s0 = now_ns();
        SB_CUDA(cudaStreamSynchronize(st));
        if (getenv("SNAPB200_DEBUG"))
            fprintf(stderr, "[snapb200] stage_indices(delta): %lld entries, %d threads, wall %.1f ms (tail sync %.1f); per-thread sums: encode %.1f ms, "
                    "cuda calls %.1f ms, slot waits %.1f ms\n", static_cast<long long>(count), threads, (now_ns() - t_all0) * 1e-6,
                    (now_ns() - t_s0) * 1e-6, g_ns_work.exchange(0) * 1e-6, g_ns_api.exchange(0) * 1e-6, g_ns_wait.exchange(0) * 1e-6);
        if (!overflow.load()) {
            c->stats.bytes_h2d_indices += shipped.load();
            return (orall.load() >> 31) == 0;
        }
        orall.store(0);     // pathological gaps: ship plain int32 instead
    }
    const int64_t per = static_cast<int64_t>(kChunkBytes / 4);
    const int64_t n_chunks = ceil_div(count, per);
    run_ring(r, n_chunks, threads, [&](int64_t j, unsigned char* slot, cudaEvent_t ev) {
        const int64_t off = j * per, len = std::min(per, count - off);
        int32_t* out = reinterpret_cast<int32_t*>(slot);
        const int64_t t_e0 = now_ns();
        if (bits == 64) {
            uint64_t o = 0;
            narrow_i64(static_cast<const int64_t*>(src) + off, out, len, o);
            if (o >> 31) orall.fetch_or(o);
        } else {
            memcpy(out, static_cast<const int32_t*>(src) + off, sizeof(int32_t) * len);
        }
        const int64_t t_e1 = now_ns();
        g_ns_work.fetch_add(t_e1 - t_e0, std::memory_order_relaxed);
        SB_CUDA(cudaMemcpyAsync(dst_dev + off, out, sizeof(int32_t) * len, cudaMemcpyHostToDevice, st));
        SB_CUDA(cudaEventRecord(ev, st));
        g_ns_api.fetch_add(now_ns() - t_e1, std::memory_order_relaxed);
    });
    const int64_t t_s0 = now_ns();
    SB_CUDA(cudaStreamSynchronize(st));
    if (getenv("SNAPB200_DEBUG"))
        fprintf(stderr, "[snapb200] stage_indices(plain): %lld entries, %d threads, wall %.1f ms (tail sync %.1f); per-thread sums: convert %.1f ms, "
                "cuda calls %.1f ms, slot waits %.1f ms\n", static_cast<long long>(count), threads, (now_ns() - t_all0) * 1e-6,
                (now_ns() - t_s0) * 1e-6, g_ns_work.exchange(0) * 1e-6, g_ns_api.exchange(0) * 1e-6, g_ns_wait.exchange(0) * 1e-6);
    c->stats.bytes_h2d_indices += 4 * count;
    return (orall.load() >> 31) == 0;
}

// threaded scan: is every stored value exactly 1?
bool host_values_all_ones(snapb200_ctx* c, const void* values, int kind, int64_t count) {
    if (count == 0) return true;
    const int threads = host_threads(c);
    const size_t es = kind_size(kind);
    const int64_t per = 4 << 20;
    const int64_t n_chunks = ceil_div(count, per);
    std::atomic<int64_t> next{0};
    std::atomic<int> not_one{0};
    auto body = [&]() {
        while (!not_one.load(std::memory_order_relaxed)) {
            const int64_t j = next.fetch_add(1);
            if (j >= n_chunks) break;
            const int64_t off = j * per, len = std::min(per, count - off);
            if (!all_ones_kind(static_cast<const unsigned char*>(values) + off * es, kind, len)) not_one.store(1);
        }
    };
    std::vector<std::thread> team;
    const int nt = static_cast<int>(std::min<int64_t>(threads, n_chunks));
    for (int t = 1; t < nt; ++t) team.emplace_back(body);
    body();
    for (auto& t : team) t.join();
    return not_one.load() == 0;
}

// values of any supported kind (host) -> f32 on the device
void stage_values(snapb200_ctx* c, const void* src, int kind, int64_t count, float* dst_dev) {
    if (count == 0) return;
    const int threads = host_threads(c);
    Ring r = ring_of(c, threads);
    const int64_t per = static_cast<int64_t>(kChunkBytes / 4);
    const int64_t n_chunks = ceil_div(count, per);
    cudaStream_t st = c->stream;
    kind_size(kind);
    run_ring(r, n_chunks, threads, [&](int64_t j, unsigned char* slot, cudaEvent_t ev) {
        const int64_t off = j * per, len = std::min(per, count - off);
        float* out = reinterpret_cast<float*>(slot);
        to_f32_kind(src, kind, off, len, out);
        SB_CUDA(cudaMemcpyAsync(dst_dev + off, out, sizeof(float) * len, cudaMemcpyHostToDevice, st));
        SB_CUDA(cudaEventRecord(ev, st));
    });
    SB_CUDA(cudaStreamSynchronize(st));
}

// device -> host.  A pinned destination takes one DMA; a pageable one (a numpy array) is filled
// through the pinned ring by the thread team, each chunk copied out as soon as its DMA lands.
// Everything queued on the context's stream before the call is waited for first.
void copy_to_host(snapb200_ctx* c, void* dst, const void* src_dev, size_t bytes) {
    if (bytes == 0) return;
    cudaStream_t st = c->stream;
    cudaPointerAttributes attr;
    const bool pinned = cudaPointerGetAttributes(&attr, dst) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (pinned || bytes < (4u << 20)) {
        SB_CUDA(cudaMemcpyAsync(dst, src_dev, bytes, cudaMemcpyDeviceToHost, st));
        SB_CUDA(cudaStreamSynchronize(st));
        return;
    }
    const int threads = host_threads(c);
    Ring r = ring_of(c, threads);
    const int64_t n_chunks = ceil_div(static_cast<int64_t>(bytes), static_cast<int64_t>(kChunkBytes));
    SB_CUDA(cudaStreamSynchronize(st));   // the producer kernels are done; the chunk copies below run concurrently
    run_ring(r, n_chunks, threads, [&](int64_t j, unsigned char* slot, cudaEvent_t ev) {
        const size_t off = static_cast<size_t>(j) * kChunkBytes, len = std::min(kChunkBytes, bytes - off);
        SB_CUDA(cudaMemcpyAsync(slot, static_cast<const unsigned char*>(src_dev) + off, len, cudaMemcpyDeviceToHost, st));
        SB_CUDA(cudaEventRecord(ev, st));
        SB_CUDA(cudaEventSynchronize(ev));
        memcpy(static_cast<unsigned char*>(dst) + off, slot, len);
    });
}

}  // namespace snapb

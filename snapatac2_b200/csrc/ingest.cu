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

#include <immintrin.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
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
                if (want >= 0) {
                    while (released[s].load(std::memory_order_acquire) != want) {
                        if (failed.load()) return;
                        std::this_thread::yield();
                    }
                    SB_CUDA(cudaEventSynchronize(r.ev[s]));
                }
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

}  // namespace

// indices (int32 or int64, pageable or pinned host memory) -> int32 on the device.
// Returns false if some 64-bit index does not fit 31 bits (the range check against m runs on the device).
bool stage_indices(snapb200_ctx* c, const void* src, int bits, int64_t count, int32_t* dst_dev) {
    if (count == 0) return true;
    const int threads = host_threads(c);
    Ring r = ring_of(c, threads);
    const int64_t per = static_cast<int64_t>(kChunkBytes / 4);
    const int64_t n_chunks = ceil_div(count, per);
    std::atomic<uint64_t> orall{0};
    cudaStream_t st = c->stream;
    run_ring(r, n_chunks, threads, [&](int64_t j, unsigned char* slot, cudaEvent_t ev) {
        const int64_t off = j * per, len = std::min(per, count - off);
        int32_t* out = reinterpret_cast<int32_t*>(slot);
        if (bits == 64) {
            uint64_t o = 0;
            narrow_i64(static_cast<const int64_t*>(src) + off, out, len, o);
            if (o >> 31) orall.fetch_or(o);
        } else {
            memcpy(out, static_cast<const int32_t*>(src) + off, sizeof(int32_t) * len);
        }
        SB_CUDA(cudaMemcpyAsync(dst_dev + off, out, sizeof(int32_t) * len, cudaMemcpyHostToDevice, st));
        SB_CUDA(cudaEventRecord(ev, st));
    });
    SB_CUDA(cudaStreamSynchronize(st));
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

// Caching device allocator behind DevBuf.
//
// cudaFree synchronises the whole device and cudaMalloc of multi-GB blocks is
// slow; a spectral() call allocates and releases dozens of temporaries (and
// bench.py repeats the call), so freed blocks are parked here and handed out
// again.  Reuse is stream ordered: a block freed while kernels still read it
// can be given to the next user on the same stream safely; every parked block
// carries an event recorded on its last stream, and a user on a different
// stream is ordered after that event (cudaStreamWaitEvent, no host stall).  On out-of-memory every parked
// block is returned to the driver and the allocation retried.
#include "common.cuh"

#include <atomic>
#include <chrono>
#include <map>
#include <mutex>
#include <unordered_map>

namespace snapb {

namespace {

struct Parked {
    void* p;
    cudaStream_t stream;   // stream the block was last used on
    cudaEvent_t ev;        // recorded on `stream` when the block was parked
};
struct Pool {
    std::mutex mu;
    std::unordered_map<void*, size_t> live;                        // ptr -> bytes
    std::map<int, std::multimap<size_t, Parked>> parked;           // device -> size -> block
};
Pool& pool() {
    static Pool* p = new Pool();   // intentionally leaked: outlives static destruction order issues
    return *p;
}
thread_local cudaStream_t t_stream = nullptr;
// diagnostics: driver allocations, trims and the host time spent in them
std::atomic<long long> g_mallocs{0}, g_trims{0}, g_reuses{0};
std::atomic<long long> g_ns{0};
struct ScopedNs {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    ~ScopedNs() {
        g_ns += std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
    }
};

size_t round_size(size_t b) {
    const size_t g = b < (1u << 20) ? 512 : (2u << 20);
    return (b + g - 1) / g * g;
}

void trim_locked(Pool& P, int dev) {
    auto it = P.parked.find(dev);
    if (it == P.parked.end()) return;
    for (auto& kv : it->second) {
        if (kv.second.ev) cudaEventDestroy(kv.second.ev);
        cudaFree(kv.second.p);
    }
    it->second.clear();
}

}  // namespace

void pool_set_stream(cudaStream_t s) { t_stream = s; }

void* pool_alloc(size_t bytes) {
    if (bytes == 0) return nullptr;
    bytes = round_size(bytes);
    int dev = 0;
    SB_CUDA(cudaGetDevice(&dev));
    Pool& P = pool();
    std::lock_guard<std::mutex> lock(P.mu);
    auto& bins = P.parked[dev];
    auto it = bins.lower_bound(bytes);
    if (it != bins.end() && it->first <= bytes + bytes / 4 + (1u << 20)) {
        Parked blk = it->second;
        const size_t sz = it->first;
        bins.erase(it);
        if (blk.ev) {
            // a block parked by another stream may still be read there: order the new user after it
            if (blk.stream != t_stream) {
                if (t_stream) cudaStreamWaitEvent(t_stream, blk.ev, 0);
                else cudaEventSynchronize(blk.ev);
            }
            cudaEventDestroy(blk.ev);
        }
        P.live[blk.p] = sz;
        ++g_reuses;
        return blk.p;
    }
    ScopedNs timer;
    ++g_mallocs;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        cudaDeviceSynchronize();
        trim_locked(P, dev);
        ++g_trims;
        e = cudaMalloc(&p, bytes);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        char buf[256];
        snprintf(buf, sizeof(buf), "device allocation of %.2f GB failed: %s", bytes / 1e9, cudaGetErrorString(e));
        throw Error(buf);
    }
    P.live[p] = bytes;
    return p;
}

void pool_free(void* p) {
    if (!p) return;
    Pool& P = pool();
    std::lock_guard<std::mutex> lock(P.mu);
    auto it = P.live.find(p);
    if (it == P.live.end()) {
        cudaFree(p);
        return;
    }
    const size_t sz = it->second;
    P.live.erase(it);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaEvent_t ev = nullptr;
    if (t_stream) {
        if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess) cudaEventRecord(ev, t_stream);
    } else {
        cudaDeviceSynchronize();   // no stream to order later users after: park the block fully quiesced
    }
    P.parked[dev].emplace(sz, Parked{p, t_stream, ev});
}

// A stream is about to be destroyed (its work has been synchronised by the caller): parked blocks
// must not refer to it any more.
void pool_forget_stream(cudaStream_t s) {
    Pool& P = pool();
    std::lock_guard<std::mutex> lock(P.mu);
    for (auto& dev : P.parked) {
        for (auto& kv : dev.second) {
            if (kv.second.stream != s) continue;
            if (kv.second.ev) {
                cudaEventSynchronize(kv.second.ev);
                cudaEventDestroy(kv.second.ev);
            }
            kv.second.ev = nullptr;
            kv.second.stream = nullptr;
        }
    }
    if (t_stream == s) t_stream = nullptr;
}

void pool_counters(long long* mallocs, long long* reuses, long long* trims, double* ms) {
    *mallocs = g_mallocs.load();
    *reuses = g_reuses.load();
    *trims = g_trims.load();
    *ms = static_cast<double>(g_ns.load()) * 1e-6;
}

// Return every parked block of the current device to the driver.
void pool_trim() {
    int dev = 0;
    cudaGetDevice(&dev);
    Pool& P = pool();
    std::lock_guard<std::mutex> lock(P.mu);
    cudaDeviceSynchronize();
    trim_locked(P, dev);
}

}  // namespace snapb

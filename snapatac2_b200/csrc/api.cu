// C ABI of libsnapb200 (see include/snapb200.h for the contract and the
// reference entry points each function stands in for).
#include "ctx.cuh"
#include "dense.cuh"

#include <string.h>
#include <algorithm>
#include <chrono>
#include <vector>

namespace snapb {

static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }

namespace {

template <typename F>
int guarded(F&& f) {
    try {
        f();
        return 0;
    } catch (const std::exception& e) {
        set_last_error(e.what());
        return 1;
    } catch (...) {
        set_last_error("unknown error");
        return 1;
    }
}

void bind(snapb200_ctx* c) {
    SB_CHECK(c != nullptr, "null context");
    SB_CUDA(cudaSetDevice(c->device));
    pool_set_stream(c->stream);
}

__global__ void narrow_i64_kernel(const int64_t* __restrict__ in, int32_t* __restrict__ out, int64_t n,
                                  int64_t limit, int* __restrict__ bad) {
    int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (; i < n; i += stride) {
        int64_t v = in[i];
        if (v < 0 || v >= limit) *bad = 1;
        out[i] = static_cast<int32_t>(v);
    }
}
__global__ void check_i32_kernel(const int32_t* __restrict__ in, int64_t n, int64_t limit, int* __restrict__ bad) {
    int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (; i < n; i += stride) {
        int32_t v = in[i];
        if (v < 0 || v >= limit) *bad = 1;
    }
}
template <typename T>
__global__ void to_f32_kernel(const T* __restrict__ in, float* __restrict__ out, int64_t n) {
    int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (; i < n; i += stride) out[i] = static_cast<float>(in[i]);
}
__global__ void widen_i32_kernel(const int32_t* __restrict__ in, int64_t* __restrict__ out, int64_t n) {
    int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}
__global__ void ones_check_kernel(const float* __restrict__ v, int64_t n, int* __restrict__ not_one) {
    int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (; i < n; i += stride)
        if (v[i] != 1.0f) *not_one = 1;
}

int stream_blocks(snapb200_ctx* c, int64_t n) {
    return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, 256), static_cast<int64_t>(c->num_sms) * 16)));
}

size_t value_size(int kind) {
    switch (kind) {
        case 1: return 4; case 2: return 8; case 3: return 4; case 4: return 4; case 5: return 8; case 6: return 8;
        case 7: case 8: return 1; case 9: case 10: return 2;
        default: throw Error("load_csr: unknown value_kind");
    }
}

// bad |= 2 if a row's column indices are not strictly increasing (the transpose and the tiled
// format build rely on sorted, duplicate-free rows); one warp per row, coalesced
__global__ void rows_sorted_kernel(const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx, int64_t nrows,
                                   int* __restrict__ bad) {
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    bool wrong = false;
    for (int64_t r = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; r < nrows; r += nwarps) {
        const int64_t s = ptr[r], e = ptr[r + 1];
        for (int64_t p = s + lane; p + 1 < e; p += 32) wrong = wrong || (idx[p] >= idx[p + 1]);
    }
    if (wrong) atomicOr(bad, 2);
}

// waits for a deferred value scan (if any); returns true if every scanned value was 1
bool join_value_scan(snapb200_ctx* c) {
    if (c->scan_pending) {
        if (c->scan_thread.joinable()) c->scan_thread.join();
        c->scan_pending = false;
    }
    return c->scan_not_one.load() == 0;
}

void load_csr(snapb200_ctx* c, int64_t n_local, int64_t n_global, int64_t row0, int64_t m, const void* indptr,
              int indptr_bits, const void* indices, int indices_bits, const void* values, int value_kind,
              int on_device) {
    SB_CHECK(n_local >= 0 && n_global >= n_local && row0 >= 0 && row0 + n_local <= n_global, "load_csr: bad shard geometry");
    SB_CHECK(m >= 1 && m < (1ll << 31), "load_csr: m must be in [1, 2^31)");
    SB_CHECK(indptr_bits == 32 || indptr_bits == 64, "load_csr: indptr_bits must be 32 or 64");
    SB_CHECK(indices_bits == 32 || indices_bits == 64, "load_csr: indices_bits must be 32 or 64");
    SB_CHECK(indptr != nullptr, "load_csr: null indptr");
    cudaStream_t st = c->stream;
    join_value_scan(c);
    c->scan_not_one.store(0);
    const auto wall0 = std::chrono::steady_clock::now();
    const bool dev = on_device != 0;
    // device-resident input was produced on the caller's stream(s), which this library's private
    // non-blocking stream does not order with: wait for the device before reading it
    if (dev) SB_CUDA(cudaDeviceSynchronize());
    SB_CUDA(cudaEventRecord(c->ev0, st));
    Csr& X = c->X;
    // the context holds no usable matrix until this load has passed validation
    c->loaded = false;
    c->prepared = false;
    c->views.clear();
    c->nnz_mode = -1;
    c->S1.clear();
    c->S2.clear();
    c->Xt.clear();
    c->xt_built = false;
    c->XtT.clear();
    c->proj_ready = false;
    X.nrows = n_local;
    X.ncols = m;
    X.ptr.alloc(n_local + 1);

    // ---- indptr (small): bring to int64 on the device, read nnz back
    const cudaMemcpyKind kind = dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (indptr_bits == 64) {
        SB_CUDA(cudaMemcpyAsync(X.ptr.p, indptr, sizeof(int64_t) * (n_local + 1), kind, st));
    } else {
        DevBuf<int32_t> tmp;
        tmp.alloc(n_local + 1);
        SB_CUDA(cudaMemcpyAsync(tmp.p, indptr, sizeof(int32_t) * (n_local + 1), kind, st));
        widen_i32_kernel<<<static_cast<unsigned>(ceil_div(n_local + 1, 256)), 256, 0, st>>>(tmp.p, X.ptr.p, n_local + 1);
        SB_LAUNCH_CHECK();
        count_launch(c);
        SB_CUDA(cudaStreamSynchronize(st));
    }
    int64_t ends[2] = {0, 0};
    SB_CUDA(cudaMemcpyAsync(&ends[0], X.ptr.p, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaMemcpyAsync(&ends[1], X.ptr.p + n_local, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    SB_CHECK(ends[0] == 0, "load_csr: indptr[0] must be 0 (shard-relative)");
    const int64_t nnz = ends[1];
    SB_CHECK(nnz >= 0, "load_csr: negative nnz");
    SB_CHECK(nnz == 0 || indices != nullptr, "load_csr: null indices");
    X.nnz = nnz;
    X.idx.alloc(std::max<int64_t>(1, nnz));
    c->stats.bytes_h2d_indices = 0;

    DevBuf<int> bad;
    bad.alloc(1);
    SB_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), st));
    // ---- column indices: int32 on the device, range-checked
    if (nnz > 0) {
        if (!dev) {
            // host arrays (pageable or pinned): narrowed / copied chunk by chunk through the pinned ring
            if (!stage_indices(c, indices, indices_bits, nnz, X.idx.p)) throw Error("load_csr: column index out of range");
        } else if (indices_bits == 32) {
            SB_CUDA(cudaMemcpyAsync(X.idx.p, indices, sizeof(int32_t) * nnz, cudaMemcpyDeviceToDevice, st));
        } else {
            narrow_i64_kernel<<<stream_blocks(c, nnz), 256, 0, st>>>(static_cast<const int64_t*>(indices), X.idx.p, nnz, m, bad.p);
            SB_LAUNCH_CHECK();
            count_launch(c);
        }
        check_i32_kernel<<<stream_blocks(c, nnz), 256, 0, st>>>(X.idx.p, nnz, m, bad.p);
        SB_LAUNCH_CHECK();
        count_launch(c);
    }
    // ---- values (optional): f32 on the device; an all-ones array is dropped (binarised input runs
    //      the pattern-only kernels).  Host values are scanned on the host first, so a binarised
    //      matrix never ships them.
    X.val.release();
    if (values != nullptr && nnz > 0) {
        value_size(value_kind);
        if (!dev && c->defer_value_scan) {
            // optimistic: load the pattern only and let a background team scan the values while the GPU
            // already works on the matrix; the caller asks for the verdict before it trusts the result
            // (snapb200_values_verdict) and ships the values then if some were not 1
            c->scan_pending = true;
            c->scan_thread = std::thread([c, values, value_kind, nnz] {
                if (!host_values_all_ones(c, values, value_kind, nnz)) c->scan_not_one.store(1);
            });
        } else if (!dev) {
            if (!host_values_all_ones(c, values, value_kind, nnz)) {
                X.val.alloc(nnz);
                stage_values(c, values, value_kind, nnz, X.val.p);
            }
        } else {
            X.val.alloc(nnz);
            const int g = stream_blocks(c, nnz);
            float* out = X.val.p;
            const void* p = values;
            switch (value_kind) {
                case 1: SB_CUDA(cudaMemcpyAsync(out, p, sizeof(float) * nnz, cudaMemcpyDeviceToDevice, st)); break;
                case 2: to_f32_kernel<double><<<g, 256, 0, st>>>(static_cast<const double*>(p), out, nnz); break;
                case 3: to_f32_kernel<uint32_t><<<g, 256, 0, st>>>(static_cast<const uint32_t*>(p), out, nnz); break;
                case 4: to_f32_kernel<int32_t><<<g, 256, 0, st>>>(static_cast<const int32_t*>(p), out, nnz); break;
                case 5: to_f32_kernel<int64_t><<<g, 256, 0, st>>>(static_cast<const int64_t*>(p), out, nnz); break;
                case 6: to_f32_kernel<uint64_t><<<g, 256, 0, st>>>(static_cast<const uint64_t*>(p), out, nnz); break;
                case 7: to_f32_kernel<uint8_t><<<g, 256, 0, st>>>(static_cast<const uint8_t*>(p), out, nnz); break;
                case 8: to_f32_kernel<int8_t><<<g, 256, 0, st>>>(static_cast<const int8_t*>(p), out, nnz); break;
                case 9: to_f32_kernel<uint16_t><<<g, 256, 0, st>>>(static_cast<const uint16_t*>(p), out, nnz); break;
                case 10: to_f32_kernel<int16_t><<<g, 256, 0, st>>>(static_cast<const int16_t*>(p), out, nnz); break;
            }
            SB_LAUNCH_CHECK();
            count_launch(c);
            DevBuf<int> not_one;
            not_one.alloc(1);
            SB_CUDA(cudaMemsetAsync(not_one.p, 0, sizeof(int), st));
            ones_check_kernel<<<stream_blocks(c, nnz), 256, 0, st>>>(X.val.p, nnz, not_one.p);
            SB_LAUNCH_CHECK();
            count_launch(c);
            int h = 0;
            SB_CUDA(cudaMemcpyAsync(&h, not_one.p, sizeof(int), cudaMemcpyDeviceToHost, st));
            SB_CUDA(cudaStreamSynchronize(st));
            if (!h) X.val.release();
        }
    }
    if (nnz > 0 && n_local > 0) {
        const int blocks = static_cast<int>(std::min<int64_t>(ceil_div(n_local, 8), static_cast<int64_t>(c->num_sms) * 16));
        rows_sorted_kernel<<<blocks, 256, 0, st>>>(X.ptr.p, X.idx.p, n_local, bad.p);
        SB_LAUNCH_CHECK();
        count_launch(c);
    }
    int hbad = 0;
    SB_CUDA(cudaMemcpyAsync(&hbad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaEventRecord(c->ev1, st));
    SB_CUDA(cudaStreamSynchronize(st));
    SB_CHECK(!(hbad & 1), "load_csr: column index out of range");
    SB_CHECK(!(hbad & 2), "load_csr: column indices must be strictly increasing within every row (sorted, no duplicates)");
    c->stats.ms_load = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count();
    c->stats.bytes_h2d = dev ? 0 : static_cast<int64_t>((indptr_bits / 8) * (n_local + 1) + (X.has_values() ? 4 * nnz : 0)) + c->stats.bytes_h2d_indices;
    c->stats.host_threads = dev ? 0 : host_threads(c);
    c->n_local = n_local;
    c->n_global = n_global;
    c->row0 = row0;
    c->m = m;
    c->loaded = true;
    c->prepared = false;
    c->stats.nnz_local = nnz;
}

// ---- block-wise load: the matrix arrives as a sequence of CSR row blocks (the chunks of a backed
//      AnnData, reference: `chunked_X` / the iteration at embedding.rs:76-84) and is assembled on the device
template <typename T>
void grow_keep(snapb200_ctx* c, DevBuf<T>& buf, int64_t used, int64_t need) {
    if (need <= buf.n) return;
    DevBuf<T> bigger;
    bigger.alloc(std::max<int64_t>(need, buf.n + buf.n / 2 + 1024));
    if (used > 0) SB_CUDA(cudaMemcpyAsync(bigger.p, buf.p, sizeof(T) * used, cudaMemcpyDeviceToDevice, c->stream));
    buf.swap(bigger);
}

void load_begin(snapb200_ctx* c, int64_t m, int64_t rows_hint, int64_t nnz_hint) {
    SB_CHECK(m >= 1 && m < (1ll << 31), "load_begin: m must be in [1, 2^31)");
    join_value_scan(c);      // a background scan of the previous matrix's values must not outlive it
    Csr& X = c->X;
    c->loaded = false;
    c->prepared = false;
    c->proj_ready = false;
    c->views.clear();
    c->nnz_mode = -1;
    c->S1.clear(); c->S2.clear(); c->Xt.clear(); c->xt_built = false; c->XtT.clear();
    X.clear();
    X.ncols = m;
    X.ptr.alloc(std::max<int64_t>(rows_hint, 1024) + 1);
    X.idx.alloc(std::max<int64_t>(nnz_hint, 1 << 20));
    SB_CUDA(cudaMemsetAsync(X.ptr.p, 0, sizeof(int64_t), c->stream));
    c->m = m;
    c->appending = true;
    c->app_rows = c->app_nnz = 0;
    c->stats.bytes_h2d = 0;
    c->stats.bytes_h2d_indices = 0;
}

void load_append(snapb200_ctx* c, int64_t n_rows, const void* indptr, int indptr_bits, const void* indices,
                 int indices_bits, const void* values, int value_kind) {
    SB_CHECK(c->appending, "load_append: call load_begin first");
    SB_CHECK(n_rows >= 0 && indptr != nullptr, "load_append: bad block");
    SB_CHECK(indptr_bits == 32 || indptr_bits == 64, "load_append: indptr_bits must be 32 or 64");
    SB_CHECK(indices_bits == 32 || indices_bits == 64, "load_append: indices_bits must be 32 or 64");
    Csr& X = c->X;
    cudaStream_t st = c->stream;
    const int64_t idx_bytes0 = c->stats.bytes_h2d_indices;
    auto at = [&](int64_t i) -> int64_t {
        return indptr_bits == 64 ? static_cast<const int64_t*>(indptr)[i] : static_cast<const int32_t*>(indptr)[i];
    };
    const int64_t first = at(0), nnz = at(n_rows) - first;
    SB_CHECK(nnz >= 0, "load_append: indptr must be non-decreasing");
    // row pointers of the block, shifted behind what is already there
    std::vector<int64_t> hp(static_cast<size_t>(n_rows));
    for (int64_t i = 0; i < n_rows; ++i) hp[i] = c->app_nnz + (at(i + 1) - first);
    grow_keep(c, X.ptr, c->app_rows + 1, c->app_rows + n_rows + 1);
    if (n_rows > 0)
        SB_CUDA(cudaMemcpyAsync(X.ptr.p + c->app_rows + 1, hp.data(), sizeof(int64_t) * n_rows, cudaMemcpyHostToDevice, st));
    grow_keep(c, X.idx, c->app_nnz, c->app_nnz + nnz);
    if (nnz > 0) {
        const unsigned char* src = static_cast<const unsigned char*>(indices) + static_cast<size_t>(first) * (indices_bits / 8);
        if (!stage_indices(c, src, indices_bits, nnz, X.idx.p + c->app_nnz)) throw Error("load_append: column index out of range");
        if (values != nullptr) {
            const size_t es = value_size(value_kind);
            const unsigned char* vsrc = static_cast<const unsigned char*>(values) + static_cast<size_t>(first) * es;
            const bool ones = host_values_all_ones(c, vsrc, value_kind, nnz);
            if (!ones && !X.has_values()) {   // first block with real values: earlier entries were all ones
                X.val.alloc(X.idx.n);
                if (c->app_nnz > 0) fill_f32(c, X.val.p, 1.f, c->app_nnz);
            }
            if (X.has_values()) {
                grow_keep(c, X.val, c->app_nnz, X.idx.n);
                stage_values(c, vsrc, value_kind, nnz, X.val.p + c->app_nnz);
            }
        } else if (X.has_values()) {
            grow_keep(c, X.val, c->app_nnz, X.idx.n);
            fill_f32(c, X.val.p + c->app_nnz, 1.f, nnz);
        }
    }
    SB_CUDA(cudaStreamSynchronize(st));   // hp is a host temporary
    c->app_rows += n_rows;
    c->app_nnz += nnz;
    c->stats.bytes_h2d += 8 * n_rows + (X.has_values() ? 4 * nnz : 0) + (c->stats.bytes_h2d_indices - idx_bytes0);
}

void load_end(snapb200_ctx* c, int64_t n_global, int64_t row0) {
    SB_CHECK(c->appending, "load_end: call load_begin first");
    Csr& X = c->X;
    cudaStream_t st = c->stream;
    const int64_t n_local = c->app_rows, nnz = c->app_nnz;
    if (n_global < 0) n_global = n_local;
    SB_CHECK(row0 >= 0 && row0 + n_local <= n_global, "load_end: bad shard geometry");
    c->appending = false;
    X.nrows = n_local;
    X.nnz = nnz;
    DevBuf<int> bad;
    bad.alloc(1);
    SB_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), st));
    if (nnz > 0) {
        check_i32_kernel<<<stream_blocks(c, nnz), 256, 0, st>>>(X.idx.p, nnz, c->m, bad.p);
        SB_LAUNCH_CHECK();
        const int blocks = static_cast<int>(std::min<int64_t>(ceil_div(n_local, 8), static_cast<int64_t>(c->num_sms) * 16));
        rows_sorted_kernel<<<blocks, 256, 0, st>>>(X.ptr.p, X.idx.p, n_local, bad.p);
        SB_LAUNCH_CHECK();
        count_launch(c, 2);
    }
    int hbad = 0;
    SB_CUDA(cudaMemcpyAsync(&hbad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    SB_CHECK(!(hbad & 1), "load_end: column index out of range");
    SB_CHECK(!(hbad & 2), "load_end: column indices must be strictly increasing within every row (sorted, no duplicates)");
    c->n_local = n_local;
    c->n_global = n_global;
    c->row0 = row0;
    c->loaded = true;
    c->stats.nnz_local = nnz;
    c->stats.host_threads = host_threads(c);
}

}  // namespace
}  // namespace snapb

using namespace snapb;

extern "C" {

const char* snapb200_last_error(void) { return g_last_error.c_str(); }
int snapb200_version(void) { return 100; }

int snapb200_create(int device, snapb200_ctx** out) {
    return guarded([&] {
        SB_CHECK(out != nullptr, "create: null out pointer");
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count == 0)
            throw Error("no CUDA device available: libsnapb200 has no CPU fallback (an NVIDIA B200 is required)");
        SB_CHECK(device >= 0 && device < count, "create: device ordinal out of range");
        SB_CUDA(cudaSetDevice(device));
        cudaDeviceProp prop;
        SB_CUDA(cudaGetDeviceProperties(&prop, device));
        if (prop.major != 10)
            throw Error(std::string("libsnapb200 is built for sm_100a only; device is ") + prop.name);
        auto* c = new snapb200_ctx();
        c->device = device;
        c->num_sms = prop.multiProcessorCount;
        SB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        SB_CUDA(cudaEventCreate(&c->ev0));
        SB_CUDA(cudaEventCreate(&c->ev1));
        *out = c;
    });
}

int snapb200_destroy(snapb200_ctx* c) {
    return guarded([&] {
        if (!c) return;
        cudaSetDevice(c->device);
        join_value_scan(c);
        cudaStreamSynchronize(c->stream);
        if (c->borrowed) {   // a view context: stream and communicator belong to its main context
            c->comm = nullptr;
            cudaStream_t shared = c->stream;
            if (c->ev0) cudaEventDestroy(c->ev0);
            if (c->ev1) cudaEventDestroy(c->ev1);
            for (auto& e : c->ring_events) cudaEventDestroy(e);
            c->ring_events.clear();
            pool_set_stream(shared);   // its buffers return to the pool ordered on the shared stream
            delete c;
            return;
        }
        comm_destroy(c);
        if (c->ev0) cudaEventDestroy(c->ev0);
        if (c->ev1) cudaEventDestroy(c->ev1);
        for (auto& e : c->ring_events) cudaEventDestroy(e);
        c->ring_events.clear();
        cudaStream_t st = c->stream;
        // everything on the stream is done: the context's buffers go back to the pool as quiesced
        // blocks (no event), and blocks parked earlier must stop referring to the stream
        pool_set_stream(nullptr);
        delete c;
        pool_forget_stream(st);
        if (st) cudaStreamDestroy(st);
    });
}

int snapb200_comm_unique_id(char id[128]) {
    return guarded([&] { comm_unique_id(id); });
}
int snapb200_comm_init(snapb200_ctx* c, int rank, int nranks, const char id[128]) {
    return guarded([&] { bind(c); comm_init(c, rank, nranks, id); });
}

int snapb200_load_csr(snapb200_ctx* c, int64_t n_local, int64_t n_global, int64_t row0, int64_t m, const void* indptr,
                      int indptr_bits, const void* indices, int indices_bits, const void* values, int value_kind,
                      int on_device) {
    return guarded([&] {
        bind(c);
        load_csr(c, n_local, n_global, row0, m, indptr, indptr_bits, indices, indices_bits, values, value_kind, on_device);
    });
}

int snapb200_load_begin(snapb200_ctx* c, int64_t m, int64_t rows_hint, int64_t nnz_hint) {
    return guarded([&] { bind(c); load_begin(c, m, rows_hint, nnz_hint); });
}
int snapb200_load_append(snapb200_ctx* c, int64_t n_rows, const void* indptr, int indptr_bits, const void* indices,
                         int indices_bits, const void* values, int value_kind) {
    return guarded([&] { bind(c); load_append(c, n_rows, indptr, indptr_bits, indices, indices_bits, values, value_kind); });
}
int snapb200_load_end(snapb200_ctx* c, int64_t n_global, int64_t row0) {
    return guarded([&] { bind(c); load_end(c, n_global, row0); });
}

int snapb200_set_defer_value_scan(snapb200_ctx* c, int on) {
    return guarded([&] {
        SB_CHECK(c != nullptr, "null context");
        c->defer_value_scan = on != 0;
    });
}

int snapb200_values_verdict(snapb200_ctx* c, int* all_ones) {
    return guarded([&] {
        SB_CHECK(c != nullptr && all_ones != nullptr, "values_verdict: null argument");
        *all_ones = join_value_scan(c) ? 1 : 0;
    });
}

int snapb200_load_values(snapb200_ctx* c, const void* values, int value_kind) {
    return guarded([&] {
        bind(c);
        SB_CHECK(c->loaded, "load_values: no matrix loaded");
        SB_CHECK(values != nullptr, "load_values: null values");
        join_value_scan(c);
        value_size(value_kind);
        Csr& X = c->X;
        if (X.nnz > 0) {
            X.val.alloc(X.nnz);
            stage_values(c, values, value_kind, X.nnz, X.val.p);
        }
        c->scan_not_one.store(0);
        c->prepared = false;
        c->proj_ready = false;
        c->views.clear();
        c->S1.clear(); c->S2.clear(); c->Xt.clear(); c->xt_built = false; c->XtT.clear();
        c->stats.bytes_h2d += 4 * X.nnz;
    });
}

int snapb200_set_geometry(snapb200_ctx* c, int64_t n_global, int64_t row0) {
    return guarded([&] {
        SB_CHECK(c != nullptr && c->loaded, "set_geometry: no matrix loaded");
        SB_CHECK(row0 >= 0 && row0 + c->n_local <= n_global, "set_geometry: bad shard geometry");
        c->n_global = n_global;
        c->row0 = row0;
        c->prepared = false;
        c->proj_ready = false;
        c->views.clear();
    });
}

int snapb200_select_features(snapb200_ctx* c, const uint8_t* keep, int64_t m) {
    return guarded([&] { bind(c); select_features(c, keep, m); });
}

int snapb200_generate(snapb200_ctx* c, int64_t n_local, int64_t n_global, int64_t row0, int64_t m, int nnz_row,
                      int n_clusters, uint64_t seed, const uint64_t* feat_cdf, const uint64_t* cluster_cdf,
                      const int64_t* block_start, const uint64_t* alpha) {
    return guarded([&] {
        bind(c);
        c->Xt.clear();
        c->xt_built = false;
        c->XtT.clear();
        c->proj_ready = false;
        generate_rows(c, n_local, n_global, row0, m, nnz_row, n_clusters, seed, feat_cdf, cluster_cdf, block_start, alpha);
    });
}

int snapb200_shape(snapb200_ctx* c, int64_t* n_local, int64_t* m, int64_t* nnz_local) {
    return guarded([&] {
        SB_CHECK(c && c->loaded, "shape: no matrix loaded");
        if (n_local) *n_local = c->n_local;
        if (m) *m = c->m;
        if (nnz_local) *nnz_local = c->X.nnz;
    });
}

int snapb200_export_csr(snapb200_ctx* c, int64_t* indptr, int32_t* indices, float* values) {
    return guarded([&] {
        bind(c);
        SB_CHECK(c->loaded, "export_csr: no matrix loaded");
        const Csr& X = c->X;
        if (indptr) SB_CUDA(cudaMemcpyAsync(indptr, X.ptr.p, sizeof(int64_t) * (X.nrows + 1), cudaMemcpyDeviceToHost, c->stream));
        if (indices && X.nnz > 0)
            SB_CUDA(cudaMemcpyAsync(indices, X.idx.p, sizeof(int32_t) * X.nnz, cudaMemcpyDeviceToHost, c->stream));
        if (values && X.nnz > 0) {
            if (X.has_values()) {
                SB_CUDA(cudaMemcpyAsync(values, X.val.p, sizeof(float) * X.nnz, cudaMemcpyDeviceToHost, c->stream));
            } else {
                SB_CUDA(cudaStreamSynchronize(c->stream));
                std::fill(values, values + X.nnz, 1.0f);
            }
        }
        SB_CUDA(cudaStreamSynchronize(c->stream));
    });
}

int snapb200_set_feature_weights(snapb200_ctx* c, const double* w, int64_t m) {
    return guarded([&] {
        SB_CHECK(c != nullptr, "null context");
        if (w == nullptr) c->user_weights.clear();
        else c->user_weights.assign(w, w + m);
        c->prepared = false;
        c->proj_ready = false;
    });
}

int snapb200_prepare(snapb200_ctx* c, double* idf_out, double* degree_out) {
    return guarded([&] { bind(c); prepare(c, idf_out, degree_out); });
}

int snapb200_attach_view(snapb200_ctx* main_ctx, snapb200_ctx* view) {
    return guarded([&] {
        bind(main_ctx);
        SB_CHECK(view != nullptr && view != main_ctx, "attach_view: need a second context");
        SB_CHECK(view->device == main_ctx->device, "attach_view: contexts must live on the same device");
        if (view->stream == main_ctx->stream) return;
        SB_CHECK(!view->borrowed, "attach_view: the context is already attached to another one");
        SB_CUDA(cudaStreamSynchronize(view->stream));
        comm_destroy(view);
        cudaStream_t old = view->stream;
        pool_forget_stream(old);
        if (old) cudaStreamDestroy(old);
        view->stream = main_ctx->stream;
        view->comm = main_ctx->comm;
        view->rank = main_ctx->rank;
        view->nranks = main_ctx->nranks;
        view->borrowed = true;
        main_ctx->views.clear();
    });
}

int snapb200_view_frobenius(snapb200_ctx* c, const int64_t* sample_rows, int64_t n_sample_local, double* snippet_sum) {
    return guarded([&] {
        bind(c);
        SB_CHECK(snippet_sum != nullptr, "view_frobenius: null output");
        *snippet_sum = view_frobenius(c, sample_rows, n_sample_local);
    });
}

int snapb200_combine_views(snapb200_ctx* main_ctx, snapb200_ctx** views, const double* view_scale, int n_views,
                           double* degree_out) {
    return guarded([&] { bind(main_ctx); combine_views(main_ctx, views, view_scale, n_views, degree_out); });
}

int snapb200_gather_rows(snapb200_ctx* src, const int64_t* rows, int64_t n_rows, snapb200_ctx* dst, int64_t n_global_dst,
                         int64_t row0_dst) {
    return guarded([&] {
        SB_CHECK(src != nullptr && dst != nullptr, "gather_rows: null context");
        bind(dst);
        gather_rows(src, rows, n_rows, dst, n_global_dst, row0_dst);
    });
}

int snapb200_get_vector(snapb200_ctx* c, int which, double* out) {
    return guarded([&] {
        bind(c);
        SB_CHECK(c->prepared || (which <= 1 && c->proj_ready), "get_vector: call prepare first");
        const double* src = nullptr;
        int64_t len = 0;
        switch (which) {
            case 0: src = c->w.p; len = c->m; break;
            case 1: src = c->rho.p; len = c->n_local; break;
            case 2: src = c->degree.p; len = c->n_local; break;
            case 3: src = c->csum.p; len = c->m; break;
            default: throw Error("get_vector: which must be 0 (weights), 1 (row norms), 2 (degrees) or 3 (column sums)");
        }
        if (len > 0) SB_CUDA(cudaMemcpyAsync(out, src, sizeof(double) * len, cudaMemcpyDeviceToHost, c->stream));
        SB_CUDA(cudaStreamSynchronize(c->stream));
    });
}

int snapb200_view_norms(snapb200_ctx* c, double* idf_out, double* rho_out) {
    return guarded([&] { bind(c); view_norms(c, idf_out, rho_out); });
}

int snapb200_prepare_projection(snapb200_ctx* c, double* idf_out, double* rho_out) {
    return guarded([&] {
        bind(c);
        prepare_projection(c);
        if (idf_out) SB_CUDA(cudaMemcpyAsync(idf_out, c->w.p, sizeof(double) * c->m, cudaMemcpyDeviceToHost, c->stream));
        if (rho_out && c->n_local > 0)
            SB_CUDA(cudaMemcpyAsync(rho_out, c->rho.p, sizeof(double) * c->n_local, cudaMemcpyDeviceToHost, c->stream));
        SB_CUDA(cudaStreamSynchronize(c->stream));
    });
}

int snapb200_project(snapb200_ctx* c, int transposed, const float* in, int k, float* out) {
    return guarded([&] {
        bind(c);
        SB_CHECK(in != nullptr && out != nullptr, "project: null buffer");
        project(c, transposed != 0, in, k, out);
    });
}

int snapb200_operator_apply(snapb200_ctx* c, const float* V, float* Y, int b) {
    return guarded([&] {
        bind(c);
        SB_CHECK(c->prepared, "operator_apply: call prepare first");
        const int64_t len = std::max<int64_t>(1, c->n_local * b);
        c->opV.ensure(len);
        c->opY.ensure(len);
        SB_CUDA(cudaMemcpyAsync(c->opV.p, V, sizeof(float) * c->n_local * b, cudaMemcpyHostToDevice, c->stream));
        operator_apply_dev(c, c->opV.p, b, c->opY.p, b, b);
        SB_CUDA(cudaMemcpyAsync(Y, c->opY.p, sizeof(float) * c->n_local * b, cudaMemcpyDeviceToHost, c->stream));
        SB_CUDA(cudaStreamSynchronize(c->stream));
    });
}

int snapb200_operator_time(snapb200_ctx* c, int b, int iters, int flush, double* ms_pass1, double* ms_comm,
                           double* ms_pass2) {
    return guarded([&] {
        bind(c);
        SB_CHECK(c->prepared, "operator_time: call prepare first");
        SB_CHECK(iters >= 1, "operator_time: iters must be >= 1");
        const int64_t len = std::max<int64_t>(1, c->n_local * b);
        c->opV.ensure(len);
        c->opY.ensure(len);
        DenseOps<8> ops;
        for (int j = 0; j < b; j += 4) {
            // fill opV with hash noise, 4 columns at a time through the B=4 generator
            DenseOps<4> o4;
            o4.random_block(c, c->opV.p + j, b, c->n_local, 99, 1000 + j);
        }
        cudaEvent_t evs[4];
        for (auto& e : evs) SB_CUDA(cudaEventCreate(&e));
        double p1 = 0, cm = 0, p2 = 0;
        for (int it = 0; it < iters + 1; ++it) {   // iteration 0 is a warm-up
            if (flush) flush_l2(c);
            operator_apply_dev(c, c->opV.p, b, c->opY.p, b, b, evs);
            SB_CUDA(cudaEventSynchronize(evs[3]));
            if (it == 0) continue;
            float a = 0.f, bb = 0.f, cc = 0.f;
            SB_CUDA(cudaEventElapsedTime(&a, evs[0], evs[1]));
            SB_CUDA(cudaEventElapsedTime(&bb, evs[1], evs[2]));
            SB_CUDA(cudaEventElapsedTime(&cc, evs[2], evs[3]));
            p1 += a; cm += bb; p2 += cc;
        }
        for (auto& e : evs) cudaEventDestroy(e);
        if (ms_pass1) *ms_pass1 = p1 / iters;
        if (ms_comm) *ms_comm = cm / iters;
        if (ms_pass2) *ms_pass2 = p2 / iters;
    });
}

int snapb200_eigsh(snapb200_ctx* c, int k, int64_t seed, double tol, int block, int max_basis, int max_ops,
                   double* evals, double* evecs, int scale_by_sqrt_eval) {
    return guarded([&] {
        bind(c);
        eigsh(c, k, seed, tol, block > 0 ? block : c->block, max_basis, max_ops, evals, evecs, scale_by_sqrt_eval != 0);
    });
}

int snapb200_get_stats(snapb200_ctx* c, snapb200_stats* out) {
    return guarded([&] {
        SB_CHECK(c && out, "get_stats: null argument");
        long long mallocs = 0, reuses = 0, trims = 0;
        double ms = 0.0;
        pool_counters(&mallocs, &reuses, &trims, &ms);
        c->stats.ms_pool = ms;
        c->stats.pool_mallocs = mallocs;
        *out = c->stats;
    });
}

int snapb200_set_spmm_mode(snapb200_ctx* c, int mode) {
    return guarded([&] {
        SB_CHECK(c != nullptr, "null context");
        SB_CHECK(mode >= 0 && mode <= 2, "set_spmm_mode: mode must be 0 (auto), 1 (csr) or 2 (tiled)");
        c->spmm_mode = mode;
        c->proj_ready = false;
        c->prepared = false;
    });
}

int snapb200_set_block(snapb200_ctx* c, int block) {
    return guarded([&] {
        SB_CHECK(c != nullptr, "null context");
        SB_CHECK(block == 4 || block == 8 || block == 16, "set_block: block width must be 4, 8 or 16");
        c->block = block;
    });
}

int snapb200_get_stream(snapb200_ctx* c, void** stream) {
    return guarded([&] {
        SB_CHECK(c && stream, "get_stream: null argument");
        *stream = reinterpret_cast<void*>(c->stream);
    });
}

// ---- test hooks (not part of the reference-facing surface) -----------------
int snapb200_knn(snapb200_ctx* c, int64_t n, int d, const double* points, int on_device, int64_t q0, int64_t nq, int k,
                 int32_t* out_indices, double* out_distances) {
    return guarded([&] {
        SB_CHECK(c != nullptr, "knn: null context");
        bind(c);
        knn(c, n, d, points, on_device, q0, nq, k, out_indices, out_distances);
    });
}

int snapb200_knn_limits(int* max_neighbors, int* max_dim) {
    if (max_neighbors) *max_neighbors = knn_max_neighbors();
    if (max_dim) *max_dim = knn_max_dim();
    return 0;
}

int snapb200_delta_selftest_host(const void* indices, int index_bits, int64_t count, int64_t* n_side) {
    int verdict = -2;
    const int rc = guarded([&] {
        SB_CHECK(index_bits == 32 || index_bits == 64, "index_bits must be 32 or 64");
        verdict = delta_selftest_host(indices, index_bits, count, n_side);
    });
    return rc != 0 ? -2 : verdict;
}

int snapb200_ortho_selftest(snapb200_ctx* c, int64_t n, int ncols, int block, double* max_err) {
    return guarded([&] { bind(c); *max_err = ortho_selftest(c, n, ncols, block); });
}

int snapb200_dense_selftest(snapb200_ctx* c, int64_t n, int ncq, int p, double* max_rel_err) {
    return guarded([&] { bind(c); *max_rel_err = dense_selftest(c, n, ncq, p); });
}
int snapb200_sym_eig(int n, double* a, double* w) {
    return guarded([&] { sym_eig(n, a, w); });
}

}  // extern "C"

// Host memory probe for the ingest path (csrc/ingest.cu): how fast can T threads stream an int64 array
// (a) read only, (b) narrowed to int32 with non-temporal stores, (c) (b) with software prefetch,
// (d) to 16-bit differences (bare loop), (e) through the library's encoder (csrc/delta_encode.h), written into
// a ring of 8 MB slots per thread as the staging team does.
// Build: g++ -O3 -std=c++17 -mavx2 -pthread scripts/host_bw_probe.cpp -o scripts/_build/host_bw_probe
#include "../snapatac2_b200/csrc/delta_encode.h"

#include <immintrin.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>

#include <atomic>
#include <chrono>
#include <thread>
#include <vector>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

template <typename F>
static double run(int threads, int64_t n, int64_t chunk, F f) {
    std::atomic<int64_t> next{0};
    const int64_t n_chunks = (n + chunk - 1) / chunk;
    const double t0 = now();
    std::vector<std::thread> team;
    for (int t = 0; t < threads; ++t)
        team.emplace_back([&, t] {
            for (;;) {
                const int64_t j = next.fetch_add(1);
                if (j >= n_chunks) break;
                f(t, j * chunk, std::min(chunk, n - j * chunk));
            }
        });
    for (auto& th : team) th.join();
    return now() - t0;
}

int main(int argc, char** argv) {
    const int64_t n = (argc > 1 ? atoll(argv[1]) : 1000) * 1000000LL;
    const int huge = argc > 2 ? atoi(argv[2]) : 0;
    int64_t* src = static_cast<int64_t*>(mmap(nullptr, n * 8, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0));
    if (huge) madvise(src, n * 8, MADV_HUGEPAGE);
    const int64_t chunk = 2 << 20;
    run(16, n, chunk, [&](int, int64_t off, int64_t len) {
        int64_t v = off * 97 % 500000;
        for (int64_t i = 0; i < len; ++i) { v += 1 + (i * 2654435761u >> 25) % 200; if (v >= 500000) v -= 500000; src[off + i] = v; }
    });
    printf("n=%lld entries (%.1f GB int64), hugepage advice=%d\n", (long long)n, n * 8e-9, huge);
    for (int threads : {8, 16}) {
        std::vector<int32_t*> outs(threads);
        for (auto& o : outs) o = static_cast<int32_t*>(aligned_alloc(64, chunk * 4 * 2));   // two slots per thread
        std::vector<uint64_t> sink(threads * 8);
        double t = run(threads, n, chunk, [&](int tid, int64_t off, int64_t len) {
            __m256i acc = _mm256_setzero_si256();
            for (int64_t i = 0; i + 4 <= len; i += 4) acc = _mm256_or_si256(acc, _mm256_loadu_si256((const __m256i*)(src + off + i)));
            alignas(32) uint64_t tmp[4]; _mm256_store_si256((__m256i*)tmp, acc); sink[tid * 8] |= tmp[0] | tmp[1] | tmp[2] | tmp[3];
        });
        printf("threads=%2d read-only        %.3f s  %.1f GB/s read\n", threads, t, n * 8e-9 / t);
        for (int pf : {0, 1}) {
            t = run(threads, n, chunk, [&](int tid, int64_t off, int64_t len) {
                const __m256i pick = _mm256_setr_epi32(0, 2, 4, 6, 0, 2, 4, 6);
                int32_t* out = outs[tid];
                const int64_t* in = src + off;
                for (int64_t i = 0; i + 8 <= len; i += 8) {
                    if (pf) _mm_prefetch((const char*)(in + i + 512), _MM_HINT_NTA);
                    const __m256i a = _mm256_loadu_si256((const __m256i*)(in + i));
                    const __m256i b = _mm256_loadu_si256((const __m256i*)(in + i + 4));
                    const __m256i lo = _mm256_permutevar8x32_epi32(a, pick), hi = _mm256_permutevar8x32_epi32(b, pick);
                    _mm256_stream_si256((__m256i*)(out + i), _mm256_permute2x128_si256(lo, hi, 0x20));
                }
                _mm_sfence();
            });
            printf("threads=%2d narrow NT pf=%d   %.3f s  %.1f GB/s read\n", threads, pf, t, n * 8e-9 / t);
        }
        t = run(threads, n, chunk, [&](int tid, int64_t off, int64_t len) {
            uint16_t* out = reinterpret_cast<uint16_t*>(outs[tid]);
            const int64_t* in = src + off;
            const __m256i pick = _mm256_setr_epi32(0, 2, 4, 6, 0, 2, 4, 6);
            for (int64_t i = 1; i + 8 <= len; i += 8) {
                const __m256i d0 = _mm256_sub_epi64(_mm256_loadu_si256((const __m256i*)(in + i)), _mm256_loadu_si256((const __m256i*)(in + i - 1)));
                const __m256i d1 = _mm256_sub_epi64(_mm256_loadu_si256((const __m256i*)(in + i + 4)), _mm256_loadu_si256((const __m256i*)(in + i + 3)));
                const __m256i v32 = _mm256_permute2x128_si256(_mm256_permutevar8x32_epi32(d0, pick), _mm256_permutevar8x32_epi32(d1, pick), 0x20);
                const __m256i v16 = _mm256_permute4x64_epi64(_mm256_packus_epi32(v32, v32), 0x08);
                _mm_storeu_si128((__m128i*)(out + i), _mm256_castsi256_si128(v16));
            }
        });
        printf("threads=%2d delta16 (cached) %.3f s  %.1f GB/s read\n", threads, t, n * 8e-9 / t);
        {
            const int64_t per = snapb::kDeltaPer;
            std::vector<unsigned char*> slots(threads * 2);
            for (auto& o : slots) { o = static_cast<unsigned char*>(aligned_alloc(64, 8 << 20)); memset(o, 0, 8 << 20); }
            std::vector<int> turn(threads, 0);
            std::atomic<int> bad{0};
            t = run(threads, n, per, [&](int tid, int64_t off, int64_t len) {
                uint64_t o = 0; size_t used = 0;
                unsigned char* slot = slots[tid * 2 + (turn[tid]++ & 1)];
                if (!snapb::encode_deltas(src + off, len, slot, 8 << 20, o, used)) bad.store(1);
            });
            printf("threads=%2d encode_deltas    %.3f s  %.1f GB/s read%s\n", threads, t, n * 8e-9 / t, bad.load() ? "  (overflow!)" : "");
            for (auto& o : slots) free(o);
        }
        for (auto& o : outs) free(o);
    }
    return 0;
}

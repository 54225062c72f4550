// Shared helpers for the snapb200 CUDA sources (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <stdexcept>

namespace snapb {

// ---------------------------------------------------------------- errors --
void set_last_error(const std::string& msg);

struct Error : public std::runtime_error {
    explicit Error(const std::string& m) : std::runtime_error(m) {}
};

#define SB_CUDA(expr)                                                          \
    do {                                                                       \
        cudaError_t _e = (expr);                                               \
        if (_e != cudaSuccess) {                                               \
            char _buf[512];                                                    \
            snprintf(_buf, sizeof(_buf), "CUDA error %s at %s:%d: %s",         \
                     cudaGetErrorName(_e), __FILE__, __LINE__,                 \
                     cudaGetErrorString(_e));                                  \
            throw ::snapb::Error(_buf);                                        \
        }                                                                      \
    } while (0)

#define SB_CHECK(cond, msg)                                                    \
    do {                                                                       \
        if (!(cond)) {                                                         \
            char _buf[512];                                                    \
            snprintf(_buf, sizeof(_buf), "%s (%s:%d)", (msg), __FILE__, __LINE__); \
            throw ::snapb::Error(_buf);                                        \
        }                                                                      \
    } while (0)

// Launch-check: catches bad configurations immediately.
#define SB_LAUNCH_CHECK() SB_CUDA(cudaGetLastError())

// ------------------------------------------------------------- constants --
constexpr int kNumSMsB200 = 148;
constexpr int kWarp = 32;

// ----------------------------------------------------- counter-based RNG --
// splitmix64 finaliser over seed*M1 + a*M2 + b*M3; mirrored bit for bit by
// snapatac2_b200/synth.py:mix64.
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t seed, uint64_t a, uint64_t b) {
    uint64_t x = seed * 0x9E3779B97F4A7C15ull + a * 0xBF58476D1CE4E5B9ull + b * 0x94D049BB133111EBull;
    x ^= x >> 30;
    x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27;
    x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return x;
}

// --------------------------------------------------------- device helpers --
#ifdef __CUDACC__
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Streaming 128-bit / 32-bit loads that do not pollute L1 (index streams are
// read exactly once per pass).
__device__ __forceinline__ int4 ld_stream_int4(const int4* p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ int ld_stream_int(const int* p) {
    int r;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ float ld_stream_float(const float* p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
#endif

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace snapb

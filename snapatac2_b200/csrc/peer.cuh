// Small all-reduce over NVLink peer memory, fused into the kernel that produces the operand.
//
// The block Lanczos step needs three tiny fp64 all-reduces (the Gram matrices of gram_kernel and
// project_chol_apply_kernel, <= 23 KB each).  Through NCCL each costs a launch and ~20-25 us of
// latency at 8 ranks -- about a quarter of the orthogonalisation time of a 1/8 shard.  Here the CTA
// that finishes a Gram kernel's fixed-order reduction also performs the exchange: it stores its
// vector straight into every peer's mailbox (plain stores to IPC-mapped peer memory, over
// NVLink / NVSwitch), publishes a sequence number, waits for the peers' sequence numbers in its own
// mailbox and adds the contributions up in rank order (bitwise reproducible, identical on every
// rank).  No second kernel, no NCCL call.
//
// Mailbox of a rank (its own device memory, cudaMalloc'ed, exported with cudaIpcGetMemHandle):
//     data  [kPeerSlots][nranks][kPeerMaxLen] doubles     written by the peers
//     flags [kPeerSlots][nranks] unsigned long long       sequence number of what data holds
// Exchange number `seq` uses slot seq % kPeerSlots.  Nobody can be more than one exchange ahead of
// the slowest rank (an exchange completes only after every rank has entered it), so a slot is never
// overwritten before its reader is done with it.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace snapb {

constexpr int kPeerMaxRanks = 8;
constexpr int kPeerSlots = 4;
constexpr int kPeerMaxLen = 3072;          // doubles per contribution (>= (168 + 16) * 16)

struct PeerBox {
    int rank = 0, nranks = 1;              // nranks == 1: no exchange
    unsigned long long seq = 0;
    double* data[kPeerMaxRanks];           // base of every rank's data area as mapped into this process
    unsigned long long* flags[kPeerMaxRanks];
    int* error = nullptr;                  // set if a peer never showed up (spin limit)
};

#ifdef __CUDACC__
__device__ __forceinline__ void st_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_volatile_f64(const double* p) {
    double v;
    asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// Called by every thread of ONE CTA: vec[0:len) (this rank's contribution, global memory, already
// written by this CTA and fenced) is replaced by the sum over the ranks, added in rank order.
__device__ __forceinline__ void peer_allreduce(const PeerBox& box, double* vec, int len) {
    if (box.nranks <= 1) return;
    const int slot = static_cast<int>(box.seq % kPeerSlots);
    const size_t mine = (static_cast<size_t>(slot) * box.nranks + box.rank) * kPeerMaxLen;
    __syncthreads();
    // my contribution into every rank's mailbox (my own included): coalesced stores over NVLink
    for (int p = 0; p < box.nranks; ++p) {
        double* dst = box.data[p] + mine;
        for (int e = threadIdx.x; e < len; e += blockDim.x) dst[e] = vec[e];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < box.nranks)
        st_sys_u64(box.flags[threadIdx.x] + static_cast<size_t>(slot) * box.nranks + box.rank, box.seq);
    // wait for everybody's sequence number in my own mailbox
    if (threadIdx.x < box.nranks) {
        const unsigned long long* f = box.flags[box.rank] + static_cast<size_t>(slot) * box.nranks + threadIdx.x;
        long long spins = 0;
        while (ld_sys_u64(f) != box.seq) {
            if (++spins > (1ll << 25)) {       // tens of seconds: a peer is gone; report instead of hanging the device
                if (box.error) *box.error = 1;
                break;
            }
        }
        __threadfence_system();
    }
    __syncthreads();
    const double* base = box.data[box.rank] + static_cast<size_t>(slot) * box.nranks * kPeerMaxLen;
    for (int e = threadIdx.x; e < len; e += blockDim.x) {
        double s = 0.0;
        for (int q = 0; q < box.nranks; ++q) s += ld_volatile_f64(base + static_cast<size_t>(q) * kPeerMaxLen + e);
        vec[e] = s;
    }
}
#endif

}  // namespace snapb

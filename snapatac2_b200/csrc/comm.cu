// NCCL plumbing: one communicator per context, all collectives on the
// context's own stream so they order with the kernels around them.
#include "ctx.cuh"
#include "peer.cuh"

#include <nccl.h>
#include <stdlib.h>
#include <string.h>

namespace snapb {

struct Comm {
    ncclComm_t comm = nullptr;
    // peer-memory mailboxes of the fused small all-reduce (peer.cuh); peer_ok == false: NCCL is used instead
    bool peer_ok = false;
    PeerBox box;
    void* my_data = nullptr;
    void* my_flags = nullptr;
    void* opened[2 * kPeerMaxRanks] = {};
    int n_opened = 0;
    int* error_flag = nullptr;
};

#define SB_NCCL(expr)                                                          \
    do {                                                                       \
        ncclResult_t _r = (expr);                                              \
        if (_r != ncclSuccess) {                                               \
            char _buf[512];                                                    \
            snprintf(_buf, sizeof(_buf), "NCCL error at %s:%d: %s", __FILE__,  \
                     __LINE__, ncclGetErrorString(_r));                        \
            throw ::snapb::Error(_buf);                                        \
        }                                                                      \
    } while (0)

void comm_unique_id(char id[128]) {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId uid;
    SB_NCCL(ncclGetUniqueId(&uid));
    memcpy(id, &uid, 128);
}

void comm_init(snapb200_ctx* c, int rank, int nranks, const char id[128]) {
    SB_CHECK(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank/nranks");
    comm_destroy(c);
    c->rank = rank;
    c->nranks = nranks;
    if (nranks == 1) return;
    ncclUniqueId uid;
    memcpy(&uid, id, 128);
    c->comm = new Comm();
    SB_CUDA(cudaSetDevice(c->device));
    SB_NCCL(ncclCommInitRank(&c->comm->comm, nranks, uid, rank));
    peer_setup(c);
}

// Mailboxes in every rank's memory, mapped into every other rank's address space (CUDA IPC).  Any failure
// (ranks in one process, no peer access, more than kPeerMaxRanks ranks, SNAPB200_NO_PEER set) leaves
// peer_ok false on EVERY rank -- the outcome is agreed by an all-reduce -- and the small all-reduces go
// through NCCL as before.
void peer_setup(snapb200_ctx* c) {
    Comm* cm = c->comm;
    const int nranks = c->nranks, rank = c->rank;
    cudaStream_t st = c->stream;
    int ok = (nranks <= kPeerMaxRanks && getenv("SNAPB200_NO_PEER") == nullptr) ? 1 : 0;
    const size_t data_bytes = sizeof(double) * kPeerSlots * nranks * kPeerMaxLen;
    const size_t flag_bytes = sizeof(unsigned long long) * kPeerSlots * nranks;
    struct Handles { cudaIpcMemHandle_t data, flags; };
    Handles mine{};
    if (ok) {
        if (cudaMalloc(&cm->my_data, data_bytes) != cudaSuccess || cudaMalloc(&cm->my_flags, flag_bytes) != cudaSuccess ||
            cudaMalloc(reinterpret_cast<void**>(&cm->error_flag), sizeof(int)) != cudaSuccess) {
            ok = 0;
        } else {
            cudaMemset(cm->my_data, 0, data_bytes);
            cudaMemset(cm->my_flags, 0, flag_bytes);
            cudaMemset(cm->error_flag, 0, sizeof(int));
            if (cudaIpcGetMemHandle(&mine.data, cm->my_data) != cudaSuccess ||
                cudaIpcGetMemHandle(&mine.flags, cm->my_flags) != cudaSuccess)
                ok = 0;
        }
        cudaGetLastError();
    }
    // everybody's handles (the collectives run even if this rank already failed: the ranks must stay in step)
    DevBuf<unsigned char> dall;
    dall.alloc(static_cast<int64_t>(sizeof(Handles)) * nranks);
    DevBuf<unsigned char> dmine;
    dmine.alloc(sizeof(Handles));
    SB_CUDA(cudaMemcpyAsync(dmine.p, &mine, sizeof(Handles), cudaMemcpyHostToDevice, st));
    SB_NCCL(ncclAllGather(dmine.p, dall.p, sizeof(Handles), ncclUint8, cm->comm, st));
    std::vector<Handles> all(nranks);
    SB_CUDA(cudaMemcpyAsync(all.data(), dall.p, sizeof(Handles) * nranks, cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    PeerBox box;
    box.rank = rank;
    box.nranks = nranks;
    box.error = cm->error_flag;
    if (ok) {
        for (int p = 0; p < nranks && ok; ++p) {
            if (p == rank) {
                box.data[p] = static_cast<double*>(cm->my_data);
                box.flags[p] = static_cast<unsigned long long*>(cm->my_flags);
                continue;
            }
            void *pd = nullptr, *pf = nullptr;
            if (cudaIpcOpenMemHandle(&pd, all[p].data, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; break; }
            cm->opened[cm->n_opened++] = pd;
            if (cudaIpcOpenMemHandle(&pf, all[p].flags, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; break; }
            cm->opened[cm->n_opened++] = pf;
            box.data[p] = static_cast<double*>(pd);
            box.flags[p] = static_cast<unsigned long long*>(pf);
        }
        cudaGetLastError();
    }
    // agree: peer path only if it works everywhere
    DevBuf<int64_t> agree;
    agree.alloc(1);
    int64_t v = ok ? 0 : 1;
    SB_CUDA(cudaMemcpyAsync(agree.p, &v, sizeof(int64_t), cudaMemcpyHostToDevice, st));
    SB_NCCL(ncclAllReduce(agree.p, agree.p, 1, ncclInt64, ncclSum, cm->comm, st));
    SB_CUDA(cudaMemcpyAsync(&v, agree.p, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    cm->peer_ok = (v == 0);
    cm->box = box;
}

bool peer_box(snapb200_ctx* c, PeerBox* out) {
    PeerBox b;
    if (c->comm == nullptr || !c->comm->peer_ok || c->nranks <= 1) {
        *out = b;
        return false;
    }
    b = c->comm->box;
    b.seq = ++c->comm->box.seq;       // every rank issues the same sequence of exchanges
    *out = b;
    return true;
}

bool peer_error(snapb200_ctx* c) {
    if (c->comm == nullptr || !c->comm->peer_ok) return false;
    int h = 0;
    cudaMemcpyAsync(&h, c->comm->error_flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream);
    cudaStreamSynchronize(c->stream);
    return h != 0;
}

void comm_destroy(snapb200_ctx* c) {
    if (c->comm) {
        for (int i = 0; i < c->comm->n_opened; ++i) cudaIpcCloseMemHandle(c->comm->opened[i]);
        if (c->comm->my_data) cudaFree(c->comm->my_data);
        if (c->comm->my_flags) cudaFree(c->comm->my_flags);
        if (c->comm->error_flag) cudaFree(c->comm->error_flag);
        if (c->comm->comm) ncclCommDestroy(c->comm->comm);
        delete c->comm;
        c->comm = nullptr;
    }
    c->rank = 0;
    c->nranks = 1;
}

void allreduce_f32(snapb200_ctx* c, float* buf, int64_t count) {
    if (c->nranks == 1 || count == 0) return;
    SB_NCCL(ncclAllReduce(buf, buf, static_cast<size_t>(count), ncclFloat32, ncclSum, c->comm->comm, c->stream));
}
void allreduce_f64(snapb200_ctx* c, double* buf, int64_t count) {
    if (c->nranks == 1 || count == 0) return;
    SB_NCCL(ncclAllReduce(buf, buf, static_cast<size_t>(count), ncclFloat64, ncclSum, c->comm->comm, c->stream));
}
void allreduce_i64(snapb200_ctx* c, int64_t* buf, int64_t count) {
    if (c->nranks == 1 || count == 0) return;
    SB_NCCL(ncclAllReduce(buf, buf, static_cast<size_t>(count), ncclInt64, ncclSum, c->comm->comm, c->stream));
}

}  // namespace snapb

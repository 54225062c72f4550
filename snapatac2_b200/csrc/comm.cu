// NCCL plumbing: one communicator per context, all collectives on the
// context's own stream so they order with the kernels around them.
#include "ctx.cuh"

#include <nccl.h>
#include <string.h>

namespace snapb {

struct Comm {
    ncclComm_t comm = nullptr;
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
}

void comm_destroy(snapb200_ctx* c) {
    if (c->comm) {
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

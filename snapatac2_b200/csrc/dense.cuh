// Host-side handles for the dense kernels in dense.cu.
#pragma once

#include "ctx.cuh"

namespace snapb {

// Dense block operations for block width B (4, 8 or 16).  The basis Q is
// fp32 row-major with leading dimension ldq; small matrices are fp64.
template <int B>
struct DenseOps {
    DevBuf<double> partial;   // per-CTA partial sums (fixed-order reduction)
    // size `partial` once for a basis of up to ld columns (no reallocation inside the solver loop)
    void reserve(snapb200_ctx* c, int64_t n, int ld);

    // H[ncq x B] = Q[:, 0:ncq]^T Z
    void gram(snapb200_ctx* c, const float* Q, int64_t ldq, int ncq, const float* Z, int64_t ldz, int64_t n, double* H);
    // G[B x B] = Z^T Z
    void zz(snapb200_ctx* c, const float* Z, int64_t ldz, int64_t n, double* G);
    // out = {R, Rinv, Rtot, flags}; G = R^T R.  ref_diag (optional B x B Gram of the block before
    // projection) lets columns whose norm collapsed be flagged as dependent.
    void chol(snapb200_ctx* c, const double* G, const double* ref_diag, double* out, bool first);
    // dst = Z * Rinv
    void apply_rinv(snapb200_ctx* c, const float* Z, int64_t ldz, const double* Rinv, int64_t n, float* dst, int64_t ldd);
    // Z -= Q[:, 0:ncq] H
    void project_out(snapb200_ctx* c, const float* Q, int64_t ldq, int ncq, const double* H, int64_t n, float* Z, int64_t ldz);
    void random_block(snapb200_ctx* c, float* Z, int64_t ldz, int64_t n, uint64_t seed, uint64_t stream);
};

// out[n x p] = Q[:, 0:ncq] S   (S fp64 row-major with leading dimension lds)
void tall_gemm_f32(snapb200_ctx* c, const float* Q, int64_t ldq, int ncq, const double* S, int lds, int p, int64_t n,
                   float* out, int64_t ldo);
void tall_gemm_f64(snapb200_ctx* c, const float* Q, int64_t ldq, int ncq, const double* S, int lds, int p, int64_t n,
                   double* out, int64_t ldo);
void copy_cols(snapb200_ctx* c, const float* src, int64_t lds, float* dst, int64_t ldd, int64_t n, int ncols);
double dense_selftest(snapb200_ctx* c, int64_t n, int ncq, int p);

// Symmetric eigen-decomposition on the host (Householder + implicit QL).
// a: n x n row-major (destroyed); on return w ascending, a holds eigenvectors
// in its columns (a[i*n + j] = component i of eigenvector j).
void sym_eig(int n, double* a, double* w);

}  // namespace snapb

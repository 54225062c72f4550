// Host-side handles for the dense kernels in dense.cu.
#pragma once

#include "ctx.cuh"
#include "peer.cuh"

namespace snapb {

// Dense block operations for block width B (4, 8 or 16).  The basis Q is
// fp32 row-major with leading dimension ldq; small matrices are fp64.
template <int B>
struct DenseOps {
    DevBuf<double> partial;     // per-CTA partial sums of the Gram kernel (fixed-order reduction)
    DevBuf<double> partial_g;   // per-CTA partial B x B Grams of project_chol_apply
    DevBuf<unsigned> counters;  // tickets of the last-CTA reductions (self-resetting)
    int pca_smem_set = 0;
    // size the partial buffers once for a basis of up to ld columns (no reallocation inside the solver loop)
    void reserve(snapb200_ctx* c, int64_t n, int ld);

    // H[(ncq + nzx) x B] = [Q[:, 0:ncq] | Zx[:, 0:nzx]]^T Z   (one kernel; Zx = Z gives Q^T Z and Z^T Z at once)
    // box (optional): the result is also summed over the ranks through the peers' mailboxes, inside the kernel
    void gram_ext(snapb200_ctx* c, const float* Q, int64_t ldq, int ncq, const float* Zx, int64_t ldzx, int nzx,
                  const float* Z, int64_t ldz, int64_t n, double* H, const PeerBox* box);
    // H[ncq x B] = Q[:, 0:ncq]^T Z
    void gram(snapb200_ctx* c, const float* Q, int64_t ldq, int ncq, const float* Z, int64_t ldz, int64_t n, double* H);
    // Z -= Q[:, 0:ncq] H
    void project_out(snapb200_ctx* c, const float* Q, int64_t ldq, int ncq, const double* H, int64_t n, float* Z, int64_t ldz);
    // Z (packed n x B) <- (Z - Q H) R1^-1 with R1 = chol(G' - H^T H), Hext = [H ; G'];  chol1 = {R1[B*B], flags[B]};
    // G3 = Gram of the result.  G0 (optional): Gram of the block before any projection (dependence test).
    void project_chol_apply(snapb200_ctx* c, const float* Q, int64_t ldq, int ncq, const double* Hext, const double* G0,
                            float* Z, int64_t n, double* chol1, double* G3, const PeerBox* box);
    // R2 = chol(G3); dst = Z R2^-1 (leading dimension ldd), Vr = rscale .* dst (if rscale);
    // out = {Rtot = R2 R1 [B*B], flags[B]}.  rows_too = false: factorisation and `out` only.
    void chol_append(snapb200_ctx* c, const double* G3, const double* chol1, const float* Z, int64_t n, float* dst,
                     int64_t ldd, const float* rscale, float* Vr, double* out, bool rows_too);
    void random_block(snapb200_ctx* c, float* Z, int64_t ldz, int64_t n, uint64_t seed, uint64_t stream);
};

// out[n x p] = Q[:, 0:ncq] S   (S fp64 row-major with leading dimension lds)
void tall_gemm_f32(snapb200_ctx* c, const float* Q, int64_t ldq, int ncq, const double* S, int lds, int p, int64_t n,
                   float* out, int64_t ldo);
void tall_gemm_f64(snapb200_ctx* c, const float* Q, int64_t ldq, int ncq, const double* S, int lds, int p, int64_t n,
                   double* out, int64_t ldo);
void copy_cols(snapb200_ctx* c, const float* src, int64_t lds, float* dst, int64_t ldd, int64_t n, int ncols);
double dense_selftest(snapb200_ctx* c, int64_t n, int ncq, int p);
double ortho_selftest(snapb200_ctx* c, int64_t n, int ncols, int block);

// Symmetric eigen-decomposition on the host (Householder + implicit QL).
// a: n x n row-major (destroyed); on return w ascending, a holds eigenvectors
// in its columns (a[i*n + j] = component i of eigenvector j).
void sym_eig(int n, double* a, double* w);

}  // namespace snapb

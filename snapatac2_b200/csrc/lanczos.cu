// Device-resident block Lanczos eigensolver with full re-orthogonalisation and
// thick restart.  Replaces the ARPACK call the reference embeds at
// snapatac2-python/src/embedding.rs:158-171:
//     eigsh(LinearOperator(v -> X~ X~^T v - dinv v), k, which='LM', tol=0)
//     followed by argsort()[::-1].
//
// The reference hands ARPACK one vector at a time (ncv = 2k+1, ~110 mat-vecs
// for k=30).  Here the operator is applied to b vectors per sweep over the
// index stream, the basis stays on the device in fp32, and only b x b / B x b
// fp64 matrices visit the host for the Rayleigh-Ritz step:
//
//   basis column 0 holds the analytically known trivial eigenvector
//   u1 = sqrt(d)/||sqrt(d)|| (lambda = 1, T[0,0] = 1), so every block is
//   orthogonalised against it and fp32 round-off of the dominant pair never
//   pollutes the O(1e-2) eigenvalues that follow;
//   each step:  Z = A Q_last;  H = Q^T Z, Z -= Q H (twice, classical
//   Gram-Schmidt on the FP64 tensor cores);  CholQR2(Z) -> next block;
//   T[:, last] = H  (the projection is formed explicitly, so a thick restart
//   only needs T = diag(theta) for the kept Ritz vectors);
//   residual of a Ritz pair = || R s_last ||  with R the CholQR factor.
#include "ctx.cuh"
#include "dense.cuh"
#include "peer.cuh"

#include <stdlib.h>
#include <algorithm>
#include <chrono>
#include <cmath>
#include <numeric>
#include <vector>

namespace snapb {

// ==========================================================================
// Host symmetric eigensolver: Householder tridiagonalisation followed by the
// implicit-shift QL iteration (the classical tred2/tql2 pair), fp64.
// ==========================================================================
void sym_eig(int n, double* a, double* w) {
    if (n == 0) return;
    std::vector<double> e(n, 0.0);
    auto A = [&](int i, int j) -> double& { return a[static_cast<size_t>(i) * n + j]; };

    // ---- reduce to tridiagonal form, accumulating the transformation in a
    for (int i = n - 1; i >= 1; --i) {
        const int l = i - 1;
        double h = 0.0, scale = 0.0;
        if (l > 0) {
            for (int k = 0; k <= l; ++k) scale += std::fabs(A(i, k));
            if (scale == 0.0) {
                e[i] = A(i, l);
            } else {
                for (int k = 0; k <= l; ++k) {
                    A(i, k) /= scale;
                    h += A(i, k) * A(i, k);
                }
                double f = A(i, l);
                double g = (f >= 0.0) ? -std::sqrt(h) : std::sqrt(h);
                e[i] = scale * g;
                h -= f * g;
                A(i, l) = f - g;
                f = 0.0;
                for (int j = 0; j <= l; ++j) {
                    A(j, i) = A(i, j) / h;
                    g = 0.0;
                    for (int k = 0; k <= j; ++k) g += A(j, k) * A(i, k);
                    for (int k = j + 1; k <= l; ++k) g += A(k, j) * A(i, k);
                    e[j] = g / h;
                    f += e[j] * A(i, j);
                }
                const double hh = f / (h + h);
                for (int j = 0; j <= l; ++j) {
                    f = A(i, j);
                    e[j] = g = e[j] - hh * f;
                    for (int k = 0; k <= j; ++k) A(j, k) -= (f * e[k] + g * A(i, k));
                }
            }
        } else {
            e[i] = A(i, l);
        }
        w[i] = h;
    }
    w[0] = 0.0;
    e[0] = 0.0;
    for (int i = 0; i < n; ++i) {
        const int l = i - 1;
        if (w[i] != 0.0) {
            for (int j = 0; j <= l; ++j) {
                double g = 0.0;
                for (int k = 0; k <= l; ++k) g += A(i, k) * A(k, j);
                for (int k = 0; k <= l; ++k) A(k, j) -= g * A(k, i);
            }
        }
        w[i] = A(i, i);
        A(i, i) = 1.0;
        for (int j = 0; j <= l; ++j) A(j, i) = A(i, j) = 0.0;
    }

    // ---- QL with implicit shifts on (w, e), rotating the columns of a
    for (int i = 1; i < n; ++i) e[i - 1] = e[i];
    e[n - 1] = 0.0;
    for (int l = 0; l < n; ++l) {
        int iter = 0, m;
        do {
            for (m = l; m < n - 1; ++m) {
                const double dd = std::fabs(w[m]) + std::fabs(w[m + 1]);
                if (std::fabs(e[m]) <= 2.3e-16 * dd) break;
            }
            if (m != l) {
                if (iter++ == 200) throw Error("sym_eig: QL iteration did not converge");
                double g = (w[l + 1] - w[l]) / (2.0 * e[l]);
                double r = std::hypot(g, 1.0);
                g = w[m] - w[l] + e[l] / (g + (g >= 0.0 ? std::fabs(r) : -std::fabs(r)));
                double s = 1.0, cth = 1.0, p = 0.0;
                int i;
                for (i = m - 1; i >= l; --i) {
                    double f = s * e[i];
                    const double b = cth * e[i];
                    e[i + 1] = (r = std::hypot(f, g));
                    if (r == 0.0) {
                        w[i + 1] -= p;
                        e[m] = 0.0;
                        break;
                    }
                    s = f / r;
                    cth = g / r;
                    g = w[i + 1] - p;
                    r = (w[i] - g) * s + 2.0 * cth * b;
                    w[i + 1] = g + (p = s * r);
                    g = cth * r - b;
                    for (int k = 0; k < n; ++k) {
                        f = A(k, i + 1);
                        A(k, i + 1) = s * A(k, i) + cth * f;
                        A(k, i) = cth * A(k, i) - s * f;
                    }
                }
                if (r == 0.0 && i >= l) continue;
                w[l] -= p;
                e[l] = g;
                e[m] = 0.0;
            }
        } while (m != l);
    }
    // ---- ascending order
    std::vector<int> ord(n);
    std::iota(ord.begin(), ord.end(), 0);
    std::sort(ord.begin(), ord.end(), [&](int x, int y) { return w[x] < w[y]; });
    std::vector<double> w2(n), a2(static_cast<size_t>(n) * n);
    for (int j = 0; j < n; ++j) {
        w2[j] = w[ord[j]];
        for (int i = 0; i < n; ++i) a2[static_cast<size_t>(i) * n + j] = A(i, ord[j]);
    }
    std::copy(w2.begin(), w2.end(), w);
    std::copy(a2.begin(), a2.end(), a);
}

namespace {

inline int round_up8(int x) { return (x + 7) / 8 * 8; }

struct HostClock {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    double ms() const {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
};

struct Ritz {
    std::vector<int> idx;        // valid basis columns, ascending
    std::vector<double> theta;   // ascending
    std::vector<double> S;       // nv x nv, columns = eigenvectors
    std::vector<int> by_mag;     // indices into theta, |theta| descending
};

// One block step on the device = the operator + ortho_block():
//   K1  Hext1 = [Q | Z]^T Z                 (projection coefficients and the block's Gram G0 at once)
//   AR  all-reduce Hext1                    (one small fp64 all-reduce)
//   P1  Z -= Q H1
//   K1  Hext2 = [Q | Z]^T Z                 (second Gram-Schmidt pass + Gram G' of the projected block)
//   AR
//   K3  Z <- (Z - Q H2) R1^-1, R1 = chol(G' - H2^T H2);  G3 = Z^T Z          (project_chol_apply)
//   AR  all-reduce G3
//   K4  R2 = chol(G3);  Q[:, nb:nb+B] = Z R2^-1;  Vr = r .* (new block)      (chol_append)
// 5 launches and 3 tiny all-reduces, no host round trip; the first version took 16 launches,
// 6 all-reduces and a stream synchronisation per step.  What the host needs (T's new columns
// H1 + H2, Rtot = R2 R1, the dependence flags) comes back in ONE asynchronous copy into a pinned
// slot and is looked at one step late: the host enqueues step j + 1 before it reads step j, so the
// GPU never waits for the Rayleigh-Ritz check.  The price is one speculative step when the check
// says "converged"; near convergence (worst residual ratio below `kSyncRatio`) the loop therefore
// switches to checking before it enqueues.
template <int B>
void eigsh_impl(snapb200_ctx* c, int k, int64_t seed, double tol, int max_basis, int max_ops, double* evals,
                double* evecs, bool scale_by_sqrt_eval) {
    HostClock wall;
    const int64_t n = c->n_local, ng = c->n_global;
    cudaStream_t st = c->stream;
    SB_CHECK(k >= 1 && k < ng, "eigsh: k must satisfy 1 <= k < n_obs");

    const int keep = round_up8(k + std::max(8, k / 2));
    int ld = max_basis > 0 ? round_up8(max_basis) : round_up8(std::max(128, 2 * keep + 2 * B));
    ld = std::max(ld, keep + 2 * B + 8);
    if (max_ops <= 0) max_ops = 1000;
    if (!(tol > 0.0)) tol = 1e-5;
    double sync_ratio = 30.0;
    if (const char* e = getenv("SNAPB200_SYNC_RATIO")) sync_ratio = atof(e);
    const bool debug = getenv("SNAPB200_DEBUG") != nullptr;

    DenseOps<B> ops;
    ops.reserve(c, n, ld);
    DevBuf<float> Q, Z, Qtmp;
    DevBuf<double> dStep, dChol1, dG3, dS, dEvec;
    Q.alloc(std::max<int64_t>(1, n * ld));
    Z.alloc(std::max<int64_t>(1, n * B));
    c->Vr.ensure(std::max<int64_t>(1, n * B));
    // device step buffer, copied to the host in one piece: Hext1 | Hext2 | Rtot | flags
    const int64_t LB = static_cast<int64_t>(ld + B) * B;
    const int64_t step_len = 2 * LB + B * B + B;
    dStep.alloc(step_len);
    dChol1.alloc(B * B + B);
    dG3.alloc(B * B);
    double* dH1 = dStep.p;
    double* dH2 = dStep.p + LB;
    double* dOut = dStep.p + 2 * LB;
    // pinned staging lives in the context (cudaMallocHost / cudaFreeHost per call cost hundreds of
    // milliseconds on a device with ~100 GB mapped); two slots: step j is read while j + 1 runs
    c->pinned.ensure(static_cast<int64_t>(sizeof(double)) * 2 * step_len);
    double* hslot[2] = {reinterpret_cast<double*>(c->pinned.p), reinterpret_cast<double*>(c->pinned.p) + step_len};

    std::vector<double> T(static_cast<size_t>(ld) * ld, 0.0);
    std::vector<char> valid(ld, 0);
    auto Tm = [&](int i, int j) -> double& { return T[static_cast<size_t>(i) * ld + j]; };

    cudaEvent_t evs[2][7];
    for (auto& es : evs)
        for (auto& e : es) SB_CUDA(cudaEventCreate(&e));
    double ms_spmm = 0.0, ms_comm = 0.0, ms_ortho = 0.0, ms_host = 0.0;

    // everything after the operator, up to (and, if `append_at` >= 0, including) the append
    auto ortho_block = [&](int nbq, int append_at) {
        // the three small all-reduces: fused into the producing kernels over peer memory (peer.cuh) when the
        // mailboxes are up, NCCL otherwise
        PeerBox box;
        bool fused = peer_box(c, &box);
        ops.gram_ext(c, Q.p, ld, nbq, Z.p, B, B, Z.p, B, n, dH1, fused ? &box : nullptr);
        if (!fused) allreduce_f64(c, dH1, static_cast<int64_t>(nbq + B) * B);
        ops.project_out(c, Q.p, ld, nbq, dH1, n, Z.p, B);
        fused = peer_box(c, &box);
        ops.gram_ext(c, Q.p, ld, nbq, Z.p, B, B, Z.p, B, n, dH2, fused ? &box : nullptr);
        if (!fused) allreduce_f64(c, dH2, static_cast<int64_t>(nbq + B) * B);
        fused = peer_box(c, &box);
        ops.project_chol_apply(c, Q.p, ld, nbq, dH2, dH1 + static_cast<int64_t>(nbq) * B, Z.p, n, dChol1.p, dG3.p,
                               fused ? &box : nullptr);
        if (!fused) allreduce_f64(c, dG3.p, B * B);
        if (append_at >= 0)
            ops.chol_append(c, dG3.p, dChol1.p, Z.p, n, Q.p + append_at, ld, c->r.p, c->Vr.p, dOut, true);
        else
            ops.chol_append(c, dG3.p, dChol1.p, Z.p, n, nullptr, 0, nullptr, nullptr, dOut, false);
    };
    auto append_block = [&](int at) {
        ops.chol_append(c, dG3.p, dChol1.p, Z.p, n, Q.p + at, ld, c->r.p, c->Vr.p, dOut, true);
    };

    // ---- basis block 0 = [u1, 0 ... 0]
    SB_CUDA(cudaMemsetAsync(Q.p, 0, sizeof(float) * static_cast<size_t>(std::max<int64_t>(1, n * ld)), st));
    copy_cols(c, c->u1.p, 1, Q.p, ld, n, 1);
    valid[0] = 1;
    Tm(0, 0) = 1.0;
    int nb = 8;

    // ---- first Krylov block from a counter-hash random start
    ops.random_block(c, Z.p, B, n, static_cast<uint64_t>(seed), 0);
    ortho_block(nb, nb);
    SB_CUDA(cudaMemcpyAsync(hslot[0], dStep.p, sizeof(double) * step_len, cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    for (int j = 0; j < B; ++j) valid[nb + j] = hslot[0][2 * LB + B * B + j] == 0.0;
    nb += B;

    Ritz ritz;
    auto solve_ritz = [&](int nbase) {
        HostClock hc;
        ritz.idx.clear();
        for (int i = 0; i < nbase; ++i)
            if (valid[i]) ritz.idx.push_back(i);
        const int nv = static_cast<int>(ritz.idx.size());
        ritz.S.assign(static_cast<size_t>(nv) * nv, 0.0);
        for (int a = 0; a < nv; ++a)
            for (int b2 = 0; b2 < nv; ++b2)
                ritz.S[static_cast<size_t>(a) * nv + b2] = 0.5 * (Tm(ritz.idx[a], ritz.idx[b2]) + Tm(ritz.idx[b2], ritz.idx[a]));
        ritz.theta.assign(nv, 0.0);
        sym_eig(nv, ritz.S.data(), ritz.theta.data());
        ritz.by_mag.resize(nv);
        std::iota(ritz.by_mag.begin(), ritz.by_mag.end(), 0);
        std::stable_sort(ritz.by_mag.begin(), ritz.by_mag.end(),
                         [&](int x, int y) { return std::fabs(ritz.theta[x]) > std::fabs(ritz.theta[y]); });
        ms_host += hc.ms();
    };

    struct Step {
        int nb = 0;           // basis width the step was orthogonalised against
        int slot = 0;
        bool appended = false;   // the new block is already in Q[:, nb : nb + B]
    };
    int64_t n_enqueued = 0;
    auto enqueue_step = [&](int nbq) {
        Step s;
        s.nb = nbq;
        s.slot = static_cast<int>(n_enqueued++ & 1);
        cudaEvent_t* ev = evs[s.slot];
        operator_apply_dev(c, Q.p + (nbq - B), ld, Z.p, B, B, ev, /*vr_ready=*/true);
        SB_CUDA(cudaEventRecord(ev[4], st));
        s.appended = nbq + B <= ld;          // room for the new block: no thick restart before the append
        ortho_block(nbq, s.appended ? nbq : -1);
        SB_CUDA(cudaEventRecord(ev[5], st));
        SB_CUDA(cudaMemcpyAsync(hslot[s.slot], dStep.p, sizeof(double) * step_len, cudaMemcpyDeviceToHost, st));
        SB_CUDA(cudaEventRecord(ev[6], st));
        return s;
    };

    int64_t n_ops = 0, n_restarts = 0, n_spec = 0;
    double max_res = 0.0, last_ratio = 1e300;
    bool have_ritz = false, converged = false, exhausted = false;
    const int min_check = (k + B - 1) / B + 2;
    int nb_final = nb;

    Step cur = enqueue_step(nb);
    while (true) {
        // speculate: enqueue the next step before looking at this one (see the comment above)
        const bool may_spec = cur.appended && last_ratio > sync_ratio && n_ops + 2 < max_ops;
        Step next;
        bool have_next = false;
        if (may_spec) {
            next = enqueue_step(cur.nb + B);
            have_next = true;
        }
        // ---- read step `cur`
        SB_CUDA(cudaEventSynchronize(evs[cur.slot][6]));
        ++n_ops;
        {
            float a = 0.f, b2 = 0.f, c2 = 0.f, d2 = 0.f;
            cudaEvent_t* ev = evs[cur.slot];
            cudaEventElapsedTime(&a, ev[0], ev[1]);
            cudaEventElapsedTime(&b2, ev[1], ev[2]);
            cudaEventElapsedTime(&c2, ev[2], ev[3]);
            cudaEventElapsedTime(&d2, ev[4], ev[5]);
            ms_spmm += a + c2;
            ms_comm += b2;
            ms_ortho += d2;
        }
        const double* hH1 = hslot[cur.slot];
        const double* hH2 = hslot[cur.slot] + LB;
        const double* Rtot = hslot[cur.slot] + 2 * LB;
        const double* flags = Rtot + B * B;
        nb = cur.nb;
        const int last0 = nb - B;   // first column of the block the operator was applied to
        // ---- T[:, last block] = H1 + H2 (explicit projection), symmetric fill
        for (int i = 0; i < nb; ++i)
            for (int j = 0; j < B; ++j) {
                const double h = hH1[static_cast<size_t>(i) * B + j] + hH2[static_cast<size_t>(i) * B + j];
                Tm(i, last0 + j) = h;
                if (i < last0) Tm(last0 + j, i) = h;
            }
        bool new_valid[B];
        bool any_new = false;
        for (int j = 0; j < B; ++j) {
            new_valid[j] = flags[j] == 0.0;
            any_new = any_new || new_valid[j];
        }
        int nv_now = 0;
        for (int i = 0; i < nb; ++i) nv_now += valid[i] ? 1 : 0;

        const bool must_restart = any_new && (nb + B > ld);
        const bool out_of_ops = n_ops >= max_ops;
        const bool forced = !any_new || out_of_ops;
        const bool due = nv_now > k && n_ops >= min_check;
        have_ritz = false;
        if (due || forced || must_restart) {
            solve_ritz(nb);
            have_ritz = true;
            const int nv = static_cast<int>(ritz.idx.size());
            if (nv >= k) {
                // residual estimates || Rtot s_last ||
                std::vector<int> lastpos(B, -1);
                for (int a = 0; a < nv; ++a)
                    if (ritz.idx[a] >= last0) lastpos[ritz.idx[a] - last0] = a;
                max_res = 0.0;
                double worst_ratio = 0.0;
                for (int q = 0; q < k; ++q) {
                    const int col = ritz.by_mag[q];
                    double r2 = 0.0;
                    for (int i = 0; i < B; ++i) {
                        double sum = 0.0;
                        for (int j = 0; j < B; ++j)
                            if (lastpos[j] >= 0) sum += Rtot[i * B + j] * ritz.S[static_cast<size_t>(lastpos[j]) * nv + col];
                        r2 += sum * sum;
                    }
                    const double res = std::sqrt(r2);
                    max_res = std::max(max_res, res);
                    worst_ratio = std::max(worst_ratio, res / (tol * std::max(std::fabs(ritz.theta[col]), 1e-300)));
                }
                last_ratio = worst_ratio;
                converged = worst_ratio <= 1.0;
                if (debug) fprintf(stderr, "[snapb200] step %lld nb=%d worst residual ratio %.3e%s\n",
                                   static_cast<long long>(n_ops), nb, worst_ratio, have_next ? " (next step in flight)" : "");
            }
        }
        if (converged || forced) {
            exhausted = !any_new;
            nb_final = nb;
            if (have_next) ++n_spec;   // its device work is discarded
            break;
        }

        int at = nb;   // where the new block goes
        if (must_restart) {
            // ---- thick restart: keep the `keep` largest-|theta| Ritz vectors (never speculated past)
            const int nv = static_cast<int>(ritz.idx.size());
            const int p = std::min(keep, nv);
            const int p8 = round_up8(p);
            std::vector<double> Smat(static_cast<size_t>(nb) * p8, 0.0);
            for (int q = 0; q < p; ++q) {
                const int col = ritz.by_mag[q];
                for (int a = 0; a < nv; ++a) Smat[static_cast<size_t>(ritz.idx[a]) * p8 + q] = ritz.S[static_cast<size_t>(a) * nv + col];
            }
            dS.ensure(static_cast<int64_t>(nb) * p8);
            Qtmp.ensure(std::max<int64_t>(1, n * p8));
            SB_CUDA(cudaMemcpyAsync(dS.p, Smat.data(), sizeof(double) * Smat.size(), cudaMemcpyHostToDevice, st));
            tall_gemm_f32(c, Q.p, ld, nb, dS.p, p8, p8, n, Qtmp.p, p8);
            copy_cols(c, Qtmp.p, p8, Q.p, ld, n, p8);
            SB_CUDA(cudaStreamSynchronize(st));   // Smat is a host temporary
            std::fill(T.begin(), T.end(), 0.0);
            std::fill(valid.begin(), valid.end(), 0);
            for (int q = 0; q < p; ++q) {
                Tm(q, q) = ritz.theta[ritz.by_mag[q]];
                valid[q] = 1;
            }
            at = p8;
            ++n_restarts;
        }
        if (!cur.appended) append_block(at);   // Q[:, at : at + B] = Z R2^-1
        for (int j = 0; j < B; ++j) valid[at + j] = new_valid[j];
        nb = at + B;
        cur = have_next ? next : enqueue_step(nb);
    }
    nb = nb_final;

    // ---- final Ritz extraction: k largest |theta|, sorted by descending value
    if (!have_ritz) solve_ritz(nb);
    const int nv = static_cast<int>(ritz.idx.size());
    SB_CHECK(nv >= k, "eigsh: Krylov space smaller than k");
    std::vector<int> sel(ritz.by_mag.begin(), ritz.by_mag.begin() + k);
    std::stable_sort(sel.begin(), sel.end(), [&](int x, int y) { return ritz.theta[x] > ritz.theta[y]; });
    const int k8 = round_up8(k);
    std::vector<double> Smat(static_cast<size_t>(nb) * k8, 0.0);
    for (int q = 0; q < k; ++q) {
        evals[q] = ritz.theta[sel[q]];
        // weighted_by_sd of the wrapper (tools/_embedding.py:286-289) folded into the rotation: columns with a
        // positive eigenvalue come out multiplied by sqrt(lambda) at no cost (the others are dropped by the caller)
        const double sc = (scale_by_sqrt_eval && evals[q] > 0.0) ? std::sqrt(evals[q]) : 1.0;
        for (int a = 0; a < nv; ++a) Smat[static_cast<size_t>(ritz.idx[a]) * k8 + q] = sc * ritz.S[static_cast<size_t>(a) * nv + sel[q]];
    }
    dS.ensure(static_cast<int64_t>(nb) * k8);
    SB_CUDA(cudaMemcpyAsync(dS.p, Smat.data(), sizeof(double) * Smat.size(), cudaMemcpyHostToDevice, st));
    if (n > 0) {
        dEvec.alloc(n * k);
        tall_gemm_f64(c, Q.p, ld, nb, dS.p, k8, k, n, dEvec.p, k);
        HostClock hx;
        copy_to_host(c, evecs, dEvec.p, sizeof(double) * static_cast<size_t>(n) * k);   // pageable numpy memory: threaded staging
        c->stats.ms_d2h = hx.ms();
    }
    SB_CUDA(cudaStreamSynchronize(st));
    for (auto& es : evs)
        for (auto& e : es) cudaEventDestroy(e);
    SB_CHECK(!peer_error(c), "eigsh: a rank never arrived at a fused all-reduce (peer mailbox timed out)");

    c->stats.ms_eigsh = wall.ms();
    c->stats.ms_spmm = ms_spmm;
    c->stats.ms_comm = ms_comm;
    c->stats.ms_ortho = ms_ortho;
    c->stats.ms_host = ms_host;
    c->stats.max_residual = max_res;
    c->stats.n_ops = n_ops;
    c->stats.n_restarts = n_restarts;
    c->stats.basis_cols = nb;
    c->stats.block = B;
    c->stats.n_spec_ops = n_spec;
    {
        PeerBox probe;
        c->stats.fused_allreduce = (c->nranks > 1 && peer_box(c, &probe)) ? 1 : 0;   // (advances the sequence on every rank alike)
    }
    // converged: every wanted pair met the tolerance, or the Krylov space is exhausted (the Ritz
    // pairs are then exact).  Not converged: max_ops reached first -- scipy's eigsh raises
    // ArpackNoConvergence there; the Python mirror does the same from this flag.
    c->stats.converged = (converged || exhausted) ? 1 : 0;
}

}  // namespace

void eigsh(snapb200_ctx* c, int k, int64_t seed, double tol, int block, int max_basis, int max_ops, double* evals,
           double* evecs, bool scale_by_sqrt_eval) {
    SB_CHECK(c->prepared, "eigsh: call prepare first");
    if (block <= 0) block = c->block;
    switch (block) {
        case 4: eigsh_impl<4>(c, k, seed, tol, max_basis, max_ops, evals, evecs, scale_by_sqrt_eval); break;
        case 8: eigsh_impl<8>(c, k, seed, tol, max_basis, max_ops, evals, evecs, scale_by_sqrt_eval); break;
        case 16: eigsh_impl<16>(c, k, seed, tol, max_basis, max_ops, evals, evecs, scale_by_sqrt_eval); break;
        default: throw Error("eigsh: block width must be 4, 8 or 16");
    }
}

}  // namespace snapb

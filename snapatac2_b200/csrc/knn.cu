// Exact k-nearest-neighbour graph of the embedding (the consumer of `X_spectral`).
//
// Reference: `snap.pp.knn(adata, n_neighbors, method="kdtree")` (preprocessing/_knn.py:53-87) ->
// `nearest_neighbour_graph` (snapatac2-core/src/utils/knn.rs:9-33): a kd-tree (crate kdtree 0.7) over the
// f64 points, per point the k nearest OTHER points (the point's own index is filtered out, knn.rs:27),
// Euclidean distance = sqrt(squared_euclidean), rows of the CSR result sorted by column (knn.rs:65).
//
// The tree is a CPU device for skipping work; on a B200 the exact answer is cheaper by brute force with
// a filter:
//   * every (query, point) pair gets a float32 lower bound of its squared distance from the dot form
//     n_i + n_j - 2 p_i.p_j (points centred, 128 x 128 tiles, 8 x 8 register blocks -- an FFMA-bound
//     SGEMM-shaped loop); the bound is made rigorous by shrinking the norms by (DP + 8) eps, which covers
//     the rounding of the operands, of the dot product and of the final subtraction;
//   * a pair whose bound is not above the query's current k-th distance goes to a shared-memory queue;
//     the owner warp of the query recomputes the distance in float64 exactly as the reference does
//     (differences, products and a left-to-right sum, no fused multiply-add: bit-identical to kdtree's
//     `squared_euclidean`) and keeps the k smallest (distance, index) pairs in shared memory.
// The result is therefore the exact graph -- the same neighbours and the same float64 distances as an
// exact search on the CPU; ties at the k-th distance are broken by the smaller index (the kd-tree's tie
// order is unspecified).  The expected number of float64 evaluations per query is ~k ln(n / k).
//
// Float64 tensor-core MMA does not apply (the filter is float32, the exact stage is scalar by
// construction of the reference's summation order); a tcgen05 tf32x3 version of the filter is the
// obvious next step and is not done here.
#include "ctx.cuh"

#include <math_constants.h>

#include <chrono>

namespace snapb {
namespace {

constexpr int kT = 128;               // queries per CTA = points per staged tile
constexpr int kKnnThreads = 256;
constexpr int kOwnerQueries = 16;     // warp w owns queries [16 w, 16 w + 16)
constexpr int kQueueCap = kOwnerQueries * kT;   // every pair of a tile can pass (first tiles)
constexpr int kKnnMaxK = 100;
constexpr int kKnnMaxDim = 64;

// column sums (for the centring; any summation order will do: the centre only tightens the filter)
template <int DP>
__global__ void __launch_bounds__(256) knn_colsum_kernel(const double* __restrict__ P, int64_t n, int d, double* __restrict__ sum) {
    double acc[DP];
#pragma unroll
    for (int k = 0; k < DP; ++k) acc[k] = 0.0;
    for (int64_t j = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; j < n; j += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const double* row = P + j * d;
#pragma unroll
        for (int k = 0; k < DP; ++k)
            if (k < d) acc[k] += row[k];
    }
    __shared__ double s_sum[kKnnMaxDim];
    if (threadIdx.x < kKnnMaxDim) s_sum[threadIdx.x] = 0.0;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < DP; ++k) {
        double v = acc[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && k < d) atomicAdd(&s_sum[k], v);
    }
    __syncthreads();
    if (threadIdx.x < d) atomicAdd(&sum[threadIdx.x], s_sum[threadIdx.x]);
}

// centred float32 copy, dimension-major (Pt[k * npad + j]), and the float64 squared norm of the centred
// point; columns j >= n are padding: zeros with a NaN norm (a NaN bound never passes the filter)
template <int DP>
__global__ void __launch_bounds__(256) knn_prep_kernel(const double* __restrict__ P, int64_t n, int64_t npad, int d,
                                                       const double* __restrict__ sum, float* __restrict__ Pt,
                                                       double* __restrict__ nrm64) {
    const int64_t j = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (j >= npad) return;
    double nn = 0.0;
    if (j < n) {
        const double* row = P + j * d;
#pragma unroll
        for (int k = 0; k < DP; ++k) {
            double a = 0.0;
            if (k < d) a = row[k] - sum[k] / static_cast<double>(n);
            nn += a * a;
            Pt[k * npad + j] = static_cast<float>(a);
        }
    } else {
#pragma unroll
        for (int k = 0; k < DP; ++k) Pt[k * npad + j] = 0.f;
        nn = CUDART_NAN;
    }
    nrm64[j] = nn;
}

struct KnnSmem {
    // byte offsets into the dynamic shared memory of knn_scan_kernel
    int tau64, nlo64, list_d, As, Bs, Bn, Tq, tauj, argmax, cnt, qcnt, list_j, queue, total;
};

inline KnnSmem knn_smem_layout(int DP, int K) {
    KnnSmem s;
    int o = 0;
    auto take = [&](int bytes) { const int at = o; o += (bytes + 15) & ~15; return at; };
    s.tau64 = take(kT * 8);
    s.nlo64 = take(kT * 8);
    s.list_d = take(kT * K * 8);
    s.As = take(DP * kT * 4);
    s.Bs = take(DP * kT * 4);
    s.Bn = take(kT * 4);
    s.Tq = take(kT * 4);
    s.tauj = take(kT * 4);
    s.argmax = take(kT * 4);
    s.cnt = take(kT * 4);
    s.qcnt = take((kKnnThreads / 32) * 4);
    s.list_j = take(kT * K * 4);
    s.queue = take((kKnnThreads / 32) * kQueueCap * 2);
    s.total = o;
    return s;
}

// One CTA = 128 consecutive queries, swept over every tile of 128 points.
template <int DP>
__global__ void __launch_bounds__(kKnnThreads, 1)
knn_scan_kernel(const float* __restrict__ Pt, const double* __restrict__ nrm64, const double* __restrict__ P64,
                int64_t n, int64_t npad, int d, int64_t q0, int64_t nq, int K, double shrink, KnnSmem L,
                int32_t* __restrict__ out_j, double* __restrict__ out_d) {
    extern __shared__ __align__(16) unsigned char smem[];
    double* tau64 = reinterpret_cast<double*>(smem + L.tau64);
    double* nlo64 = reinterpret_cast<double*>(smem + L.nlo64);
    double* list_d = reinterpret_cast<double*>(smem + L.list_d);
    float* As = reinterpret_cast<float*>(smem + L.As);
    float* Bs = reinterpret_cast<float*>(smem + L.Bs);
    float* Bn = reinterpret_cast<float*>(smem + L.Bn);
    float* Tq = reinterpret_cast<float*>(smem + L.Tq);
    int* tauj = reinterpret_cast<int*>(smem + L.tauj);
    int* argmax = reinterpret_cast<int*>(smem + L.argmax);
    int* cnt = reinterpret_cast<int*>(smem + L.cnt);
    int* qcnt = reinterpret_cast<int*>(smem + L.qcnt);
    int* list_j = reinterpret_cast<int*>(smem + L.list_j);
    unsigned short* queue = reinterpret_cast<unsigned short*>(smem + L.queue);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ty = tid >> 4, tx = tid & 15;
    const int64_t qlocal0 = static_cast<int64_t>(blockIdx.x) * kT;     // first query of this CTA, relative to q0
    const int64_t qbase = q0 + qlocal0;                                 // ... as a point index

    // ---- the query tile and the per-query state
    for (int e = tid; e < DP * kT; e += kKnnThreads) {
        const int k = e >> 7, q = e & (kT - 1);
        const int64_t i = qbase + q;
        As[e] = (qlocal0 + q < nq && i < n) ? Pt[k * npad + i] : 0.f;
    }
    if (tid < kT) {
        const int64_t i = qbase + tid;
        const bool valid = qlocal0 + tid < nq && i < n;
        tau64[tid] = CUDART_INF;
        tauj[tid] = 0x7fffffff;
        argmax[tid] = 0;
        cnt[tid] = 0;
        nlo64[tid] = valid ? nrm64[i] * shrink : 0.0;
        Tq[tid] = valid ? CUDART_INF_F : -CUDART_INF_F;    // a query outside the range accepts nothing
    }
    if (tid < kKnnThreads / 32) qcnt[tid] = 0;

    const int64_t n_tiles = npad / kT;
    constexpr int kPre = DP / 8;          // float4 per thread of one point tile (DP * 128 / 4 / 256)
    float4 pre[kPre];
    float pre_n = 0.f;
    auto fetch = [&](int64_t t) {
        const int64_t tile0 = t * kT;
#pragma unroll
        for (int r = 0; r < kPre; ++r) {
            const int e4 = tid + kKnnThreads * r;           // float4 index inside the tile: 32 per dimension
            const int k = e4 >> 5, c4 = e4 & 31;
            pre[r] = *reinterpret_cast<const float4*>(Pt + k * npad + tile0 + c4 * 4);
        }
        if (tid < kT) pre_n = __double2float_rd(nrm64[tile0 + tid] * shrink);
    };
    fetch(0);
    __syncthreads();

    for (int64_t t = 0; t < n_tiles; ++t) {
        const int64_t tile0 = t * kT;
#pragma unroll
        for (int r = 0; r < kPre; ++r) {
            const int e4 = tid + kKnnThreads * r;
            *reinterpret_cast<float4*>(Bs + e4 * 4) = pre[r];
        }
        if (tid < kT) Bn[tid] = pre_n;
        __syncthreads();
        if (t + 1 < n_tiles) fetch(t + 1);

        // ---- 8 x 8 dot products per thread
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
#pragma unroll 4
        for (int k = 0; k < DP; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(As + k * kT + ty * 4);
            const float4 a1 = *reinterpret_cast<const float4*>(As + k * kT + 64 + ty * 4);
            const float4 b0 = *reinterpret_cast<const float4*>(Bs + k * kT + tx * 4);
            const float4 b1 = *reinterpret_cast<const float4*>(Bs + k * kT + 64 + tx * 4);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        // ---- filter: lower bound of the squared distance against the query's current k-th distance
        {
            const float4 n0 = *reinterpret_cast<const float4*>(Bn + tx * 4);
            const float4 n1 = *reinterpret_cast<const float4*>(Bn + 64 + tx * 4);
            const float bn[8] = {n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, n1.z, n1.w};
            const float4 t0 = *reinterpret_cast<const float4*>(Tq + ty * 4);
            const float4 t1 = *reinterpret_cast<const float4*>(Tq + 64 + ty * 4);
            const float tq[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
            bool any = false;                      // the common case: nothing passes, one branch for the 64 pairs
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) any |= fmaf(-2.f, acc[i][j], bn[j]) <= tq[i];
            if (any) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int q = (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (fmaf(-2.f, acc[i][j], bn[j]) <= tq[i]) {
                            const int p = (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
                            const int owner = q / kOwnerQueries;
                            const int pos = atomicAdd(&qcnt[owner], 1);
                            queue[owner * kQueueCap + pos] = static_cast<unsigned short>((q << 7) | p);
                        }
                    }
                }
            }
        }
        __syncthreads();

        // ---- the owner warp: exact float64 distances, 32 candidates at a time, then serial insertion
        const int ne = qcnt[warp];
        for (int base = 0; base < ne; base += 32) {
            const int e = base + lane;
            const bool have = e < ne;
            const int code = have ? queue[warp * kQueueCap + e] : 0;
            const int q = code >> 7, p = code & (kT - 1);
            const int64_t i = qbase + q, j = tile0 + p;
            const bool ok = have && j < n && j != i;
            double d2 = CUDART_INF;
            if (ok) {
                const double* x = P64 + i * d;
                const double* y = P64 + j * d;
                double s = 0.0;
                for (int k = 0; k < d; ++k) {             // the reference's fold: ((x - y) * (x - y)) summed left to right
                    const double df = __dsub_rn(x[k], y[k]);
                    s = __dadd_rn(s, __dmul_rn(df, df));
                }
                d2 = s;
            }
            const int jj32 = static_cast<int>(j);
            unsigned pend = __ballot_sync(0xffffffffu, ok && (d2 < tau64[q] || (d2 == tau64[q] && jj32 < tauj[q])));
            while (pend) {
                const int src = __ffs(pend) - 1;
                pend &= pend - 1;
                const int qq = __shfl_sync(0xffffffffu, q, src);
                const int jn = __shfl_sync(0xffffffffu, jj32, src);
                const double dd = __shfl_sync(0xffffffffu, d2, src);
                const double tcur = tau64[qq];
                if (!(dd < tcur || (dd == tcur && jn < tauj[qq]))) continue;      // an earlier insertion tightened the bound
                const int c = cnt[qq];
                const int slot = c < K ? c : argmax[qq];
                __syncwarp();
                if (lane == 0) {
                    list_d[qq * K + slot] = dd;
                    list_j[qq * K + slot] = jn;
                    if (c < K) cnt[qq] = c + 1;
                }
                __syncwarp();
                if (c + 1 >= K) {
                    // the list is full: its largest (distance, index) is the new bound
                    double bd = -1.0;
                    int bj = -1, bs = 0;
                    for (int s = lane; s < K; s += 32) {
                        const double v = list_d[qq * K + s];
                        const int vj = list_j[qq * K + s];
                        if (v > bd || (v == bd && vj > bj)) { bd = v; bj = vj; bs = s; }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const double od = __shfl_xor_sync(0xffffffffu, bd, o);
                        const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
                        const int os = __shfl_xor_sync(0xffffffffu, bs, o);
                        if (od > bd || (od == bd && oj > bj)) { bd = od; bj = oj; bs = os; }
                    }
                    if (lane == 0) {
                        tau64[qq] = bd;
                        tauj[qq] = bj;
                        argmax[qq] = bs;
                        Tq[qq] = __double2float_ru(bd - nlo64[qq]);
                    }
                    __syncwarp();
                }
            }
        }
        if (lane == 0) qcnt[warp] = 0;
        __syncthreads();
    }

    // ---- rows of the result: sorted by column, distances as square roots (knn.rs:27, :65)
    for (int q = warp * kOwnerQueries; q < (warp + 1) * kOwnerQueries; ++q) {
        if (qlocal0 + q >= nq || cnt[q] != K) continue;
        for (int s = lane; s < K; s += 32) {
            const int j = list_j[q * K + s];
            int rank = 0;
            for (int u = 0; u < K; ++u) rank += list_j[q * K + u] < j;
            out_j[(qlocal0 + q) * K + rank] = j;
            out_d[(qlocal0 + q) * K + rank] = sqrt(list_d[q * K + s]);
        }
    }
}

template <int DP>
void knn_run(snapb200_ctx* c, const double* P64, int64_t n, int64_t npad, int d, int64_t q0, int64_t nq, int K,
             float* Pt, double* nrm64, double* colsum, int32_t* out_j, double* out_d) {
    cudaStream_t st = c->stream;
    SB_CUDA(cudaMemsetAsync(colsum, 0, sizeof(double) * kKnnMaxDim, st));
    knn_colsum_kernel<DP><<<c->num_sms * 2, 256, 0, st>>>(P64, n, d, colsum);
    SB_LAUNCH_CHECK();
    knn_prep_kernel<DP><<<static_cast<unsigned>(ceil_div(npad, 256)), 256, 0, st>>>(P64, n, npad, d, colsum, Pt, nrm64);
    SB_LAUNCH_CHECK();
    const KnnSmem L = knn_smem_layout(DP, K);
    SB_CHECK(L.total <= 232448, "knn: n_neighbors does not fit the shared-memory lists (at most 100; 84 with more than 32 dimensions)");
    SB_CUDA(cudaFuncSetAttribute(knn_scan_kernel<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    const double shrink = 1.0 - static_cast<double>(DP + 8) * 5.9604644775390625e-08;   // (DP + 8) * 2^-24
    knn_scan_kernel<DP><<<static_cast<unsigned>(ceil_div(nq, kT)), kKnnThreads, L.total, st>>>(
        Pt, nrm64, P64, n, npad, d, q0, nq, K, shrink, L, out_j, out_d);
    SB_LAUNCH_CHECK();
    count_launch(c);
    count_launch(c);
    count_launch(c);
}

}  // namespace

int knn_max_neighbors() { return kKnnMaxK; }
int knn_max_dim() { return kKnnMaxDim; }

// points: n x d float64 row-major (host, or device when on_device); queries are the points
// [q0, q0 + nq); out_indices / out_distances: nq x min(k, n - 1) host arrays, rows sorted by index.
void knn(snapb200_ctx* c, int64_t n, int d, const double* points, int on_device, int64_t q0, int64_t nq, int k,
         int32_t* out_indices, double* out_distances) {
    SB_CHECK(n >= 1 && d >= 1, "knn: the matrix is empty");
    SB_CHECK(d <= kKnnMaxDim, "knn: at most 64 dimensions (use_dims)");
    SB_CHECK(n < (static_cast<int64_t>(1) << 31), "knn: at most 2^31 - 1 points");
    SB_CHECK(k >= 1, "knn: n_neighbors must be positive");
    SB_CHECK(q0 >= 0 && nq >= 0 && q0 + nq <= n, "knn: query range outside the points");
    const int K = static_cast<int>(std::min<int64_t>(k, n - 1));
    SB_CHECK(K <= kKnnMaxK, "knn: at most 100 neighbours");
    if (K == 0 || nq == 0) return;
    SB_CHECK(points != nullptr && out_indices != nullptr && out_distances != nullptr, "knn: null argument");
    const auto wall0 = std::chrono::steady_clock::now();
    cudaStream_t st = c->stream;
    const int64_t npad = ceil_div(n, kT) * kT;
    DevBuf<double> P64, nrm64, colsum, out_d;
    DevBuf<float> Pt;
    DevBuf<int32_t> out_j;
    const double* Pdev = points;
    if (!on_device) {
        P64.alloc(n * d);
        SB_CUDA(cudaMemcpyAsync(P64.p, points, sizeof(double) * n * d, cudaMemcpyHostToDevice, st));
        Pdev = P64.p;
    }
    const int DP = d <= 8 ? 8 : d <= 16 ? 16 : d <= 32 ? 32 : 64;
    Pt.alloc(npad * DP);
    nrm64.alloc(npad);
    colsum.alloc(kKnnMaxDim);
    out_j.alloc(nq * K);
    out_d.alloc(nq * K);
    SB_CUDA(cudaEventRecord(c->ev0, st));
    switch (DP) {
        case 8: knn_run<8>(c, Pdev, n, npad, d, q0, nq, K, Pt.p, nrm64.p, colsum.p, out_j.p, out_d.p); break;
        case 16: knn_run<16>(c, Pdev, n, npad, d, q0, nq, K, Pt.p, nrm64.p, colsum.p, out_j.p, out_d.p); break;
        case 32: knn_run<32>(c, Pdev, n, npad, d, q0, nq, K, Pt.p, nrm64.p, colsum.p, out_j.p, out_d.p); break;
        default: knn_run<64>(c, Pdev, n, npad, d, q0, nq, K, Pt.p, nrm64.p, colsum.p, out_j.p, out_d.p); break;
    }
    SB_CUDA(cudaEventRecord(c->ev1, st));
    copy_to_host(c, out_indices, out_j.p, sizeof(int32_t) * nq * K);
    copy_to_host(c, out_distances, out_d.p, sizeof(double) * nq * K);
    SB_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f;
    SB_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->stats.ms_knn = ms;
    c->stats.ms_knn_wall = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count();
}

}  // namespace snapb

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
// The filter has two implementations: FFMA (below) and mma.sync tf32 x 3 (further down, the default where it fits);
// the exact stage is scalar by construction of the reference's summation order.  A tcgen05 version of the filter
// is the obvious next step and is not done here.
#include "ctx.cuh"

#include <cuda_pipeline.h>
#include <math_constants.h>

#include <chrono>

namespace snapb {
namespace {

constexpr int kT = 128;               // queries per CTA = points per staged tile
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

__device__ __forceinline__ float to_tf32(float x) {
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// Tensor-core operand layout (Ptf): per tile of 128 points one contiguous block of DP x 128 x 2 floats, ordered
// [k / 8][point / 8][lane][hi b0, hi b1, lo b0, lo b1] where lane = (point % 8) * 4 + k % 4 holds the B fragment of
// mma.m16n8k8 (b0: k % 8 < 4, b1: k % 8 >= 4): a warp reads its fragment, hi and lo, with one 16-byte load per lane.
template <int DP>
__device__ __forceinline__ int64_t knn_frag_index(int64_t j, int k) {
    const int64_t tile = j >> 7;
    const int pl = static_cast<int>(j & 127), kk = k & 7;
    return (((tile * (DP / 8) + (k >> 3)) * 16 + (pl >> 3)) * 32 + (pl & 7) * 4 + (kk & 3)) * 4 + (kk >> 2);
}

// Per point: the centred float32 copy, dimension-major (Pt[k * npad + j]); the float64 squared norm of the
// centred point and its shrunk float32 lower bound; the original float64 row padded to DP terms (P64p, the
// exact stage reads it with compile-time trip counts: trailing (0 - 0)^2 terms add +0.0, which changes no sum).
// Columns j >= n are padding: zeros with NaN norms (a NaN bound never passes the filter).
template <int DP>
__global__ void __launch_bounds__(256) knn_prep_kernel(const double* __restrict__ P, int64_t n, int64_t npad, int d,
                                                       const double* __restrict__ sum, double shrink, float* __restrict__ Pt,
                                                       double* __restrict__ nrm64, float* __restrict__ nlo32,
                                                       double* __restrict__ P64p, int* __restrict__ bad,
                                                       float* __restrict__ Ptf) {
    const int64_t j = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (j >= npad) return;
    double nn = 0.0;
    if (j < n) {
        const double* row = P + j * d;
#pragma unroll
        for (int k = 0; k < DP; ++k) {
            double v = 0.0, a = 0.0;
            if (k < d) {
                v = row[k];
                a = v - sum[k] / static_cast<double>(n);
            }
            nn += a * a;
            Pt[k * npad + j] = static_cast<float>(a);
            P64p[j * DP + k] = v;
            if (Ptf != nullptr) {          // tf32 split for the tensor-core filter: a = hi + lo + O(2^-22 |a|)
                const float hi = to_tf32(static_cast<float>(a));
                const int64_t at = knn_frag_index<DP>(j, k);
                Ptf[at] = hi;
                Ptf[at + 2] = to_tf32(static_cast<float>(a - static_cast<double>(hi)));
            }
        }
        // NaN / inf coordinates (the reference's kd-tree refuses them: NonFiniteCoordinate), or a spread float32 cannot hold
        if (!(nn <= 3.0e38)) *bad = 1;
    } else {
#pragma unroll
        for (int k = 0; k < DP; ++k) {
            Pt[k * npad + j] = 0.f;
            P64p[j * DP + k] = 0.0;
            if (Ptf != nullptr) {
                const int64_t at = knn_frag_index<DP>(j, k);
                Ptf[at] = 0.f;
                Ptf[at + 2] = 0.f;
            }
        }
        nn = CUDART_NAN;
    }
    nrm64[j] = nn;
    nlo32[j] = __double2float_rd(nn * shrink);
}

constexpr int kProducers = 8;          // warps 0-7: the float32 filter
constexpr int kConsumers = 8;          // warps 8-15: exact distances + lists; warp 8 + c owns queries [16 c, 16 c + 16)
constexpr int kScanThreads = 32 * (kProducers + kConsumers);
constexpr int kConsQueries = kT / kConsumers;
constexpr int kQueueSlots = 512;       // per consumer and buffer; a tile that passes more flags the query for a full re-check

struct KnnSmem {
    // byte offsets into the dynamic shared memory of knn_scan_kernel
    int nlo64, list_d, q64, As, Bs, Bn, Tq, qcnt, redo, list_j, queue, total;
};

// `stage_q`: the padded float64 rows of the CTA's queries are kept in shared memory (row stride DP + 2 doubles)
inline KnnSmem knn_smem_layout(int DP, int K, bool stage_q) {
    KnnSmem s;
    int o = 0;
    auto take = [&](int bytes) { const int at = o; o += (bytes + 15) & ~15; return at; };
    s.nlo64 = take(kT * 8);
    s.list_d = take(kT * K * 8);
    s.q64 = take(stage_q ? kT * (DP + 2) * 8 : 0);
    s.As = take(DP * kT * 4);
    s.Bs = take(2 * DP * kT * 4);
    s.Bn = take(2 * kT * 4);
    s.Tq = take(kT * 4);
    s.qcnt = take(2 * kConsumers * 4);
    s.redo = take(2 * kConsumers * 4);
    s.list_j = take(kT * K * 4);
    s.queue = take(2 * kConsumers * kQueueSlots * 2);
    s.total = o;
    return s;
}

// kdtree 0.7 `squared_euclidean`: ((x - y) * (x - y)) summed left to right, one rounding per operation.
// Both rows are padded to DP terms and 16-byte aligned: the point row is fetched with DP / 2 independent
// 16-byte loads (one round trip), the arithmetic keeps the reference's order.
template <int DP>
__device__ __forceinline__ double exact_d2(const double* __restrict__ x, const double* __restrict__ y) {
    double2 yv[DP / 2];
#pragma unroll
    for (int u = 0; u < DP / 2; ++u) yv[u] = __ldg(reinterpret_cast<const double2*>(y) + u);
    double s = 0.0;
#pragma unroll
    for (int u = 0; u < DP / 2; ++u) {
        const double2 xv = *(reinterpret_cast<const double2*>(x) + u);
        const double d0 = __dsub_rn(xv.x, yv[u].x);
        s = __dadd_rn(s, __dmul_rn(d0, d0));
        const double d1 = __dsub_rn(xv.y, yv[u].y);
        s = __dadd_rn(s, __dmul_rn(d1, d1));
    }
    return s;
}

// Per-query lists of the CTA's 128 queries (shared memory, touched by the consumer warps only): K entries
// (distance^2, index) kept in ascending order, empty slots = (+inf, INT_MAX); the K-th is the bound.
struct KnnLists {
    double* list_d;    // [q][K]
    int* list_j;       // [q][K]
    float* Tq;         // float32 bound the filter compares with (rounded up)
    const double* nlo64;
    int K;
};

__device__ __forceinline__ bool knn_before(double a, int aj, double b, int bj) { return a < b || (a == b && aj < bj); }

// One warp offers up to 32 candidates (one per lane: `ok`, query q, point j, exact distance^2 d2) to the lists
// of its own queries.  Insertions are serial, in lane order; one insertion is a parallel shift: every lane
// owns the slots lane, lane + 32, ... and takes its left neighbour's entry where that one moves up.
__device__ __forceinline__ void knn_offer(const KnnLists& S, bool ok, int q, int j, double d2, int lane) {
    const int K = S.K;
    unsigned pend = __ballot_sync(0xffffffffu, ok && knn_before(d2, j, S.list_d[q * K + K - 1], S.list_j[q * K + K - 1]));
    while (pend) {
        const int src = __ffs(pend) - 1;
        pend &= pend - 1;
        const int qq = __shfl_sync(0xffffffffu, q, src);
        const int jn = __shfl_sync(0xffffffffu, j, src);
        const double dd = __shfl_sync(0xffffffffu, d2, src);
        double* ld = S.list_d + qq * K;
        int* lj = S.list_j + qq * K;
        if (!knn_before(dd, jn, ld[K - 1], lj[K - 1])) continue;      // an earlier insertion tightened the bound
        double nv[4];
        int nj[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int s = lane + 32 * r;
            if (s < K) {
                const double v = ld[s];
                const int vj = lj[s];
                const bool left_moves = s > 0 && knn_before(dd, jn, ld[s - 1], lj[s - 1]);     // left neighbour is behind the newcomer
                if (left_moves) { nv[r] = ld[s - 1]; nj[r] = lj[s - 1]; }
                else if (knn_before(dd, jn, v, vj)) { nv[r] = dd; nj[r] = jn; }
                else { nv[r] = v; nj[r] = vj; }
            }
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int s = lane + 32 * r;
            if (s < K) { ld[s] = nv[r]; lj[s] = nj[r]; }
        }
        __syncwarp();
        if (lane == 0) *reinterpret_cast<volatile float*>(S.Tq + qq) = __double2float_ru(ld[K - 1] - S.nlo64[qq]);
    }
}

// One CTA = 128 consecutive queries, swept over every tile of 128 points.  Eight producer warps run the
// float32 filter of tile t (and fetch tile t + 1 with asynchronous copies) while eight consumer warps settle
// what tile t - 1 let through (exact distances, lists); one barrier per tile hands the candidate queue over.
// The bound the producers read may therefore be one tile old -- it only ever decreases, so a stale bound
// lets a few more candidates through, never fewer.
template <int DP, bool DBG, int UNROLL>
__global__ void __launch_bounds__(kScanThreads, 1)
knn_scan_kernel(const float* __restrict__ Pt, const double* __restrict__ nrm64, const float* __restrict__ nlo32,
                const double* __restrict__ P64p, int64_t n, int64_t npad, int64_t q0, int64_t nq, int K, double shrink,
                KnnSmem L, int32_t* __restrict__ out_j, double* __restrict__ out_d, int probe, int stage_q,
                unsigned long long* __restrict__ dbg) {
    extern __shared__ __align__(16) unsigned char smem[];
    double* nlo64 = reinterpret_cast<double*>(smem + L.nlo64);
    double* list_d = reinterpret_cast<double*>(smem + L.list_d);
    double* Q64 = reinterpret_cast<double*>(smem + L.q64);
    float* As = reinterpret_cast<float*>(smem + L.As);
    float* Bs = reinterpret_cast<float*>(smem + L.Bs);
    float* Bn = reinterpret_cast<float*>(smem + L.Bn);
    float* Tq = reinterpret_cast<float*>(smem + L.Tq);
    int* qcnt = reinterpret_cast<int*>(smem + L.qcnt);
    unsigned* redo = reinterpret_cast<unsigned*>(smem + L.redo);
    int* list_j = reinterpret_cast<int*>(smem + L.list_j);
    unsigned short* queue = reinterpret_cast<unsigned short*>(smem + L.queue);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool producer = warp < kProducers;
    const int ty = (tid >> 4) & 15, tx = tid & 15;
    const int64_t qlocal0 = static_cast<int64_t>(blockIdx.x) * kT;     // first query of this CTA, relative to q0
    const int64_t qbase = q0 + qlocal0;                                 // ... as a point index

    // ---- the query tile and the per-query state
    for (int e = tid; e < DP * kT; e += kScanThreads) {
        const int k = e >> 7, q = e & (kT - 1);
        const int64_t i = qbase + q;
        As[e] = (qlocal0 + q < nq && i < n) ? Pt[k * npad + i] : 0.f;
    }
    if (tid < kT) {
        const int64_t i = qbase + tid;
        const bool valid = qlocal0 + tid < nq && i < n;
        nlo64[tid] = valid ? nrm64[i] * shrink : 0.0;
        Tq[tid] = valid ? CUDART_INF_F : -CUDART_INF_F;    // a query outside the range accepts nothing
    }
    if (tid < 2 * kConsumers) { qcnt[tid] = 0; redo[tid] = 0u; }
    for (int e = tid; e < kT * K; e += kScanThreads) {
        list_d[e] = CUDART_INF;
        list_j[e] = 0x7fffffff;
    }
    constexpr int kQStride = DP + 2;
    if (stage_q) {
        for (int e = tid; e < kT * DP; e += kScanThreads) {
            const int q = e / DP, k = e % DP;
            Q64[q * kQStride + k] = (qbase + q < npad) ? P64p[(qbase + q) * DP + k] : 0.0;
        }
    }
    const double* qrows = stage_q ? Q64 : P64p + qbase * DP;     // row q of the CTA's queries: qrows + q * qstride
    const int qstride = stage_q ? kQStride : DP;

    const int64_t n_tiles = npad / kT;
    // tile t -> Bs[t & 1], Bn[t & 1]: 16-byte asynchronous copies issued by the producer threads
    auto fetch = [&](int64_t t) {
        const int64_t tile0 = t * kT;
        float* B = Bs + (t & 1) * DP * kT;
#pragma unroll
        for (int r = 0; r < DP / 8; ++r) {
            const int e4 = tid + 32 * kProducers * r;       // float4 index inside the tile: 32 per dimension
            const int k = e4 >> 5, c4 = e4 & 31;
            __pipeline_memcpy_async(B + e4 * 4, Pt + k * npad + tile0 + c4 * 4, 16);
        }
        if (tid < kT / 4) __pipeline_memcpy_async(Bn + (t & 1) * kT + tid * 4, nlo32 + tile0 + tid * 4, 16);
        __pipeline_commit();
    };
    if (producer) {
        fetch(0);
        __pipeline_wait_prior(0);
    }
    __syncthreads();

    const KnnLists S{list_d, list_j, Tq, nlo64, K};

    // iteration t: producers filter tile t into queue[t & 1]; consumers settle queue[(t - 1) & 1] (tile t - 1)
    long long c_work = 0, c_wait = 0, c_cand = 0, c_redo = 0;      // (probe 4: cycle accounting of warp 0 / warp 8)
    for (int64_t t = 0; t <= n_tiles; ++t) {
        const int buf = static_cast<int>(t & 1);
        const long long t_begin = DBG ? clock64() : 0;
        if (producer) {
            if (t < n_tiles) {
                if (t + 1 < n_tiles) fetch(t + 1);          // the other buffer: last read in iteration t - 1
                const float* B = Bs + buf * DP * kT;
                // ---- 8 x 8 dot products per thread
                float acc[8][8];
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
#pragma unroll UNROLL
                for (int k = 0; k < DP; ++k) {
                    const float4 a0 = *reinterpret_cast<const float4*>(As + k * kT + ty * 4);
                    const float4 a1 = *reinterpret_cast<const float4*>(As + k * kT + 64 + ty * 4);
                    const float4 b0 = *reinterpret_cast<const float4*>(B + k * kT + tx * 4);
                    const float4 b1 = *reinterpret_cast<const float4*>(B + k * kT + 64 + tx * 4);
                    const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                    const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
                }
                // ---- filter: lower bound of the squared distance against the query's current K-th distance
                const float4 n0 = *reinterpret_cast<const float4*>(Bn + buf * kT + tx * 4);
                const float4 n1 = *reinterpret_cast<const float4*>(Bn + buf * kT + 64 + tx * 4);
                const float bn[8] = {n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, n1.z, n1.w};
                float tq[8];
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    tq[i] = *reinterpret_cast<volatile float*>(Tq + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4)));
                bool any = false;                      // the common case: nothing passes, one branch for the 64 pairs
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) any |= fmaf(-2.f, acc[i][j], bn[j]) <= tq[i];
                if (any && probe != 1) {      // (probe 1: timing of the bare filter loop, results are meaningless)
                    // which of the 64 pairs: two 32-bit masks built without branches, then one push per set bit
                    // (64 separately guarded pushes cost ~3000 cycles of convergence barriers per warp and tile)
                    unsigned m0 = 0u, m1 = 0u;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            m0 |= (fmaf(-2.f, acc[i][j], bn[j]) <= tq[i] ? 1u : 0u) << (i * 8 + j);
                            m1 |= (fmaf(-2.f, acc[i + 4][j], bn[j]) <= tq[i + 4] ? 1u : 0u) << (i * 8 + j);
                        }
#pragma unroll 1
                    for (int half = 0; half < 2; ++half) {
                        unsigned m = half ? m1 : m0;
                        while (m) {
                            const int bit = __ffs(m) - 1;
                            m &= m - 1;
                            const int q = half * 64 + ty * 4 + (bit >> 3);
                            const int jj = bit & 7;
                            const int p = (jj >> 2) * 64 + tx * 4 + (jj & 3);
                            const int owner = q / kConsQueries;
                            const int pos = atomicAdd(&qcnt[buf * kConsumers + owner], 1);
                            if (pos < kQueueSlots) queue[(buf * kConsumers + owner) * kQueueSlots + pos] = static_cast<unsigned short>((q << 7) | p);
                            else atomicOr(&redo[buf * kConsumers + owner], 1u << (q % kConsQueries));     // no room: re-check the whole tile for q
                        }
                    }
                }
                __pipeline_wait_prior(0);
            }
        } else if (t > 0) {
            // ---- consumer warp: exact float64 distances, 32 candidates at a time, then serial insertion
            const int cw = warp - kProducers;
            const int pb = buf ^ 1;
            const int64_t tile0 = (t - 1) * kT;
            const int ne = min(qcnt[pb * kConsumers + cw], kQueueSlots);
            const unsigned flagged = redo[pb * kConsumers + cw];
            if (DBG) {
                c_cand += ne;
                c_redo += __popc(flagged);
            }
            const unsigned short* Q = queue + (pb * kConsumers + cw) * kQueueSlots;
            for (int base = 0; base < ne; base += 32) {
                const int e = base + lane;
                const bool have = e < ne;
                const int code = have ? Q[e] : (cw * kConsQueries) << 7;
                const int q = code >> 7, p = code & (kT - 1);
                const int64_t i = qbase + q, j = tile0 + p;
                const bool ok = have && j < n && j != i && !((flagged >> (q % kConsQueries)) & 1u);
                double d2 = CUDART_INF;
                if (ok) d2 = exact_d2<DP>(qrows + q * qstride, P64p + j * DP);
                knn_offer(S, ok, q, static_cast<int>(j), d2, lane);
            }
            // queries whose candidates did not all fit the queue: every point of the tile, exactly
            for (unsigned f = flagged; f; f &= f - 1) {
                const int q = cw * kConsQueries + (__ffs(f) - 1);
                const int64_t i = qbase + q;
                for (int p = lane; p < kT; p += 32) {
                    const int64_t j = tile0 + p;
                    const bool ok = j < n && j != i;
                    double d2 = CUDART_INF;
                    if (ok) d2 = exact_d2<DP>(qrows + q * qstride, P64p + j * DP);
                    knn_offer(S, ok, q, static_cast<int>(j), d2, lane);
                }
            }
            __syncwarp();
            if (lane == 0) {
                qcnt[pb * kConsumers + cw] = 0;
                redo[pb * kConsumers + cw] = 0u;
            }
        }
        const long long t_mid = DBG ? clock64() : 0;
        __syncthreads();
        if (DBG) {
            c_work += t_mid - t_begin;
            c_wait += clock64() - t_mid;
        }
    }
    if (DBG && lane == 0) {
        if (warp == 0) { atomicAdd(dbg + 0, static_cast<unsigned long long>(c_work)); atomicAdd(dbg + 1, static_cast<unsigned long long>(c_wait)); }
        if (warp == kProducers) { atomicAdd(dbg + 2, static_cast<unsigned long long>(c_work)); atomicAdd(dbg + 3, static_cast<unsigned long long>(c_wait)); }
        if (!producer) { atomicAdd(dbg + 4, static_cast<unsigned long long>(c_cand)); atomicAdd(dbg + 5, static_cast<unsigned long long>(c_redo)); }
        if (warp == 0) atomicAdd(dbg + 6, 1ull);
    }

    // ---- rows of the result: sorted by column, distances as square roots (knn.rs:27, :65)
    if (!producer) {
        const int cw = warp - kProducers;
        for (int q = cw * kConsQueries; q < (cw + 1) * kConsQueries; ++q) {
            if (qlocal0 + q >= nq || list_j[q * K + K - 1] == 0x7fffffff) continue;
            for (int s = lane; s < K; s += 32) {
                const int j = list_j[q * K + s];
                int rank = 0;
                for (int u = 0; u < K; ++u) rank += list_j[q * K + u] < j;
                out_j[(qlocal0 + q) * K + rank] = j;
                out_d[(qlocal0 + q) * K + rank] = sqrt(list_d[q * K + s]);
            }
        }
    }
}

// ---- tensor-core variant of the filter (the default where it fits; SNAPB200_KNN_MMA=0 turns it off) ----------
// The dot products of a tile through `mma.sync.m16n8k8` on tf32 operands, three passes per product
// (lo*hi + hi*lo + hi*hi with a = hi + lo split in the prep kernel), fp32 accumulation.  The hardware's
// accumulation order and rounding are not documented, so the norms are shrunk by 2^-14 instead of (DP + 8) 2^-24:
// ~40x what the split (2^-21 relative) and ~100 fp32 accumulations can lose, still ~1 % of a typical k-th
// distance.  Everything behind the filter (queue, exact float64 stage, lists) is the code above, unchanged.
// A warp owns 64 queries x 32 points of the tile: 4 x 4 fragments of 16 x 8, 64 accumulator registers.
struct KnnSmemM {
    int nlo64, list_d, Af, Bf, Bn, Tq, qcnt, redo, list_j, queue, total;
};

inline KnnSmemM knn_smem_layout_mma(int DP, int K) {
    KnnSmemM s;
    int o = 0;
    auto take = [&](int bytes) { const int at = o; o += (bytes + 15) & ~15; return at; };
    s.nlo64 = take(kT * 8);
    s.list_d = take(kT * K * 8);
    s.Af = take(DP * kT * 2 * 4);          // [k / 8][query / 16][lane][hi a0..a3, lo a0..a3]
    s.Bf = take(2 * DP * kT * 2 * 4);      // two buffers, each a verbatim copy of a tile of Ptf
    s.Bn = take(2 * kT * 4);
    s.Tq = take(kT * 4);
    s.qcnt = take(2 * kConsumers * 4);
    s.redo = take(2 * kConsumers * 4);
    s.list_j = take(kT * K * 4);
    s.queue = take(2 * kConsumers * kQueueSlots * 2);
    s.total = o;
    return s;
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int DP, bool DBG>
__global__ void __launch_bounds__(kScanThreads, 1)
knn_scan_mma_kernel(const float* __restrict__ Ptf, const double* __restrict__ nrm64,
                    const float* __restrict__ nlo32, const double* __restrict__ P64p, int64_t n, int64_t npad, int64_t q0,
                    int64_t nq, int K, double shrink, KnnSmemM L, int32_t* __restrict__ out_j, double* __restrict__ out_d,
                    unsigned long long* __restrict__ dbg) {
    extern __shared__ __align__(16) unsigned char smem[];
    double* nlo64 = reinterpret_cast<double*>(smem + L.nlo64);
    double* list_d = reinterpret_cast<double*>(smem + L.list_d);
    float* Af = reinterpret_cast<float*>(smem + L.Af);
    float* Bf = reinterpret_cast<float*>(smem + L.Bf);
    float* Bn = reinterpret_cast<float*>(smem + L.Bn);
    float* Tq = reinterpret_cast<float*>(smem + L.Tq);
    int* qcnt = reinterpret_cast<int*>(smem + L.qcnt);
    unsigned* redo = reinterpret_cast<unsigned*>(smem + L.redo);
    int* list_j = reinterpret_cast<int*>(smem + L.list_j);
    unsigned short* queue = reinterpret_cast<unsigned short*>(smem + L.queue);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool producer = warp < kProducers;
    const int64_t qlocal0 = static_cast<int64_t>(blockIdx.x) * kT;
    const int64_t qbase = q0 + qlocal0;

    for (int e = tid; e < DP * kT; e += kScanThreads) {
        const int k = e >> 7, q = e & (kT - 1);
        const int64_t i = qbase + q;
        const bool valid = qlocal0 + q < nq && i < n;
        // A fragment of mma.m16n8k8 (row = query, col = k): a0 (g, t), a1 (g + 8, t), a2 (g, t + 4), a3 (g + 8, t + 4)
        const int kk = k & 7, qq = q & 15;
        const int dst = ((((k >> 3) * 8 + (q >> 4)) * 32 + (qq & 7) * 4 + (kk & 3)) * 8) + (kk >> 2) * 2 + (qq >> 3);
        const int64_t src = knn_frag_index<DP>(valid ? i : 0, k);
        Af[dst] = valid ? Ptf[src] : 0.f;
        Af[dst + 4] = valid ? Ptf[src + 2] : 0.f;
    }
    if (tid < kT) {
        const int64_t i = qbase + tid;
        const bool valid = qlocal0 + tid < nq && i < n;
        nlo64[tid] = valid ? nrm64[i] * shrink : 0.0;
        Tq[tid] = valid ? CUDART_INF_F : -CUDART_INF_F;
    }
    if (tid < 2 * kConsumers) { qcnt[tid] = 0; redo[tid] = 0u; }
    for (int e = tid; e < kT * K; e += kScanThreads) {
        list_d[e] = CUDART_INF;
        list_j[e] = 0x7fffffff;
    }
    const double* qrows = P64p + qbase * DP;

    const int64_t n_tiles = npad / kT;
    constexpr int kTileFloats = DP * kT * 2;
    auto fetch = [&](int64_t t) {
        const int64_t tile0 = t * kT;
        float* dst = Bf + (t & 1) * kTileFloats;
        const float* src = Ptf + t * kTileFloats;
#pragma unroll
        for (int r = 0; r < DP / 4; ++r) {
            const int e4 = tid + 32 * kProducers * r;       // 16-byte chunk of the tile: DP * 64 of them
            __pipeline_memcpy_async(dst + e4 * 4, src + e4 * 4, 16);
        }
        if (tid < kT / 4) __pipeline_memcpy_async(Bn + (t & 1) * kT + tid * 4, nlo32 + tile0 + tid * 4, 16);
        __pipeline_commit();
    };
    if (producer) {
        fetch(0);
        __pipeline_wait_prior(0);
    }
    __syncthreads();

    const KnnLists S{list_d, list_j, Tq, nlo64, K};
    const int g = lane >> 2, t4 = lane & 3;            // fragment coordinates of this lane
    const int qw = (warp & 1) * 64, pw = ((warp >> 1) & 3) * 32;

    long long c_work = 0, c_wait = 0, c_cand = 0, c_redo = 0;
    for (int64_t t = 0; t <= n_tiles; ++t) {
        const int buf = static_cast<int>(t & 1);
        const long long t_begin = DBG ? clock64() : 0;
        if (producer) {
            if (t < n_tiles) {
                if (t + 1 < n_tiles) fetch(t + 1);
                const float* bf = Bf + buf * kTileFloats;
                float c[4][4][4];
#pragma unroll
                for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                        for (int r = 0; r < 4; ++r) c[mt][nt][r] = 0.f;
#pragma unroll
                for (int ks = 0; ks < DP / 8; ++ks) {
                    unsigned bhi[4][2], blo[4][2];
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) {
                        const float4 v = *reinterpret_cast<const float4*>(bf + ((ks * 16 + (pw >> 3) + nt) * 32 + lane) * 4);
                        bhi[nt][0] = __float_as_uint(v.x);
                        bhi[nt][1] = __float_as_uint(v.y);
                        blo[nt][0] = __float_as_uint(v.z);
                        blo[nt][1] = __float_as_uint(v.w);
                    }
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt) {
                        const float* ap = Af + ((ks * 8 + (qw >> 4) + mt) * 32 + lane) * 8;
                        const float4 h = *reinterpret_cast<const float4*>(ap);
                        const float4 l = *reinterpret_cast<const float4*>(ap + 4);
                        const unsigned ahi[4] = {__float_as_uint(h.x), __float_as_uint(h.y), __float_as_uint(h.z), __float_as_uint(h.w)};
                        const unsigned alo[4] = {__float_as_uint(l.x), __float_as_uint(l.y), __float_as_uint(l.z), __float_as_uint(l.w)};
#pragma unroll
                        for (int nt = 0; nt < 4; ++nt) {
                            mma_tf32(c[mt][nt], alo, bhi[nt]);
                            mma_tf32(c[mt][nt], ahi, blo[nt]);
                            mma_tf32(c[mt][nt], ahi, bhi[nt]);
                        }
                    }
                }
                // ---- filter.  Fragment element r of (mt, nt): query qw + 16 mt + g + 8 (r >> 1), point pw + 8 nt + 2 t4 + (r & 1)
                float bn[4][2], tq[4][2];
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    const float2 v = *reinterpret_cast<const float2*>(Bn + buf * kT + pw + nt * 8 + 2 * t4);
                    bn[nt][0] = v.x;
                    bn[nt][1] = v.y;
                }
#pragma unroll
                for (int mt = 0; mt < 4; ++mt) {
                    tq[mt][0] = *reinterpret_cast<volatile float*>(Tq + qw + mt * 16 + g);
                    tq[mt][1] = *reinterpret_cast<volatile float*>(Tq + qw + mt * 16 + g + 8);
                }
                bool any = false;
#pragma unroll
                for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                        for (int r = 0; r < 4; ++r) any |= fmaf(-2.f, c[mt][nt][r], bn[nt][r & 1]) <= tq[mt][r >> 1];
                if (any) {
                    unsigned m0 = 0u, m1 = 0u;        // bit (2 (mt & 1) + (r >> 1)) * 8 + 2 nt + (r & 1); m0: mt < 2, m1: mt >= 2
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                            for (int r = 0; r < 4; ++r) {
                                const int bit = (2 * mt + (r >> 1)) * 8 + 2 * nt + (r & 1);
                                m0 |= (fmaf(-2.f, c[mt][nt][r], bn[nt][r & 1]) <= tq[mt][r >> 1] ? 1u : 0u) << bit;
                                m1 |= (fmaf(-2.f, c[mt + 2][nt][r], bn[nt][r & 1]) <= tq[mt + 2][r >> 1] ? 1u : 0u) << bit;
                            }
#pragma unroll 1
                    for (int half = 0; half < 2; ++half) {
                        unsigned m = half ? m1 : m0;
                        while (m) {
                            const int bit = __ffs(m) - 1;
                            m &= m - 1;
                            const int row = bit >> 3, col = bit & 7;
                            const int q = qw + (2 * half + (row >> 1)) * 16 + g + 8 * (row & 1);
                            const int p = pw + (col >> 1) * 8 + 2 * t4 + (col & 1);
                            const int owner = q / kConsQueries;
                            const int pos = atomicAdd(&qcnt[buf * kConsumers + owner], 1);
                            if (pos < kQueueSlots) queue[(buf * kConsumers + owner) * kQueueSlots + pos] = static_cast<unsigned short>((q << 7) | p);
                            else atomicOr(&redo[buf * kConsumers + owner], 1u << (q % kConsQueries));
                        }
                    }
                }
                __pipeline_wait_prior(0);
            }
        } else if (t > 0) {
            const int cw = warp - kProducers;
            const int pb = buf ^ 1;
            const int64_t tile0 = (t - 1) * kT;
            const int ne = min(qcnt[pb * kConsumers + cw], kQueueSlots);
            const unsigned flagged = redo[pb * kConsumers + cw];
            if (DBG) {
                c_cand += ne;
                c_redo += __popc(flagged);
            }
            const unsigned short* Q = queue + (pb * kConsumers + cw) * kQueueSlots;
            for (int base = 0; base < ne; base += 32) {
                const int e = base + lane;
                const bool have = e < ne;
                const int code = have ? Q[e] : (cw * kConsQueries) << 7;
                const int q = code >> 7, p = code & (kT - 1);
                const int64_t i = qbase + q, j = tile0 + p;
                const bool ok = have && j < n && j != i && !((flagged >> (q % kConsQueries)) & 1u);
                double d2 = CUDART_INF;
                if (ok) d2 = exact_d2<DP>(qrows + q * DP, P64p + j * DP);
                knn_offer(S, ok, q, static_cast<int>(j), d2, lane);
            }
            for (unsigned f = flagged; f; f &= f - 1) {
                const int q = cw * kConsQueries + (__ffs(f) - 1);
                const int64_t i = qbase + q;
                for (int p = lane; p < kT; p += 32) {
                    const int64_t j = tile0 + p;
                    const bool ok = j < n && j != i;
                    double d2 = CUDART_INF;
                    if (ok) d2 = exact_d2<DP>(qrows + q * DP, P64p + j * DP);
                    knn_offer(S, ok, q, static_cast<int>(j), d2, lane);
                }
            }
            __syncwarp();
            if (lane == 0) {
                qcnt[pb * kConsumers + cw] = 0;
                redo[pb * kConsumers + cw] = 0u;
            }
        }
        const long long t_mid = DBG ? clock64() : 0;
        __syncthreads();
        if (DBG) {
            c_work += t_mid - t_begin;
            c_wait += clock64() - t_mid;
        }
    }
    if (DBG && lane == 0) {
        if (warp == 0) { atomicAdd(dbg + 0, static_cast<unsigned long long>(c_work)); atomicAdd(dbg + 1, static_cast<unsigned long long>(c_wait)); }
        if (warp == kProducers) { atomicAdd(dbg + 2, static_cast<unsigned long long>(c_work)); atomicAdd(dbg + 3, static_cast<unsigned long long>(c_wait)); }
        if (!producer) { atomicAdd(dbg + 4, static_cast<unsigned long long>(c_cand)); atomicAdd(dbg + 5, static_cast<unsigned long long>(c_redo)); }
        if (warp == 0) atomicAdd(dbg + 6, 1ull);
    }

    if (!producer) {
        const int cw = warp - kProducers;
        for (int q = cw * kConsQueries; q < (cw + 1) * kConsQueries; ++q) {
            if (qlocal0 + q >= nq || list_j[q * K + K - 1] == 0x7fffffff) continue;
            for (int s = lane; s < K; s += 32) {
                const int j = list_j[q * K + s];
                int rank = 0;
                for (int u = 0; u < K; ++u) rank += list_j[q * K + u] < j;
                out_j[(qlocal0 + q) * K + rank] = j;
                out_d[(qlocal0 + q) * K + rank] = sqrt(list_d[q * K + s]);
            }
        }
    }
}

template <int DP>
void knn_run(snapb200_ctx* c, const double* P64, int64_t n, int64_t npad, int d, int64_t q0, int64_t nq, int K,
             float* Pt, double* nrm64, float* nlo32, double* P64p, double* colsum, int32_t* out_j, double* out_d) {
    cudaStream_t st = c->stream;
    DevBuf<int> bad;
    bad.alloc(1);
    SB_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), st));
    constexpr int kSmemMax = 232448;
    const int probe = getenv("SNAPB200_KNN_PROBE") ? atoi(getenv("SNAPB200_KNN_PROBE")) : 0;
    // tensor-core filter (default; SNAPB200_KNN_MMA=0 selects the FFMA filter): at most 32 dimensions and lists
    // that leave room for the split operand tiles (K <= 74 at DP = 32), otherwise the FFMA filter
    const KnnSmemM LM = knn_smem_layout_mma(DP, K);
    const bool mma_off = getenv("SNAPB200_KNN_MMA") != nullptr && atoi(getenv("SNAPB200_KNN_MMA")) == 0;
    const bool use_mma = DP <= 32 && LM.total <= kSmemMax && !mma_off && probe != 1;
    DevBuf<float> Ptf;
    if (use_mma) Ptf.alloc(npad * DP * 2);
    const double shrink = use_mma ? 1.0 - 6.103515625e-05                                           // 2^-14
                                  : 1.0 - static_cast<double>(DP + 8) * 5.9604644775390625e-08;     // (DP + 8) * 2^-24
    SB_CUDA(cudaMemsetAsync(colsum, 0, sizeof(double) * kKnnMaxDim, st));
    knn_colsum_kernel<DP><<<c->num_sms * 2, 256, 0, st>>>(P64, n, d, colsum);
    SB_LAUNCH_CHECK();
    knn_prep_kernel<DP><<<static_cast<unsigned>(ceil_div(npad, 256)), 256, 0, st>>>(P64, n, npad, d, colsum, shrink, Pt, nrm64, nlo32, P64p, bad.p,
                                                                                     Ptf.p);
    SB_LAUNCH_CHECK();
    int hbad = 0;
    SB_CUDA(cudaMemcpyAsync(&hbad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    SB_CHECK(hbad == 0, "knn: the points contain a non-finite coordinate (or span more than float32 can hold)");
    bool stage_q = true;
    KnnSmem L = knn_smem_layout(DP, K, true);
    if (L.total > kSmemMax) {
        stage_q = false;
        L = knn_smem_layout(DP, K, false);
    }
    SB_CHECK(L.total <= kSmemMax, "knn: n_neighbors does not fit the shared-memory lists (at most 100; 74 with more than 32 dimensions)");
    DevBuf<unsigned long long> dbg;
    const unsigned grid = static_cast<unsigned>(ceil_div(nq, kT));
    static const int unroll = getenv("SNAPB200_KNN_UNROLL") ? atoi(getenv("SNAPB200_KNN_UNROLL")) : 8;
    auto launch = [&](auto kernel, unsigned long long* dbg_ptr) {
        SB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
        kernel<<<grid, kScanThreads, L.total, st>>>(Pt, nrm64, nlo32, P64p, n, npad, q0, nq, K, shrink, L, out_j, out_d, probe,
                                                    stage_q ? 1 : 0, dbg_ptr);
    };
    if (probe == 4) {
        dbg.alloc(8);
        SB_CUDA(cudaMemsetAsync(dbg.p, 0, 64, st));
    }
    if constexpr (DP <= 32) {
        if (use_mma) {
            auto launch_mma = [&](auto kernel, unsigned long long* dbg_ptr) {
                SB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LM.total));
                kernel<<<grid, kScanThreads, LM.total, st>>>(Ptf.p, nrm64, nlo32, P64p, n, npad, q0, nq, K, shrink, LM, out_j, out_d, dbg_ptr);
            };
            if (probe == 4) launch_mma(knn_scan_mma_kernel<DP, true>, dbg.p);
            else launch_mma(knn_scan_mma_kernel<DP, false>, nullptr);
        }
    }
    if (use_mma) {
        stage_q = false;
    } else if (probe == 4) {
        launch(knn_scan_kernel<DP, true, 4>, dbg.p);
    } else if (unroll == 4) {
        launch(knn_scan_kernel<DP, false, 4>, nullptr);
    } else {
        launch(knn_scan_kernel<DP, false, 8>, nullptr);     // 3 % faster than 4 at 1M points (profiles/README.md)
    }
    SB_LAUNCH_CHECK();
    if (probe == 4) {
        unsigned long long h[8];
        SB_CUDA(cudaMemcpyAsync(h, dbg.p, 64, cudaMemcpyDeviceToHost, st));
        SB_CUDA(cudaStreamSynchronize(st));
        const double ctas = static_cast<double>(h[6]), tiles = ctas * static_cast<double>(npad / kT + 1);
        fprintf(stderr, "[snapb200] knn probe: %.0f CTAs x %lld tiles; per CTA-tile cycles: producer work %.0f wait %.0f | consumer(0) work %.0f wait %.0f | "
                "queued candidates %.2f, flagged queries %.4f (stage_q=%d)\n", ctas, static_cast<long long>(npad / kT), h[0] / tiles, h[1] / tiles,
                h[2] / tiles, h[3] / tiles, h[4] / tiles, h[5] / tiles, stage_q ? 1 : 0);
    }
    c->stats.knn_mma = use_mma ? 1 : 0;
    count_launch(c, 3);
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
    DevBuf<double> P64, P64p, nrm64, colsum, out_d;
    DevBuf<float> Pt, nlo32;
    DevBuf<int32_t> out_j;
    const double* Pdev = points;
    if (!on_device) {
        P64.alloc(n * d);
        SB_CUDA(cudaMemcpyAsync(P64.p, points, sizeof(double) * n * d, cudaMemcpyHostToDevice, st));
        Pdev = P64.p;
    }
    const int DP = d <= 8 ? 8 : d <= 16 ? 16 : d <= 32 ? 32 : 64;
    Pt.alloc(npad * DP);
    P64p.alloc(npad * DP);
    nrm64.alloc(npad);
    nlo32.alloc(npad);
    colsum.alloc(kKnnMaxDim);
    out_j.alloc(nq * K);
    out_d.alloc(nq * K);
    SB_CUDA(cudaEventRecord(c->ev0, st));
    switch (DP) {
        case 8: knn_run<8>(c, Pdev, n, npad, d, q0, nq, K, Pt.p, nrm64.p, nlo32.p, P64p.p, colsum.p, out_j.p, out_d.p); break;
        case 16: knn_run<16>(c, Pdev, n, npad, d, q0, nq, K, Pt.p, nrm64.p, nlo32.p, P64p.p, colsum.p, out_j.p, out_d.p); break;
        case 32: knn_run<32>(c, Pdev, n, npad, d, q0, nq, K, Pt.p, nrm64.p, nlo32.p, P64p.p, colsum.p, out_j.p, out_d.p); break;
        default: knn_run<64>(c, Pdev, n, npad, d, q0, nq, K, Pt.p, nrm64.p, nlo32.p, P64p.p, colsum.p, out_j.p, out_d.p); break;
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

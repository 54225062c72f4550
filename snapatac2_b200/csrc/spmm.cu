// The matrix-free operator  Y = X~ (X~^T V) - dinv .* V  on a block of b vectors
// (reference: the closure f(v) at snapatac2-python/src/embedding.rs:162-163,
// applied one vector at a time by ARPACK; here b vectors per sweep).
//
// Factored form:  X~ = diag(r) P diag(w)  with P the stored
// pattern (times the raw values when the matrix is not binarised),
// r_i = sqrt(dinv_i)/rho_i.  Hence
//   pass 1:  W  = w^2 .* ( P^T (r .* V) )      gather over the feature-major copy
//   (all-reduce of W over the row shards)
//   pass 2:  Y  = r .* ( P W ) - dinv .* V      gather over the cell-major rows
// Both passes are gathers (no atomics), so results are bitwise reproducible.
// The only large traffic is the int32 index stream: 4 B/nnz per pass.
#include "ctx.cuh"
#include "dense.cuh"

#include <algorithm>

namespace snapb {

namespace {

template <int B>
__global__ void scale_rows_kernel(const float* __restrict__ V, int64_t ldv, const float* __restrict__ r, int64_t n,
                                  float* __restrict__ out) {
    int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= n * B) return;
    int64_t i = t / B;
    int k = static_cast<int>(t - i * B);
    out[t] = r[i] * V[i * ldv + k];
}

// out[row, 0:B] = rowscale[row] * sum_p val_p * in[idx_p, 0:B]  - subscale[row] * sub[row, 0:B]
// one warp per row; `in` is packed (leading dimension B).
template <int B, bool HAS_VAL, bool HAS_SUB>
__global__ void __launch_bounds__(256)
gather_rows_kernel(const int64_t* __restrict__ ptr, const int32_t* __restrict__ idx, const float* __restrict__ val,
                   const float* __restrict__ in, const float* __restrict__ rowscale, int64_t nrows,
                   float* __restrict__ out, int64_t ldo, const float* __restrict__ subscale,
                   const float* __restrict__ sub, int64_t lds) {
    constexpr int Q = B / 4;  // float4 per dense row
    const float4* __restrict__ in4 = reinterpret_cast<const float4*>(in);
    int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;

    for (int64_t row = warp; row < nrows; row += nwarps) {
        const int64_t s = ptr[row], e = ptr[row + 1];
        float4 acc[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);

        int64_t p = s + lane;
        // 4 independent gathers in flight per lane
        for (; p + 96 < e; p += 128) {
            int j[4];
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                j[u] = ld_stream_int(idx + p + 32 * u);
                v[u] = HAS_VAL ? ld_stream_float(val + p + 32 * u) : 1.f;
            }
            float4 x[4][Q];
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int q = 0; q < Q; ++q) x[u][q] = __ldg(in4 + static_cast<int64_t>(j[u]) * Q + q);
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    if (HAS_VAL) {
                        acc[q].x = fmaf(v[u], x[u][q].x, acc[q].x); acc[q].y = fmaf(v[u], x[u][q].y, acc[q].y);
                        acc[q].z = fmaf(v[u], x[u][q].z, acc[q].z); acc[q].w = fmaf(v[u], x[u][q].w, acc[q].w);
                    } else {
                        acc[q].x += x[u][q].x; acc[q].y += x[u][q].y; acc[q].z += x[u][q].z; acc[q].w += x[u][q].w;
                    }
                }
        }
        for (; p < e; p += 32) {
            int j = ld_stream_int(idx + p);
            float v = HAS_VAL ? ld_stream_float(val + p) : 1.f;
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                float4 x = __ldg(in4 + static_cast<int64_t>(j) * Q + q);
                acc[q].x = fmaf(v, x.x, acc[q].x); acc[q].y = fmaf(v, x.y, acc[q].y);
                acc[q].z = fmaf(v, x.z, acc[q].z); acc[q].w = fmaf(v, x.w, acc[q].w);
            }
        }
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            acc[q].x = warp_sum(acc[q].x); acc[q].y = warp_sum(acc[q].y);
            acc[q].z = warp_sum(acc[q].z); acc[q].w = warp_sum(acc[q].w);
        }
        if (lane == 0) {
            const float rs = rowscale[row];
            float* o = out + row * ldo;
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                float4 y = make_float4(rs * acc[q].x, rs * acc[q].y, rs * acc[q].z, rs * acc[q].w);
                if (HAS_SUB) {
                    const float ss = subscale[row];
                    const float* sp = sub + row * lds + 4 * q;
                    y.x -= ss * sp[0]; y.y -= ss * sp[1]; y.z -= ss * sp[2]; y.w -= ss * sp[3];
                }
                *reinterpret_cast<float4*>(o + 4 * q) = y;
            }
        }
    }
}

// Shared-memory tiled path (b = 4 or 8): both passes run over the sliced-ELL copies.
template <int B>
void apply_tiled(snapb200_ctx* c, const float* V, int64_t ldv, float* Y, int64_t ldy, cudaEvent_t* evs, bool vr_ready,
                 const float* subscale, const float* sub, int64_t lds) {
    SB_CHECK(ldy == B, "tiled operator: Y must be packed (leading dimension b)");
    const int64_t n = c->n_local, m = c->m;
    cudaStream_t st = c->stream;
    c->Vr.ensure(std::max<int64_t>(1, n * B));
    c->W.ensure(m * B);
    if (n > 0 && !vr_ready) {
        scale_rows_kernel<B><<<static_cast<unsigned>(ceil_div(n * B, 256)), 256, 0, st>>>(V, ldv, c->r.p, n, c->Vr.p);
        SB_LAUNCH_CHECK();
        count_launch(c);
    }
    if (evs) SB_CUDA(cudaEventRecord(evs[0], st));
    // pass 1: W = w^2 .* P^T (r V)
    sell_spmm(c, c->S1, c->Vr.p, c->W.p, c->w2.p, nullptr, nullptr, 0);
    if (evs) SB_CUDA(cudaEventRecord(evs[1], st));
    allreduce_f32(c, c->W.p, m * B);
    if (evs) SB_CUDA(cudaEventRecord(evs[2], st));
    // pass 2: Y = r .* (P W) - dinv .* V
    sell_spmm(c, c->S2, c->W.p, Y, c->r.p, subscale, sub, lds);
    if (evs) SB_CUDA(cudaEventRecord(evs[3], st));
}

inline int grid_rows(snapb200_ctx* c, int64_t nrows) {
    return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(ceil_div(nrows, 8), static_cast<int64_t>(c->num_sms) * 16)));
}

template <int B>
void apply_impl(snapb200_ctx* c, const float* V, int64_t ldv, float* Y, int64_t ldy, cudaEvent_t* evs, bool vr_ready,
                const float* subscale, const float* sub, int64_t lds) {
    const int64_t n = c->n_local, m = c->m;
    cudaStream_t st = c->stream;
    c->Vr.ensure(std::max<int64_t>(1, n * B));
    c->W.ensure(m * B);
    if (n > 0 && !vr_ready) {
        scale_rows_kernel<B><<<static_cast<unsigned>(ceil_div(n * B, 256)), 256, 0, st>>>(V, ldv, c->r.p, n, c->Vr.p);
        SB_LAUNCH_CHECK();
        count_launch(c);
    }
    ensure_xt(c);   // the CSR feature-major copy is built on first use
    if (evs) SB_CUDA(cudaEventRecord(evs[0], st));
    // pass 1: W = w^2 .* P^T (r V)
    if (c->Xt.has_values())
        gather_rows_kernel<B, true, false><<<grid_rows(c, m), 256, 0, st>>>(
            c->Xt.ptr.p, c->Xt.idx.p, c->Xt.val.p, c->Vr.p, c->w2.p, m, c->W.p, B, nullptr, nullptr, 0);
    else
        gather_rows_kernel<B, false, false><<<grid_rows(c, m), 256, 0, st>>>(
            c->Xt.ptr.p, c->Xt.idx.p, nullptr, c->Vr.p, c->w2.p, m, c->W.p, B, nullptr, nullptr, 0);
    SB_LAUNCH_CHECK();
    count_launch(c);
    if (evs) SB_CUDA(cudaEventRecord(evs[1], st));
    allreduce_f32(c, c->W.p, m * B);
    if (evs) SB_CUDA(cudaEventRecord(evs[2], st));
    // pass 2: Y = r .* (P W) - dinv .* V
    if (n > 0) {
        if (c->X.has_values())
            gather_rows_kernel<B, true, true><<<grid_rows(c, n), 256, 0, st>>>(
                c->X.ptr.p, c->X.idx.p, c->X.val.p, c->W.p, c->r.p, n, Y, ldy, subscale, sub, lds);
        else
            gather_rows_kernel<B, false, true><<<grid_rows(c, n), 256, 0, st>>>(
                c->X.ptr.p, c->X.idx.p, nullptr, c->W.p, c->r.p, n, Y, ldy, subscale, sub, lds);
        SB_LAUNCH_CHECK();
        count_launch(c);
    }
    if (evs) SB_CUDA(cudaEventRecord(evs[3], st));
}

}  // namespace

// one view's two passes:  Y = r .* (P (w^2 .* P^T (r .* V))) - subscale .* sub
static void apply_view(snapb200_ctx* c, const float* V, int64_t ldv, float* Y, int64_t ldy, int b, cudaEvent_t* evs,
                       bool vr_ready, const float* subscale, const float* sub, int64_t lds) {
    if (use_tiled(c, b)) {
        ensure_tiled(c, b);   // no-op when prepare() already built the copies for this width
        c->stats.spmm_tiled = 1;
        if (b == 8) apply_tiled<8>(c, V, ldv, Y, ldy, evs, vr_ready, subscale, sub, lds);
        else apply_tiled<4>(c, V, ldv, Y, ldy, evs, vr_ready, subscale, sub, lds);
        return;
    }
    c->stats.spmm_tiled = 0;
    switch (b) {
        case 4: apply_impl<4>(c, V, ldv, Y, ldy, evs, vr_ready, subscale, sub, lds); break;
        case 8: apply_impl<8>(c, V, ldv, Y, ldy, evs, vr_ready, subscale, sub, lds); break;
        case 16: apply_impl<16>(c, V, ldv, Y, ldy, evs, vr_ready, subscale, sub, lds); break;
        default: throw Error("operator: block width must be 4, 8 or 16");
    }
}

void operator_apply_dev(snapb200_ctx* c, const float* V, int64_t ldv, float* Y, int64_t ldy, int b, cudaEvent_t* evs,
                        bool vr_ready) {
    SB_CHECK(c->prepared, "operator: call prepare first");
    SB_CHECK(ldy % 4 == 0 && (reinterpret_cast<uintptr_t>(Y) & 15) == 0, "operator: Y must be 16-byte aligned");
    if (c->views.empty()) {
        apply_view(c, V, ldv, Y, ldy, b, evs, vr_ready, c->dinv.p, V, ldv);
        return;
    }
    // multi-view: Y = sum_v X~_v X~_v^T V - dinv .* V; the first view subtracts dinv .* V, the others
    // accumulate onto Y (subscale = -1, sub = Y)
    if (evs) {
        SB_CUDA(cudaEventRecord(evs[0], c->stream));
    }
    apply_view(c, V, ldv, Y, ldy, b, nullptr, vr_ready, c->dinv.p, V, ldv);
    for (snapb200_ctx* v : c->views) {
        apply_view(v, V, ldv, Y, ldy, b, nullptr, false, c->neg_one.p, Y, ldy);
    }
    if (evs) {
        SB_CUDA(cudaEventRecord(evs[1], c->stream));
        SB_CUDA(cudaEventRecord(evs[2], c->stream));
        SB_CUDA(cudaEventRecord(evs[3], c->stream));
    }
}

// --------------------------------------------------------------------------
// Products with the row-normalised, feature-weighted matrix  Xhat = diag(1/rho) P diag(w)
// on k dense columns (the Nystrom extension, embedding.rs:194-267:  q = sample @ (seed.T @ evecs)).
//   project:    out[n x k] = Xhat   in[m x k]     (needs prepare_projection or prepare)
//   project_t:  out[m x k] = Xhat^T in[n x k]     (needs prepare; all-reduced over the row shards)
// The columns go through the same SpMM kernels as the operator, b at a time.
// --------------------------------------------------------------------------
namespace {

// out[row, 0:B] = scale[row] * in[row * ld + col0 + (0:B)]   (columns past ncols read as zero)
template <int B>
__global__ void pack_scaled_kernel(const float* __restrict__ in, int64_t ld, int col0, int ncols,
                                   const float* __restrict__ scale, int64_t rows, float* __restrict__ out) {
    const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= rows * B) return;
    const int64_t i = t / B;
    const int k = static_cast<int>(t - i * B);
    out[t] = (col0 + k < ncols) ? scale[i] * in[i * ld + col0 + k] : 0.f;
}

__global__ void to_float_kernel(const double* __restrict__ in, float* __restrict__ out, int64_t n, int recip) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = static_cast<float>(recip ? 1.0 / in[i] : in[i]);
}

template <int B>
void product_block(snapb200_ctx* c, bool transposed, const float* in_b, float* out_b) {
    const int64_t n = c->n_local, m = c->m;
    cudaStream_t st = c->stream;
    const bool tiled = use_tiled(c, B) && (transposed ? c->S1.built && c->S1.b == B : c->S2.built && c->S2.b == B);
    if (!transposed) {
        if (n == 0) return;
        if (tiled) {
            sell_spmm(c, c->S2, in_b, out_b, c->rhoinv_f.p, nullptr, nullptr, 0);
        } else if (c->X.has_values()) {
            gather_rows_kernel<B, true, false><<<grid_rows(c, n), 256, 0, st>>>(
                c->X.ptr.p, c->X.idx.p, c->X.val.p, in_b, c->rhoinv_f.p, n, out_b, B, nullptr, nullptr, 0);
        } else {
            gather_rows_kernel<B, false, false><<<grid_rows(c, n), 256, 0, st>>>(
                c->X.ptr.p, c->X.idx.p, nullptr, in_b, c->rhoinv_f.p, n, out_b, B, nullptr, nullptr, 0);
        }
    } else {
        if (tiled) {
            sell_spmm(c, c->S1, in_b, out_b, c->w_f.p, nullptr, nullptr, 0);
        } else {
            ensure_xt(c);
            if (c->Xt.has_values())
                gather_rows_kernel<B, true, false><<<grid_rows(c, m), 256, 0, st>>>(
                    c->Xt.ptr.p, c->Xt.idx.p, c->Xt.val.p, in_b, c->w_f.p, m, out_b, B, nullptr, nullptr, 0);
            else
                gather_rows_kernel<B, false, false><<<grid_rows(c, m), 256, 0, st>>>(
                    c->Xt.ptr.p, c->Xt.idx.p, nullptr, in_b, c->w_f.p, m, out_b, B, nullptr, nullptr, 0);
        }
        allreduce_f32(c, out_b, m * B);
    }
    SB_LAUNCH_CHECK();
    count_launch(c);
}

template <int B>
void product_impl(snapb200_ctx* c, bool transposed, const float* in_host, int k, float* out_host) {
    const int64_t n = c->n_local, m = c->m;
    const int64_t rin = transposed ? n : m, rout = transposed ? m : n;
    cudaStream_t st = c->stream;
    DevBuf<float> din, dout, bin, bout;
    din.alloc(std::max<int64_t>(1, rin * k));
    dout.alloc(std::max<int64_t>(1, rout * k));
    bin.alloc(std::max<int64_t>(1, rin * B));
    bout.alloc(std::max<int64_t>(1, rout * B));
    if (rin > 0) SB_CUDA(cudaMemcpyAsync(din.p, in_host, sizeof(float) * rin * k, cudaMemcpyHostToDevice, st));
    const float* in_scale = transposed ? c->rhoinv_f.p : c->w_f.p;
    for (int j0 = 0; j0 < k; j0 += B) {
        const int nb = std::min(B, k - j0);
        if (rin > 0) {
            pack_scaled_kernel<B><<<static_cast<unsigned>(ceil_div(rin * B, 256)), 256, 0, st>>>(din.p, k, j0, k, in_scale, rin,
                                                                                             bin.p);
            SB_LAUNCH_CHECK();
        }
        product_block<B>(c, transposed, bin.p, bout.p);
        if (rout > 0) copy_cols(c, bout.p, B, dout.p + j0, k, rout, nb);
    }
    if (rout > 0) SB_CUDA(cudaMemcpyAsync(out_host, dout.p, sizeof(float) * rout * k, cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
}

}  // namespace

void prepare_projection(snapb200_ctx* c) {
    SB_CHECK(c->loaded, "prepare_projection: no matrix loaded");
    const int64_t n = c->n_local, m = c->m;
    cudaStream_t st = c->stream;
    if (!c->prepared) {
        decide_spmm_mode(c);
        // IDF (or user) weights and row norms only; no transpose (view_norms leaves them on the device)
        c->w.alloc(m);
        c->rho.alloc(std::max<int64_t>(1, n));
        weights_and_norms(c, c->w.p, c->rho.p);
        c->S1.clear();
        c->S2.clear();
        if (use_tiled(c, c->block)) sell_build(c, c->X, c->S2, c->block);
    }
    c->w_f.alloc(m);
    c->rhoinv_f.alloc(std::max<int64_t>(1, n));
    to_float_kernel<<<static_cast<unsigned>(ceil_div(m, 256)), 256, 0, st>>>(c->w.p, c->w_f.p, m, 0);
    SB_LAUNCH_CHECK();
    if (n > 0) {
        to_float_kernel<<<static_cast<unsigned>(ceil_div(n, 256)), 256, 0, st>>>(c->rho.p, c->rhoinv_f.p, n, 1);
        SB_LAUNCH_CHECK();
    }
    SB_CUDA(cudaStreamSynchronize(st));
    c->proj_ready = true;
}

void project(snapb200_ctx* c, bool transposed, const float* in_host, int k, float* out_host) {
    SB_CHECK(k >= 1, "project: k must be positive");
    // (the transposed product runs over the feature-major tiled copy when prepare() built it, otherwise over
    //  a CSR transpose built on first use)
    if (!c->proj_ready) prepare_projection(c);
    // column blocks of the width the tiled copies were built for (8 on the CSR path)
    const int b = (c->S2.built && (c->S2.b == 4 || c->S2.b == 8)) ? c->S2.b : 8;
    if (b == 4) product_impl<4>(c, transposed, in_host, k, out_host);
    else product_impl<8>(c, transposed, in_host, k, out_host);
}

}  // namespace snapb

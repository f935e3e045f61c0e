// clr_pixel_dots: per-pixel contractions over the channel axis,
//     dots[b,q,p] = sum_c op(x[b,c,p], V[q][c]),   op = x*v  or  (x-v)^2,   (+ sumsq[b,p] = sum_c x^2)
// with a fused epilogue.  One read of the feature map.
//
// Users
//   - adjoint of the pooling w.r.t. soft predictions (utils/Utils.py:114-130 under autograd;
//     dL/dw_r[b,p] = sum_c G[r][c] (x[b,c,p] - mu_r[c]))
//   - pixel<->prototype L2 distance / cosine similarity (Trainer_prototype.py:98-116, utils/Utils.py:86-88)
//   - the prototype-guided discriminative hinge (Trainer_prototype_mt bytecode L454-474), whose
//     margin argument d_obj - d_bck is affine in x:  -(2/C) x.(P_obj-P_bck) + (|P_obj|^2-|P_bck|^2)/C.
//
// Bound: HBM.  Algorithmic bytes = 4*B*C*HW (+ 4*B*Q*HW outputs).
//
//   item   = (b, tile of 16*VEC pixels), all C channels -> C rows of 64*VEC bytes
//   warp   = two half-warps on two different channels, 16 lanes x VEC pixels each, so every
//            128-bit load instruction covers 2 channel rows; warp w owns channels {2w, 2w+1} mod 16
//   reduce = lanes l and l+16 by one shuffle, the 8 warps through shared memory; no atomics
//   grid   = persistent with contiguous item ranges
#include "clr_common.cuh"
#include "clr_internal.h"

namespace clr {

enum { DOTS_OP_DOT = 0, DOTS_OP_SQDIFF = 1 };
enum { DOTS_EPI_AFFINE = 0, DOTS_EPI_HINGE = 1, DOTS_EPI_SQRT = 2, DOTS_EPI_COSINE = 3 };

constexpr int kDotsMaxQ = 16;
constexpr int kDotsUnroll = 8;

struct DotsParams {
    const float* feat;
    const float* V;        // [Q][C]
    float* out;            // [B,Q,HW]
    float* sumsq;          // [B,HW] or null
    const float* y;        // hinge: labels [B,Q,HW]
    float* coef;           // hinge: d(loss*npx)/d(delta) [B,Q,HW]
    float* partials;       // hinge: [gridDim.x][1+Q] loss numerator and coefficient sums per CTA
    const float* beta_dev; // optional device-side beta[Q] (overrides beta[])
    const float* vnorm_dev;// cosine: device-side max(|V_0|, eps)
    float alpha[kDotsMaxQ];
    float beta[kDotsMaxQ];
    float margin;
    int B, C, HW, Q, epi;
    int tilesPerSample, total;
};

template <int QT, int VEC, int OP, bool SUMSQ>
__global__ void __launch_bounds__(kThreads, 3) pixel_dots_kernel(const DotsParams p) {
    pdl_wait();
    constexpr int TP = 16 * VEC;                 // pixels per tile
    constexpr int NA = QT + (SUMSQ ? 1 : 0);     // accumulator rows
    extern __shared__ __align__(16) float smem[];
    float* Vs = smem;                            // [QT][C]
    float* red = smem + (size_t)QT * p.C;        // [kWarps][NA][TP]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int half = lane >> 4, pl = lane & 15;

    for (int i = tid; i < QT * p.C; i += kThreads) {
        const int q = i / p.C;
        Vs[i] = (q < p.Q) ? p.V[i] : 0.f;
    }
    __syncthreads();

    int begin, end;
    partition(p.total, gridDim.x, blockIdx.x, begin, end);
    float hinge_loss = 0.f;          // per-thread running sums (hinge epilogue)
    float hinge_n[QT];
#pragma unroll
    for (int q = 0; q < QT; ++q) hinge_n[q] = 0.f;

    for (int it = begin; it < end; ++it) {
        const int b = it / p.tilesPerSample, tile = it - b * p.tilesPerSample;
        const int px = tile * TP + pl * VEC;
        const bool ok = px < p.HW;
        float acc[NA][VEC];
#pragma unroll
        for (int a = 0; a < NA; ++a)
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[a][v] = 0.f;

        const int cbase = warp * 2 + half;
        const float* xb = p.feat + (size_t)b * p.C * p.HW + px;
        for (int c0 = cbase; c0 < p.C; c0 += 16 * kDotsUnroll) {
            Pack<VEC> x[kDotsUnroll];
#pragma unroll
            for (int u = 0; u < kDotsUnroll; ++u) {
                const int c = c0 + 16 * u;
                if (ok && c < p.C) x[u] = ld_stream<VEC>(xb + (size_t)c * p.HW);
                else {
#pragma unroll
                    for (int v = 0; v < VEC; ++v) x[u].v[v] = 0.f;
                }
            }
#pragma unroll
            for (int u = 0; u < kDotsUnroll; ++u) {
                const int c = c0 + 16 * u;
                if (c < p.C) {
#pragma unroll
                    for (int q = 0; q < QT; ++q) {
                        const float vq = Vs[q * p.C + c];
#pragma unroll
                        for (int v = 0; v < VEC; ++v) {
                            if (OP == DOTS_OP_DOT) acc[q][v] = fmaf(x[u].v[v], vq, acc[q][v]);
                            else { const float dlt = x[u].v[v] - vq; acc[q][v] = fmaf(dlt, dlt, acc[q][v]); }
                        }
                    }
                    if (SUMSQ) {
#pragma unroll
                        for (int v = 0; v < VEC; ++v) acc[QT][v] = fmaf(x[u].v[v], x[u].v[v], acc[QT][v]);
                    }
                }
            }
        }
        // half-warps hold different channels of the same pixels
#pragma unroll
        for (int a = 0; a < NA; ++a)
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[a][v] += __shfl_xor_sync(0xffffffffu, acc[a][v], 16);
        __syncthreads();   // previous item's epilogue readers are done with `red`
        if (half == 0) {
#pragma unroll
            for (int a = 0; a < NA; ++a)
#pragma unroll
                for (int v = 0; v < VEC; ++v) red[(warp * NA + a) * TP + pl * VEC + v] = acc[a][v];
        }
        __syncthreads();
        // ---- epilogue: thread e -> (row a, pixel j) --------------------------------------------
        for (int e = tid; e < NA * TP; e += kThreads) {
            const int a = e / TP, j = e - a * TP;
            const int pxo = tile * TP + j;
            if (pxo >= p.HW) continue;
            float s = 0.f;
#pragma unroll
            for (int wq = 0; wq < kWarps; ++wq) s += red[(wq * NA + a) * TP + j];
            if (SUMSQ && a == QT) {
                if (p.sumsq) p.sumsq[(size_t)b * p.HW + pxo] = s;
                continue;
            }
            if (a >= p.Q) continue;
            const size_t o = ((size_t)b * p.Q + a) * p.HW + pxo;
            if (p.epi == DOTS_EPI_AFFINE) {
                p.out[o] = fmaf(p.alpha[a], s, p.beta_dev ? p.beta_dev[a] : p.beta[a]);
            } else if (p.epi == DOTS_EPI_SQRT) {
                p.out[o] = sqrtf(s);
            } else if (p.epi == DOTS_EPI_COSINE) {
                float ss = 0.f;
#pragma unroll
                for (int wq = 0; wq < kWarps; ++wq) ss += red[(wq * NA + (NA - 1)) * TP + j];
                p.out[o] = s / (fmaxf(sqrtf(ss), 1e-8f) * __ldg(p.vnorm_dev));
            } else {  // DOTS_EPI_HINGE
                const float delta = fmaf(p.alpha[a], s, p.beta_dev ? p.beta_dev[a] : p.beta[a]);
                const float yv = p.y[o];
                const float ho = delta + p.margin, hb = p.margin - delta;
                hinge_loss += yv * fmaxf(ho, 0.f) + (1.f - yv) * fmaxf(hb, 0.f);
                const float cf = (ho > 0.f ? yv : 0.f) - (hb > 0.f ? (1.f - yv) : 0.f);
                p.coef[o] = cf;
                if (p.out) p.out[o] = delta;
                // a is not compile-time here; QT is small
#pragma unroll
                for (int q = 0; q < QT; ++q) if (q == a) hinge_n[q] += cf;
            }
        }
    }
    if (p.epi == DOTS_EPI_HINGE) {
        // deterministic per-CTA partial: warp shuffle, then shared memory in warp order
        float v0 = warp_sum(hinge_loss);
        __shared__ float wred[kWarps][1 + kDotsMaxQ];
        if (lane == 0) wred[warp][0] = v0;
#pragma unroll
        for (int q = 0; q < QT; ++q) {
            const float t = warp_sum(hinge_n[q]);
            if (lane == 0) wred[warp][1 + q] = t;
        }
        __syncthreads();
        if (tid <= p.Q) {
            float s = 0.f;
            for (int wq = 0; wq < kWarps; ++wq) s += wred[wq][tid];
            p.partials[(size_t)blockIdx.x * (1 + p.Q) + tid] = s;
        }
    }
}

template <int QT, int VEC, int OP, bool SUMSQ>
static int launch_dots(DotsParams& p, int* grid_out, cudaStream_t st) {
    constexpr int TP = 16 * VEC, NA = QT + (SUMSQ ? 1 : 0);
    auto kern = pixel_dots_kernel<QT, VEC, OP, SUMSQ>;
    const size_t smem = sizeof(float) * ((size_t)QT * p.C + (size_t)kWarps * NA * TP);
    if (smem > (size_t)device_facts().max_smem_optin) return CLR_ERR_UNSUPPORTED;
    CLR_RETURN_IF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CLR_RETURN_IF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, smem));
    if (occ < 1) occ = 1;
    p.tilesPerSample = (p.HW + TP - 1) / TP;
    const long long total = (long long)p.B * p.tilesPerSample;
    if (total > 0x3fffffff) return CLR_ERR_UNSUPPORTED;
    p.total = (int)total;
    int grid = device_facts().sms * occ;
    if (grid > p.total) grid = p.total;
    if (grid_out) {
        // hinge partials are sized by the caller for `*grid_out` CTAs at most
        if (*grid_out > 0 && grid > *grid_out) grid = *grid_out;
        *grid_out = grid;
    }
    clr::launch_k(kern, grid, kThreads, smem, st, p);
    return launch_status();
}

template <int VEC, int OP, bool SUMSQ>
static int dispatch_q(DotsParams& p, int* grid_out, cudaStream_t st) {
    if (p.Q <= 1) return launch_dots<1, VEC, OP, SUMSQ>(p, grid_out, st);
    if (p.Q <= 2) return launch_dots<2, VEC, OP, SUMSQ>(p, grid_out, st);
    if (p.Q <= 4) return launch_dots<4, VEC, OP, SUMSQ>(p, grid_out, st);
    if (p.Q <= 8) return launch_dots<8, VEC, OP, SUMSQ>(p, grid_out, st);
    return launch_dots<16, VEC, OP, SUMSQ>(p, grid_out, st);
}

int pixel_dots_impl(DotsParams& p, int op, bool want_sumsq, int* grid_out, cudaStream_t st) {
    CLR_CHECK_ARG(p.feat && p.V && p.B > 0 && p.C > 0 && p.HW > 0 && p.Q >= 1 && p.Q <= kDotsMaxQ);
    if (!aligned4(p.feat)) return CLR_ERR_ALIGN;
    const bool vec4 = (p.HW % 4 == 0) && aligned16(p.feat);
    if (op == DOTS_OP_DOT) {
        if (want_sumsq) return vec4 ? dispatch_q<4, DOTS_OP_DOT, true>(p, grid_out, st) : dispatch_q<1, DOTS_OP_DOT, true>(p, grid_out, st);
        return vec4 ? dispatch_q<4, DOTS_OP_DOT, false>(p, grid_out, st) : dispatch_q<1, DOTS_OP_DOT, false>(p, grid_out, st);
    }
    return vec4 ? dispatch_q<4, DOTS_OP_SQDIFF, false>(p, grid_out, st) : dispatch_q<1, DOTS_OP_SQDIFF, false>(p, grid_out, st);
}

// g/N tables for the adjoint w.r.t. the weights: V[q][c] and beta[q] = -sum_c V[q][c]*mu[q][c] are
// built on the device so the call stays asynchronous.
__global__ void bwd_w_tables_kernel(const float* __restrict__ g, const float* __restrict__ sums, int K, int C,
                                    int fmt, float scale, float* __restrict__ V, float* __restrict__ beta) {
    pdl_wait();
    // one CTA per output row q
    const int q = blockIdx.x;
    const int R = 2 * K;
    double acc = 0.0;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float v;
        double cst;
        if (fmt == CLR_W_COMPLEMENT) {
            const float No = sums[(size_t)q * (C + 1) + C], Nb = sums[(size_t)(K + q) * (C + 1) + C];
            const float Go = scale * g[(size_t)q * C + c] / No, Gb = scale * g[(size_t)(K + q) * C + c] / Nb;
            v = Go - Gb;
            cst = (double)Go * (sums[(size_t)q * (C + 1) + c] / No) - (double)Gb * (sums[(size_t)(K + q) * (C + 1) + c] / Nb);
        } else {
            const float N = sums[(size_t)q * (C + 1) + C];
            v = scale * g[(size_t)q * C + c] / N;
            cst = (double)v * (sums[(size_t)q * (C + 1) + c] / N);
        }
        V[(size_t)q * C + c] = v;
        acc += cst;
    }
    (void)R;
    acc = warp_sum(acc);
    __shared__ double sh[32];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
        beta[q] = (float)(-t);
    }
}

__global__ void vec_norm_kernel(const float* __restrict__ v, int C, float eps, float* __restrict__ out) {
    pdl_wait();
    double acc = 0.0;
    for (int c = threadIdx.x; c < C; c += blockDim.x) acc += (double)v[c] * v[c];
    acc = warp_sum(acc);
    __shared__ double sh[8];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
        out[0] = fmaxf((float)sqrt(t), eps);
    }
}

// global min / max of a small map (two stages, fixed order), then (x - min) / (max - min) in place
__global__ void __launch_bounds__(256) minmax_partial_kernel(const float* __restrict__ x, size_t n, float* __restrict__ part) {
    pdl_wait();
    float lo = INFINITY, hi = -INFINITY;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = x[i];
        lo = fminf(lo, v); hi = fmaxf(hi, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    __shared__ float sl[8], shh[8];
    if ((threadIdx.x & 31) == 0) { sl[threadIdx.x >> 5] = lo; shh[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) { lo = fminf(lo, sl[i]); hi = fmaxf(hi, shh[i]); }
        part[2 * blockIdx.x] = fminf(lo, sl[0]);
        part[2 * blockIdx.x + 1] = fmaxf(hi, shh[0]);
    }
}
__global__ void __launch_bounds__(256) minmax_apply_kernel(float* __restrict__ x, size_t n, const float* __restrict__ part, int nparts) {
    pdl_wait();
    float lo = INFINITY, hi = -INFINITY;
    for (int i = 0; i < nparts; ++i) { lo = fminf(lo, part[2 * i]); hi = fmaxf(hi, part[2 * i + 1]); }
    const float den = hi - lo;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        x[i] = (x[i] - lo) / den;
}

int disc_fwd_impl(const float* xs, const float* ys, int B, int C, int HW, int K,
                  const float* disc_vec, const float* disc_beta, float margin,
                  float* coef, float* delta, float* partials, int partials_cap, int* nparts, cudaStream_t st) {
    CLR_CHECK_ARG(xs && ys && disc_vec && disc_beta && coef && partials && nparts && partials_cap > 0);
    CLR_CHECK_ARG(K >= 1 && K <= CLR_MAX_K);
    DotsParams p{};
    p.feat = xs; p.V = disc_vec; p.out = delta; p.y = ys; p.coef = coef; p.partials = partials;
    p.beta_dev = disc_beta; p.margin = margin;
    p.B = B; p.C = C; p.HW = HW; p.Q = K; p.epi = DOTS_EPI_HINGE;
    for (int q = 0; q < kDotsMaxQ; ++q) { p.alpha[q] = -2.0f / (float)C; p.beta[q] = 0.f; }
    int grid = partials_cap;
    const int rc = pixel_dots_impl(p, DOTS_OP_DOT, false, &grid, st);
    *nparts = grid;
    return rc;
}

}  // namespace clr

extern "C" {

int clr_disc_partials_cap(void) { return 148 * 16; }

int clr_disc_fwd(const float* xs, const float* ys, int B, int C, int HW, int K,
                 const float* disc_vec, const float* disc_beta, float margin,
                 float* coef, float* delta, float* partials, int partials_cap, int* nparts, clr_stream_t stream) {
    return clr::disc_fwd_impl(xs, ys, B, C, HW, K, disc_vec, disc_beta, margin, coef, delta, partials, partials_cap,
                              nparts, static_cast<cudaStream_t>(stream));
}

int clr_proto_distance(const float* feat, int B, int C, int HW, const float* protos, int Q,
                       float* dist, clr_stream_t stream) {
    if (!dist) return CLR_ERR_BAD_ARG;
    clr::DotsParams p{};
    p.feat = feat; p.V = protos; p.out = dist;
    p.B = B; p.C = C; p.HW = HW; p.Q = Q; p.epi = clr::DOTS_EPI_SQRT;
    return clr::pixel_dots_impl(p, clr::DOTS_OP_SQDIFF, false, nullptr, static_cast<cudaStream_t>(stream));
}

int clr_proto_cosine(const float* feat, int B, int C, int HW, const float* proto, float* ws4, float* out,
                     clr_stream_t stream) {
    if (!out || !ws4 || !proto || C < 1) return CLR_ERR_BAD_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    clr::launch_k(clr::vec_norm_kernel, 1, 256, 0, st, proto, C, 1e-8f, ws4);
    clr::DotsParams p{};
    p.feat = feat; p.V = proto; p.out = out; p.vnorm_dev = ws4;
    p.B = B; p.C = C; p.HW = HW; p.Q = 1; p.epi = clr::DOTS_EPI_COSINE;
    return clr::pixel_dots_impl(p, clr::DOTS_OP_DOT, true, nullptr, st);
}

int clr_minmax_normalize(float* x, size_t n, float* ws /*>= 2*256 floats*/, clr_stream_t stream) {
    if (!x || !ws || n == 0) return CLR_ERR_BAD_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int blocks = (int)((n + 255) / 256 < 256 ? (n + 255) / 256 : 256);
    clr::launch_k(clr::minmax_partial_kernel, blocks, 256, 0, st, x, n, ws);
    clr::launch_k(clr::minmax_apply_kernel, blocks, 256, 0, st, x, n, ws, blocks);
    return clr::launch_status();
}

int clr_pixel_dots(const float* feat, int B, int C, int HW, const float* V, int Q,
                   float* dots, float* sumsq, clr_stream_t stream) {
    if (!dots) return CLR_ERR_BAD_ARG;
    clr::DotsParams p{};
    p.feat = feat; p.V = V; p.out = dots; p.sumsq = sumsq;
    p.B = B; p.C = C; p.HW = HW; p.Q = Q; p.epi = clr::DOTS_EPI_AFFINE;
    for (int q = 0; q < clr::kDotsMaxQ; ++q) { p.alpha[q] = 1.f; p.beta[q] = 0.f; }
    return clr::pixel_dots_impl(p, clr::DOTS_OP_DOT, sumsq != nullptr, nullptr, static_cast<cudaStream_t>(stream));
}

size_t clr_pool_bwd_w_ws_bytes(int C, int K, int fmt) {
    const int Q = fmt == CLR_W_COMPLEMENT ? K : 2 * K;
    return sizeof(float) * ((size_t)Q * C + Q);
}

int clr_pool_bwd_w(const float* feat, int fmt, int B, int C, int HW, int K,
                   const float* g, const float* sums, float scale,
                   void* ws, size_t ws_bytes, float* grad_w, clr_stream_t stream) {
    if (!feat || !g || !sums || !ws || !grad_w || K < 1 || K > CLR_MAX_K) return CLR_ERR_BAD_ARG;
    if (fmt != CLR_W_COMPLEMENT && fmt != CLR_W_EXPLICIT) return CLR_ERR_BAD_ARG;
    const int Q = fmt == CLR_W_COMPLEMENT ? K : 2 * K;
    const size_t need = sizeof(float) * ((size_t)Q * C + Q);
    if (ws_bytes < need) return CLR_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* V = static_cast<float*>(ws);
    float* beta = V + (size_t)Q * C;
    clr::launch_k(clr::bwd_w_tables_kernel, Q, 256, 0, st, g, sums, K, C, fmt, scale, V, beta);
    clr::DotsParams p{};
    p.feat = feat; p.V = V; p.out = grad_w; p.B = B; p.C = C; p.HW = HW; p.Q = Q; p.epi = clr::DOTS_EPI_AFFINE;
    p.beta_dev = beta;
    for (int q = 0; q < clr::kDotsMaxQ; ++q) { p.alpha[q] = 1.f; p.beta[q] = 0.f; }
    return clr::pixel_dots_impl(p, clr::DOTS_OP_DOT, false, nullptr, st);
}

}  // extern "C"

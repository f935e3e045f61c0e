// TransNorm: the domain-split batch normalisation with an adaptive channel weight that the
// reference builds into DeepLab with --use_TN (networks/sync_batchnorm/batchnorm.py:439-493 training, :494-521 eval;
// selected at networks/deeplabv3.py:17-23).
//
//   training: the batch is split in two halves (source = x[:B/2], target = x[B/2:]); every half is batch-normalised
//             with its OWN per-channel statistics (biased variance for the normalisation, unbiased for the running
//             estimates, F.batch_norm semantics) and the shared affine (gamma, beta); then every channel is scaled by
//             1 + alpha[c], alpha = C * prob / sum(prob), prob = 1 / (1 + |mu_s/sqrt(var_s+eps) - mu_t/sqrt(var_t+eps)|)
//             with the UNBIASED variances (torch.var default, :474-476); alpha is detached (:493).
//   eval    : normalise with the target running statistics; alpha from the two sets of running statistics.
//
// The reference spends, per layer: two cuDNN batch norms, a concatenation, two permute().contiguous() transposed
// copies, four mean/var reductions and ~12 small launches (about 9 passes over the activation forward).  Here:
//
//   forward : tn_stats_kernel (ONE read of x: shifted sums per (domain, channel))  ->  tn_fwd_finalize_kernel (O(C))
//             ->  tn_apply_kernel (one read, one write)
//   backward: tn_bwd_reduce_kernel (one read of gy and x)  ->  tn_bwd_finalize_kernel (O(C))  ->  tn_apply_kernel
//             (reads gy and x, writes gx)
//
// Bound: HBM (2 reads + 1 write forward = the minimum for a batch-statistics normalisation whose output depends on
// the whole batch; 4 reads + 1 write backward).  No atomics: per-CTA fp64 partials combined in a fixed order.
// Variance numerics: sums of (x - pivot) and (x - pivot)^2 with the pivot = first element of the (domain, channel)
// slice ("shifted data"), fp32 per plane segment, fp64 across segments / threads / CTAs.
#include "clr_common.cuh"
#include "clr_internal.h"
#include <math.h>

namespace clr {

constexpr int kTnThreads = 256;
constexpr int kTnMaxSplit = 64;
constexpr int kTnMaxC = 8192;          // prob[] lives in shared memory in the finalize kernel
constexpr int kTnChunk = 4096;         // floats per CTA in the apply kernel

struct TnGeom {
    int B, nb0, C, HW, S;              // nb0 = B/2 source samples, the rest target (batchnorm.py:452-454)
    int vec;                           // 1: HW % 4 == 0 and all planes 16-byte aligned
};

static int tn_splits(int C, int HW) {
    // enough CTAs for ~4 waves of 148 SMs, but no CTA with less than 2048 pixels per sample
    int s = (4 * 148 + 2 * C - 1) / (2 * C);
    const int cap = HW / 2048 > 1 ? HW / 2048 : 1;
    if (s > cap) s = cap;
    if (s > kTnMaxSplit) s = kTnMaxSplit;
    return s < 1 ? 1 : s;
}

// pixel range of split s: boundaries are multiples of 4 pixels
__device__ __forceinline__ void tn_range(int HW, int S, int s, int& p0, int& p1) {
    const int quads = (HW + 3) / 4, per = (quads + S - 1) / S;
    p0 = s * per * 4;
    p1 = (s + 1) * per * 4;
    if (p0 > HW) p0 = HW;
    if (p1 > HW) p1 = HW;
}

__device__ __forceinline__ void tn_block_store2(double a, double b, double* out) {
    a = warp_sum(a);
    b = warp_sum(b);
    __shared__ double sh[2][kTnThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sh[0][warp] = a; sh[1][warp] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t0 = 0.0, t1 = 0.0;
#pragma unroll
        for (int w = 0; w < kTnThreads / 32; ++w) { t0 += sh[0][w]; t1 += sh[1][w]; }
        out[0] = t0;
        out[1] = t1;
    }
}

// grid (S, C, 2): CTA = (pixel split, channel, domain).  BWD = false: partial = { sum (x - pivot), sum (x - pivot)^2 };
// BWD = true: partial = { sum g, sum g * (x - mean) } with the saved mean as the shift.
template <int VEC, bool BWD>
__global__ void __launch_bounds__(kTnThreads, 4) tn_reduce_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                                  const float* __restrict__ save, const TnGeom q,
                                                                  double* __restrict__ partial) {
    kernel_begin(TR_OTHER);
    const int s = blockIdx.x, c = blockIdx.y, d = blockIdx.z;
    const int b0 = d ? q.nb0 : 0, b1 = d ? q.B : q.nb0;
    int p0, p1;
    tn_range(q.HW, q.S, s, p0, p1);
    const float shift = BWD ? __ldg(save + (size_t)d * q.C + c) : __ldg(x + ((size_t)b0 * q.C + c) * q.HW);
    double a0 = 0.0, a1 = 0.0;
    for (int b = b0; b < b1; ++b) {
        const size_t base = ((size_t)b * q.C + c) * q.HW;
        const float* xp = x + base;
        const float* gp = BWD ? g + base : nullptr;
        float f0 = 0.f, f1 = 0.f;
        if (VEC == 4) {
            constexpr int U = 4;
            for (int p = p0 + 4 * threadIdx.x; p < p1; p += 4 * kTnThreads * U) {
                Pack<4> xv[U], gv[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int pp = p + u * 4 * kTnThreads;
                    if (pp < p1) {
                        xv[u] = ld_stream<4>(xp + pp);
                        if (BWD) gv[u] = ld_stream<4>(gp + pp);
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (p + u * 4 * kTnThreads >= p1) continue;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float t = xv[u].v[i] - shift;
                        if (BWD) { f0 += gv[u].v[i]; f1 = fmaf(gv[u].v[i], t, f1); }
                        else { f0 += t; f1 = fmaf(t, t, f1); }
                    }
                }
            }
        } else {
            for (int p = p0 + threadIdx.x; p < p1; p += kTnThreads) {
                const float t = __ldg(xp + p) - shift;
                if (BWD) { const float gg = __ldg(gp + p); f0 += gg; f1 = fmaf(gg, t, f1); }
                else { f0 += t; f1 = fmaf(t, t, f1); }
            }
        }
        a0 += (double)f0;
        a1 += (double)f1;
    }
    tn_block_store2(a0, a1, partial + (((size_t)d * q.C + c) * q.S + s) * 2);
}

struct TnFwdFin {
    const double* partial;   // [2][C][S][2]
    const float* x;          // pivots
    const float* weight; const float* bias;     // nullable (affine = False)
    float* rm[2]; float* rv[2];                 // running estimates (source, target), nullable; updated in place
    float momentum, eps;
    float* save;             // [5][C]: mean_s, mean_t, rstd_s, rstd_t, alpha
    float* coef;             // [2][C][4]: { mean, a = gamma * rstd * (1 + alpha), b = beta * (1 + alpha), 0 }
    int eval;                // 1: statistics = running estimates (normalise with the TARGET ones, batchnorm.py:497-509)
};

// fp32 arithmetic in the reference's operation order (no contraction) for dis / prob / alpha (batchnorm.py:481-487)
__device__ __forceinline__ float tn_ratio(float mean, float var, float eps) { return __fdiv_rn(mean, __fsqrt_rn(__fadd_rn(var, eps))); }

__global__ void __launch_bounds__(512) tn_fwd_finalize_kernel(const TnFwdFin f, const TnGeom q) {
    kernel_begin(TR_OTHER);
    __shared__ float prob[kTnMaxC];
    __shared__ float total;
    const int C = q.C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float mean[2], var_u[2], var_b[2];
        for (int d = 0; d < 2; ++d) {
            if (f.eval) {
                mean[d] = f.rm[d][c];
                var_u[d] = var_b[d] = f.rv[d][c];
                continue;
            }
            const int nb = d ? q.B - q.nb0 : q.nb0, b0 = d ? q.nb0 : 0;
            const double n = (double)nb * (double)q.HW;
            const double* pp = f.partial + ((size_t)d * C + c) * q.S * 2;
            double s1 = 0.0, s2 = 0.0;
            for (int s = 0; s < q.S; ++s) { s1 += pp[2 * s]; s2 += pp[2 * s + 1]; }
            const double pivot = (double)f.x[((size_t)b0 * C + c) * q.HW];
            const double m = s1 / n;
            double ss = s2 - s1 * m;                  // sum (x - mean)^2
            if (ss < 0.0) ss = 0.0;
            mean[d] = (float)(pivot + m);
            var_b[d] = (float)(ss / n);
            var_u[d] = (float)(ss / (n - 1.0));       // n = 1 -> inf/NaN like torch.var
            if (f.rm[d]) f.rm[d][c] = __fadd_rn(__fmul_rn(1.0f - f.momentum, f.rm[d][c]), __fmul_rn(f.momentum, mean[d]));
            if (f.rv[d]) f.rv[d][c] = __fadd_rn(__fmul_rn(1.0f - f.momentum, f.rv[d][c]), __fmul_rn(f.momentum, var_u[d]));
        }
        const float dis = fabsf(__fsub_rn(tn_ratio(mean[0], var_u[0], f.eps), tn_ratio(mean[1], var_u[1], f.eps)));
        prob[c] = __fdiv_rn(1.0f, __fadd_rn(1.0f, dis));
        for (int d = 0; d < 2; ++d) {
            const int ds = f.eval ? 1 : d;            // eval: every sample is normalised with the target estimates
            const float rstd = __frsqrt_rn(__fadd_rn(var_b[ds], f.eps));
            f.save[(size_t)d * C + c] = mean[ds];
            f.save[(size_t)(2 + d) * C + c] = rstd;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {       // Python's sum(prob): sequential fp32 adds starting from 0 (batchnorm.py:487)
        float t = 0.f;
        for (int c = 0; c < C; ++c) t = __fadd_rn(t, prob[c]);
        total = t;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float alpha = __fdiv_rn(__fmul_rn((float)C, prob[c]), total);
        f.save[(size_t)4 * C + c] = alpha;
        const float qq = 1.0f + alpha;
        const float gamma = f.weight ? f.weight[c] : 1.f, beta = f.bias ? f.bias[c] : 0.f;
        for (int d = 0; d < 2; ++d) {
            float* o = f.coef + ((size_t)d * C + c) * 4;
            o[0] = f.save[(size_t)d * C + c];
            o[1] = gamma * f.save[(size_t)(2 + d) * C + c] * qq;
            o[2] = beta * qq;
            o[3] = 0.f;
        }
    }
}

struct TnBwdFin {
    const double* partial;   // [2][C][S][2]: sum g, sum g (x - mean)
    const float* weight;     // nullable
    const float* save;       // [5][C]
    float* coef;             // [2][C][4]: { mean, A, Bc, Cc }: gx = A g + Bc (x - mean) + Cc
    float* gweight; float* gbias;    // nullable
    int eval;                // 1: the statistics were constants (running estimates): no batch-statistics terms
};

__global__ void __launch_bounds__(256) tn_bwd_finalize_kernel(const TnBwdFin f, const TnGeom q) {
    kernel_begin(TR_OTHER);
    const int C = q.C;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double gamma = f.weight ? (double)f.weight[c] : 1.0;
    const double qq = 1.0 + (double)f.save[(size_t)4 * C + c];
    double gw = 0.0, gb = 0.0;
    for (int d = 0; d < 2; ++d) {
        const int nb = d ? q.B - q.nb0 : q.nb0;
        const double n = (double)nb * (double)q.HW;
        const double* pp = f.partial + ((size_t)d * C + c) * q.S * 2;
        double sg = 0.0, sgx = 0.0;
        for (int s = 0; s < q.S; ++s) { sg += pp[2 * s]; sgx += pp[2 * s + 1]; }
        const double mean = (double)f.save[(size_t)d * C + c], rstd = (double)f.save[(size_t)(2 + d) * C + c];
        // dz = (1 + alpha) g; gx = gamma rstd (dz - mean(dz) - xhat mean(dz xhat)), xhat = (x - mean) rstd
        float* o = f.coef + ((size_t)d * C + c) * 4;
        o[0] = (float)mean;
        o[1] = (float)(gamma * rstd * qq);
        o[2] = f.eval ? 0.f : (float)(-gamma * qq * rstd * rstd * rstd * sgx / n);
        o[3] = f.eval ? 0.f : (float)(-gamma * rstd * qq * sg / n);
        gw += qq * rstd * sgx;
        gb += qq * sg;
    }
    if (f.gweight) f.gweight[c] = (float)gw;
    if (f.gbias) f.gbias[c] = (float)gb;
}

// grid (chunks, C, B).  BWD = false: y = (x - mean) a + b;  BWD = true: gx = A g + Bc (x - mean) + Cc
template <int VEC, bool BWD>
__global__ void __launch_bounds__(kTnThreads, 4) tn_apply_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                                 const float* __restrict__ coef, const TnGeom q,
                                                                 float* __restrict__ out) {
    kernel_begin(TR_OTHER);
    const int c = blockIdx.y, b = blockIdx.z, d = b >= q.nb0 ? 1 : 0;
    const float4 k = __ldcg(reinterpret_cast<const float4*>(coef) + (size_t)d * q.C + c);
    const size_t base = ((size_t)b * q.C + c) * q.HW;
    const int p0 = blockIdx.x * kTnChunk;
    const int p1 = p0 + kTnChunk < q.HW ? p0 + kTnChunk : q.HW;
    if (VEC == 4) {
        constexpr int U = kTnChunk / (4 * kTnThreads);
        Pack<4> xv[U], gv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int p = p0 + 4 * (u * kTnThreads + threadIdx.x);
            if (p < p1) {
                xv[u] = ld_stream<4>(x + base + p);
                if (BWD) gv[u] = ld_stream<4>(g + base + p);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int p = p0 + 4 * (u * kTnThreads + threadIdx.x);
            if (p >= p1) continue;
            Pack<4> r;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float t = xv[u].v[i] - k.x;
                r.v[i] = BWD ? fmaf(k.y, gv[u].v[i], fmaf(k.z, t, k.w)) : fmaf(t, k.y, k.z);
            }
            st_stream<4>(out + base + p, r);
        }
    } else {
        for (int p = p0 + threadIdx.x; p < p1; p += kTnThreads) {
            const float t = __ldg(x + base + p) - k.x;
            out[base + p] = BWD ? fmaf(k.y, __ldg(g + base + p), fmaf(k.z, t, k.w)) : fmaf(t, k.y, k.z);
        }
    }
}

static int tn_geom(const void* x, const void* y, const void* z, int B, int C, int HW, TnGeom* q) {
    CLR_CHECK_ARG(x && B >= 2 && C >= 1 && HW >= 1);
    if (C > kTnMaxC || C > 65535 || B > 65535) return CLR_ERR_UNSUPPORTED;
    if (!aligned4(x) || (y && !aligned4(y)) || (z && !aligned4(z))) return CLR_ERR_ALIGN;
    q->B = B; q->nb0 = B / 2; q->C = C; q->HW = HW; q->S = tn_splits(C, HW);
    q->vec = (HW % 4 == 0) && aligned16(x) && (!y || aligned16(y)) && (!z || aligned16(z));
    return CLR_OK;
}
static size_t tn_ws(int C) {
    return sizeof(double) * 2 * (size_t)C * kTnMaxSplit * 2 + sizeof(float) * 2 * (size_t)C * 4;
}
static dim3 tn_apply_grid(const TnGeom& q) { return dim3((unsigned)((q.HW + kTnChunk - 1) / kTnChunk), (unsigned)q.C, (unsigned)q.B); }

}  // namespace clr

extern "C" {

size_t clr_tn_ws_bytes(int C) { return C > 0 ? clr::tn_ws(C) : 0; }

int clr_tn_fwd(const float* x, int B, int C, int HW, const float* weight, const float* bias,
               float* running_mean_s, float* running_var_s, float* running_mean_t, float* running_var_t,
               float momentum, float eps, void* ws, size_t ws_bytes, float* y, float* save, clr_stream_t stream) {
    using namespace clr;
    TnGeom q;
    CLR_CHECK_ARG(y && save && ws);
    int rc = tn_geom(x, y, nullptr, B, C, HW, &q);
    if (rc != CLR_OK) return rc;
    if (ws_bytes < tn_ws(C)) return CLR_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double* partial = static_cast<double*>(ws);
    float* coef = reinterpret_cast<float*>(partial + 2 * (size_t)C * kTnMaxSplit * 2);
    const dim3 rg((unsigned)q.S, (unsigned)C, 2);
    if (q.vec) launch_k(tn_reduce_kernel<4, false>, rg, dim3(kTnThreads), 0, st, x, (const float*)nullptr, (const float*)nullptr, q, partial);
    else launch_k(tn_reduce_kernel<1, false>, rg, dim3(kTnThreads), 0, st, x, (const float*)nullptr, (const float*)nullptr, q, partial);
    TnFwdFin f{partial, x, weight, bias, {running_mean_s, running_mean_t}, {running_var_s, running_var_t}, momentum, eps, save, coef, 0};
    launch_k(tn_fwd_finalize_kernel, dim3(1), dim3(512), 0, st, f, q);
    if (q.vec) launch_k(tn_apply_kernel<4, false>, tn_apply_grid(q), dim3(kTnThreads), 0, st, x, (const float*)nullptr, (const float*)coef, q, y);
    else launch_k(tn_apply_kernel<1, false>, tn_apply_grid(q), dim3(kTnThreads), 0, st, x, (const float*)nullptr, (const float*)coef, q, y);
    return launch_status();
}

int clr_tn_eval(const float* x, int B, int C, int HW, const float* weight, const float* bias,
                const float* running_mean_s, const float* running_var_s, const float* running_mean_t,
                const float* running_var_t, float eps, void* ws, size_t ws_bytes, float* y, float* save,
                clr_stream_t stream) {
    using namespace clr;
    TnGeom q;
    CLR_CHECK_ARG(y && save && ws && running_mean_s && running_var_s && running_mean_t && running_var_t);
    int rc = tn_geom(x, y, nullptr, B < 2 ? 2 : B, C, HW, &q);
    if (rc != CLR_OK) return rc;
    CLR_CHECK_ARG(B >= 1);
    q.B = B; q.nb0 = B / 2;
    if (ws_bytes < tn_ws(C)) return CLR_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double* partial = static_cast<double*>(ws);
    float* coef = reinterpret_cast<float*>(partial + 2 * (size_t)C * kTnMaxSplit * 2);
    TnFwdFin f{partial, x, weight, bias, {const_cast<float*>(running_mean_s), const_cast<float*>(running_mean_t)},
               {const_cast<float*>(running_var_s), const_cast<float*>(running_var_t)}, 0.f, eps, save, coef, 1};
    launch_k(tn_fwd_finalize_kernel, dim3(1), dim3(512), 0, st, f, q);
    if (q.vec) launch_k(tn_apply_kernel<4, false>, tn_apply_grid(q), dim3(kTnThreads), 0, st, x, (const float*)nullptr, (const float*)coef, q, y);
    else launch_k(tn_apply_kernel<1, false>, tn_apply_grid(q), dim3(kTnThreads), 0, st, x, (const float*)nullptr, (const float*)coef, q, y);
    return launch_status();
}

int clr_tn_bwd(const float* x, const float* gy, int B, int C, int HW, const float* weight, const float* save,
               int eval_mode, void* ws, size_t ws_bytes, float* gx, float* gweight, float* gbias, clr_stream_t stream) {
    using namespace clr;
    TnGeom q;
    CLR_CHECK_ARG(gy && gx && save && ws && B >= (eval_mode ? 1 : 2));
    int rc = tn_geom(x, gy, gx, B < 2 ? 2 : B, C, HW, &q);
    if (rc != CLR_OK) return rc;
    q.B = B; q.nb0 = B / 2;
    if (ws_bytes < tn_ws(C)) return CLR_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double* partial = static_cast<double*>(ws);
    float* coef = reinterpret_cast<float*>(partial + 2 * (size_t)C * kTnMaxSplit * 2);
    const dim3 rg((unsigned)q.S, (unsigned)C, 2);
    if (q.vec) launch_k(tn_reduce_kernel<4, true>, rg, dim3(kTnThreads), 0, st, x, gy, save, q, partial);
    else launch_k(tn_reduce_kernel<1, true>, rg, dim3(kTnThreads), 0, st, x, gy, save, q, partial);
    TnBwdFin f{partial, weight, save, coef, gweight, gbias, eval_mode ? 1 : 0};
    launch_k(tn_bwd_finalize_kernel, dim3((unsigned)((C + 255) / 256)), dim3(256), 0, st, f, q);
    if (q.vec) launch_k(tn_apply_kernel<4, true>, tn_apply_grid(q), dim3(kTnThreads), 0, st, x, gy, (const float*)coef, q, gx);
    else launch_k(tn_apply_kernel<1, true>, tn_apply_grid(q), dim3(kTnThreads), 0, st, x, gy, (const float*)coef, q, gx);
    return launch_status();
}

}  // extern "C"

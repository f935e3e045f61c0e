// Image-resolution elementwise glue around the CLR block, each a chain of separate ATen
// launches plus .item() syncs in the reference:
//
//   clr_seg_loss_fwd/bwd : loss_seg = BCELoss(sigmoid(oS), target_map) + MSELoss(sigmoid(boundaryS), target_boundary)
//                          (Trainer_prototype_full.py:292-294), both means; one launch forward (+ a 1-CTA final sum),
//                          one launch backward writing d/d oS and d/d boundaryS (the upstream gradient is read from
//                          device memory: no host round trip).
//   clr_entropy_fwd/bwd  : uncertainty_map = -sigmoid(o) * log(sigmoid(o) + smooth)   (:452, :481, :500), forward and
//                          the adjoint the adversarial loss back-propagates into the generator.
//
// Bound: HBM (2 reads forward, 2 reads + 1 write backward per term), 2-3 MUFU per element.
// BCE follows ATen: -(y * max(log q, -100) + (1-y) * max(log(1-q), -100)) with q = sigmoid(o) ROUNDED to fp32, i.e.
// the wrong-side term saturates at 100 once q rounds to 1 (o >= 24 ln 2) or 0 (exp(-o) overflows, o <= -88.72), and
// for o > 9 ATen's own expression is evaluated (its 1 - q is quantised there); elsewhere the log terms are
// softplus(-/+ o) = max(-/+o, 0) + log1p(exp(-|o|)) -- one exponential and one logarithm for both terms.
// Backward = ATen's binary_cross_entropy_backward chained with sigmoid'.
#include "clr_common.cuh"
#include "clr_internal.h"
#include <math.h>

namespace clr {

__device__ __forceinline__ float g_ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float g_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float g_lg2(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;

__device__ __forceinline__ float sigmoid_g(float x) { return g_rcp(1.0f + g_ex2(-kLog2e * x)); }

__device__ __noinline__ float nl1q_aten(float o) { return -fmaxf(log1pf(-(1.0f / (1.0f + expf(-o)))), -100.0f); }
__device__ __forceinline__ float bce_elem(float o, float y) {
    const float L = kLn2 * g_lg2(1.0f + g_ex2(-kLog2e * fabsf(o)));     // log1p(exp(-|o|))
    float nlq = fmaxf(-o, 0.f) + L;                                       // -log q      = softplus(-o)
    float nl1q = fmaxf(o, 0.f) + L;                                       // -log(1 - q) = softplus(o)
    nlq = (o <= -88.7228f) ? 100.0f : fminf(nlq, 100.0f);
    // Above o ~ 9 ATen's 1 - q keeps only a few significant bits (q is rounded to fp32 first) and its log1p(-q) drifts
    // by up to 2 % from softplus(o) before saturating at 100: reproduce the reference's value there (rare path)
    if (o > 9.0f) nl1q = nl1q_aten(o);
    return y * nlq + (1.0f - y) * nl1q;
}
__device__ __forceinline__ float bce_grad_elem(float o, float y) {
    const float q = sigmoid_g(o);
    const float qq = (1.0f - q) * q;
    return (q - y) / fmaxf(qq, 1e-12f) * qq;
}
__device__ __forceinline__ float mse_elem(float o, float t) { const float d = sigmoid_g(o) - t; return d * d; }
__device__ __forceinline__ float mse_grad_elem(float o, float t) {
    const float q = sigmoid_g(o);
    return 2.0f * (q - t) * q * (1.0f - q);
}

constexpr int kGlueThreads = 256;
constexpr int kGlueMaxBlocks = 2048;

struct SegArgs {
    const float* o1; const float* y1; size_t n1;       // BCE term
    const float* o2; const float* y2; size_t n2;       // MSE term
    double* partial;                                   // fwd: [grid][2]
    const float* gup; float gscale; float* g1; float* g2;   // bwd
};

// flat grid-stride over the VEC-wide vectors of both terms (term 1 first), U independent load pairs in flight
template <int VEC, bool BWD>
__global__ void __launch_bounds__(kGlueThreads, 4) seg_loss_kernel(const SegArgs a) {
    kernel_begin(TR_OTHER);
    constexpr int U = 4;
    const size_t v1 = a.n1 / VEC, v2 = a.n2 / VEC, total = v1 + v2;
    const size_t nthreads = (size_t)gridDim.x * kGlueThreads;
    float s1 = 0.f, s2 = 0.f;
    float c1 = 0.f, c2 = 0.f;
    if (BWD) {
        const float up = a.gup ? a.gscale * __ldg(a.gup) : a.gscale;
        c1 = up / (float)a.n1;
        c2 = a.n2 ? up / (float)a.n2 : 0.f;
    }
    for (size_t v0 = (size_t)blockIdx.x * kGlueThreads + threadIdx.x; v0 < total; v0 += U * nthreads) {
        Pack<VEC> o[U], y[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t v = v0 + u * nthreads;
            if (v < total) {
                const bool t1 = v < v1;
                const size_t e = (t1 ? v : v - v1) * VEC;
                o[u] = ld_stream<VEC>((t1 ? a.o1 : a.o2) + e);
                y[u] = ld_stream<VEC>((t1 ? a.y1 : a.y2) + e);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t v = v0 + u * nthreads;
            if (v >= total) continue;
            const bool t1 = v < v1;
            if (BWD) {
                Pack<VEC> g;
#pragma unroll
                for (int i = 0; i < VEC; ++i)
                    g.v[i] = t1 ? c1 * bce_grad_elem(o[u].v[i], y[u].v[i]) : c2 * mse_grad_elem(o[u].v[i], y[u].v[i]);
                st_stream<VEC>((t1 ? a.g1 : a.g2) + (t1 ? v : v - v1) * VEC, g);
            } else {
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    if (t1) s1 += bce_elem(o[u].v[i], y[u].v[i]);
                    else s2 += mse_elem(o[u].v[i], y[u].v[i]);
                }
            }
        }
    }
    if (!BWD) {
        double d1 = warp_sum((double)s1), d2 = warp_sum((double)s2);
        __shared__ double sh[2][kGlueThreads / 32];
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (lane == 0) { sh[0][warp] = d1; sh[1][warp] = d2; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double t1 = 0.0, t2 = 0.0;
            for (int w = 0; w < kGlueThreads / 32; ++w) { t1 += sh[0][w]; t2 += sh[1][w]; }
            a.partial[2 * blockIdx.x] = t1;
            a.partial[2 * blockIdx.x + 1] = t2;
        }
    }
}

__global__ void __launch_bounds__(256) seg_final_kernel(const double* __restrict__ partial, int nblk, double n1, double n2,
                                                        float* __restrict__ out) {
    kernel_begin(TR_OTHER);
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < nblk; i += blockDim.x) { a += partial[2 * i]; b += partial[2 * i + 1]; }
    a = warp_sum(a);
    b = warp_sum(b);
    __shared__ double sh[2][8];
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = a; sh[1][threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        a = 0.0; b = 0.0;
        for (int w = 0; w < 8; ++w) { a += sh[0][w]; b += sh[1][w]; }
        const float l1 = (float)(a / n1), l2 = n2 > 0 ? (float)(b / n2) : 0.f;
        out[0] = l1; out[1] = l2; out[2] = l1 + l2; out[3] = 0.f;      // :292-294
    }
}

template <bool BWD>
static int seg_grid(bool vec4, int* grid, size_t total_vec) {
    int occ = 0;
    if (vec4) CLR_RETURN_IF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, seg_loss_kernel<4, BWD>, kGlueThreads, 0));
    else CLR_RETURN_IF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, seg_loss_kernel<1, BWD>, kGlueThreads, 0));
    if (occ < 1) occ = 1;
    long long want = (long long)device_facts().sms * occ;
    const long long need = (long long)((total_vec + kGlueThreads - 1) / kGlueThreads);
    if (want > need) want = need;
    if (want > kGlueMaxBlocks) want = kGlueMaxBlocks;
    *grid = (int)(want < 1 ? 1 : want);
    return CLR_OK;
}

// uncertainty map -sigmoid(o) log(sigmoid(o) + smooth) and its adjoint
template <int VEC, bool BWD>
__global__ void __launch_bounds__(kGlueThreads, 4) entropy_kernel(const float* __restrict__ o, const float* __restrict__ gout,
                                                                  size_t n, float smooth, float* __restrict__ out) {
    kernel_begin(TR_OTHER);
    const size_t nv = n / VEC, nthreads = (size_t)gridDim.x * kGlueThreads;
    constexpr int U = 4;
    for (size_t v0 = (size_t)blockIdx.x * kGlueThreads + threadIdx.x; v0 < nv; v0 += U * nthreads) {
        Pack<VEC> x[U], g[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t v = v0 + u * nthreads;
            if (v < nv) {
                x[u] = ld_stream<VEC>(o + v * VEC);
                if (BWD) g[u] = ld_stream<VEC>(gout + v * VEC);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t v = v0 + u * nthreads;
            if (v >= nv) continue;
            Pack<VEC> r;
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const float q = sigmoid_g(x[u].v[i]);
                const float lg = kLn2 * g_lg2(q + smooth);
                if (BWD) r.v[i] = -g[u].v[i] * q * (1.0f - q) * (lg + q * g_rcp(q + smooth));
                else r.v[i] = -q * lg;
            }
            st_stream<VEC>(out + v * VEC, r);
        }
    }
}

template <bool BWD>
static int launch_entropy(const float* o, const float* gout, size_t n, float smooth, float* out, cudaStream_t st) {
    if (!o || !out || n == 0 || (BWD && !gout)) return CLR_ERR_BAD_ARG;
    if (!aligned4(o) || !aligned4(out)) return CLR_ERR_ALIGN;
    const bool vec4 = (n % 4 == 0) && aligned16(o) && aligned16(out) && (!BWD || aligned16(gout));
    const size_t nv = vec4 ? n / 4 : n;
    long long blocks = (long long)((nv + kGlueThreads - 1) / kGlueThreads);
    const long long cap = (long long)device_facts().sms * 8;
    if (blocks > cap) blocks = cap;
    if (vec4) launch_k(entropy_kernel<4, BWD>, (unsigned)blocks, kGlueThreads, 0, st, o, gout, n, smooth, out);
    else launch_k(entropy_kernel<1, BWD>, (unsigned)blocks, kGlueThreads, 0, st, o, gout, n, smooth, out);
    return launch_status();
}

// ---------------------------------------------------------------------------------------------------------------
// Validation counts: per class the 2x2 confusion matrix of (sigmoid(logit) > thr) against the
// binary ground truth -- everything utils/metrics.py:118-168 derives Dice@0.75, pixel accuracy and IoU from, as exact
// integers, without the full-resolution device->host copy and the NumPy passes of the reference.
// counts[k][2*gt + pred]  (the reference's genConfusionMatrix index, utils/metrics.py:38-43).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool above_thr(float z, float thr, float zthr, float guard) {
    const float d = z - zthr;
    if (fabsf(d) > guard) return d > 0.f;
    return 1.0f / (1.0f + expf(-z)) > thr;           // ATen's sigmoid, exact decision inside the guard band
}

__global__ void __launch_bounds__(256) seg_counts_kernel(const float* __restrict__ logits, const float* __restrict__ target,
                                                         int K, size_t HW, float thr, float zthr, float guard,
                                                         unsigned long long* __restrict__ counts) {
    kernel_begin(TR_OTHER);
    const int plane = blockIdx.y;                     // b*K + k
    const int k = plane % K;
    const float* z = logits + (size_t)plane * HW;
    const float* t = target + (size_t)plane * HW;
    unsigned int c[4] = {0u, 0u, 0u, 0u};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (size_t)gridDim.x * blockDim.x) {
        const int p = above_thr(__ldg(z + i), thr, zthr, guard) ? 1 : 0;
        const int g = (__ldg(t + i) != 0.f) ? 1 : 0;
        ++c[2 * g + p];
    }
    __shared__ unsigned int sh[4][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        unsigned int v = c[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) sh[j][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        unsigned long long s = 0;
        for (int w = 0; w < 8; ++w) s += sh[threadIdx.x][w];
        if (s) atomicAdd(counts + (size_t)k * 4 + threadIdx.x, s);      // integer atomics: order-independent, exact
    }
}

}  // namespace clr

extern "C" {

size_t clr_seg_loss_ws_bytes(void) { return sizeof(double) * 2 * clr::kGlueMaxBlocks; }

int clr_seg_loss_fwd(const float* oS, const float* target_map, size_t n1, const float* boundaryS, const float* target_boundary,
                     size_t n2, void* ws, size_t ws_bytes, float* out, clr_stream_t stream) {
    if (!oS || !target_map || n1 == 0 || !ws || !out || (n2 && (!boundaryS || !target_boundary))) return CLR_ERR_BAD_ARG;
    if (ws_bytes < clr_seg_loss_ws_bytes()) return CLR_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool vec4 = (n1 % 4 == 0) && (n2 % 4 == 0) && clr::aligned16(oS) && clr::aligned16(target_map) &&
                      (!n2 || (clr::aligned16(boundaryS) && clr::aligned16(target_boundary)));
    clr::SegArgs a{oS, target_map, n1, boundaryS, target_boundary, n2, static_cast<double*>(ws), nullptr, 0.f, nullptr, nullptr};
    int grid = 1;
    const size_t tv = vec4 ? (n1 + n2) / 4 : n1 + n2;
    const int rc = clr::seg_grid<false>(vec4, &grid, tv);
    if (rc != CLR_OK) return rc;
    if (vec4) clr::launch_k(clr::seg_loss_kernel<4, false>, grid, clr::kGlueThreads, 0, st, a);
    else clr::launch_k(clr::seg_loss_kernel<1, false>, grid, clr::kGlueThreads, 0, st, a);
    clr::launch_k(clr::seg_final_kernel, 1, 256, 0, st, static_cast<const double*>(ws), grid, (double)n1, (double)n2, out);
    return clr::launch_status();
}

int clr_seg_loss_bwd(const float* oS, const float* target_map, size_t n1, const float* boundaryS, const float* target_boundary,
                     size_t n2, const float* gup_dev, float gscale, float* g_oS, float* g_boundaryS, clr_stream_t stream) {
    if (!oS || !target_map || n1 == 0 || !g_oS || (n2 && (!boundaryS || !target_boundary || !g_boundaryS))) return CLR_ERR_BAD_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool vec4 = (n1 % 4 == 0) && (n2 % 4 == 0) && clr::aligned16(oS) && clr::aligned16(target_map) && clr::aligned16(g_oS) &&
                      (!n2 || (clr::aligned16(boundaryS) && clr::aligned16(target_boundary) && clr::aligned16(g_boundaryS)));
    clr::SegArgs a{oS, target_map, n1, boundaryS, target_boundary, n2, nullptr, gup_dev, gscale, g_oS, g_boundaryS};
    int grid = 1;
    const size_t tv = vec4 ? (n1 + n2) / 4 : n1 + n2;
    const int rc = clr::seg_grid<true>(vec4, &grid, tv);
    if (rc != CLR_OK) return rc;
    if (vec4) clr::launch_k(clr::seg_loss_kernel<4, true>, grid, clr::kGlueThreads, 0, st, a);
    else clr::launch_k(clr::seg_loss_kernel<1, true>, grid, clr::kGlueThreads, 0, st, a);
    return clr::launch_status();
}

int clr_entropy_fwd(const float* o, size_t n, float smooth, float* out, clr_stream_t stream) {
    return clr::launch_entropy<false>(o, nullptr, n, smooth, out, static_cast<cudaStream_t>(stream));
}

int clr_entropy_bwd(const float* o, const float* gout, size_t n, float smooth, float* gin, clr_stream_t stream) {
    return clr::launch_entropy<true>(o, gout, n, smooth, gin, static_cast<cudaStream_t>(stream));
}

int clr_seg_counts(const float* logits, const float* target, int B, int K, size_t HW, float thr,
                   unsigned long long* counts, clr_stream_t stream) {
    if (!logits || !target || !counts || B < 1 || K < 1 || HW == 0 || (long long)B * K > 65535) return CLR_ERR_BAD_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CLR_RETURN_IF_CUDA(cudaMemsetAsync(counts, 0, sizeof(unsigned long long) * 4 * (size_t)K, st));
    float zthr = 0.f, guard = 3.0e38f;
    const double t = (double)thr;
    if (t > 0.0 && t < 1.0) {
        zthr = (float)log(t / (1.0 - t));
        const double g = 4e-6 / (t * (1.0 - t));
        guard = (float)(g > 1e-4 ? g : 1e-4);
    }
    unsigned gx = (unsigned)((HW + 256 * 8 - 1) / (256 * 8));
    if (gx < 1) gx = 1;
    if (gx > 64) gx = 64;
    clr::launch_k(clr::seg_counts_kernel, dim3(gx, (unsigned)(B * K)), 256, 0, st, logits, target, K, HW, thr, zthr, guard, counts);
    return clr::launch_status();
}

}  // extern "C"

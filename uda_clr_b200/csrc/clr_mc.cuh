// MC-dropout statistics device code shared by mc_stats.cu and the fused [pooling | MC statistics] launch (pool_fwd.cu):
// the two sigmoids of a logit from one exponential, the per-position std / mean over the T passes (streaming form and the
// ATen-exact form that the mask guard band and `mc_precise` use).  See mc_stats.cu for the reference lines.
#pragma once
#include "clr_common.cuh"

namespace clr {

__device__ __forceinline__ float sigmoid_aten(float x) { return 1.0f / (1.0f + expf(-x)); }

// ---- torch.std(dim=0) / torch.mean(dim=0) over the T MC passes, in ATen's CUDA evaluation order ------------------
// (utils/Utils.py:166, 168 run ATen's generic reduction: ReduceMomentKernel.cu / ReduceOps via Reduce.cuh.)  For an
// outer reduction with fewer than 64 values per output ONE thread reduces an output, with `vt0` interleaved
// accumulators that are combined at the end (element t goes to accumulator t % vt0): vt0 = 2 Welford accumulators for
// std, 4 plain sums for mean.  The uncertainty mask `std_small < 0.04` (utils/Utils.py:197-200) is an integer output:
// it must equal eager torch on the same device, so pixels near the threshold are re-evaluated with exactly this order
// (guard band, retrify epilogues below); `mc_precise` = 1 computes the whole maps this way (slow, tests).
// Roundings are pinned with intrinsics to the contraction nvcc applies to ATen's expressions (WelfordOps::reduce:
// `m2 + delta * (x - new_mean)` -> fma; ::combine: `a.mean + delta * nb_over_n` -> fma, `a.m2 + b.m2 + delta * delta *
// a.nf * nb_over_n` -> fma of the last product into the sum).  Verified bit for bit against torch 2.11 on the B200 for
// T = 2..20 (tools/aten_order_probe.py, profiles/r02_aten_order_probe.json: the no-fma forms and vt0 = 1 / 4 all
// mismatch, this form has 0 mismatches in 4.7 M values) and by tests/test_gpu_step.py on every GPU test run.
struct WelfordAcc { float mean, m2, nf; };
__device__ __forceinline__ void welford_push(WelfordAcc& a, float x) {
    a.nf += 1.0f;
    const float delta = __fsub_rn(x, a.mean);
    a.mean = __fadd_rn(a.mean, __fdiv_rn(delta, a.nf));
    a.m2 = __fmaf_rn(delta, __fsub_rn(x, a.mean), a.m2);
}
__device__ __forceinline__ WelfordAcc welford_merge(const WelfordAcc& a, const WelfordAcc& b) {
    if (a.nf == 0.f) return b;
    if (b.nf == 0.f) return a;
    const float delta = __fsub_rn(b.mean, a.mean);
    const float n = __fadd_rn(a.nf, b.nf);
    const float nb_over_n = __fdiv_rn(b.nf, n);
    WelfordAcc r;
    r.mean = __fmaf_rn(delta, nb_over_n, a.mean);
    r.m2 = __fmaf_rn(__fmul_rn(__fmul_rn(delta, delta), a.nf), nb_over_n, __fadd_rn(a.m2, b.m2));
    r.nf = n;
    return r;
}
__device__ __forceinline__ float sigmoid_half_aten(float p) { return sigmoid_aten(__fmul_rn(p, 0.5f)); }   // preds / 2.0 (:165)
// std_T(sigmoid(p/2)) (unbiased) at position i of preds [T][n].  Eight loads in flight per round (a guard-band pixel sits
// on the kernel's critical path: its T logits must not be T dependent DRAM round trips).
static __device__ __noinline__ float std_aten_at(const float* __restrict__ preds, int T, size_t n, size_t i) {
    WelfordAcc a0{0.f, 0.f, 0.f}, a1{0.f, 0.f, 0.f};      // even / odd passes
    for (int t0 = 0; t0 < T; t0 += 8) {
        float x[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) x[u] = (t0 + u < T) ? __ldg(preds + (size_t)(t0 + u) * n + i) : 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (t0 + u < T) {
                const float sgm = sigmoid_half_aten(x[u]);
                if (u & 1) welford_push(a1, sgm); else welford_push(a0, sgm);
            }
        }
    }
    const WelfordAcc r = welford_merge(a0, a1);
    const float divisor = r.nf > 1.0f ? __fsub_rn(r.nf, 1.0f) : 0.0f;      // correction = 1; T = 1 -> 0/0 = NaN like torch
    return __fsqrt_rn(__fdiv_rn(r.m2, divisor));
}
// mean_T(sigmoid(p)) at position i: 4 interleaved partial sums, ((s0 + s1) + s2) + s3, times factor = n_out / numel
static __device__ __noinline__ float mean_aten_at(const float* __restrict__ preds, int T, size_t n, size_t i, float factor) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int t0 = 0; t0 < T; t0 += 8) {
        float x[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) x[u] = (t0 + u < T) ? __ldg(preds + (size_t)(t0 + u) * n + i) : 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (t0 + u < T) acc[u & 3] = __fadd_rn(acc[u & 3], sigmoid_aten(x[u]));
    }
    return __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(acc[0], acc[1]), acc[2]), acc[3]), factor);
}
static inline float mean_factor_aten(size_t n_out, int T) { return (float)n_out / (float)(n_out * (size_t)T); }

// The two sigmoids of one MC logit from ONE exponential:  u = e^{-p/2}:  sigmoid(p/2) = 1/(1+u),
// sigmoid(p) = 1/(1+u^2).  PRECISE keeps expf + IEEE division for sigmoid(p/2) (ATen's expression, bit for
// bit); the default uses ex2.approx / rcp.approx (2 MUFU per logit, a few ulp) because with two precise
// sigmoids per logit the pass is instruction-bound at ~1 TB/s instead of HBM-bound (profiles/r01).
template <bool PRECISE>
__device__ __forceinline__ void mc_sigmoids(float p, float& s_half, float& s_full) {
    if (PRECISE) {
        const float u = expf(-(p / 2.0f));
        s_half = 1.0f / (1.0f + u);
        s_full = 1.0f / (1.0f + expf(-p));
    } else {
        // 2 MUFU per logit: u = 2^(-p/2 * log2 e);  r = 1 / ((1+u)(1+u^2));  s_half = r (1+u^2), s_full = r (1+u).
        // The clamp keeps (1+u)(1+u^2) finite; below -55 both sigmoids are < 2e-12 anyway.
        float u, r;
        const float pc = fmaxf(p, -55.0f);
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(u) : "f"(pc * -0.72134752044448170368f));
        const float a = 1.0f + u, b = fmaf(u, u, 1.0f);
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a * b));
        s_half = __fmul_rn(r, b);      // pinned (never contracted into a later add / subtract): the value is also stored
        s_full = r * a;                // and re-used by the variance pass, and every instantiation must round alike
    }
}

// std (unbiased, of sigmoid(p/2)) and mean (of sigmoid(p)) over the T MC passes for the VEC positions starting at i.
// preds is [T][n].  TT > 0: T <= TT, values kept in registers (two-pass variance); TT == 0: any T, Welford.
// EXACT: T == TT is known at compile time (the reference's T = 8): no per-pass predicates -- they were ~10 % of the
// kernel's instructions (32 BRA + 26 ISETP per thread in the ncu source page) in a pass that is issue / MUFU co-limited.
struct McAten { float factor; };     // PRECISE instantiations only: ATen's mean factor (n_out / numel as float)

template <int VEC, int TT, bool PRECISE, bool EXACT = false>
__device__ __forceinline__ void mc_vec_stats(const float* __restrict__ preds, int T_rt, size_t n, size_t i, Pack<VEC>& s, Pack<VEC>& m,
                                             const McAten aten = McAten{0.f}) {
    const int T = EXACT ? TT : T_rt;
    if constexpr (PRECISE) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            s.v[v] = std_aten_at(preds, T, n, i + v);
            m.v[v] = mean_aten_at(preds, T, n, i + v, aten.factor);
        }
        return;
    }
    float mean_h[VEC], m2[VEC], mean_f[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) { mean_h[v] = 0.f; m2[v] = 0.f; mean_f[v] = 0.f; }
    if constexpr (TT > 0) {
        Pack<VEC> x[TT];
#pragma unroll
        for (int t = 0; t < TT; ++t)
            if (t < T) x[t] = ld_stream<VEC>(preds + (size_t)t * n + i);     // all T loads in flight
#pragma unroll
        for (int t = 0; t < TT; ++t) {
            if (t < T) {
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    float sh, sf;
                    mc_sigmoids<PRECISE>(x[t].v[v], sh, sf);                   // utils/Utils.py:164-165
                    x[t].v[v] = sh;
                    mean_f[v] += sf;
                    mean_h[v] += sh;
                }
            }
        }
        const float invT = 1.0f / (float)T;
#pragma unroll
        for (int v = 0; v < VEC; ++v) mean_h[v] = __fmul_rn(mean_h[v], invT);   // pinned: never contracted into the subtraction
                                                                                 // below, so every instantiation rounds alike
#pragma unroll
        for (int t = 0; t < TT; ++t) {
            if (t < T) {
#pragma unroll
                for (int v = 0; v < VEC; ++v) { const float dl = x[t].v[v] - mean_h[v]; m2[v] = fmaf(dl, dl, m2[v]); }
            }
        }
    } else {
        for (int t = 0; t < T; ++t) {
            const Pack<VEC> x = ld_stream<VEC>(preds + (size_t)t * n + i);
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                float av, sf;
                mc_sigmoids<PRECISE>(x.v[v], av, sf);
                mean_f[v] += sf;
                const float dl = av - mean_h[v];
                mean_h[v] += dl / (float)(t + 1);
                m2[v] = fmaf(dl, av - mean_h[v], m2[v]);
            }
        }
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        s.v[v] = sqrtf(m2[v] / (float)(T - 1));   // unbiased (torch.std default, :166); T = 1 -> NaN like torch
        m.v[v] = mean_f[v] / (float)T;            // :168
    }
}


}  // namespace clr

// The O(K*C) glue between the streaming passes -- everything the reference does with ~40 tiny ATen
// launches and two .item() host syncs (Trainer_prototype_full.py:335-355, 378-398, 428-449):
//
// clr_align_finalize : mu = S/N (utils/Utils.py:127-130) -> EMA with the stored prototypes
//                      (first step: copy; later (1-d)*stored + d*cur; stored <- detached result)
//                      -> intra / inter losses (:428-444) -> dL/d(cur prototypes) for both domains
//                      -> the discriminative pass's contraction vectors D_k = P_obj,k - P_bck,k and
//                         offsets beta_k = (|P_obj,k|^2 - |P_bck,k|^2)/C.
// clr_disc_finalize  : loss_disc and its prototype gradients from the active-set sums
//                      (Trainer_prototype_mt bytecode L454-474), the direct-gradient
//                      table for the backward write, and the step total.
// One CTA each; reductions in fp64, fixed order.
#include "clr_common.cuh"
#include "clr_internal.h"
#include "clr_finish.cuh"

namespace clr {

// One thread per (class k, channel c): it owns BOTH rows k (obj) and K+k (bck) of both domains, so the EMA
// state can be updated in place with no cross-thread hazard and no second pass.
__global__ void __launch_bounds__(1024) align_finalize_kernel(
    const float* __restrict__ sums_s, const float* __restrict__ sums_t, int K, int C,
    float* __restrict__ stored_s, float* __restrict__ stored_t, int first_s, int first_t, double decay,
    float w_intra, float w_inter,
    float* __restrict__ P_s, float* __restrict__ P_t, float* __restrict__ cur_s, float* __restrict__ cur_t,
    float* __restrict__ g_s, float* __restrict__ g_t,
    float* __restrict__ disc_vec, float* __restrict__ disc_beta, float* __restrict__ losses) {
    kernel_begin(TR_ALIGN);
    __shared__ double sh[(2 + CLR_MAX_K) * 32];
    const float d = (float)decay, omd = (float)(1.0 - decay);   // the reference forms (1 - decay) in double, then casts
    const float ds = first_s ? 1.f : d, dt = first_t ? 1.f : d;
    const float invC = 1.0f / (float)C;
    double acc[2 + CLR_MAX_K];   // intra, inter, nb_0 .. nb_{K-1}
#pragma unroll
    for (int i = 0; i < 2 + CLR_MAX_K; ++i) acc[i] = 0.0;
    for (int i = threadIdx.x; i < K * C; i += blockDim.x) {
        const int k = i / C, c = i - k * C;
        float ps[2], pt[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = k + h * K;
            const size_t e = (size_t)r * C + c;
            const float cs = sums_s[(size_t)r * (C + 1) + c] / sums_s[(size_t)r * (C + 1) + C];   // utils/Utils.py:127-130
            const float ct = sums_t[(size_t)r * (C + 1) + c] / sums_t[(size_t)r * (C + 1) + C];
            ps[h] = first_s ? cs : __fadd_rn(__fmul_rn(omd, stored_s[e]), __fmul_rn(d, cs));
            pt[h] = first_t ? ct : __fadd_rn(__fmul_rn(omd, stored_t[e]), __fmul_rn(d, ct));
            if (cur_s) cur_s[e] = cs;
            if (cur_t) cur_t[e] = ct;
            P_s[e] = ps[h]; P_t[e] = pt[h];
            stored_s[e] = ps[h]; stored_t[e] = pt[h];     // .detach() copies (Trainer_prototype_full.py:341-344)
            const double df = (double)ps[h] - (double)pt[h];
            acc[0] += df * df;
        }
        const float dob = ps[0] - ps[1];                    // P_obj,k - P_bck,k
        acc[1] += (double)dob * dob;
        const float gsep = ds * w_inter * 2.0f * dob * invC;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const size_t e = (size_t)(k + h * K) * C + c;
            const float gi = w_intra * 2.0f * (ps[h] - pt[h]) * invC;
            g_s[e] = ds * gi + (h == 0 ? gsep : -gsep);
            g_t[e] = -dt * gi;
        }
        if (disc_vec) disc_vec[i] = dob;
#pragma unroll
        for (int kk = 0; kk < CLR_MAX_K; ++kk)
            if (kk == k) acc[2 + kk] += (double)ps[0] * ps[0] - (double)ps[1] * ps[1];
    }
    block_sum_n<2 + CLR_MAX_K>(acc, sh);
    if (threadIdx.x == 0) {
        losses[0] = (float)(acc[0] / C);
        losses[1] = (float)(acc[1] / C);
    }
    if (disc_beta && threadIdx.x == 0) {
#pragma unroll
        for (int kk = 0; kk < CLR_MAX_K; ++kk)
            if (kk < K) disc_beta[kk] = (float)(acc[2 + kk] / C);
    }
    trace_exit(TR_ALIGN);
}

// packed2 layout: [K][C+1] active-set sums (col C = n_k) | loss numerator | cons num | cons den | pad
__global__ void __launch_bounds__(256) disc_finalize_kernel(
    float* __restrict__ packed2, const float* __restrict__ P_s, int K, int C, double npx, float w_disc,
    float ema_factor, float gscale, float* __restrict__ g_s, float* __restrict__ xtab,
    float w_intra, float w_inter, float w_aug, float aug_weight, int use_disc, int use_cons,
    float* __restrict__ losses, PackSrc ps) {
    kernel_begin(TR_DISC_FIN);
    float* tail = packed2 + (size_t)K * (C + 1);
    __shared__ double shp[3 * 32];
    __shared__ float tail_s[3];
    if (ps.hinge || ps.cons) {
        double v[3] = {0.0, 0.0, 0.0};
        if (ps.hinge)
            for (int i = threadIdx.x; i < ps.n_hinge; i += blockDim.x) v[0] += (double)ps.hinge[(size_t)i * ps.hinge_stride];
        if (ps.cons) {
            const double2* c2 = reinterpret_cast<const double2*>(ps.cons);
            int i = threadIdx.x;
            for (; i + 3 * (int)blockDim.x < ps.n_cons; i += 4 * blockDim.x) {
                double2 t[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) t[u] = c2[i + u * blockDim.x];
#pragma unroll
                for (int u = 0; u < 4; ++u) { v[1] += t[u].x; v[2] += t[u].y; }
            }
            for (; i < ps.n_cons; i += blockDim.x) { const double2 t = c2[i]; v[1] += t.x; v[2] += t.y; }
        }
        block_sum_n<3>(v, shp);
        if (threadIdx.x == 0) {
            tail[0] = tail_s[0] = (float)v[0]; tail[1] = tail_s[1] = (float)v[1]; tail[2] = tail_s[2] = (float)v[2];
            tail[3] = 0.f;
        }
    } else if (threadIdx.x == 0) {
        tail_s[0] = tail[0]; tail_s[1] = tail[1]; tail_s[2] = tail[2];
    }
    __syncthreads();
    if (use_disc) {
        const float coef = (float)(2.0 / ((double)C * npx));
        for (int i = threadIdx.x; i < K * C; i += blockDim.x) {
            const int k = i / C, c = i - k * C;
            const float nk = packed2[(size_t)k * (C + 1) + C];
            const float A = packed2[(size_t)k * (C + 1) + c];
            const float po = P_s[(size_t)k * C + c], pb = P_s[(size_t)(K + k) * C + c];
            g_s[(size_t)k * C + c] += ema_factor * w_disc * coef * (nk * po - A);
            g_s[(size_t)(K + k) * C + c] -= ema_factor * w_disc * coef * (nk * pb - A);
            xtab[i] = -gscale * w_disc * coef * (po - pb);
        }
    }
    if (threadIdx.x == 0) {
        const float disc = use_disc ? (float)((double)tail_s[0] / npx) : 0.f;
        const float aug = use_cons ? (float)((double)tail_s[1] / (double)tail_s[2] * (double)aug_weight) : 0.f;
        losses[2] = disc;
        losses[3] = aug;
        losses[4] = w_intra * losses[0] + w_inter * losses[1] + w_disc * disc + w_aug * aug;
        losses[5] = 0.f; losses[6] = 0.f; losses[7] = 0.f;
    }
    trace_exit(TR_DISC_FIN);
}

// Sum the hinge per-CTA partials and the consistency per-CTA partials into the tail of packed2 (fixed order).
__global__ void __launch_bounds__(256) step_pack_kernel(const float* __restrict__ hinge_partials, int n_hinge, int hinge_stride,
                                                        const double* __restrict__ cons_partials, int n_cons,
                                                        float* __restrict__ tail) {
    kernel_begin(TR_PACK);
    __shared__ double sh[3 * 32];
    double v[3] = {0.0, 0.0, 0.0};
    if (hinge_partials)
        for (int i = threadIdx.x; i < n_hinge; i += blockDim.x) v[0] += (double)hinge_partials[(size_t)i * hinge_stride];
    if (cons_partials)
        for (int i = threadIdx.x; i < n_cons; i += blockDim.x) { v[1] += cons_partials[2 * i]; v[2] += cons_partials[2 * i + 1]; }
    block_sum_n<3>(v, sh);
    if (threadIdx.x == 0) { tail[0] = (float)v[0]; tail[1] = (float)v[1]; tail[2] = (float)v[2]; tail[3] = 0.f; }
    trace_exit(TR_PACK);
}

void launch_step_pack(const float* hinge_partials, int n_hinge, int hinge_stride,
                      const double* cons_partials, int n_cons, float* tail, cudaStream_t st) {
    clr::launch_k(step_pack_kernel, 1, 256, 0, st, hinge_partials, n_hinge, hinge_stride, cons_partials, n_cons, tail);
}

int disc_finalize_impl(float* packed2, const float* P_s, int K, int C, double npx, float w_disc,
                       float ema_factor, float gscale, float* g_s, float* xtab,
                       float w_intra, float w_inter, float w_aug, float aug_weight, int use_disc, int use_cons,
                       float* losses, const float* hinge, int n_hinge, int hinge_stride,
                       const double* cons, int n_cons, cudaStream_t stream) {
    if (!packed2 || !losses || K < 1 || K > CLR_MAX_K || C < 1) return CLR_ERR_BAD_ARG;
    if (use_disc && (!P_s || !g_s || !xtab || npx <= 0)) return CLR_ERR_BAD_ARG;
    PackSrc ps{hinge, n_hinge, hinge_stride, cons, n_cons};
    clr::launch_k(disc_finalize_kernel, 1, 256, 0, stream, packed2, P_s, K, C, npx, w_disc, ema_factor, gscale, g_s, xtab, w_intra, w_inter, w_aug, aug_weight,
        use_disc, use_cons, losses, ps);
    return launch_status();
}

// Stand-alone launches of the merged reduce + finalize bodies (clr_finish.cuh); the fused step co-schedules them with
// the consistency pass / the target-gradient write instead (cons.cu, pool_bwd.cu).
__global__ void __launch_bounds__(kThreads) pool_finish_kernel(const PoolFinishParams p) {
    const int tr = p.mode == 1 ? TR_FIN_S : TR_ALIGN;
    if (p.done_fin || p.done_all) kernel_begin_late_trigger(tr); else kernel_begin(tr);
    pool_finish_body(p, blockIdx.x, gridDim.x);
    if (p.done_all) cta_signal(p.early_signal ? nullptr : p.done_fin, p.done_all);   // (early: done_fin was bumped inside the body)
    trace_exit(tr);
}
__global__ void __launch_bounds__(kThreads) disc_finish_kernel(const DiscFinishParams p) {
    // gate_signal: the backward launch behind this one skips its griddepcontrol.wait (its source CTAs wait on the gate
    // instead), so it must not be launched before everything older has completed: trigger AFTER the wait
    if (p.gate_signal) kernel_begin_late_trigger(TR_DISC_FIN); else kernel_begin(TR_DISC_FIN);
    disc_finish_body(p, blockIdx.x, gridDim.x);
    if (p.gate_signal) cta_signal(p.gate_signal, nullptr);
    trace_exit(TR_DISC_FIN);
}
int pool_finish_launch(const PoolFinishParams& p, cudaStream_t st) {
    clr::launch_k(pool_finish_kernel, pool_finish_ctas(p.C), kThreads, 0, st, p);
    return launch_status();
}
int disc_finish_launch(const DiscFinishParams& p, cudaStream_t st) {
    clr::launch_k(disc_finish_kernel, disc_finish_ctas(p.C), kThreads, 0, st, p);
    return launch_status();
}

}  // namespace clr

extern "C" {

int clr_align_finalize(const float* sums_s, const float* sums_t, int K, int C,
                       float* stored_s, float* stored_t, int first_s, int first_t, double decay,
                       float w_intra, float w_inter, float* P_s, float* P_t, float* cur_s, float* cur_t,
                       float* g_s, float* g_t, float* disc_vec, float* disc_beta, float* losses,
                       clr_stream_t stream) {
    if (!sums_s || !sums_t || !stored_s || !stored_t || !P_s || !P_t || !g_s || !g_t || !losses ||
        K < 1 || K > CLR_MAX_K || C < 1)
        return CLR_ERR_BAD_ARG;
    clr::launch_k(clr::align_finalize_kernel, 1, 1024, 0, static_cast<cudaStream_t>(stream), sums_s, sums_t, K, C, stored_s, stored_t, first_s, first_t, decay, w_intra, w_inter,
        P_s, P_t, cur_s, cur_t, g_s, g_t, disc_vec, disc_beta, losses);
    return clr::launch_status();
}

int clr_disc_finalize(float* packed2, const float* P_s, int K, int C, double npx, float w_disc,
                      float ema_factor, float gscale, float* g_s, float* xtab,
                      float w_intra, float w_inter, float w_aug, float aug_weight, int use_disc, int use_cons,
                      float* losses, clr_stream_t stream) {
    return clr::disc_finalize_impl(packed2, P_s, K, C, npx, w_disc, ema_factor, gscale, g_s, xtab, w_intra, w_inter,
                                   w_aug, aug_weight, use_disc, use_cons, losses, nullptr, 0, 0, nullptr, 0,
                                   static_cast<cudaStream_t>(stream));
}

}  // extern "C"

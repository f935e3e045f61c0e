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

namespace clr {

// Block-wide sums of NV doubles at once (one barrier): result valid in THREAD 0 only.
template <int NV>
__device__ __forceinline__ void block_sum_n(double (&v)[NV], double* sh /*[NV][32]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) sh[i * 32 + warp] = v[i];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = warp_sum(lane < nw ? sh[i * 32 + lane] : 0.0);
    }
}

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
struct PackSrc {   // per-CTA partials still to be summed (single-GPU path: no exchange between pack and finalize)
    const float* hinge; int n_hinge, hinge_stride;
    const double* cons; int n_cons;
};

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

// ------------------------------------------------------------------------------------------------------------------
// Single-GPU fused step: "finish" kernels = partial reduce + finalize in ONE launch each (no exchange point between
// them when nothing is sharded), which takes two kernel boundaries and two single-CTA latency chains off the
// critical path of every step.  Same arithmetic as the separate kernels above.
// ------------------------------------------------------------------------------------------------------------------


// Sum of col[(sl0 + i*step) * stride] over the slots in [.., s_end): rounds of U predicated loads, all in flight at once
// (a plain remainder loop would serialise one L2 round trip per slot); fp64 accumulation in slot order.
template <int U>
__device__ __forceinline__ double strided_slot_sum(const float* __restrict__ col, int sl, int s_end, size_t stride, int step) {
    double s = 0.0;
    for (; sl < s_end; sl += U * step) {
        float v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = (sl + u * step < s_end) ? col[(size_t)(sl + u * step) * stride] : 0.f;
#pragma unroll
        for (int u = 0; u < U; ++u) s += (double)v[u];
    }
    return s;
}

// CTA = 8 channels; warp w = (domain d, row r) reduces column block [c0, c0+8) of partial[d][slot][r][.] over the
// slots with 4 slot-lanes per channel (fp64, fixed order) plus its weight-sum column; then K*8 threads do the
// align_finalize arithmetic for the CTA's channels; the loss terms are combined across CTAs by the last CTA to
// finish (per-CTA fp64 partials summed in CTA order -> deterministic).  `counter` is zeroed by the pooling kernel
// of the same step (stream order), so no initialisation contract leaks into the ABI.
struct PoolFinishParams {
    const float* partial[2];   // [slots][R][C+1]
    float* sums[2];            // packed sums out
    float* stored[2];
    float* P[2];
    float* g[2];
    int slots[2];
    int first[2];
    int K, C;
    float d, omd, w_intra, w_inter;
    float* disc_vec;
    float* disc_beta;
    float* losses;
    double* loss_partial;      // [grid][2 + CLR_MAX_K]
    unsigned int* counter;
};

__global__ void __launch_bounds__(1024) pool_finish_kernel(const PoolFinishParams p) {
    kernel_begin(TR_ALIGN);
    constexpr int NL = 2 + CLR_MAX_K;
    const int K = p.K, R = 2 * K, C = p.C, n = R * (C + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int d = warp / R, r = warp - d * R;          // blockDim.x = 2 * R * 32
    const int ch = lane >> 2, sl0 = lane & 3;
    const int c = blockIdx.x * 8 + ch;
    __shared__ float S[2][2 * CLR_MAX_K][8];
    __shared__ float Nn[2][2 * CLR_MAX_K];
    __shared__ double lp[8 * CLR_MAX_K][NL];
    __shared__ bool is_last;
    {
        const float* part = p.partial[d];
        const int slots = p.slots[d];
        double s = 0.0;
        if (c < C) s = strided_slot_sum<8>(part + (size_t)r * (C + 1) + c, sl0, slots, (size_t)n, 4);
        double nn = strided_slot_sum<4>(part + (size_t)r * (C + 1) + C, lane, slots, (size_t)n, 32);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        nn = warp_sum(nn);
        if (sl0 == 0 && c < C) { S[d][r][ch] = (float)s; p.sums[d][(size_t)r * (C + 1) + c] = (float)s; }
        if (lane == 0) { Nn[d][r] = (float)nn; if (blockIdx.x == 0) p.sums[d][(size_t)r * (C + 1) + C] = (float)nn; }
    }
    __syncthreads();
    if (tid < K * 8) {
        const int k = tid >> 3, j = tid & 7, cc = blockIdx.x * 8 + j;
        double acc[NL];
#pragma unroll
        for (int i = 0; i < NL; ++i) acc[i] = 0.0;
        if (cc < C) {
            const float dd = p.d, omd = p.omd, invC = 1.0f / (float)C;
            const float ds = p.first[0] ? 1.f : dd, dt = p.first[1] ? 1.f : dd;
            float ps[2], pt[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int rr = k + h * K;
                const size_t e = (size_t)rr * C + cc;
                const float cs = S[0][rr][j] / Nn[0][rr];                         // utils/Utils.py:127-130
                const float ct = S[1][rr][j] / Nn[1][rr];
                ps[h] = p.first[0] ? cs : __fadd_rn(__fmul_rn(omd, p.stored[0][e]), __fmul_rn(dd, cs));
                pt[h] = p.first[1] ? ct : __fadd_rn(__fmul_rn(omd, p.stored[1][e]), __fmul_rn(dd, ct));
                p.P[0][e] = ps[h]; p.P[1][e] = pt[h];
                p.stored[0][e] = ps[h]; p.stored[1][e] = pt[h];                  // .detach() copies (Trainer_prototype_full.py:341-344)
                const double df = (double)ps[h] - (double)pt[h];
                acc[0] += df * df;
            }
            const float dob = ps[0] - ps[1];
            acc[1] += (double)dob * dob;
            const float gsep = ds * p.w_inter * 2.0f * dob * invC;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const size_t e = (size_t)(k + h * K) * C + cc;
                const float gi = p.w_intra * 2.0f * (ps[h] - pt[h]) * invC;
                p.g[0][e] = ds * gi + (h == 0 ? gsep : -gsep);
                p.g[1][e] = -dt * gi;
            }
            if (p.disc_vec) p.disc_vec[(size_t)k * C + cc] = dob;
#pragma unroll
            for (int kk = 0; kk < CLR_MAX_K; ++kk)
                if (kk == k) acc[2 + kk] = (double)ps[0] * ps[0] - (double)ps[1] * ps[1];
        }
#pragma unroll
        for (int i = 0; i < NL; ++i) lp[tid][i] = acc[i];
    }
    __syncthreads();
    if (tid < NL) {
        double t = 0.0;
        for (int e = 0; e < K * 8; ++e) t += lp[e][tid];
        p.loss_partial[(size_t)blockIdx.x * NL + tid] = t;
        __threadfence();
    }
    __syncthreads();
    if (tid == 0) {
        const unsigned prev = atomicAdd(p.counter, 1u);
        is_last = (prev == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        // warp i sums value i over the CTAs: lanes take CTAs lane, lane+32, .. (loads in flight together), fixed order
        if (warp < 2 + K) {
            double t = 0.0;
            for (unsigned b = lane; b < gridDim.x; b += 32) t += __ldcg(p.loss_partial + (size_t)b * NL + warp);
            t = warp_sum(t);
            if (lane == 0) {
                if (warp < 2) p.losses[warp] = (float)(t / C);
                else if (p.disc_beta) p.disc_beta[warp - 2] = (float)(t / C);
            }
        }
        if (tid == 0) *p.counter = 0u;
    }
    trace_exit(TR_ALIGN);
}

int pool_finish_impl(const float* partial_s, int slots_s, const float* partial_t, int slots_t, float* sums_s, float* sums_t,
                     int K, int C, float* stored_s, float* stored_t, int first_s, int first_t, double decay,
                     float w_intra, float w_inter, float* P_s, float* P_t, float* g_s, float* g_t,
                     float* disc_vec, float* disc_beta, float* losses, double* loss_partial, unsigned int* counter,
                     cudaStream_t st) {
    if (!partial_s || !partial_t || !sums_s || !sums_t || !stored_s || !stored_t || !P_s || !P_t || !g_s || !g_t || !losses ||
        !loss_partial || !counter || K < 1 || K > CLR_MAX_K || C < 1 || slots_s < 1 || slots_t < 1)
        return CLR_ERR_BAD_ARG;
    PoolFinishParams p{};
    p.partial[0] = partial_s; p.partial[1] = partial_t; p.sums[0] = sums_s; p.sums[1] = sums_t;
    p.stored[0] = stored_s; p.stored[1] = stored_t; p.P[0] = P_s; p.P[1] = P_t; p.g[0] = g_s; p.g[1] = g_t;
    p.slots[0] = slots_s; p.slots[1] = slots_t; p.first[0] = first_s; p.first[1] = first_t;
    p.K = K; p.C = C; p.d = (float)decay; p.omd = (float)(1.0 - decay); p.w_intra = w_intra; p.w_inter = w_inter;
    p.disc_vec = disc_vec; p.disc_beta = disc_beta; p.losses = losses; p.loss_partial = loss_partial; p.counter = counter;
    clr::launch_k(pool_finish_kernel, (C + 7) / 8, 2 * 2 * K * 32, 0, st, p);
    return launch_status();
}
int pool_finish_max_ctas(int C) { return (C + 7) / 8; }

// CTA = 8 channels; warp w = (class k, slot quarter q) reduces the per-CTA partials of the fused discriminative kernel
// ([slots][K][C+1]) for the CTA's channels, then K*8 threads apply the disc_finalize arithmetic to them.  One extra CTA
// folds the hinge / consistency per-CTA partials into the loss tail and writes the step totals.  No cross-CTA step.
struct DiscFinishParams {
    const float* partial; int slots;
    float* packed2; const float* P_s; float* g_s; float* xtab; float* losses;
    int K, C; double npx;
    float w_disc, ema_factor, gscale, w_intra, w_inter, w_aug, aug_weight;
    int use_cons;
    PackSrc ps;
};

__global__ void __launch_bounds__(1024) disc_finish_kernel(const DiscFinishParams p) {
    kernel_begin(TR_DISC_FIN);
    const int K = p.K, C = p.C, n = K * (C + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int k = warp >> 2, q = warp & 3;             // blockDim.x = K * 4 * 32
    const int ch = lane >> 2, sl0 = lane & 3;
    const int c = blockIdx.x * 8 + ch;
    __shared__ double Sq[CLR_MAX_K][4][8];
    __shared__ double Nq[CLR_MAX_K][4];
    __shared__ double shp[3 * 32];
    const int per = (p.slots + 3) / 4;
    const int s_begin = q * per, s_end = (s_begin + per) < p.slots ? (s_begin + per) : p.slots;
    {
        double s = 0.0;
        if (c < C) s = strided_slot_sum<10>(p.partial + (size_t)k * (C + 1) + c, s_begin + sl0, s_end, (size_t)n, 4);
        double nn = strided_slot_sum<4>(p.partial + (size_t)k * (C + 1) + C, s_begin + lane, s_end, (size_t)n, 32);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        nn = warp_sum(nn);
        if (sl0 == 0) Sq[k][q][ch] = s;
        if (lane == 0) Nq[k][q] = nn;
    }
    // the extra last CTA owns no channels: it folds the hinge + consistency per-CTA partials (uniform branch)
    const bool loss_cta = blockIdx.x == gridDim.x - 1;
    double v[3] = {0.0, 0.0, 0.0};
    if (loss_cta) {
        if (p.ps.hinge)
            for (int i = tid; i < p.ps.n_hinge; i += blockDim.x) v[0] += (double)p.ps.hinge[(size_t)i * p.ps.hinge_stride];
        if (p.ps.cons) {
            const double2* c2 = reinterpret_cast<const double2*>(p.ps.cons);
            int i = tid;
            for (; i + 3 * (int)blockDim.x < p.ps.n_cons; i += 4 * blockDim.x) {
                double2 t[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) t[u] = c2[i + u * blockDim.x];
#pragma unroll
                for (int u = 0; u < 4; ++u) { v[1] += t[u].x; v[2] += t[u].y; }
            }
            for (; i < p.ps.n_cons; i += blockDim.x) { const double2 t = c2[i]; v[1] += t.x; v[2] += t.y; }
        }
        block_sum_n<3>(v, shp);       // contains a __syncthreads
    } else {
        __syncthreads();
    }
    __syncthreads();
    float* tail = p.packed2 + (size_t)K * (C + 1);
    if (tid < K * 8) {
        const int kk = tid >> 3, j = tid & 7, cc = blockIdx.x * 8 + j;
        const float nk = (float)(((Nq[kk][0] + Nq[kk][1]) + Nq[kk][2]) + Nq[kk][3]);
        if (cc < C) {
            const float A = (float)(((Sq[kk][0][j] + Sq[kk][1][j]) + Sq[kk][2][j]) + Sq[kk][3][j]);
            p.packed2[(size_t)kk * (C + 1) + cc] = A;
            const float coef = (float)(2.0 / ((double)C * p.npx));
            const float po = p.P_s[(size_t)kk * C + cc], pb = p.P_s[(size_t)(K + kk) * C + cc];
            p.g_s[(size_t)kk * C + cc] += p.ema_factor * p.w_disc * coef * (nk * po - A);
            p.g_s[(size_t)(K + kk) * C + cc] -= p.ema_factor * p.w_disc * coef * (nk * pb - A);
            p.xtab[(size_t)kk * C + cc] = -p.gscale * p.w_disc * coef * (po - pb);
        }
        if (blockIdx.x == 0 && j == 0) p.packed2[(size_t)kk * (C + 1) + C] = nk;
    }
    if (loss_cta && tid == 0) {
        tail[0] = (float)v[0]; tail[1] = (float)v[1]; tail[2] = (float)v[2]; tail[3] = 0.f;
        const float disc = (float)((double)(float)v[0] / p.npx);
        const float aug = p.use_cons ? (float)((double)(float)v[1] / (double)(float)v[2] * (double)p.aug_weight) : 0.f;
        p.losses[2] = disc;
        p.losses[3] = aug;
        p.losses[4] = p.w_intra * p.losses[0] + p.w_inter * p.losses[1] + p.w_disc * disc + p.w_aug * aug;
        p.losses[5] = 0.f; p.losses[6] = 0.f; p.losses[7] = 0.f;
    }
    trace_exit(TR_DISC_FIN);
}

int disc_finish_impl(const float* partial, int slots, float* packed2, const float* P_s, int K, int C, double npx,
                     float w_disc, float ema_factor, float gscale, float* g_s, float* xtab,
                     float w_intra, float w_inter, float w_aug, float aug_weight, int use_cons, float* losses,
                     const float* hinge, int n_hinge, int hinge_stride, const double* cons, int n_cons, cudaStream_t st) {
    if (!partial || slots < 1 || !packed2 || !P_s || !g_s || !xtab || !losses || K < 1 || K > CLR_MAX_K || C < 1 || npx <= 0)
        return CLR_ERR_BAD_ARG;
    DiscFinishParams p{};
    p.partial = partial; p.slots = slots; p.packed2 = packed2; p.P_s = P_s; p.g_s = g_s; p.xtab = xtab; p.losses = losses;
    p.K = K; p.C = C; p.npx = npx; p.w_disc = w_disc; p.ema_factor = ema_factor; p.gscale = gscale;
    p.w_intra = w_intra; p.w_inter = w_inter; p.w_aug = w_aug; p.aug_weight = aug_weight; p.use_cons = use_cons;
    p.ps = PackSrc{hinge, n_hinge, hinge_stride, cons, n_cons};
    clr::launch_k(disc_finish_kernel, (C + 7) / 8 + 1, K * 4 * 32, 0, st, p);
    return launch_status();
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

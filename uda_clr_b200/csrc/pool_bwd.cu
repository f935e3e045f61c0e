// clr_pool_bwd: adjoint of the class-wise pooling w.r.t. the feature map.
//
//   grad[b,c,p] = scale * sum_r (g[r][c] / N_r) * w_r[b,p]   (+ sum_k xtab[k][c] * xcoef[b,k,p])
//
// This is the closed form of autograd's DivBackward -> SumBackward(expand) -> MulBackward chain for
// utils/Utils.py:114-130: a rank-2K outer product, so the feature map is never
// re-read.  The optional (xtab, xcoef) term is the discriminative hinge's direct gradient
// (d(d_obj - d_bck)/dx = 2 (P_bck - P_obj)/C, independent of x), folded into the same write.
//
// Bound: HBM, write-only.  Algorithmic bytes per domain = 4*B*C*HW (grad) + 4*B*Q*HW (coefficient planes).
//
//   item   = (domain, b, pixel chunk of PX pixels, group of 8 channels) -> 8 rows of PX floats
//   grid   = persistent, contiguous item ranges (so the chunk's coefficient planes are staged in
//            shared memory once per (b, chunk)); the per-channel tables T[q][c] are built once per
//            CTA in shared memory from g and N and read back as warp-wide broadcasts.
//   stores = 128-bit, evict-first (st.global.cs).
#include "clr_common.cuh"

namespace clr {

struct BwdDom {
    const float* w;
    const float* g;       // [R][C]
    const float* sums;    // [R][C+1]
    const float* xcoef;   // [B,Kx,HW] or null
    const float* xtab;    // [Kx][C] or null
    float* grad;
    float scale;
    int fmt, B, Kx, items;
};

struct BwdParams {
    BwdDom dom[2];
    int ndom, C, HW, K, nChunk, nGroup, total;
    int Cpad;   // table row stride (multiple of 4)
};

constexpr int kBwdCG = 8;

template <int VEC>
struct BwdCfg {
    static constexpr int REPS = 2;
    static constexpr int PX = kThreads * VEC * REPS;
};

// Shared memory: T[1 + 3K][Cpad] | coef[3K][PX]
template <int VEC>
__global__ void __launch_bounds__(kThreads, 2) pool_bwd_kernel(const BwdParams p) {
    constexpr int REPS = BwdCfg<VEC>::REPS, PX = BwdCfg<VEC>::PX;
    extern __shared__ __align__(16) float smem[];
    const int K = p.K, R = 2 * K;
    float* T = smem;
    float* coef = smem + (size_t)(1 + 3 * K) * p.Cpad;
    const int tid = threadIdx.x;
    int begin, end;
    partition(p.total, gridDim.x, blockIdx.x, begin, end);

    int cur_dom = -1, cur_slot = -1, Q = 0;
    for (int it = begin; it < end; ++it) {
        const int d = (p.ndom > 1 && it >= p.dom[0].items) ? 1 : 0;
        const BwdDom& D = p.dom[d];
        const int local = it - (d ? p.dom[0].items : 0);
        const int slot = local / p.nGroup, grp = local - slot * p.nGroup;
        const int b = slot / p.nChunk, chunk = slot - b * p.nChunk;
        const int px0 = chunk * PX;
        const int QW = (D.fmt == CLR_W_COMPLEMENT) ? K : R;

        if (d != cur_dom) {
            // ---- per-domain tables: T[0] = constant term, T[1+q] = coefficient of plane q -------
            __syncthreads();
            cur_dom = d; cur_slot = -1;
            Q = QW + D.Kx;
            for (int c = tid; c < p.Cpad; c += kThreads) {
                float t0 = 0.f;
                if (c < p.C) {
                    if (D.fmt == CLR_W_COMPLEMENT) {
                        for (int k = 0; k < K; ++k) {
                            const float go = D.scale * D.g[(size_t)k * p.C + c] / D.sums[(size_t)k * (p.C + 1) + p.C];
                            const float gb = D.scale * D.g[(size_t)(K + k) * p.C + c] / D.sums[(size_t)(K + k) * (p.C + 1) + p.C];
                            t0 += gb;
                            T[(size_t)(1 + k) * p.Cpad + c] = go - gb;
                        }
                    } else {
                        for (int r = 0; r < R; ++r)
                            T[(size_t)(1 + r) * p.Cpad + c] = D.scale * D.g[(size_t)r * p.C + c] / D.sums[(size_t)r * (p.C + 1) + p.C];
                    }
                    for (int k = 0; k < D.Kx; ++k) T[(size_t)(1 + QW + k) * p.Cpad + c] = D.xtab[(size_t)k * p.C + c];
                } else {
                    for (int q = 0; q < Q; ++q) T[(size_t)(1 + q) * p.Cpad + c] = 0.f;
                }
                T[c] = t0;
            }
        }
        if (slot != cur_slot) {
            // ---- stage the chunk's coefficient planes ---------------------------------------------
            __syncthreads();
            cur_slot = slot;
            for (int q = 0; q < Q; ++q) {
                const float* src = (q < QW) ? D.w + ((size_t)b * QW + q) * p.HW
                                            : D.xcoef + ((size_t)b * D.Kx + (q - QW)) * p.HW;
#pragma unroll
                for (int rep = 0; rep < REPS; ++rep) {
                    const int off = (rep * kThreads + tid) * VEC;
                    Pack<VEC> v;
                    if (px0 + off < p.HW) v = ld_keep<VEC>(src + px0 + off);
                    else {
#pragma unroll
                        for (int i = 0; i < VEC; ++i) v.v[i] = 0.f;
                    }
                    st_keep<VEC>(coef + (size_t)q * PX + off, v);
                }
            }
            __syncthreads();
        }

        const int c0 = grp * kBwdCG;
        float* gb = D.grad + ((size_t)b * p.C + c0) * p.HW + px0;
#pragma unroll
        for (int rep = 0; rep < REPS; ++rep) {
            const int off = (rep * kThreads + tid) * VEC;
            if (px0 + off >= p.HW) continue;
            Pack<VEC> out[kBwdCG];
            {
                const float4 a = *reinterpret_cast<const float4*>(T + c0);
                const float4 bq = *reinterpret_cast<const float4*>(T + c0 + 4);
                const float t[8] = {a.x, a.y, a.z, a.w, bq.x, bq.y, bq.z, bq.w};
#pragma unroll
                for (int j = 0; j < kBwdCG; ++j)
#pragma unroll
                    for (int i = 0; i < VEC; ++i) out[j].v[i] = t[j];
            }
            for (int q = 0; q < Q; ++q) {
                Pack<VEC> cf;
                if constexpr (VEC == 4) {
                    const float4 t4 = *reinterpret_cast<const float4*>(coef + (size_t)q * PX + off);
                    cf.v[0] = t4.x; cf.v[1] = t4.y; cf.v[2] = t4.z; cf.v[3] = t4.w;
                } else {
                    cf.v[0] = coef[(size_t)q * PX + off];
                }
                const float* Tq = T + (size_t)(1 + q) * p.Cpad + c0;
                const float4 a = *reinterpret_cast<const float4*>(Tq);
                const float4 bq = *reinterpret_cast<const float4*>(Tq + 4);
                const float t[8] = {a.x, a.y, a.z, a.w, bq.x, bq.y, bq.z, bq.w};
#pragma unroll
                for (int j = 0; j < kBwdCG; ++j)
#pragma unroll
                    for (int i = 0; i < VEC; ++i) out[j].v[i] = fmaf(t[j], cf.v[i], out[j].v[i]);
            }
#pragma unroll
            for (int j = 0; j < kBwdCG; ++j)
                if (c0 + j < p.C) st_stream<VEC>(gb + (size_t)j * p.HW + off, out[j]);
        }
    }
}

int pool_bwd_impl(const BwdDom* doms, int ndom, int C, int HW, int K, cudaStream_t st) {
    CLR_CHECK_ARG(ndom >= 1 && ndom <= 2 && C > 0 && HW > 0 && K >= 1 && K <= CLR_MAX_K);
    bool vec4 = (HW % 4 == 0);
    for (int d = 0; d < ndom; ++d) {
        const BwdDom& D = doms[d];
        CLR_CHECK_ARG(D.w && D.g && D.sums && D.grad && D.B > 0);
        CLR_CHECK_ARG(D.fmt == CLR_W_COMPLEMENT || D.fmt == CLR_W_EXPLICIT);
        CLR_CHECK_ARG(D.Kx >= 0 && D.Kx <= K && (D.Kx == 0 || (D.xcoef && D.xtab)));
        if (!aligned4(D.w) || !aligned4(D.grad)) return CLR_ERR_ALIGN;
        vec4 = vec4 && aligned16(D.w) && aligned16(D.grad) && (D.Kx == 0 || aligned16(D.xcoef));
    }
    BwdParams p{};
    p.ndom = ndom; p.C = C; p.HW = HW; p.K = K;
    const int px = vec4 ? BwdCfg<4>::PX : BwdCfg<1>::PX;
    p.nChunk = (HW + px - 1) / px;
    p.nGroup = (C + kBwdCG - 1) / kBwdCG;
    p.Cpad = p.nGroup * kBwdCG;
    long long total = 0;
    for (int d = 0; d < ndom; ++d) {
        p.dom[d] = doms[d];
        const long long items = (long long)doms[d].B * p.nChunk * p.nGroup;
        if (items > 0x3fffffff) return CLR_ERR_UNSUPPORTED;
        p.dom[d].items = (int)items;
        total += items;
    }
    if (total > 0x3fffffff) return CLR_ERR_UNSUPPORTED;
    p.total = (int)total;
    const size_t smem = sizeof(float) * ((size_t)(1 + 3 * K) * p.Cpad + (size_t)3 * K * px);
    if (smem > (size_t)device_facts().max_smem_optin) return CLR_ERR_UNSUPPORTED;
    auto kern = vec4 ? pool_bwd_kernel<4> : pool_bwd_kernel<1>;
    CLR_RETURN_IF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CLR_RETURN_IF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, smem));
    if (occ < 1) occ = 1;
    int grid = device_facts().sms * occ;
    if (grid > p.total) grid = p.total;
    kern<<<grid, kThreads, smem, st>>>(p);
    return launch_status();
}

}  // namespace clr

extern "C" {

int clr_pool_bwd(const float* w, int fmt, int B, int C, int HW, int K,
                 const float* g, const float* sums, float scale,
                 const float* xcoef, const float* xtab, int Kx,
                 float* grad, clr_stream_t stream) {
    clr::BwdDom d{w, g, sums, xcoef, xtab, grad, scale, fmt, B, Kx, 0};
    return clr::pool_bwd_impl(&d, 1, C, HW, K, static_cast<cudaStream_t>(stream));
}

int clr_pool_bwd2(const float* w0, int fmt0, int B0, const float* g0, const float* sums0, float scale0,
                  const float* xcoef0, const float* xtab0, int Kx0, float* grad0,
                  const float* w1, int fmt1, int B1, const float* g1, const float* sums1, float scale1,
                  float* grad1, int C, int HW, int K, clr_stream_t stream) {
    clr::BwdDom d[2] = {{w0, g0, sums0, xcoef0, xtab0, grad0, scale0, fmt0, B0, Kx0, 0},
                        {w1, g1, sums1, nullptr, nullptr, grad1, scale1, fmt1, B1, 0, 0}};
    return clr::pool_bwd_impl(d, 2, C, HW, K, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

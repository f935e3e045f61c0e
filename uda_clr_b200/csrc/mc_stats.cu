// MC-dropout statistics and the retrify weights of gen_prototype_retrify (utils/Utils.py:159-223).
//
// clr_mc_stats      : preds [T*B,K,Hi,Wi] logits -> std_map = std_T(sigmoid(p/2)) (unbiased) and
//                     prediction = mean_T(sigmoid(p)), both [B,K,Hi,Wi]  (utils/Utils.py:164-168).
//                     One read of preds, two small writes.  The reference also averages the
//                     [T*B,305,128,128] feature stack (:169, 1.28 GB) and uses only its size: not read here.
// clr_retrify_weights: bilinear(align_corners) 4-tap gathers of both maps at feature resolution
//                     (:170-171), pseudo = sigmoid(oT_before) > 0.75 (:173-177), mask = std_small < 0.04
//                     (:188-200) -> explicit weight planes [B,2K,H,W] (:207-223) + mask_k in {0,2} (:205-206).
//
// Integer-valued decisions follow ATen's CUDA arithmetic: sigmoid = 1/(1+expf(-x)) in fp32 with IEEE
// division (no fast-math), thresholds compared in fp32.
#include "clr_common.cuh"

namespace clr {

__device__ __forceinline__ float sigmoid_aten(float x) { return 1.0f / (1.0f + expf(-x)); }

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
        s_half = r * b;
        s_full = r * a;
    }
}

template <int VEC, int TT, bool PRECISE>
__global__ void __launch_bounds__(256) mc_stats_kernel(const float* __restrict__ preds, int T, size_t n,
                                                       float* __restrict__ std_map, float* __restrict__ pred_mean) {
    kernel_begin(TR_MC_STATS);
    // n = B*K*Hi*Wi positions; preds is [T][n].  TT > 0: T <= TT, values kept in registers (two-pass
    // variance); TT == 0: any T, Welford.
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (i >= n) { trace_exit(TR_MC_STATS); return; }
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
        for (int v = 0; v < VEC; ++v) mean_h[v] *= invT;
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
    Pack<VEC> s, m;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        s.v[v] = sqrtf(m2[v] / (float)(T - 1));   // unbiased (torch.std default, :166); T = 1 -> NaN like torch
        m.v[v] = mean_f[v] / (float)T;            // :168
    }
    st_keep<VEC>(std_map + i, s);
    st_keep<VEC>(pred_mean + i, m);
    trace_exit(TR_MC_STATS);
}

template <int VEC, bool PRECISE>
static void launch_mc(const float* preds, int T, size_t n, float* std_map, float* pred_mean, cudaStream_t st) {
    const size_t threads = (n + VEC - 1) / VEC;
    const unsigned blocks = (unsigned)((threads + 255) / 256);
    if (T <= 8) launch_k(mc_stats_kernel<VEC, 8, PRECISE>, blocks, 256, 0, st, preds, T, n, std_map, pred_mean);
    else if (T <= 16) launch_k(mc_stats_kernel<VEC, 16, PRECISE>, blocks, 256, 0, st, preds, T, n, std_map, pred_mean);
    else launch_k(mc_stats_kernel<VEC, 0, PRECISE>, blocks, 256, 0, st, preds, T, n, std_map, pred_mean);
}

// upsample_bilinear2d (align_corners=True) source coordinates, as ATen computes them in fp32
struct Tap { int i0, i1; float l0, l1; };
__device__ __forceinline__ Tap bilinear_tap(int dst, int n_in, float scale) {
    Tap t;
    const float src = scale * (float)dst;
    t.i0 = (int)src;
    t.i1 = t.i0 + ((t.i0 < n_in - 1) ? 1 : 0);
    t.l1 = src - (float)t.i0;
    t.l0 = 1.0f - t.l1;
    return t;
}
__device__ __forceinline__ float bilinear_at(const float* __restrict__ plane, int Wi, const Tap& h, const Tap& w) {
    const float* r0 = plane + (size_t)h.i0 * Wi;
    const float* r1 = plane + (size_t)h.i1 * Wi;
    return h.l0 * (w.l0 * __ldg(r0 + w.i0) + w.l1 * __ldg(r0 + w.i1)) +
           h.l1 * (w.l0 * __ldg(r1 + w.i0) + w.l1 * __ldg(r1 + w.i1));
}

__global__ void __launch_bounds__(256) retrify_weights_kernel(
    const float* __restrict__ oT_before, const float* __restrict__ pred_mean, const float* __restrict__ std_map,
    int B, int K, int H, int W, int Hi, int Wi, float pseudo_thr, float std_thr,
    float* __restrict__ weights /*[B,2K,H,W]*/, float* __restrict__ masks /*[B,K,H,W]*/,
    float* __restrict__ pseudo_out /*[B,K,H,W] or null*/, float* __restrict__ small_out /*[2][B,K,H,W] or null*/) {
    kernel_begin(TR_RETRIFY);
    // grid.y = b*K + k (one plane), grid.x covers the plane: no 64-bit divisions
    const size_t n = (size_t)B * K * H * W;
    const int pixi = blockIdx.x * blockDim.x + threadIdx.x;
    if (pixi >= H * W) { trace_exit(TR_RETRIFY); return; }
    const int bk = blockIdx.y;
    const int b = bk / K, k = bk - b * K;
    const int y = pixi / W, x = pixi - y * W;
    const size_t i = (size_t)bk * H * W + pixi;
    const float sh = H > 1 ? (float)(Hi - 1) / (float)(H - 1) : 0.f;
    const float sw = W > 1 ? (float)(Wi - 1) / (float)(W - 1) : 0.f;
    const Tap th = bilinear_tap(y, Hi, sh), tw = bilinear_tap(x, Wi, sw);
    const size_t plane = ((size_t)b * K + k) * Hi * Wi;
    const float ps = bilinear_at(pred_mean + plane, Wi, th, tw);
    const float ss = bilinear_at(std_map + plane, Wi, th, tw);
    const bool pseudo = sigmoid_aten(oT_before[i]) > pseudo_thr;
    const bool m = ss < std_thr;
    const size_t hw = (size_t)H * W, pix = (size_t)y * W + x;
    weights[((size_t)b * 2 * K + k) * hw + pix] = (pseudo && m) ? ps : 0.f;
    weights[((size_t)b * 2 * K + K + k) * hw + pix] = (!pseudo && m) ? (1.0f - ps) : 0.f;
    masks[i] = m ? 2.0f : 0.f;
    if (pseudo_out) pseudo_out[i] = pseudo ? 1.0f : 0.f;
    if (small_out) { small_out[i] = ps; small_out[n + i] = ss; }
    trace_exit(TR_RETRIFY);
}

}  // namespace clr

extern "C" {

int clr_mc_stats(const float* preds, int T, int B, int K, int Hi, int Wi,
                 float* std_map, float* pred_mean, clr_stream_t stream) {
    if (!preds || !std_map || !pred_mean || T < 1 || B < 1 || K < 1 || Hi < 1 || Wi < 1) return CLR_ERR_BAD_ARG;
    const size_t n = (size_t)B * K * Hi * Wi;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool vec4 = (n % 4 == 0) && clr::aligned16(preds) && clr::aligned16(std_map) && clr::aligned16(pred_mean);
    const bool precise = clr::tunables().mc_precise != 0;
    if (vec4) {
        if (precise) clr::launch_mc<4, true>(preds, T, n, std_map, pred_mean, st);
        else clr::launch_mc<4, false>(preds, T, n, std_map, pred_mean, st);
    } else {
        if (precise) clr::launch_mc<1, true>(preds, T, n, std_map, pred_mean, st);
        else clr::launch_mc<1, false>(preds, T, n, std_map, pred_mean, st);
    }
    return clr::launch_status();
}

int clr_retrify_weights(const float* oT_before, const float* pred_mean, const float* std_map,
                        int B, int K, int H, int W, int Hi, int Wi, float pseudo_thr, float std_thr,
                        float* weights, float* masks, float* pseudo_out, float* small_out, clr_stream_t stream) {
    if (!oT_before || !pred_mean || !std_map || !weights || !masks || B < 1 || K < 1 || K > CLR_MAX_K ||
        H < 1 || W < 1 || Hi < 1 || Wi < 1)
        return CLR_ERR_BAD_ARG;
    if ((long long)H * W > 0x7fffff00LL || (long long)B * K > 65535) return CLR_ERR_UNSUPPORTED;
    clr::launch_k(clr::retrify_weights_kernel, dim3((unsigned)((H * W + 255) / 256), (unsigned)(B * K)), 256, 0, static_cast<cudaStream_t>(stream), oT_before, pred_mean, std_map, B, K, H, W, Hi, Wi, pseudo_thr, std_thr, weights, masks, pseudo_out, small_out);
    return clr::launch_status();
}

}  // extern "C"

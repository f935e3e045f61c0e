// Small kernels that complete rows of SURVEY.md 8(a)/(f) around the hot path:
//
//   clr_ema_rows          A7: `objective_vectors[id] = obj * (1 - 0.001) + 0.001 * v` unless `v.sum() == 0`
//                         (Trainer_prototype.py:117-123), R stored vectors per launch, the zero test on the device
//                         (the reference pays a `.item()` host sync per vector).
//   clr_label_downsample  8(f) rank 2, third op: `F.interpolate(target_map, size=(H, W), mode='nearest')`
//                         (Trainer_prototype_full.py:329-330) -- ATen's source index floor(dst * scale) with the fp32
//                         scale in/out, min'ed to the last element.
//   clr_mc_accumulate /   8(f) rank 1: MC statistics WITHOUT the [T*B,K,Hi,Wi] staging buffer of the trainer's loop
//   clr_mc_finalize       (Trainer_prototype_full.py:359-368 fills preds_trg and a dead 1.28 GB features_trg): every MC
//                         forward hands its logits to clr_mc_accumulate, which keeps 4 running maps per position --
//                         pivot = sigmoid(p_0/2) of the first pass, sum (x - pivot), sum (x - pivot)^2 of
//                         x = sigmoid(p/2), and sum sigmoid(p) -- and clr_mc_finalize turns them into std_map
//                         (unbiased) and the mean prediction.  Shifted sums: the variance (std ~ 0.04 around values
//                         ~ 0.5) would lose 3 digits in plain sum / sum-of-squares form.
//                         Traffic: T * (Li + 8 Li) here against T * Li + 2 Li for one read of staged logits -- the gain
//                         is memory (no staging buffer), not time; measured in profiles/r02_*.  The knife-edge guard of
//                         the uncertainty mask needs the raw logits and is not available on this path.
#include "clr_common.cuh"
#include "clr_internal.h"

namespace clr {

// ---------------------------------------------------------------------------------------------------------------- A7
__global__ void __launch_bounds__(256) ema_rows_kernel(const float* __restrict__ v, const float* __restrict__ stored, int C,
                                                       float keep, float rate, float* __restrict__ out) {
    kernel_begin(TR_OTHER);
    const int r = blockIdx.x;
    const float* vr = v + (size_t)r * C;
    double s = 0.0;
    for (int c = threadIdx.x; c < C; c += blockDim.x) s += (double)vr[c];
    s = warp_sum(s);
    __shared__ double sh[8];
    __shared__ int skip;
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
        skip = (t == 0.0) ? 1 : 0;                       // `if vector.sum().item() == 0: return` (:118-119)
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float o = stored[(size_t)r * C + c];
        // obj * (1 - 0.001) + 0.001 * v in ATen's fp32 order: two rounded products, one rounded sum (:121)
        out[(size_t)r * C + c] = skip ? o : __fadd_rn(__fmul_rn(o, keep), __fmul_rn(rate, vr[c]));
    }
}

// ------------------------------------------------------------------------------------------------ nearest down-sample
__global__ void __launch_bounds__(256) label_downsample_kernel(const float* __restrict__ src, int planes, int Hi, int Wi, int H, int W,
                                                              float sh, float sw, float* __restrict__ dst) {
    kernel_begin(TR_OTHER);
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    const int plane = blockIdx.y;
    if (pix >= H * W) return;
    const int y = pix / W, x = pix - y * W;
    // ATen upsample_nearest2d: src = min(floor(dst * scale), in - 1), scale = (float)in / out
    int sy = (int)floorf((float)y * sh), sx = (int)floorf((float)x * sw);
    sy = sy < Hi - 1 ? sy : Hi - 1;
    sx = sx < Wi - 1 ? sx : Wi - 1;
    dst[(size_t)plane * H * W + pix] = __ldg(src + ((size_t)plane * Hi + sy) * Wi + sx);
}

// ------------------------------------------------------------------------------------------------ MC accumulation
__device__ __forceinline__ void acc_sigmoids(float p, float& s_half, float& s_full) {
    // one exponential for both sigmoids (as clr_mc_stats): u = e^{-p/2}; sigmoid(p/2) = 1/(1+u), sigmoid(p) = 1/(1+u^2)
    float u, r;
    const float pc = fmaxf(p, -55.0f);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(u) : "f"(pc * -0.72134752044448170368f));
    const float a = 1.0f + u, b = fmaf(u, u, 1.0f);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a * b));
    s_half = r * b;
    s_full = r * a;
}

// state = [4][n]: pivot | sum d | sum d^2 | sum sigmoid(p);  logits = [passes][n]
template <int VEC>
__global__ void __launch_bounds__(256) mc_accumulate_kernel(const float* __restrict__ logits, int passes, size_t n, int first,
                                                            float* __restrict__ state) {
    kernel_begin(TR_OTHER);
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (i >= n) return;
    Pack<VEC> pv, s1, s2, sf;
    if (first) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) { s1.v[v] = 0.f; s2.v[v] = 0.f; sf.v[v] = 0.f; }
    } else {
        pv = ld_stream<VEC>(state + i);
        s1 = ld_stream<VEC>(state + n + i);
        s2 = ld_stream<VEC>(state + 2 * n + i);
        sf = ld_stream<VEC>(state + 3 * n + i);
    }
    for (int t = 0; t < passes; ++t) {
        const Pack<VEC> x = ld_stream<VEC>(logits + (size_t)t * n + i);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            float sh, sfl;
            acc_sigmoids(x.v[v], sh, sfl);
            if (first && t == 0) pv.v[v] = sh;
            const float d = sh - pv.v[v];
            s1.v[v] += d;
            s2.v[v] = fmaf(d, d, s2.v[v]);
            sf.v[v] += sfl;
        }
    }
    if (first) st_keep<VEC>(state + i, pv);
    st_keep<VEC>(state + n + i, s1);
    st_keep<VEC>(state + 2 * n + i, s2);
    st_keep<VEC>(state + 3 * n + i, sf);
}

template <int VEC>
__global__ void __launch_bounds__(256) mc_finalize_kernel(const float* __restrict__ state, int T, size_t n,
                                                          float* __restrict__ std_map, float* __restrict__ pred_mean) {
    kernel_begin(TR_OTHER);
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (i >= n) return;
    const Pack<VEC> s1 = ld_stream<VEC>(state + n + i), s2 = ld_stream<VEC>(state + 2 * n + i), sf = ld_stream<VEC>(state + 3 * n + i);
    Pack<VEC> sd, mn;
    const float invT = 1.0f / (float)T;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        // sum (x - mean)^2 = sum d^2 - (sum d)^2 / T  (d = x - pivot: both terms are O(T * var), no cancellation of O(1) values)
        const float m2 = fmaxf(s2.v[v] - s1.v[v] * s1.v[v] * invT, 0.f);
        sd.v[v] = sqrtf(m2 / (float)(T - 1));                 // unbiased (torch.std default, utils/Utils.py:166); T = 1 -> NaN
        mn.v[v] = sf.v[v] * invT;                             // :168
    }
    st_keep<VEC>(std_map + i, sd);
    st_keep<VEC>(pred_mean + i, mn);
}

}  // namespace clr

extern "C" {

int clr_ema_rows(const float* v, const float* stored, int R, int C, float rate, float* out, clr_stream_t stream) {
    if (!v || !stored || !out || R < 1 || C < 1 || R > 65535) return CLR_ERR_BAD_ARG;
    // Python evaluates (1 - 0.001) in double; ATen then multiplies by that scalar cast to fp32
    const float keep = (float)(1.0 - (double)rate);
    clr::launch_k(clr::ema_rows_kernel, R, 256, 0, static_cast<cudaStream_t>(stream), v, stored, C, keep, rate, out);
    return clr::launch_status();
}

int clr_label_downsample(const float* src, int planes, int Hi, int Wi, int H, int W, float* dst, clr_stream_t stream) {
    if (!src || !dst || planes < 1 || Hi < 1 || Wi < 1 || H < 1 || W < 1) return CLR_ERR_BAD_ARG;
    if (planes > 65535 || (long long)H * W > 0x7fffff00LL) return CLR_ERR_UNSUPPORTED;
    const float sh = (float)Hi / (float)H, sw = (float)Wi / (float)W;
    clr::launch_k(clr::label_downsample_kernel, dim3((unsigned)((H * W + 255) / 256), (unsigned)planes), 256, 0,
                  static_cast<cudaStream_t>(stream), src, planes, Hi, Wi, H, W, sh, sw, dst);
    return clr::launch_status();
}

size_t clr_mc_state_floats(int B, int K, int Hi, int Wi) {
    if (B < 1 || K < 1 || Hi < 1 || Wi < 1) return 0;
    return (size_t)4 * B * K * Hi * Wi;
}

int clr_mc_accumulate(const float* logits, int passes, int B, int K, int Hi, int Wi, int first, float* state, clr_stream_t stream) {
    if (!logits || !state || passes < 1 || B < 1 || K < 1 || Hi < 1 || Wi < 1) return CLR_ERR_BAD_ARG;
    const size_t n = (size_t)B * K * Hi * Wi;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool vec4 = (n % 4 == 0) && clr::aligned16(logits) && clr::aligned16(state);
    if (vec4) clr::launch_k(clr::mc_accumulate_kernel<4>, (unsigned)((n / 4 + 255) / 256), 256, 0, st, logits, passes, n, first, state);
    else clr::launch_k(clr::mc_accumulate_kernel<1>, (unsigned)((n + 255) / 256), 256, 0, st, logits, passes, n, first, state);
    return clr::launch_status();
}

int clr_mc_finalize(const float* state, int T, int B, int K, int Hi, int Wi, float* std_map, float* pred_mean, clr_stream_t stream) {
    if (!state || !std_map || !pred_mean || T < 1 || B < 1 || K < 1 || Hi < 1 || Wi < 1) return CLR_ERR_BAD_ARG;
    const size_t n = (size_t)B * K * Hi * Wi;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool vec4 = (n % 4 == 0) && clr::aligned16(state) && clr::aligned16(std_map) && clr::aligned16(pred_mean);
    if (vec4) clr::launch_k(clr::mc_finalize_kernel<4>, (unsigned)((n / 4 + 255) / 256), 256, 0, st, state, T, n, std_map, pred_mean);
    else clr::launch_k(clr::mc_finalize_kernel<1>, (unsigned)((n + 255) / 256), 256, 0, st, state, T, n, std_map, pred_mean);
    return clr::launch_status();
}

}  // extern "C"

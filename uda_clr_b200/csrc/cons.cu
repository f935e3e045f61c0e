// Augmented-consistency loss (Trainer_prototype_mt.cpython-38.pyc L502-561, restated from the bytecode):
//
//   y    = sigmoid(oT) > thr                      (thr = (0.85 + 0.25*sigmoid_rampup(epoch,200)) * ln 2, fp32 compare)
//   l    = BCELoss(reduction='none')(sigmoid(oT_aug), y)      (log terms clamped at -100, as ATen)
//   m    = nearest_upsample(cat(mask_0, mask_1))  ([B,K,H,W] in {0,2} -> [B,K,Hi,Wi])
//   loss = sum(m*l) / sum(m) * aug_weight
//
// clr_cons_fwd: one read of oT, oT_aug (+ the small masks), per-CTA fp64 partials, fixed-order final sum.
// clr_cons_bwd: d loss / d oT_aug = gscale * aug_weight * m/sum(m) * (q-y)/max(q(1-q),1e-12) * q(1-q)
//               (ATen's binary_cross_entropy_backward chained with sigmoid').  Elementwise, one write.
// Bound: HBM; algorithmic bytes fwd = 2*4*B*K*Hi*Wi (+masks), bwd = 2 reads + 1 write of the same size.
#include "clr_common.cuh"
#include "clr_internal.h"

namespace clr {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

struct ConsGeom {
    int B, K, Hi, Wi, H, W;
    float sh, sw;   // nearest scales: (float)H/Hi, (float)W/Wi
};

__device__ __forceinline__ float mask_at(const float* __restrict__ masks, const ConsGeom& g, size_t i) {
    const int x = (int)(i % g.Wi);
    const int y = (int)((i / g.Wi) % g.Hi);
    const size_t bk = i / ((size_t)g.Wi * g.Hi);
    int sy = (int)floorf((float)y * g.sh); if (sy > g.H - 1) sy = g.H - 1;
    int sx = (int)floorf((float)x * g.sw); if (sx > g.W - 1) sx = g.W - 1;
    return __ldg(masks + (bk * g.H + sy) * g.W + sx);
}

__global__ void __launch_bounds__(256) cons_fwd_kernel(const float* __restrict__ oT, const float* __restrict__ oT_aug,
                                                       const float* __restrict__ masks, ConsGeom g, float thr,
                                                       size_t n, double* __restrict__ partial) {
    double num = 0.0, den = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float m = mask_at(masks, g, i);
        const float y = sigmoidf_(__ldg(oT + i)) > thr ? 1.0f : 0.0f;
        const float q = sigmoidf_(__ldg(oT_aug + i));
        const float lq = fmaxf(logf(q), -100.0f);
        const float l1q = fmaxf(log1pf(-q), -100.0f);
        const float l = (y - 1.0f) * l1q - y * lq;
        num += (double)(m * l);
        den += (double)m;
    }
    num = warp_sum(num);
    den = warp_sum(den);
    __shared__ double sh[2][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sh[0][warp] = num; sh[1][warp] = den; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < 8; ++w) { a += sh[0][w]; b += sh[1][w]; }
        partial[2 * blockIdx.x] = a;
        partial[2 * blockIdx.x + 1] = b;
    }
}

__global__ void cons_final_kernel(const double* __restrict__ partial, int nblk, float aug_weight, float* __restrict__ stats) {
    // one warp, fixed order
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < nblk; i += 32) { a += partial[2 * i]; b += partial[2 * i + 1]; }
    a = warp_sum(a);
    b = warp_sum(b);
    if (threadIdx.x == 0) {
        stats[0] = (float)a;
        stats[1] = (float)b;
        stats[2] = (float)(a / b * (double)aug_weight);
        stats[3] = 0.f;
    }
}

__global__ void __launch_bounds__(256) cons_bwd_kernel(const float* __restrict__ oT, const float* __restrict__ oT_aug,
                                                       const float* __restrict__ masks, ConsGeom g, float thr,
                                                       const float* __restrict__ stats, const float* __restrict__ gscale_dev,
                                                       float gscale, size_t n, float* __restrict__ grad) {
    const float coef = (gscale_dev ? gscale * __ldg(gscale_dev) : gscale) / __ldg(stats + 1);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float m = mask_at(masks, g, i);
        const float y = sigmoidf_(__ldg(oT + i)) > thr ? 1.0f : 0.0f;
        const float q = sigmoidf_(__ldg(oT_aug + i));
        const float qq = (1.0f - q) * q;
        const float gq = coef * m * (q - y) / fmaxf(qq, 1e-12f);
        grad[i] = gq * qq;
    }
}

static ConsGeom make_geom(int B, int K, int Hi, int Wi, int H, int W) {
    ConsGeom g{B, K, Hi, Wi, H, W, (float)H / (float)Hi, (float)W / (float)Wi};
    return g;
}

constexpr int kConsMaxBlocks = 148 * 8;

int cons_fwd_partials(const float* oT, const float* oT_aug, const float* masks, int B, int K, int Hi, int Wi,
                      int H, int W, float threshold, double* partial, int* nblocks, cudaStream_t st) {
    const size_t n = (size_t)B * K * Hi * Wi;
    const int blocks = (int)((n + 255) / 256 < (size_t)kConsMaxBlocks ? (n + 255) / 256 : (size_t)kConsMaxBlocks);
    clr::count_launch(); cons_fwd_kernel<<<blocks, 256, 0, st>>>(oT, oT_aug, masks, make_geom(B, K, Hi, Wi, H, W), threshold, n, partial);
    *nblocks = blocks;
    return launch_status();
}

}  // namespace clr

extern "C" {

size_t clr_cons_ws_bytes(void) { return sizeof(double) * 2 * clr::kConsMaxBlocks; }

int clr_cons_fwd(const float* oT, const float* oT_aug, const float* masks, int B, int K, int Hi, int Wi, int H, int W,
                 float threshold, float aug_weight, void* ws, size_t ws_bytes, float* stats, clr_stream_t stream) {
    if (!oT || !oT_aug || !masks || !ws || !stats || B < 1 || K < 1 || Hi < 1 || Wi < 1 || H < 1 || W < 1) return CLR_ERR_BAD_ARG;
    if (ws_bytes < clr_cons_ws_bytes()) return CLR_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int blocks = 0;
    const int rc = clr::cons_fwd_partials(oT, oT_aug, masks, B, K, Hi, Wi, H, W, threshold, static_cast<double*>(ws), &blocks, st);
    if (rc != CLR_OK) return rc;
    clr::count_launch(); clr::cons_final_kernel<<<1, 32, 0, st>>>(static_cast<const double*>(ws), blocks, aug_weight, stats);
    return clr::launch_status();
}

int clr_cons_bwd(const float* oT, const float* oT_aug, const float* masks, int B, int K, int Hi, int Wi, int H, int W,
                 float threshold, float aug_weight, const float* stats, const float* gscale_dev, float gscale,
                 float* grad_oT_aug, clr_stream_t stream) {
    if (!oT || !oT_aug || !masks || !stats || !grad_oT_aug || B < 1 || K < 1 || Hi < 1 || Wi < 1 || H < 1 || W < 1)
        return CLR_ERR_BAD_ARG;
    const size_t n = (size_t)B * K * Hi * Wi;
    int blocks = (int)((n + 255) / 256 < (size_t)clr::kConsMaxBlocks * 2 ? (n + 255) / 256 : (size_t)clr::kConsMaxBlocks * 2);
    clr::count_launch(); clr::cons_bwd_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        oT, oT_aug, masks, clr::make_geom(B, K, Hi, Wi, H, W), threshold, stats, gscale_dev, gscale * aug_weight, n, grad_oT_aug);
    return clr::launch_status();
}

}  // extern "C"

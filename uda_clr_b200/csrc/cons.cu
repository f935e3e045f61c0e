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

__device__ __forceinline__ float sigmoid_aten(float x) { return 1.0f / (1.0f + expf(-x)); }   // label decision: exact
__device__ __forceinline__ float sigmoid_fast(float x) { return __frcp_rn(1.0f + __expf(-x)); }

struct ConsGeom {
    int BK, Hi, Wi, H, W;
    float sh, sw;   // nearest scales: (float)H/Hi, (float)W/Wi
    int rows_per_cta;
};

// F.interpolate(mode='nearest') source index: min(floor(dst * scale), in - 1)
__device__ __forceinline__ int nearest_src(int dst, float scale, int n_in) {
    const int s = (int)floorf((float)dst * scale);
    return s < n_in - 1 ? s : n_in - 1;
}

// Exact label decision sigmoid(z) > thr at the cost of the fast sigmoid: the fast value (error < 1e-6) decides
// unless it lies within 1e-5 of the threshold, in which case ATen's exact expression is evaluated.
__device__ __forceinline__ bool label_above(float z, float thr) {
    const float qf = sigmoid_fast(z);
    if (fabsf(qf - thr) > 1e-5f) return qf > thr;
    return sigmoid_aten(z) > thr;
}

// One element of the masked BCE: returns m*l and m; `y` decided exactly, the log terms through fast intrinsics.
__device__ __forceinline__ void cons_elem(float zt, float za, float m, float thr, float& ml, float& q_out, float& y_out) {
    const float y = label_above(zt, thr) ? 1.0f : 0.0f;
    const float q = sigmoid_fast(za);
    const float lq = fmaxf(__logf(q), -100.0f);
    const float l1q = fmaxf(__logf(1.0f - q), -100.0f);   // ATen: log1p(-q); identical after fp32 rounding of q
    ml = m * ((y - 1.0f) * l1q - y * lq);
    q_out = q; y_out = y;
}

// grid = (row blocks, B*K planes); each CTA covers `rows_per_cta` image rows of one plane.
template <int VEC, bool BWD>
__global__ void __launch_bounds__(256) cons_kernel(const float* __restrict__ oT, const float* __restrict__ oT_aug,
                                                   const float* __restrict__ masks, ConsGeom g, float thr,
                                                   double* __restrict__ partial,
                                                   const float* __restrict__ stats, const float* __restrict__ gscale_dev,
                                                   float gscale, float* __restrict__ grad) {
    pdl_wait();
    const int bk = blockIdx.y;
    const int y0 = blockIdx.x * g.rows_per_cta;
    const int wv = g.Wi / VEC;                       // vectors per row
    const int nvec = g.rows_per_cta * wv;
    const size_t plane = (size_t)bk * g.Hi * g.Wi;
    const float* mplane = masks + (size_t)bk * g.H * g.W;
    float coef = 0.f;
    if (BWD) coef = (gscale_dev ? gscale * __ldg(gscale_dev) : gscale) / __ldg(stats + 1);
    float num = 0.f, den = 0.f;
    constexpr int U = 4;   // independent vector loads in flight per thread
    for (int base = threadIdx.x; base < nvec; base += U * blockDim.x) {
        Pack<VEC> zt[U], za[U];
        int yy[U], xv_[U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int idx = base + u * blockDim.x;
            const int ry = idx / wv;
            xv_[u] = idx - ry * wv;
            yy[u] = y0 + ry;
            ok[u] = idx < nvec && yy[u] < g.Hi;
            if (ok[u]) {
                const size_t off = plane + (size_t)yy[u] * g.Wi + (size_t)xv_[u] * VEC;
                zt[u] = ld_stream<VEC>(oT + off);
                za[u] = ld_stream<VEC>(oT_aug + off);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!ok[u]) continue;
            const int sy = nearest_src(yy[u], g.sh, g.H);
            const float* mrow = mplane + (size_t)sy * g.W;
            // exact VEC:1 upsampling (the reference's 512 -> 128 case): the whole vector shares one mask pixel
            const bool shared_mask = (g.Wi == VEC * g.W);
            const float m_shared = shared_mask ? __ldg(mrow + xv_[u]) : 0.f;
            Pack<VEC> go;
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const float m = shared_mask ? m_shared : __ldg(mrow + nearest_src(xv_[u] * VEC + v, g.sw, g.W));
                float ml, q, yv;
                cons_elem(zt[u].v[v], za[u].v[v], m, thr, ml, q, yv);
                num += ml; den += m;
                if (BWD) {
                    // ATen binary_cross_entropy_backward: (q - y) / max((1-q) q, 1e-12), chained with sigmoid' = q (1-q)
                    const float qq = (1.0f - q) * q;
                    go.v[v] = coef * m * (q - yv) / fmaxf(qq, 1e-12f) * qq;
                }
            }
            if (BWD) st_stream<VEC>(grad + plane + (size_t)yy[u] * g.Wi + (size_t)xv_[u] * VEC, go);
        }
    }
    if (!BWD) {
        double dn = warp_sum((double)num), dd = warp_sum((double)den);
        __shared__ double sh[2][8];
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (lane == 0) { sh[0][warp] = dn; sh[1][warp] = dd; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double a = 0.0, b = 0.0;
            for (int w = 0; w < 8; ++w) { a += sh[0][w]; b += sh[1][w]; }
            const int cta = blockIdx.y * gridDim.x + blockIdx.x;
            partial[2 * cta] = a;
            partial[2 * cta + 1] = b;
        }
    }
}

__global__ void __launch_bounds__(256) cons_final_kernel(const double* __restrict__ partial, int nblk, float aug_weight,
                                                         float* __restrict__ stats) {
    pdl_wait();
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
        stats[0] = (float)a;
        stats[1] = (float)b;
        stats[2] = (float)(a / b * (double)aug_weight);
        stats[3] = 0.f;
    }
}

constexpr int kConsMaxBlocks = 2048;

// rows per CTA so that the grid stays <= kConsMaxBlocks CTAs and every CTA has >= 4 vectors per thread
static ConsGeom make_geom(int B, int K, int Hi, int Wi, int H, int W, dim3& grid) {
    ConsGeom g{B * K, Hi, Wi, H, W, (float)H / (float)Hi, (float)W / (float)Wi, 1};
    int rows = (4 * 256 * 4 + Wi - 1) / Wi;
    if (rows < 1) rows = 1;
    while ((long long)((Hi + rows - 1) / rows) * B * K > kConsMaxBlocks && rows < Hi) rows *= 2;
    g.rows_per_cta = rows;
    grid = dim3((unsigned)((Hi + rows - 1) / rows), (unsigned)(B * K));
    return g;
}

int cons_fwd_partials(const float* oT, const float* oT_aug, const float* masks, int B, int K, int Hi, int Wi,
                      int H, int W, float threshold, double* partial, int* nblocks, cudaStream_t st) {
    dim3 grid;
    const ConsGeom g = make_geom(B, K, Hi, Wi, H, W, grid);
    if ((long long)grid.x * grid.y > kConsMaxBlocks || grid.y > 65535) return CLR_ERR_UNSUPPORTED;
    const bool vec4 = (Wi % 4 == 0) && aligned16(oT) && aligned16(oT_aug);
    if (vec4) launch_k(cons_kernel<4, false>, grid, 256, 0, st, oT, oT_aug, masks, g, threshold, partial, nullptr, nullptr, 0.f, nullptr);
    else launch_k(cons_kernel<1, false>, grid, 256, 0, st, oT, oT_aug, masks, g, threshold, partial, nullptr, nullptr, 0.f, nullptr);
    *nblocks = (int)(grid.x * grid.y);
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
    clr::launch_k(clr::cons_final_kernel, 1, 256, 0, st, static_cast<const double*>(ws), blocks, aug_weight, stats);
    return clr::launch_status();
}

int clr_cons_bwd(const float* oT, const float* oT_aug, const float* masks, int B, int K, int Hi, int Wi, int H, int W,
                 float threshold, float aug_weight, const float* stats, const float* gscale_dev, float gscale,
                 float* grad_oT_aug, clr_stream_t stream) {
    if (!oT || !oT_aug || !masks || !stats || !grad_oT_aug || B < 1 || K < 1 || Hi < 1 || Wi < 1 || H < 1 || W < 1)
        return CLR_ERR_BAD_ARG;
    dim3 grid;
    const clr::ConsGeom g = clr::make_geom(B, K, Hi, Wi, H, W, grid);
    if (grid.y > 65535) return CLR_ERR_UNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool vec4 = (Wi % 4 == 0) && clr::aligned16(oT) && clr::aligned16(oT_aug) && clr::aligned16(grad_oT_aug);
    if (vec4) clr::launch_k(clr::cons_kernel<4, true>, grid, 256, 0, st, oT, oT_aug, masks, g, threshold, nullptr, stats, gscale_dev, gscale * aug_weight, grad_oT_aug);
    else clr::launch_k(clr::cons_kernel<1, true>, grid, 256, 0, st, oT, oT_aug, masks, g, threshold, nullptr, stats, gscale_dev, gscale * aug_weight, grad_oT_aug);
    return clr::launch_status();
}

}  // extern "C"

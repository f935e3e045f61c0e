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
__device__ __forceinline__ float ex2_approx(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float lg2_approx(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sigmoid_fast(float x) { return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x)); }

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

// Exact label decision sigmoid(z) > thr for the price of one compare: sigmoid is monotone with slope thr(1-thr) at
// the threshold, so |z - logit(thr)| > guard implies |sigmoid(z) - thr| >> the fp32 evaluation error and the sign
// of (z - logit(thr)) IS the decision; inside the guard band ATen's exact expression is evaluated.
struct LabelRule { float thr, zthr, guard; };
__device__ __forceinline__ bool label_above(float z, const LabelRule& r) {
    const float d = z - r.zthr;
    if (fabsf(d) > r.guard) return d > 0.f;
    return sigmoid_aten(z) > r.thr;
}

// One element of the masked BCE: returns m*l.  y is 0/1, so only one of ATen's two log terms is non-zero:
// l = -max(log(y ? q : 1-q), -100) with q = sigmoid(z) (fast intrinsics; ATen: log(q) / log1p(-q)).
__device__ __forceinline__ void cons_elem(float zt, float za, float m, const LabelRule& r, float& ml, float& q_out, float& y_out) {
    const bool yb = label_above(zt, r);
    const float q = sigmoid_fast(za);
    const float t = yb ? q : 1.0f - q;
    ml = -m * fmaxf(0.6931471805599453f * lg2_approx(t), -100.0f);
    q_out = q; y_out = yb ? 1.0f : 0.0f;
}

static LabelRule make_rule(float thr) {
    LabelRule r;
    r.thr = thr;
    const double t = (double)thr;
    if (t > 0.0 && t < 1.0) {
        r.zthr = (float)log(t / (1.0 - t));
        const double g = 4e-6 / (t * (1.0 - t));
        r.guard = (float)(g > 1e-4 ? g : 1e-4);
    } else {
        r.zthr = 0.f;
        r.guard = 3.0e38f;   // degenerate threshold: always take the exact path
    }
    return r;
}

constexpr int kConsTX = 64, kConsTY = 4;   // block = 64 vector columns x 4 rows

// grid = (row blocks, B*K planes); each CTA covers `rows_per_cta` image rows of one plane; thread (tx, ty) walks
// vector columns tx, tx+64, .. of rows ty, ty+4, .. -- no divisions, up to 4 independent 128-bit load pairs in flight.
template <int VEC, bool BWD>
__global__ void __launch_bounds__(kConsTX * kConsTY) cons_kernel(const float* __restrict__ oT, const float* __restrict__ oT_aug,
                                                                 const float* __restrict__ masks, ConsGeom g, LabelRule thr,
                                                                 double* __restrict__ partial,
                                                                 const float* __restrict__ stats, const float* __restrict__ gscale_dev,
                                                                 float gscale, float* __restrict__ grad) {
    pdl_wait();
    const int bk = blockIdx.y;
    const int y0 = blockIdx.x * g.rows_per_cta;
    const int wv = g.Wi / VEC;                       // vectors per row
    const size_t plane = (size_t)bk * g.Hi * g.Wi;
    const float* mplane = masks + (size_t)bk * g.H * g.W;
    const bool shared_mask = (g.Wi == VEC * g.W);    // exact VEC:1 upsampling: a vector shares one mask pixel
    float coef = 0.f;
    if (BWD) coef = (gscale_dev ? gscale * __ldg(gscale_dev) : gscale) / __ldg(stats + 1);
    float num = 0.f, den = 0.f;
    constexpr int U = 4;
    const int tx = threadIdx.x, ty = threadIdx.y;
    for (int ry = ty; ry < g.rows_per_cta; ry += kConsTY) {
        const int y = y0 + ry;
        if (y >= g.Hi) break;
        const int sy = nearest_src(y, g.sh, g.H);
        const float* mrow = mplane + (size_t)sy * g.W;
        const size_t rowoff = plane + (size_t)y * g.Wi;
        for (int xb = tx; xb < wv; xb += U * kConsTX) {
            Pack<VEC> zt[U], za[U];
            float msh[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int xv = xb + u * kConsTX;
                if (xv < wv) {
                    zt[u] = ld_stream<VEC>(oT + rowoff + (size_t)xv * VEC);
                    za[u] = ld_stream<VEC>(oT_aug + rowoff + (size_t)xv * VEC);
                    msh[u] = shared_mask ? __ldg(mrow + xv) : 0.f;
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int xv = xb + u * kConsTX;
                if (xv >= wv) continue;
                Pack<VEC> go;
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    const float m = shared_mask ? msh[u] : __ldg(mrow + nearest_src(xv * VEC + v, g.sw, g.W));
                    float ml, q, yv;
                    cons_elem(zt[u].v[v], za[u].v[v], m, thr, ml, q, yv);
                    num += ml; den += m;
                    if (BWD) {
                        // ATen binary_cross_entropy_backward: (q - y) / max((1-q) q, 1e-12), chained with sigmoid' = q (1-q)
                        const float qq = (1.0f - q) * q;
                        go.v[v] = coef * m * (q - yv) / fmaxf(qq, 1e-12f) * qq;
                    }
                }
                if (BWD) st_stream<VEC>(grad + rowoff + (size_t)xv * VEC, go);
            }
        }
    }
    if (!BWD) {
        double dn = warp_sum((double)num), dd = warp_sum((double)den);
        __shared__ double sh[2][8];
        const int t = ty * kConsTX + tx, lane = t & 31, warp = t >> 5;
        if (lane == 0) { sh[0][warp] = dn; sh[1][warp] = dd; }
        __syncthreads();
        if (t == 0) {
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
    if (vec4) launch_k(cons_kernel<4, false>, grid, dim3(kConsTX, kConsTY), 0, st, oT, oT_aug, masks, g, make_rule(threshold), partial, nullptr, nullptr, 0.f, nullptr);
    else launch_k(cons_kernel<1, false>, grid, dim3(kConsTX, kConsTY), 0, st, oT, oT_aug, masks, g, make_rule(threshold), partial, nullptr, nullptr, 0.f, nullptr);
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
    if (vec4) clr::launch_k(clr::cons_kernel<4, true>, grid, dim3(clr::kConsTX, clr::kConsTY), 0, st, oT, oT_aug, masks, g, clr::make_rule(threshold), nullptr, stats, gscale_dev, gscale * aug_weight, grad_oT_aug);
    else clr::launch_k(clr::cons_kernel<1, true>, grid, dim3(clr::kConsTX, clr::kConsTY), 0, st, oT, oT_aug, masks, g, clr::make_rule(threshold), nullptr, stats, gscale_dev, gscale * aug_weight, grad_oT_aug);
    return clr::launch_status();
}

}  // extern "C"

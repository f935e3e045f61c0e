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
#include "clr_finish.cuh"

namespace clr {

__device__ __forceinline__ float sigmoid_aten(float x) { return 1.0f / (1.0f + expf(-x)); }   // label decision: exact
__device__ __forceinline__ float ex2_approx(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float lg2_approx(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sigmoid_fast(float x) { return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x)); }

struct ConsGeom {
    int BK, Hi, Wi, H, W;
    float sh, sw;   // nearest scales: (float)H/Hi, (float)W/Wi
    unsigned total_vec, wv;      // vectors in all planes / per image row
    int wv_shift, hi_shift;      // log2 when a power of two (shifts instead of integer divisions), else -1
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

// Forward element: l = BCE(sigmoid(za), y) = -log(y ? q : 1-q) = softplus(-/+ za) = ln2 * lg2(1 + 2^(-/+ za*log2 e)):
// two MUFU ops per element instead of three (exp, reciprocal, log).  ATen evaluates log(q) / log(1-q) on the ROUNDED
// fp32 q and clamps the log at -100: q rounds to exactly 1 (0) once |za| >= 24 ln 2, where the wrong-side term becomes
// 100 -- reproduced by the explicit saturation test; in the narrow band below it ATen's own value carries the
// quantisation error of 1-q, the softplus form is the mathematically exact one.
__device__ __noinline__ float cons_loss_elem_aten(float za, bool yb) {
    const float q = sigmoid_aten(za);
    return -fmaxf(yb ? logf(q) : log1pf(-q), -100.0f);
}
__device__ __forceinline__ float cons_loss_elem(float zt, float za, const LabelRule& r) {
    const bool yb = label_above(zt, r);
    const float x = (yb ? -1.4426950408889634f : 1.4426950408889634f) * za;    // wrong-side exponent, base 2
    const float l = 0.6931471805599453f * lg2_approx(1.0f + ex2_approx(x));
    // confidently wrong (|za| > 9): ATen's 1 - q (or q) keeps only a few significant bits before it saturates, and its
    // log drifts by up to 2 % from the softplus -- evaluate the reference's own expression (rare, out of line)
    if (x > 13.0f) return cons_loss_elem_aten(za, yb);
    return l;
}

// Backward element: (q - y) / max(q (1-q), 1e-12) * q (1-q)   (ATen binary_cross_entropy_backward x sigmoid')
__device__ __forceinline__ float cons_grad_elem(float zt, float za, const LabelRule& r) {
    const float y = label_above(zt, r) ? 1.0f : 0.0f;
    const float q = sigmoid_fast(za);
    const float qq = (1.0f - q) * q;
    return (q - y) / fmaxf(qq, 1e-12f) * qq;
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

constexpr int kConsThreads = 256;

// Persistent flat grid-stride kernel over the VEC-wide vectors of [B*K, Hi, Wi]: the grid is exactly the resident CTA
// capacity (no partial last wave -- the 2-D grid this replaces ran 2.3 waves), every thread keeps U = 4 independent
// 128-bit load pairs in flight, and the (plane, row, column) of a vector costs two shifts (two 32-bit divisions when
// the image size is not a power of two).
struct ConsArgs {
    const float* oT; const float* oT_aug; const float* masks;
    ConsGeom g; LabelRule thr;
    double* partial;                                   // forward: [ctas][2] per-CTA { sum m*l, sum m }
    const float* stats; const float* gscale_dev; float gscale; float* grad;    // backward
};

// `cta` of `ncta` CTAs (the body may share a launch with other work, see pool_finish_cons_kernel)
template <int VEC, bool BWD>
__device__ __forceinline__ void cons_body(const ConsArgs& a, const unsigned cta, const unsigned ncta) {
    const float* __restrict__ oT = a.oT; const float* __restrict__ oT_aug = a.oT_aug; const float* __restrict__ masks = a.masks;
    const ConsGeom& g = a.g; const LabelRule& thr = a.thr;
    double* __restrict__ partial = a.partial; const float* __restrict__ stats = a.stats;
    const float* __restrict__ gscale_dev = a.gscale_dev; const float gscale = a.gscale; float* __restrict__ grad = a.grad;
    const bool shared_mask = (g.Wi == VEC * g.W);    // exact VEC:1 upsampling: a vector shares one mask pixel
    float coef = 0.f;
    if (BWD) coef = (gscale_dev ? gscale * __ldg(gscale_dev) : gscale) / __ldg(stats + 1);
    float num = 0.f, den = 0.f;
    constexpr int U = 4;
    const unsigned nthreads = ncta * kConsThreads;
    for (unsigned v0 = cta * kConsThreads + threadIdx.x; v0 < g.total_vec; v0 += U * nthreads) {
        Pack<VEC> zt[U], za[U];
        const float* mrow[U];
        unsigned xv[U];
        float msh[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned v = v0 + u * nthreads;
            if (v < g.total_vec) {
                zt[u] = ld_stream<VEC>(oT + (size_t)v * VEC);
                za[u] = ld_stream<VEC>(oT_aug + (size_t)v * VEC);
                const unsigned row = g.wv_shift >= 0 ? (v >> g.wv_shift) : (v / g.wv);
                xv[u] = v - row * g.wv;
                const unsigned plane = g.hi_shift >= 0 ? (row >> g.hi_shift) : (row / (unsigned)g.Hi);
                const int y = (int)(row - plane * (unsigned)g.Hi);
                mrow[u] = masks + ((size_t)plane * g.H + nearest_src(y, g.sh, g.H)) * g.W;
                msh[u] = shared_mask ? __ldg(mrow[u] + xv[u]) : 0.f;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned v = v0 + u * nthreads;
            if (v >= g.total_vec) continue;
            Pack<VEC> go;
            float lv[VEC], mv[VEC];
            bool rare = false;
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                const float m = shared_mask ? msh[u] : __ldg(mrow[u] + nearest_src((int)(xv[u] * VEC + e), g.sw, g.W));
                if (BWD) {
                    go.v[e] = coef * m * cons_grad_elem(zt[u].v[e], za[u].v[e], thr);
                } else {
                    // branch-free common case; `rare` collects the elements that need the exact expressions (label
                    // inside the guard band, or confidently wrong |za| > 9): ONE branch per vector re-does them
                    const float d = zt[u].v[e] - thr.zthr;
                    const float x = (d > 0.f ? -1.4426950408889634f : 1.4426950408889634f) * za[u].v[e];
                    lv[e] = 0.6931471805599453f * lg2_approx(1.0f + ex2_approx(x));
                    mv[e] = m;
                    rare |= (fabsf(d) <= thr.guard) | (x > 13.0f);
                    den += m;
                }
            }
            if (BWD) {
                st_stream<VEC>(grad + (size_t)v * VEC, go);
            } else {
                if (rare) {
#pragma unroll
                    for (int e = 0; e < VEC; ++e) lv[e] = cons_loss_elem(zt[u].v[e], za[u].v[e], thr);
                }
#pragma unroll
                for (int e = 0; e < VEC; ++e) num = fmaf(mv[e], lv[e], num);
            }
        }
    }
    if (!BWD) {
        double dn = warp_sum((double)num), dd = warp_sum((double)den);
        __shared__ double sh[2][kConsThreads / 32];
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (lane == 0) { sh[0][warp] = dn; sh[1][warp] = dd; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double sa = 0.0, sb = 0.0;
            for (int w = 0; w < kConsThreads / 32; ++w) { sa += sh[0][w]; sb += sh[1][w]; }
            partial[2 * cta] = sa;
            partial[2 * cta + 1] = sb;
        }
    }
}

template <int VEC, bool BWD>
__global__ void __launch_bounds__(kConsThreads, 3) cons_kernel(const ConsArgs a) {
    kernel_begin(BWD ? TR_CONS_BWD : TR_CONS);
    cons_body<VEC, BWD>(a, blockIdx.x, gridDim.x);
    trace_exit(BWD ? TR_CONS_BWD : TR_CONS);
}

// Horizontal fusion for the fused step: CTAs [0, n_fin) run the pooling finish (partial reduce + EMA + alignment losses:
// a latency chain of a few KB) while the remaining CTAs stream the consistency pass, which does not depend on it.
template <int VEC>
__global__ void __launch_bounds__(kConsThreads, 3) pool_finish_cons_kernel(const PoolFinishParams f, const int n_fin, const ConsArgs a) {
    const bool flag_producer = f.done_all != nullptr;      // grid-uniform
    if ((int)blockIdx.x < n_fin) {
        if (flag_producer) kernel_begin_late_trigger(TR_ALIGN); else kernel_begin(TR_ALIGN);
        pool_finish_body(f, blockIdx.x, n_fin);
        if (f.done_all) cta_signal(f.early_signal ? nullptr : f.done_fin, f.done_all);   // (early: done_fin was bumped inside the body)
        trace_exit(TR_ALIGN);
    } else {
        if (flag_producer) kernel_begin_late_trigger(TR_CONS); else kernel_begin(TR_CONS);
        cons_body<VEC, false>(a, blockIdx.x - n_fin, gridDim.x - n_fin);
        if (f.done_all) cta_signal(nullptr, f.done_all);
        trace_exit(TR_CONS);
    }
}

__global__ void __launch_bounds__(256) cons_final_kernel(const double* __restrict__ partial, int nblk, float aug_weight,
                                                         float* __restrict__ stats) {
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
        stats[0] = (float)a;
        stats[1] = (float)b;
        stats[2] = (float)(a / b * (double)aug_weight);
        stats[3] = 0.f;
    }
}

constexpr int kConsMaxBlocks = 2048;

static int log2_exact(unsigned x) {
    if (x == 0 || (x & (x - 1))) return -1;
    int s = 0;
    while ((1u << s) != x) ++s;
    return s;
}

static int make_geom(int B, int K, int Hi, int Wi, int H, int W, int vec, ConsGeom& g) {
    const unsigned long long total = (unsigned long long)B * K * Hi * (unsigned long long)(Wi / vec);
    if (total > 0x7fffffffull) return CLR_ERR_UNSUPPORTED;
    g = ConsGeom{B * K, Hi, Wi, H, W, (float)H / (float)Hi, (float)W / (float)Wi, (unsigned)total, (unsigned)(Wi / vec),
                 log2_exact((unsigned)(Wi / vec)), log2_exact((unsigned)Hi)};
    return CLR_OK;
}

template <int VEC, bool BWD>
static int cons_grid(const ConsGeom& g, int& grid) {
    int occ = 0;
    { const int rc = kernel_occupancy(reinterpret_cast<const void*>(cons_kernel<VEC, BWD>), kConsThreads, 0, &occ); if (rc != CLR_OK) return rc; }
    if (occ < 1) occ = 1;
    long long want = (long long)device_facts().sms * occ;
    const long long need = ((long long)g.total_vec + kConsThreads - 1) / kConsThreads;
    if (want > need) want = need;
    if (want > kConsMaxBlocks) want = kConsMaxBlocks;
    grid = (int)(want < 1 ? 1 : want);
    return CLR_OK;
}

int cons_fwd_partials(const float* oT, const float* oT_aug, const float* masks, int B, int K, int Hi, int Wi,
                      int H, int W, float threshold, double* partial, int* nblocks, cudaStream_t st,
                      const PoolFinishParams* fused_finish) {
    const bool vec4 = (Wi % 4 == 0) && aligned16(oT) && aligned16(oT_aug);
    ConsArgs a{};
    int rc = make_geom(B, K, Hi, Wi, H, W, vec4 ? 4 : 1, a.g);
    if (rc != CLR_OK) return rc;
    int grid = 1;
    rc = vec4 ? cons_grid<4, false>(a.g, grid) : cons_grid<1, false>(a.g, grid);
    if (rc != CLR_OK) return rc;
    a.oT = oT; a.oT_aug = oT_aug; a.masks = masks; a.thr = make_rule(threshold); a.partial = partial;
    if (fused_finish) {
        const int n_fin = pool_finish_ctas(fused_finish->C);
        if (grid > n_fin + 1) grid -= n_fin;          // keep the launch at one resident wave
        if (vec4) launch_k(pool_finish_cons_kernel<4>, n_fin + grid, kConsThreads, 0, st, *fused_finish, n_fin, a);
        else launch_k(pool_finish_cons_kernel<1>, n_fin + grid, kConsThreads, 0, st, *fused_finish, n_fin, a);
    } else {
        if (vec4) launch_k(cons_kernel<4, false>, grid, kConsThreads, 0, st, a);
        else launch_k(cons_kernel<1, false>, grid, kConsThreads, 0, st, a);
    }
    *nblocks = grid;
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
    const int rc = clr::cons_fwd_partials(oT, oT_aug, masks, B, K, Hi, Wi, H, W, threshold, static_cast<double*>(ws), &blocks, st, nullptr);
    if (rc != CLR_OK) return rc;
    clr::launch_k(clr::cons_final_kernel, 1, 256, 0, st, static_cast<const double*>(ws), blocks, aug_weight, stats);
    return clr::launch_status();
}

int clr_cons_bwd(const float* oT, const float* oT_aug, const float* masks, int B, int K, int Hi, int Wi, int H, int W,
                 float threshold, float aug_weight, const float* stats, const float* gscale_dev, float gscale,
                 float* grad_oT_aug, clr_stream_t stream) {
    if (!oT || !oT_aug || !masks || !stats || !grad_oT_aug || B < 1 || K < 1 || Hi < 1 || Wi < 1 || H < 1 || W < 1)
        return CLR_ERR_BAD_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool vec4 = (Wi % 4 == 0) && clr::aligned16(oT) && clr::aligned16(oT_aug) && clr::aligned16(grad_oT_aug);
    clr::ConsGeom g;
    int rc = clr::make_geom(B, K, Hi, Wi, H, W, vec4 ? 4 : 1, g);
    if (rc != CLR_OK) return rc;
    int grid = 1;
    rc = vec4 ? clr::cons_grid<4, true>(g, grid) : clr::cons_grid<1, true>(g, grid);
    if (rc != CLR_OK) return rc;
    clr::ConsArgs a{oT, oT_aug, masks, g, clr::make_rule(threshold), nullptr, stats, gscale_dev, gscale * aug_weight, grad_oT_aug};
    if (vec4) clr::launch_k(clr::cons_kernel<4, true>, grid, clr::kConsThreads, 0, st, a);
    else clr::launch_k(clr::cons_kernel<1, true>, grid, clr::kConsThreads, 0, st, a);
    return clr::launch_status();
}

}  // extern "C"

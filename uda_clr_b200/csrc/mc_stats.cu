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
#include "clr_internal.h"
#include "clr_mc.cuh"

namespace clr {

// The full-resolution mean map exists only to be down-sampled (utils/Utils.py:170): the fused step reads it at the two
// bilinear source rows of every feature row and nowhere else.  With `rows.H` > 0 the kernel writes pred_mean only on those
// rows (half of them at the reference's 512 -> 128): 8.4 MB less write traffic per step.  Same source-row formula as
// bilinear_tap (ATen's); only for power-of-two image sizes and scale >= 1, otherwise every row is written.
struct McTapRows { int H, Hi, wi_shift, hi_shift; float sh, inv_sh; };
__device__ __forceinline__ bool is_tap_row(const McTapRows& g, size_t i) {
    const int r = (int)((i >> g.wi_shift) & (size_t)(g.Hi - 1));
    const int yc = (int)((float)r * g.inv_sh);
    bool tap = false;
#pragma unroll
    for (int dy = -2; dy <= 2; ++dy) {       // r = i0(y) or i0(y) + 1  =>  y within [r/sh - 1, r/sh + 1]; +-1 for the fp32 quotient
        const int y = yc + dy;
        if (y >= 0 && y < g.H) {
            const int i0 = (int)(g.sh * (float)y);
            const int i1 = i0 + ((i0 < g.Hi - 1) ? 1 : 0);
            tap = tap || r == i0 || r == i1;
        }
    }
    return tap;
}

template <int VEC, int TT, bool PRECISE, bool EXACT = false>
__global__ void __launch_bounds__(256) mc_stats_kernel(const float* __restrict__ preds, int T, size_t n,
                                                       float* __restrict__ std_map, float* __restrict__ pred_mean,
                                                       const McAten aten, const int nowait, const McTapRows rows) {
    // nowait (fused step, schedule 2): the launch in front of this one is the source half of the pooling finish -- a few
    // CTAs whose results this kernel does not read, and which trigger their dependents only AFTER their own
    // griddepcontrol.wait (everything older in the stream has completed by then).  Skipping the wait lets the whole
    // machine stream the MC logits while that latency chain (and, sharded, its cross-GPU exchange) runs.
    trace_enter(TR_MC_STATS);
    pdl_trigger();
    if (!nowait) pdl_wait();
    trace_ready(TR_MC_STATS);
    // n = B*K*Hi*Wi positions
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (i >= n) { trace_exit(TR_MC_STATS); return; }
    Pack<VEC> s, m;
    mc_vec_stats<VEC, TT, PRECISE, EXACT>(preds, T, n, i, s, m, aten);
    st_keep<VEC>(std_map + i, s);
    if (rows.H == 0 || is_tap_row(rows, i)) st_keep<VEC>(pred_mean + i, m);
    trace_exit(TR_MC_STATS);
}

template <int VEC, bool PRECISE>
static void launch_mc(const float* preds, int T, size_t n, float* std_map, float* pred_mean, cudaStream_t st, int nowait,
                      const McTapRows rows) {
    const size_t threads = (n + VEC - 1) / VEC;
    const unsigned blocks = (unsigned)((threads + 255) / 256);
    const McAten aten{mean_factor_aten(n, T)};
    if (PRECISE) launch_k(mc_stats_kernel<VEC, 0, PRECISE>, blocks, 256, 0, st, preds, T, n, std_map, pred_mean, aten, nowait, rows);
    else if (T == 8 && !tunables().mc_generic) launch_k(mc_stats_kernel<VEC, 8, PRECISE, true>, blocks, 256, 0, st, preds, T, n, std_map, pred_mean, aten, nowait, rows);
    else if (T <= 8) launch_k(mc_stats_kernel<VEC, 8, PRECISE>, blocks, 256, 0, st, preds, T, n, std_map, pred_mean, aten, nowait, rows);
    else if (T <= 16) launch_k(mc_stats_kernel<VEC, 16, PRECISE>, blocks, 256, 0, st, preds, T, n, std_map, pred_mean, aten, nowait, rows);
    else launch_k(mc_stats_kernel<VEC, 0, PRECISE>, blocks, 256, 0, st, preds, T, n, std_map, pred_mean, aten, nowait, rows);
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
// h0 * (w0 * a + w1 * b) + h1 * (w0 * c + w1 * d) with the contraction nvcc applies to that expression in ATen's
// upsample_bilinear2d kernel pinned explicitly, so every caller rounds identically
__device__ __forceinline__ float bilinear_mix(float a, float b, float c, float d, const Tap& h, const Tap& w) {
    const float top = __fmaf_rn(w.l0, a, __fmul_rn(w.l1, b));
    const float bot = __fmaf_rn(w.l0, c, __fmul_rn(w.l1, d));
    return __fmaf_rn(h.l0, top, __fmul_rn(h.l1, bot));
}
__device__ __forceinline__ float bilinear_at(const float* __restrict__ plane, int Wi, const Tap& h, const Tap& w) {
    const float* r0 = plane + (size_t)h.i0 * Wi;
    const float* r1 = plane + (size_t)h.i1 * Wi;
    return bilinear_mix(__ldg(r0 + w.i0), __ldg(r0 + w.i1), __ldg(r1 + w.i0), __ldg(r1 + w.i1), h, w);
}

// Guard band of the uncertainty mask: the streaming statistics use approximate sigmoids and a two-pass variance (error
// of std_small <= ~3e-7); a pixel whose value lies within kMaskBand of the threshold is decided by an
// exact re-evaluation instead -- its 4 bilinear taps recomputed from the MC logits in ATen's order -- so that mask_k is
// bit-identical to eager torch on the same device (utils/Utils.py:166, 171, 197-200) at streaming speed.
// About one pixel in 2000 is in the band, and it sits on the kernel's critical path, so the WARP re-evaluates it: up to 8
// flagged pixels per round, 4 lanes per pixel (one bilinear tap each, all T logits of a tap in flight), the owner lane
// gathers its 4 taps by shuffle.  Must be called by all 32 lanes (converged); `need` / geometry are per lane.
constexpr float kMaskBand = 1e-5f;
__device__ __forceinline__ float std_small_exact_warp(bool need, float ss, const float* __restrict__ preds, int T, size_t n,
                                                      size_t plane, int Wi, const Tap& h, const Tap& w) {
    const unsigned lane = threadIdx.x & 31u;
    unsigned pending = __ballot_sync(0xffffffffu, need);
    if (T <= 8) {
        // the reference's T = 8: one flagged pixel per round, lane = (tap, MC pass): ONE load and ONE exact sigmoid per lane,
        // then every lane of a tap group runs the (cheap) Welford recurrence over the group's 8 values by shuffle
        while (pending) {                               // warp-uniform
            const unsigned src = __ffs(pending) - 1u;
            pending &= pending - 1u;
            const unsigned long long pl = __shfl_sync(0xffffffffu, (unsigned long long)plane, src);
            const int hi0 = __shfl_sync(0xffffffffu, h.i0, src), hi1 = __shfl_sync(0xffffffffu, h.i1, src);
            const int wi0 = __shfl_sync(0xffffffffu, w.i0, src), wi1 = __shfl_sync(0xffffffffu, w.i1, src);
            const unsigned tap = lane >> 3, t = lane & 7u;
            float x = 0.f;
            if ((int)t < T)
                x = sigmoid_half_aten(__ldg(preds + (size_t)t * n + (size_t)pl + (size_t)((tap & 2u) ? hi1 : hi0) * Wi + ((tap & 1u) ? wi1 : wi0)));
            WelfordAcc a0{0.f, 0.f, 0.f}, a1{0.f, 0.f, 0.f};
#pragma unroll
            for (int tt = 0; tt < 8; ++tt) {
                const float v = __shfl_sync(0xffffffffu, x, (lane & 24u) + tt);
                if (tt < T) { if (tt & 1) welford_push(a1, v); else welford_push(a0, v); }
            }
            const WelfordAcc r = welford_merge(a0, a1);
            const float divisor = r.nf > 1.0f ? __fsub_rn(r.nf, 1.0f) : 0.0f;
            const float val = __fsqrt_rn(__fdiv_rn(r.m2, divisor));
            const float a = __shfl_sync(0xffffffffu, val, 0), b = __shfl_sync(0xffffffffu, val, 8);
            const float c = __shfl_sync(0xffffffffu, val, 16), d = __shfl_sync(0xffffffffu, val, 24);
            if (lane == src) ss = bilinear_mix(a, b, c, d, h, w);
        }
        return ss;
    }
    while (pending) {                                   // warp-uniform
        const unsigned slot = lane >> 2, tap = lane & 3u;
        const unsigned src = __fns(pending, 0, (int)slot + 1);             // lane of the slot-th flagged pixel, or ~0u
        const unsigned sl = src < 32u ? src : 0u;
        const unsigned long long pl = __shfl_sync(0xffffffffu, (unsigned long long)plane, sl);
        const int hi0 = __shfl_sync(0xffffffffu, h.i0, sl), hi1 = __shfl_sync(0xffffffffu, h.i1, sl);
        const int wi0 = __shfl_sync(0xffffffffu, w.i0, sl), wi1 = __shfl_sync(0xffffffffu, w.i1, sl);
        float val = 0.f;
        if (src < 32u) val = std_aten_at(preds, T, n, (size_t)pl + (size_t)((tap & 2u) ? hi1 : hi0) * Wi + ((tap & 1u) ? wi1 : wi0));
        const unsigned rank = __popc(pending & ((1u << lane) - 1u));       // this lane's position among the flagged ones
        const unsigned base = (rank & 7u) * 4u;
        const float a = __shfl_sync(0xffffffffu, val, base), b = __shfl_sync(0xffffffffu, val, base + 1);
        const float c = __shfl_sync(0xffffffffu, val, base + 2), d = __shfl_sync(0xffffffffu, val, base + 3);
        if (((pending >> lane) & 1u) && rank < 8u) ss = bilinear_mix(a, b, c, d, h, w);
#pragma unroll
        for (int i = 0; i < 8; ++i) pending &= pending - 1u;               // the 8 lowest flagged lanes are done
    }
    return ss;
}

__global__ void __launch_bounds__(256) retrify_weights_kernel(
    const float* __restrict__ oT_before, const float* __restrict__ pred_mean, const float* __restrict__ std_map,
    const float* __restrict__ preds /*[T][B,K,Hi,Wi] or null: no guard band*/, int T,
    int B, int K, int H, int W, int Hi, int Wi, float pseudo_thr, float std_thr,
    float* __restrict__ weights /*[B,2K,H,W]*/, float* __restrict__ masks /*[B,K,H,W]*/,
    float* __restrict__ pseudo_out /*[B,K,H,W] or null*/, float* __restrict__ small_out /*[2][B,K,H,W] or null*/, const int nowait) {
    // nowait (schedule 3): the launch in front is the source half of the pooling finish, whose late trigger already implies
    // that the MC statistics this kernel reads are complete; skipping the wait keeps that latency chain off the critical path
    trace_enter(TR_RETRIFY);
    pdl_trigger();
    if (!nowait) pdl_wait();
    trace_ready(TR_RETRIFY);
    // grid.y = b*K + k (one plane), grid.x covers the plane: no 64-bit divisions.  No early return: the guard band below
    // is evaluated by whole warps.
    const size_t n = (size_t)B * K * H * W;
    const int pixi = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = pixi < H * W;
    const int pc = valid ? pixi : 0;
    const int bk = blockIdx.y;
    const int b = bk / K, k = bk - b * K;
    const int y = pc / W, x = pc - y * W;
    const size_t i = (size_t)bk * H * W + pc;
    const float sh = H > 1 ? (float)(Hi - 1) / (float)(H - 1) : 0.f;
    const float sw = W > 1 ? (float)(Wi - 1) / (float)(W - 1) : 0.f;
    const Tap th = bilinear_tap(y, Hi, sh), tw = bilinear_tap(x, Wi, sw);
    const size_t plane = ((size_t)b * K + k) * Hi * Wi;
    const float ps = bilinear_at(pred_mean + plane, Wi, th, tw);
    float ss = bilinear_at(std_map + plane, Wi, th, tw);
    const bool pseudo = sigmoid_aten(oT_before[i]) > pseudo_thr;
    if (preds != nullptr)      // kernel-uniform
        ss = std_small_exact_warp(valid && fabsf(ss - std_thr) < kMaskBand, ss, preds, T, (size_t)B * K * Hi * Wi, plane, Wi, th, tw);
    const bool m = ss < std_thr;
    if (valid) {
        const size_t hw = (size_t)H * W, pix = (size_t)y * W + x;
        weights[((size_t)b * 2 * K + k) * hw + pix] = (pseudo && m) ? ps : 0.f;
        weights[((size_t)b * 2 * K + K + k) * hw + pix] = (!pseudo && m) ? (1.0f - ps) : 0.f;
        masks[i] = m ? 2.0f : 0.f;
        if (pseudo_out) pseudo_out[i] = pseudo ? 1.0f : 0.f;
        if (small_out) { small_out[i] = ps; small_out[n + i] = ss; }
    }
    trace_exit(TR_RETRIFY);
}

// ---------------------------------------------------------------------------------------------------------------
// MC statistics + retrify weights in ONE pass (the fused step): CTA = (plane b*K+k, feature row y).  It owns the image
// rows [r0(y), r0(y+1)) with r0(y) = (int)(sh*y) -- exactly ATen's align_corners source row of feature row y -- so the
// two bilinear tap rows r0(y), r0(y)+1 of its feature row are rows it computes itself (needs sh >= 2): their std / mean
// values are kept in shared memory, and after one barrier W threads evaluate the bilinear taps, the pseudo-label and
// the uncertainty mask and write the 2K weight planes.  The rows in between are computed and stored to std_map only.
// Saves the retrify launch, its re-read of the two maps and the write of pred_mean (never needed at full resolution:
// utils/Utils.py:170 only down-samples it).
// ---------------------------------------------------------------------------------------------------------------
struct McRetrifyParams {
    const float* preds; const float* oT_before;
    float* std_map; float* pred_mean /*nullable*/; float* weights; float* masks;
    int T, B, K, H, W, Hi, Wi;
    int parts, cw;            // the image rows are split into `parts` column blocks of cw floats, one CTA each
    float sh, sw, pseudo_thr, std_thr;
    size_t n;
    McAten aten;
};

template <int TT, bool PRECISE>
__global__ void __launch_bounds__(256, 4) mc_retrify_kernel(const McRetrifyParams p) {
    kernel_begin(TR_MC_STATS);
    extern __shared__ __align__(16) float taps[];      // [2 maps: std, mean][2 rows][cw]
    const int y = blockIdx.x / p.parts, part = blockIdx.x - y * p.parts, bk = blockIdx.y;
    const int c0 = part * p.cw;                        // first image column of this CTA
    const int r0 = (int)(p.sh * (float)y);
    const int r1 = (y == p.H - 1) ? p.Hi : (int)(p.sh * (float)(y + 1));
    const int i1 = r0 + ((r0 < p.Hi - 1) ? 1 : 0);
    const int wv = p.cw >> 2;
    const int total = (r1 - r0) * wv;
    const size_t plane = (size_t)bk * p.Hi * p.Wi;
    const size_t hw = (size_t)p.H * p.W;
    // the feature pixels whose taps fall into this column block: x in [x_lo, x_hi) (host checked: no tap pair straddles)
    const Tap th = bilinear_tap(y, p.Hi, p.sh);          // th.i0 == r0, th.i1 == i1
    // this thread's feature-row logit for the epilogue: issued now so its latency hides behind the streaming phase
    const int xg = (int)threadIdx.x;                     // candidate feature column handled in the epilogue (round 0)
    float o_pre = 0.f;
    if (xg < p.W) o_pre = __ldg(p.oT_before + (size_t)bk * hw + (size_t)y * p.W + xg);
    for (int v = threadIdx.x; v < total; v += 256) {
        const int rr = v / wv, xv = v - rr * wv;
        const int row = r0 + rr;
        const size_t i = plane + (size_t)row * p.Wi + c0 + 4 * xv;
        Pack<4> s, m;
        mc_vec_stats<4, TT, PRECISE>(p.preds, p.T, p.n, i, s, m, p.aten);
        st_keep<4>(p.std_map + i, s);
        if (p.pred_mean) st_keep<4>(p.pred_mean + i, m);
        if (row == r0) { st_keep<4>(taps + 4 * xv, s); st_keep<4>(taps + 2 * p.cw + 4 * xv, m); }
        if (row == i1) { st_keep<4>(taps + p.cw + 4 * xv, s); st_keep<4>(taps + 3 * p.cw + 4 * xv, m); }
    }
    __syncthreads();
    const int b = bk / p.K, k = bk - b * p.K;
    const float* s0 = taps;                 // std, row i0
    const float* s1 = taps + p.cw;          // std, row i1
    const float* m0 = taps + 2 * p.cw;      // mean, row i0
    const float* m1 = taps + 3 * p.cw;      // mean, row i1
    for (int x0 = 0; x0 < p.W; x0 += 256) {                 // CTA-uniform trip count: the guard band is evaluated by whole warps
        const int x = x0 + (int)threadIdx.x;
        const int xc = x < p.W ? x : 0;
        const Tap tw = bilinear_tap(xc, p.Wi, p.sw);
        const bool mine = x < p.W && tw.i0 >= c0 && tw.i0 < c0 + p.cw;    // else: past the row / another column block's pixel
        const int a0 = mine ? tw.i0 - c0 : 0, a1 = mine ? tw.i1 - c0 : 0;
        const float ps = bilinear_mix(m0[a0], m0[a1], m1[a0], m1[a1], th, tw);
        float ss = bilinear_mix(s0[a0], s0[a1], s1[a0], s1[a1], th, tw);
        if (!PRECISE) ss = std_small_exact_warp(mine && fabsf(ss - p.std_thr) < kMaskBand, ss, p.preds, p.T, p.n, plane, p.Wi, th, tw);
        if (!mine) continue;
        const size_t pix = (size_t)y * p.W + x;
        const size_t i = (size_t)bk * hw + pix;
        const float o = (x == xg) ? o_pre : __ldg(p.oT_before + i);
        const bool pseudo = sigmoid_aten(o) > p.pseudo_thr;
        const bool m = ss < p.std_thr;
        p.weights[((size_t)b * 2 * p.K + k) * hw + pix] = (pseudo && m) ? ps : 0.f;
        p.weights[((size_t)b * 2 * p.K + p.K + k) * hw + pix] = (!pseudo && m) ? (1.0f - ps) : 0.f;
        p.masks[i] = m ? 2.0f : 0.f;
    }
    trace_exit(TR_MC_STATS);
}

// host mirror of bilinear_tap's source index (same fp32 expression)
static inline int tap_i0(int dst, float scale) { return (int)(scale * (float)dst); }

// Largest split of the image rows into column blocks such that (a) blocks are whole float4 vectors, (b) no feature
// pixel's tap pair (i0, i0+1) straddles a block boundary, (c) a typical CTA still has a full 256 vectors of work.
static int choose_parts(int W, int Wi, float sw, int rows_per_cta) {
    for (int parts = 8; parts > 1; parts >>= 1) {
        if (Wi % (4 * parts)) continue;
        const int cw = Wi / parts;
        if ((long long)rows_per_cta * (cw / 4) < 256) continue;
        bool ok = true;
        for (int x = 0; x < W && ok; ++x) {
            const int i0 = tap_i0(x, sw), i1 = i0 + ((i0 < Wi - 1) ? 1 : 0);
            ok = (i0 / cw) == (i1 / cw);
        }
        if (ok) return parts;
    }
    return 1;
}

// CLR_ERR_UNSUPPORTED when the geometry does not allow the fusion (the caller then runs the two kernels).
int mc_retrify_fused(const float* preds, const float* oT_before, int T, int B, int K, int H, int W, int Hi, int Wi,
                     float pseudo_thr, float std_thr, float* std_map, float* pred_mean, float* weights, float* masks,
                     cudaStream_t st) {
    if (!preds || !oT_before || !std_map || !weights || !masks || T < 1 || B < 1 || K < 1 || K > CLR_MAX_K) return CLR_ERR_BAD_ARG;
    if (H < 2 || W < 1 || Hi - 1 < 2 * (H - 1) || Wi % 4 != 0 || (long long)B * K > 65535) return CLR_ERR_UNSUPPORTED;
    if (!aligned16(preds) || !aligned16(std_map) || (pred_mean && !aligned16(pred_mean))) return CLR_ERR_UNSUPPORTED;
    const float sh_f = (float)(Hi - 1) / (float)(H - 1), sw_f = W > 1 ? (float)(Wi - 1) / (float)(W - 1) : 0.f;
    const int parts = tunables().mc_split ? choose_parts(W, Wi, sw_f, (int)sh_f) : 1;   // column split measured slower (shorter DRAM runs): opt-in
    const size_t smem = sizeof(float) * 4 * (size_t)(Wi / parts);
    if (smem > 48 * 1024 || (long long)H * parts > 0x7fffffffLL) return CLR_ERR_UNSUPPORTED;
    McRetrifyParams p{};
    p.preds = preds; p.oT_before = oT_before; p.std_map = std_map; p.pred_mean = pred_mean; p.weights = weights; p.masks = masks;
    p.T = T; p.B = B; p.K = K; p.H = H; p.W = W; p.Hi = Hi; p.Wi = Wi;
    p.sh = sh_f; p.sw = sw_f; p.parts = parts; p.cw = Wi / parts;
    p.pseudo_thr = pseudo_thr; p.std_thr = std_thr;
    p.n = (size_t)B * K * Hi * Wi;
    p.aten = McAten{mean_factor_aten(p.n, T)};
    const dim3 grid((unsigned)(H * parts), (unsigned)(B * K));
    const bool precise = tunables().mc_precise != 0;
    if (precise) launch_k(mc_retrify_kernel<0, true>, grid, 256, smem, st, p);
    else if (T <= 8) launch_k(mc_retrify_kernel<8, false>, grid, 256, smem, st, p);
    else if (T <= 16) launch_k(mc_retrify_kernel<16, false>, grid, 256, smem, st, p);
    else launch_k(mc_retrify_kernel<0, false>, grid, 256, smem, st, p);
    return launch_status();
}

static int log2_pow2(int x) {
    if (x <= 0 || (x & (x - 1))) return -1;
    int s = 0;
    while ((1 << s) != x) ++s;
    return s;
}

// tap_H > 0: the caller only ever reads pred_mean at the bilinear source rows of a tap_H-row feature map (fused step)
int mc_stats_impl(const float* preds, int T, int B, int K, int Hi, int Wi, float* std_map, float* pred_mean, cudaStream_t st,
                  bool nowait, int tap_H) {
    if (!preds || !std_map || !pred_mean || T < 1 || B < 1 || K < 1 || Hi < 1 || Wi < 1) return CLR_ERR_BAD_ARG;
    const size_t n = (size_t)B * K * Hi * Wi;
    const bool vec4 = (n % 4 == 0) && aligned16(preds) && aligned16(std_map) && aligned16(pred_mean);
    const bool precise = tunables().mc_precise != 0;
    const int nw = nowait ? 1 : 0;
    McTapRows rows{};
    const int ws = log2_pow2(Wi), hs = log2_pow2(Hi);
    if (tap_H > 1 && ws >= 2 && hs >= 0 && Hi >= tap_H && !tunables().mc_all_rows) {
        const float sh = (float)(Hi - 1) / (float)(tap_H - 1);      // bilinear_tap's scale
        if (sh >= 1.0f) rows = McTapRows{tap_H, Hi, ws, hs, sh, 1.0f / sh};
    }
    if (vec4) {
        if (precise) launch_mc<4, true>(preds, T, n, std_map, pred_mean, st, nw, rows);
        else launch_mc<4, false>(preds, T, n, std_map, pred_mean, st, nw, rows);
    } else {
        rows = McTapRows{};          // (a scalar thread's position is not row-aligned: write everything)
        if (precise) launch_mc<1, true>(preds, T, n, std_map, pred_mean, st, nw, rows);
        else launch_mc<1, false>(preds, T, n, std_map, pred_mean, st, nw, rows);
    }
    return launch_status();
}

int retrify_weights_impl(const float* oT_before, const float* pred_mean, const float* std_map,
                         const float* preds, int T,
                         int B, int K, int H, int W, int Hi, int Wi, float pseudo_thr, float std_thr,
                         float* weights, float* masks, float* pseudo_out, float* small_out, cudaStream_t stream, bool nowait) {
    if (!oT_before || !pred_mean || !std_map || !weights || !masks || B < 1 || K < 1 || K > CLR_MAX_K ||
        H < 1 || W < 1 || Hi < 1 || Wi < 1 || (preds && T < 1))
        return CLR_ERR_BAD_ARG;
    if ((long long)H * W > 0x7fffff00LL || (long long)B * K > 65535) return CLR_ERR_UNSUPPORTED;
    clr::launch_k(clr::retrify_weights_kernel, dim3((unsigned)((H * W + 255) / 256), (unsigned)(B * K)), 256, 0, stream, oT_before, pred_mean, std_map, preds, T, B, K, H, W, Hi, Wi, pseudo_thr, std_thr, weights, masks, pseudo_out, small_out, nowait ? 1 : 0);
    return clr::launch_status();
}

}  // namespace clr

extern "C" {

int clr_mc_stats(const float* preds, int T, int B, int K, int Hi, int Wi,
                 float* std_map, float* pred_mean, clr_stream_t stream) {
    return clr::mc_stats_impl(preds, T, B, K, Hi, Wi, std_map, pred_mean, static_cast<cudaStream_t>(stream), false, 0);
}

int clr_retrify_weights(const float* oT_before, const float* pred_mean, const float* std_map,
                        const float* preds, int T,
                        int B, int K, int H, int W, int Hi, int Wi, float pseudo_thr, float std_thr,
                        float* weights, float* masks, float* pseudo_out, float* small_out, clr_stream_t stream) {
    return clr::retrify_weights_impl(oT_before, pred_mean, std_map, preds, T, B, K, H, W, Hi, Wi, pseudo_thr, std_thr, weights, masks,
                                     pseudo_out, small_out, static_cast<cudaStream_t>(stream), false);
}



int clr_mc_retrify(const float* preds, const float* oT_before, int T, int B, int K, int H, int W, int Hi, int Wi,
                   float pseudo_thr, float std_thr, float* std_map, float* pred_mean, float* weights, float* masks,
                   clr_stream_t stream) {
    return clr::mc_retrify_fused(preds, oT_before, T, B, K, H, W, Hi, Wi, pseudo_thr, std_thr, std_map, pred_mean, weights,
                                 masks, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

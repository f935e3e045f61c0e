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
//   CTA    = (domain, b, block of 256*VEC pixels, span of 32 channels): no cross-thread traffic at all
//   thread = VEC pixels: loads its Q coefficient values once (kept in registers), then walks the
//            span's channels; the per-channel table T[q][c] (built per CTA in shared memory from g and N)
//            is read as a warp-wide broadcast; one 128-bit evict-first store per channel.
//   grid   = every CTA writes the same 128 KB, all are short; no persistent scheduling needed.
#include "clr_common.cuh"
#include "clr_internal.h"
#include "clr_finish.cuh"

namespace clr {

struct BwdDom {
    const float* w;
    const float* g;       // [R][C]
    const float* sums;    // [R][C+1]
    const float* xcoef;   // [B,Kx,HW] or null
    const float* xtab;    // [Kx][C] or null
    float* grad;
    const float* scale_dev;   // optional device scalar multiplied into `scale` (upstream dL/dtotal)
    float scale;
    int fmt, B, Kx, ctas;
    int nrows;            // explicit format: number of weight planes (2K for prototypes, any 1..16 for row pooling)
    int per_sample;       // 1: sums is [B][nrows][C+1] and the divisor is N_b[r] + n_add (bmm-style pooling)
    float n_add;
};

struct BwdParams {
    BwdDom dom[2];
    int ndom, C, HW, K, nPx, nSpan;
    int trace_id;
    // merged backward launch ([disc finish | gradient of xt | gradient of xs]): the CTAs of domain `gate_dom` need the
    // finish CTAs' output (g of that domain, xtab) and wait until `gate` (bumped once by every finish CTA, zeroed by the
    // step's pooling kernel) reaches gate_n; they then read those vectors through coherent loads only.
    unsigned int* gate; unsigned int gate_n; int gate_dom; float* gate_err;
    // the finish runs as its OWN launch in front of this one (sharded step: a grid that touched peer memory pays a ~3.5 us
    // longer end-of-grid flush -- kept out of the big backward grid, whose completion the next step waits for): this kernel
    // then skips griddepcontrol.wait; the finish triggered it only after its own wait, the gated CTAs wait on the gate
    int nowait;
};

constexpr int kBwdSpan = 32;   // channels per CTA

template <int QT, int VEC>
__device__ __forceinline__ void pool_bwd_body(const BwdParams& p, int bid) {
    __shared__ __align__(16) float T[(1 + QT) * kBwdSpan];
    const int tid = threadIdx.x;
    const int d = (p.ndom > 1 && bid >= p.dom[0].ctas) ? 1 : 0;
    if (d) bid -= p.dom[0].ctas;
    const BwdDom& D = p.dom[d];
    const int span = bid % p.nSpan;
    const int rest = bid / p.nSpan;
    const int pb = rest % p.nPx, b = rest / p.nPx;
    const int K = p.K;
    const int QW = (D.fmt == CLR_W_COMPLEMENT) ? K : D.nrows;
    const int Q = QW + D.Kx;
    const int c0 = span * kBwdSpan;

    // ---- coefficient values of this thread's pixels (issued first: independent of the table) -----
    const int px = (pb * kThreads + tid) * VEC;
    const bool ok = px < p.HW;
    Pack<VEC> cf[QT];
#pragma unroll
    for (int q = 0; q < QT; ++q) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) cf[q].v[v] = 0.f;
        if (ok && q < Q) {
            const float* src = (q < QW) ? D.w + ((size_t)b * QW + q) * p.HW
                                        : D.xcoef + ((size_t)b * D.Kx + (q - QW)) * p.HW;
            cf[q] = ld_keep<VEC>(src + px);
        }
    }
    const bool gated = p.gate != nullptr && d == p.gate_dom;      // CTA-uniform
    __shared__ int gate_bad;
    if (gated) {
        // The finish CTAs have lower block indices than every gated CTA of the same grid and CTAs are dispatched in index
        // order, so they are resident (or done) before a gated CTA can spin.  That order is de-facto hardware behaviour, not
        // a documented guarantee: if the gate is ever missed (~2 s) the CTA sets losses[7] and writes NaN gradients -- loud,
        // never silently wrong -- and "bwd_merge_off" = 1 selects the two-launch form that needs no gate.
        if (tid == 0) {
            gate_bad = spin_until_at_least(p.gate, p.gate_n) ? 0 : 1;
            if (gate_bad && p.gate_err) *p.gate_err = 1.f;
        }
        __syncthreads();
    }
    const float poison = (gated && gate_bad) ? __int_as_float(0x7fc00000) : 1.0f;
    // vectors the finish CTAs of the same launch may have written: coherent loads when gated
    auto ldv = [gated](const float* q_) -> float { return gated ? __ldcg(q_) : __ldg(q_); };
    // ---- per-CTA table for channels c0 .. c0+31: T[0] constant term, T[1+q] coefficient of plane q.
    //      One thread per (row q, channel j) entry so the g / N loads and divides are a single parallel round.
    {
        const int j = tid & (kBwdSpan - 1), q = tid / kBwdSpan;     // q in [0, 8): rows 0..QT handled in rounds
        const int c = c0 + j;
        const float scale = poison * (D.scale_dev ? D.scale * __ldg(D.scale_dev) : D.scale);
        const float* sums = D.per_sample ? D.sums + (size_t)b * D.nrows * (p.C + 1) : D.sums;
        const float nadd = D.n_add;
        for (int row = q; row <= QT; row += kThreads / kBwdSpan) {
            float t = 0.f;
            if (c < p.C) {
                if (row == 0) {
                    if (D.fmt == CLR_W_COMPLEMENT)
                        for (int k = 0; k < K; ++k)
                            t += scale * ldv(D.g + (size_t)(K + k) * p.C + c) / (__ldg(sums + (size_t)(K + k) * (p.C + 1) + p.C) + nadd);
                } else if (row - 1 < QW) {
                    const int r = row - 1;
                    const float gr = scale * ldv(D.g + (size_t)r * p.C + c) / (__ldg(sums + (size_t)r * (p.C + 1) + p.C) + nadd);
                    if (D.fmt == CLR_W_COMPLEMENT) {
                        const float gb = scale * ldv(D.g + (size_t)(K + r) * p.C + c) / (__ldg(sums + (size_t)(K + r) * (p.C + 1) + p.C) + nadd);
                        t = gr - gb;
                    } else {
                        t = gr;
                    }
                } else if (row - 1 < Q) {
                    t = poison * (D.scale_dev ? __ldg(D.scale_dev) : 1.f) * ldv(D.xtab + (size_t)(row - 1 - QW) * p.C + c);
                }
            }
            T[row * kBwdSpan + j] = t;
        }
    }
    __syncthreads();
    if (!ok) return;

    float* gp = D.grad + ((size_t)b * p.C + c0) * p.HW + px;
    const int nc = (p.C - c0) < kBwdSpan ? (p.C - c0) : kBwdSpan;
#pragma unroll 4
    for (int j = 0; j < nc; ++j) {
        Pack<VEC> o;
        const float t0 = T[j];
#pragma unroll
        for (int v = 0; v < VEC; ++v) o.v[v] = t0;
#pragma unroll
        for (int q = 0; q < QT; ++q) {
            const float t = T[(1 + q) * kBwdSpan + j];
#pragma unroll
            for (int v = 0; v < VEC; ++v) o.v[v] = fmaf(t, cf[q].v[v], o.v[v]);
        }
        st_stream<VEC>(gp + (size_t)j * p.HW, o);
    }
}

// resident CTAs per SM by register need: QT coefficient planes x VEC pixels live in registers (QT = 16 -> 64 of them)
constexpr int bwd_min_blocks(int QT) { return QT <= 4 ? 4 : (QT <= 8 ? 3 : 2); }

template <int QT, int VEC>
__global__ void __launch_bounds__(kThreads, bwd_min_blocks(QT)) pool_bwd_kernel(const BwdParams p) {
    const int tr = (p.ndom > 1 && p.gate) ? ((int)blockIdx.x >= p.dom[0].ctas ? TR_BWD_S : TR_BWD_T) : p.trace_id;
    trace_enter(tr);
    pdl_trigger();
    if (!p.nowait) pdl_wait();
    trace_ready(tr);
    pool_bwd_body<QT, VEC>(p, blockIdx.x);
    trace_exit(tr);
}

// Horizontal fusion for the fused step: CTAs [0, n_fin) run the discriminative finish (partial reduce + prototype
// gradients + direct-gradient table + step totals: a latency chain) while the remaining CTAs write the gradient of the
// TARGET features, which depends on the alignment term only.  The source-gradient write follows as the next launch.
template <int QT, int VEC>
__global__ void __launch_bounds__(kThreads, bwd_min_blocks(QT)) bwd_finish_kernel(const BwdParams p, const DiscFinishParams f, const int n_fin) {
    if ((int)blockIdx.x < n_fin) {
        kernel_begin(TR_DISC_FIN);
        disc_finish_body(f, blockIdx.x, n_fin);
        if (p.gate) cta_signal(p.gate, nullptr);
        trace_exit(TR_DISC_FIN);
    } else {
        // merged launch (two domains): target CTAs first, then the gated source CTAs; each keeps its own trace slot
        const int bid = blockIdx.x - n_fin;
        const int tr = p.ndom > 1 ? (bid >= p.dom[0].ctas ? TR_BWD_S : TR_BWD_T) : p.trace_id;
        kernel_begin(tr);
        pool_bwd_body<QT, VEC>(p, bid);
        trace_exit(tr);
    }
}

template <int QT>
static int launch_bwd(const BwdParams& p, bool vec4, int ctas, cudaStream_t st, const DiscFinishParams* fin) {
    if (fin) {
        const int n_fin = disc_finish_ctas(fin->C);
        if (vec4) { clr::launch_k(bwd_finish_kernel<QT, 4>, ctas + n_fin, kThreads, 0, st, p, *fin, n_fin); }
        else { clr::launch_k(bwd_finish_kernel<QT, 1>, ctas + n_fin, kThreads, 0, st, p, *fin, n_fin); }
        return launch_status();
    }
    if (vec4) { clr::launch_k(pool_bwd_kernel<QT, 4>, ctas, kThreads, 0, st, p); }
    else { clr::launch_k(pool_bwd_kernel<QT, 1>, ctas, kThreads, 0, st, p); }
    return launch_status();
}

int pool_bwd_impl(const BwdDom* doms, int ndom, int C, int HW, int K, cudaStream_t st, const DiscFinishParams* fin = nullptr,
                  int trace_id = TR_BWD_BOTH, unsigned int* gate = nullptr, int gate_dom = 0, float* gate_err = nullptr,
                  unsigned int gate_n_ext = 0) {
    CLR_CHECK_ARG(ndom >= 1 && ndom <= 2 && C > 0 && HW > 0 && K >= 1 && K <= CLR_MAX_K);
    bool vec4 = (HW % 4 == 0);
    int Qmax = 0;
    for (int d = 0; d < ndom; ++d) {
        const BwdDom& D = doms[d];
        CLR_CHECK_ARG(D.w && D.g && D.sums && D.grad && D.B > 0);
        CLR_CHECK_ARG(D.fmt == CLR_W_COMPLEMENT || D.fmt == CLR_W_EXPLICIT);
        CLR_CHECK_ARG(D.Kx >= 0 && D.Kx <= K && (D.Kx == 0 || (D.xcoef && D.xtab)));
        if (!aligned4(D.w) || !aligned4(D.grad)) return CLR_ERR_ALIGN;
        vec4 = vec4 && aligned16(D.w) && aligned16(D.grad) && (D.Kx == 0 || aligned16(D.xcoef));
        const int Q = (D.fmt == CLR_W_COMPLEMENT ? K : D.nrows) + D.Kx;
        CLR_CHECK_ARG(D.fmt == CLR_W_COMPLEMENT || (D.nrows >= 1 && D.nrows <= 2 * CLR_MAX_K));
        if (Q > Qmax) Qmax = Q;
    }
    BwdParams p{};
    p.trace_id = trace_id;
    p.ndom = ndom; p.C = C; p.HW = HW; p.K = K;
    if (gate && fin) { p.gate = gate; p.gate_n = (unsigned int)disc_finish_ctas(fin->C); p.gate_dom = gate_dom; p.gate_err = gate_err; }
    else if (gate && gate_n_ext) { p.gate = gate; p.gate_n = gate_n_ext; p.gate_dom = gate_dom; p.gate_err = gate_err; p.nowait = 1; }
    const int pxb = kThreads * (vec4 ? 4 : 1);
    p.nPx = (HW + pxb - 1) / pxb;
    p.nSpan = (C + kBwdSpan - 1) / kBwdSpan;
    long long total = 0;
    for (int d = 0; d < ndom; ++d) {
        p.dom[d] = doms[d];
        const long long ctas = (long long)doms[d].B * p.nPx * p.nSpan;
        if (ctas > 0x3fffffff) return CLR_ERR_UNSUPPORTED;
        p.dom[d].ctas = (int)ctas;
        total += ctas;
    }
    if (total > 0x7fffffff) return CLR_ERR_UNSUPPORTED;
    const int ctas = (int)total;
    if (Qmax <= 2) return launch_bwd<2>(p, vec4, ctas, st, fin);
    if (Qmax <= 3) return launch_bwd<3>(p, vec4, ctas, st, fin);
    if (Qmax <= 4) return launch_bwd<4>(p, vec4, ctas, st, fin);
    if (Qmax <= 6) return launch_bwd<6>(p, vec4, ctas, st, fin);
    if (Qmax <= 8) return launch_bwd<8>(p, vec4, ctas, st, fin);
    if (Qmax <= 12) return launch_bwd<12>(p, vec4, ctas, st, fin);
    if (Qmax <= 16) return launch_bwd<16>(p, vec4, ctas, st, fin);
    return launch_bwd<24>(p, vec4, ctas, st, fin);
}

// one domain; `f` != NULL: the disc finish body rides as the first CTAs of the launch.  `source` picks the trace slot.
int pool_bwd_one(const clr_bwd_dom* dom, int C, int HW, int K, const DiscFinishParams* f, bool source, cudaStream_t st) {
    if (!dom) return CLR_ERR_BAD_ARG;
    BwdDom d{dom->w, dom->g, dom->sums, dom->xcoef, dom->xtab, dom->grad, dom->scale_dev, dom->scale, dom->fmt, dom->B, dom->Kx,
             0, 2 * K, 0, 0.f};
    return pool_bwd_impl(&d, 1, C, HW, K, st, f, source ? TR_BWD_S : TR_BWD_T);
}

// [disc finish | gradient of `first` | gradient of `gated`] in ONE launch: the CTAs of `gated` (the source features, whose
// table needs the finish's output) wait on `gate`, a counter zeroed earlier in the same step (step.cu).
int pool_bwd_merged(const clr_bwd_dom* first, const clr_bwd_dom* gated, int C, int HW, int K, const DiscFinishParams* f,
                    unsigned int* gate, float* gate_err, cudaStream_t st) {
    if (!first || !gated || !f || !gate) return CLR_ERR_BAD_ARG;
    BwdDom d[2];
    const clr_bwd_dom* src[2] = {first, gated};
    for (int i = 0; i < 2; ++i)
        d[i] = BwdDom{src[i]->w, src[i]->g, src[i]->sums, src[i]->xcoef, src[i]->xtab, src[i]->grad, src[i]->scale_dev,
                      src[i]->scale, src[i]->fmt, src[i]->B, src[i]->Kx, 0, 2 * K, 0, 0.f};
    return pool_bwd_impl(d, 2, C, HW, K, st, f, TR_BWD_BOTH, gate, 1, gate_err);
}

// The same two gradient maps behind a disc finish that was launched on its own (gate_signal = gate): [gradient of `first` |
// gradient of `gated`], no griddepcontrol.wait, the gated CTAs wait for `gate_n` finish CTAs.
int pool_bwd_gated(const clr_bwd_dom* first, const clr_bwd_dom* gated, int C, int HW, int K, unsigned int* gate, unsigned int gate_n,
                   float* gate_err, cudaStream_t st) {
    if (!first || !gated || !gate || !gate_n) return CLR_ERR_BAD_ARG;
    BwdDom d[2];
    const clr_bwd_dom* src[2] = {first, gated};
    for (int i = 0; i < 2; ++i)
        d[i] = BwdDom{src[i]->w, src[i]->g, src[i]->sums, src[i]->xcoef, src[i]->xtab, src[i]->grad, src[i]->scale_dev,
                      src[i]->scale, src[i]->fmt, src[i]->B, src[i]->Kx, 0, 2 * K, 0, 0.f};
    return pool_bwd_impl(d, 2, C, HW, K, st, nullptr, TR_BWD_BOTH, gate, 1, gate_err, gate_n);
}

}  // namespace clr

extern "C" {

int clr_pool_bwd(const float* w, int fmt, int B, int C, int HW, int K,
                 const float* g, const float* sums, float scale,
                 const float* xcoef, const float* xtab, int Kx,
                 float* grad, clr_stream_t stream) {
    clr::BwdDom d{w, g, sums, xcoef, xtab, grad, nullptr, scale, fmt, B, Kx, 0, 2 * K, 0, 0.f};
    return clr::pool_bwd_impl(&d, 1, C, HW, K, static_cast<cudaStream_t>(stream));
}

int clr_pool_bwd_multi(const clr_bwd_dom* doms, int ndom, int C, int HW, int K, clr_stream_t stream) {
    if (!doms || ndom < 1 || ndom > 2) return CLR_ERR_BAD_ARG;
    clr::BwdDom d[2];
    for (int i = 0; i < ndom; ++i)
        d[i] = clr::BwdDom{doms[i].w, doms[i].g, doms[i].sums, doms[i].xcoef, doms[i].xtab, doms[i].grad,
                           doms[i].scale_dev, doms[i].scale, doms[i].fmt, doms[i].B, doms[i].Kx, 0, 2 * K, 0, 0.f};
    return clr::pool_bwd_impl(d, ndom, C, HW, K, static_cast<cudaStream_t>(stream));
}

/* bmm-style (per-sample normalised) pooling, adjoint w.r.t. the features (Trainer_prototype.py:364-383 under autograd):
 * grad[b,c,p] = scale * sum_r g[r][c] / (N_b[r] + n_add) * rows[b,r,p]   (scale carries the 1/B of the batch mean) */
int clr_pool_bwd_ps(const float* rows, int B, int C, int HW, int R, const float* g, const float* sums_b, float n_add,
                    float scale, float* grad, clr_stream_t stream) {
    if (R < 1 || R > 2 * CLR_MAX_K) return CLR_ERR_BAD_ARG;
    clr::BwdDom d{rows, g, sums_b, nullptr, nullptr, grad, nullptr, scale, CLR_W_EXPLICIT, B, 0, 0, R, 1, n_add};
    return clr::pool_bwd_impl(&d, 1, C, HW, /*K (unused for explicit rows)*/ 1, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

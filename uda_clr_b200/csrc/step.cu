// The fused CLR step: host-side orchestration of the kernels (no device code of its own).
//
// What the reference runs as ~150 eager ATen launches and ~25 passes over each [B,C,H,W] feature map per
// step (Trainer_prototype_full.py:328-449 plus the two bytecode-only losses) becomes:
//   forward : 1 read of xs + 1 read of xt (pooling, one launch) + 1 more read of xs (discriminative)
//   backward: 1 write of gxs + 1 write of gxt (one launch)
// plus O(K*C) glue kernels, with no host synchronisation (.item()) anywhere.
#include "clr_common.cuh"
#include "clr_internal.h"
#include "clr_finish.cuh"
#include <string.h>

namespace clr {

struct StepWs {
    char* pool;        size_t pool_bytes;
    char* rows;        size_t rows_bytes;
    float* hinge;      size_t hinge_bytes;
    double* cons;      size_t cons_bytes;
    double* fin;       size_t fin_bytes;     // pool_finish: per-CTA loss partials + the last-CTA counter
    size_t total;
};

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

static StepWs carve(const clr_step_args* a) {
    StepWs w{};
    const int HW = a->H * a->W;
    w.pool_bytes = align_up(pool_partial_bytes(a->B_s, a->C, HW, 2 * a->K) + pool_partial_bytes(a->B_t, a->C, HW, 2 * a->K));
    // discriminative pass: fused-kernel partials ([parts][K][C+1] + [parts]) or, for the two-pass fallback,
    // the row-pooling partials; the larger of the two
    size_t rows = a->use_disc ? pool_partial_bytes(a->B_s, a->C, HW, a->K) : 0;
    if (a->use_disc && clr_disc_fused_ws_bytes(a->C, a->K) > rows) rows = clr_disc_fused_ws_bytes(a->C, a->K);
    w.rows_bytes = align_up(rows);
    w.hinge_bytes = a->use_disc ? align_up(sizeof(float) * (size_t)clr_disc_partials_cap() * (1 + a->K)) : 0;
    w.cons_bytes = a->use_cons ? align_up(clr_cons_ws_bytes()) : 0;
    // pool finish: per-CTA loss partials of the two halves (source / target, schedule 2; the one-body form uses the first) + the
    // step's 8 counter words (last-CTA counters, completion counters, gate; all zeroed by the step's first pooling kernel)
    w.fin_bytes = align_up(sizeof(double) * 2 * (size_t)pool_finish_ctas(a->C) * (2 + CLR_MAX_K) + 64);
    char* base = static_cast<char*>(a->ws);
    size_t off = 0;
    w.pool = base + off; off += w.pool_bytes;
    w.rows = base + off; off += w.rows_bytes;
    w.hinge = reinterpret_cast<float*>(base + off); off += w.hinge_bytes;
    w.cons = reinterpret_cast<double*>(base + off); off += w.cons_bytes;
    w.fin = reinterpret_cast<double*>(base + off); off += w.fin_bytes;
    w.total = off;
    return w;
}

static int check_args(const clr_step_args* a) {
    CLR_CHECK_ARG(a && a->B_s > 0 && a->B_t > 0 && a->C > 0 && a->H > 0 && a->W > 0 && a->K >= 1 && a->K <= CLR_MAX_K);
    CLR_CHECK_ARG(a->xs && a->ys && a->xt && a->packed1 && a->packed2 && a->P_s && a->P_t && a->g_s && a->g_t &&
                  a->losses && a->stored_s && a->stored_t && a->ws);
    if (a->use_retrify)
        CLR_CHECK_ARG(a->oT_before && a->preds && a->std_map && a->pred_mean && a->wt_retrify && a->masks &&
                      a->Hi > 0 && a->Wi > 0 && a->T >= 1);
    else
        CLR_CHECK_ARG(a->wt != nullptr);
    if (a->use_disc) CLR_CHECK_ARG(a->disc_coef && a->disc_vec && a->disc_beta && a->xtab && a->npx_global > 0);
    if (a->use_cons) CLR_CHECK_ARG(a->oT && a->oT_aug && a->Hi > 0 && a->Wi > 0 && (a->use_retrify || a->masks));
    if (a->world > 1) {
        CLR_CHECK_ARG(a->world <= CLR_MAX_WORLD && a->rank >= 0 && a->rank < a->world && a->seq != 0);
        for (int q = 0; q < a->world; ++q) CLR_CHECK_ARG(a->peer_rx[q] != nullptr);
    }
    if (a->ws_bytes < carve(a).total) return CLR_ERR_WORKSPACE;
    return CLR_OK;
}

// counter words: [0] last-CTA counter (one-body finish / target half), [1] done_fin, [2] done_all of the [finish | consistency]
// launch, [3] gate of the merged backward, [4] last-CTA counter of the source half, [5] source half: vectors published, [6] source half: CTAs finished
static unsigned int* step_counters(const StepWs& w, int C) {
    return reinterpret_cast<unsigned int*>(w.fin + 2 * (size_t)pool_finish_ctas(C) * (2 + CLR_MAX_K));
}
static const float* target_weights(const clr_step_args* a) { return a->use_retrify ? a->wt_retrify : a->wt; }
static int target_fmt(const clr_step_args* a) { return a->use_retrify ? CLR_W_EXPLICIT : a->wt_fmt; }

}  // namespace clr

extern "C" {

size_t clr_step_ws_bytes(const clr_step_args* a) {
    if (!a) return 0;
    clr_step_args tmp = *a;
    tmp.ws = nullptr;
    return clr::carve(&tmp).total;
}

int clr_step_fwd_a(const clr_step_args* a, clr_stream_t stream) {
    int rc = clr::check_args(a);
    if (rc != CLR_OK) return rc;
    const clr::StepWs w = clr::carve(a);
    const int HW = a->H * a->W, R = 2 * a->K;
    if (a->use_retrify) {
        rc = clr_mc_stats(a->preds, a->T, a->B_t, a->K, a->Hi, a->Wi, a->std_map, a->pred_mean, stream);
        if (rc != CLR_OK) return rc;
        rc = clr_retrify_weights(a->oT_before, a->pred_mean, a->std_map, a->preds, a->T, a->B_t, a->K, a->H, a->W, a->Hi, a->Wi,
                                 a->pseudo_thr, a->std_thr, a->wt_retrify, a->masks, nullptr, nullptr, stream);
        if (rc != CLR_OK) return rc;
    }
    float* sums_s = a->packed1;
    float* sums_t = a->packed1 + (size_t)R * (a->C + 1);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (a->ev_pool_begin) cudaEventRecord(static_cast<cudaEvent_t>(a->ev_pool_begin), st);
    // target first, source last (an evict-last L2 policy on xs for the discriminative re-read was measured slower:
    // 134 MB > the 126 MB L2)
    rc = clr::pool_fwd_impl(a->xt, clr::target_weights(a), clr::target_fmt(a), a->B_t, sums_t,
                            a->xs, a->ys, CLR_W_COMPLEMENT, a->B_s, sums_s,
                            a->C, HW, R, w.pool, w.pool_bytes, st, 0, 0);
    if (a->ev_pool_end) cudaEventRecord(static_cast<cudaEvent_t>(a->ev_pool_end), st);
    return rc;
}

namespace clr {
struct PendingPack { const float* hinge; int n_hinge, hinge_stride; const double* cons; int n_cons; };
static int step_fwd_b_impl(const clr_step_args* a, clr_stream_t stream, PendingPack* defer);
}  // namespace clr

int clr_step_fwd_b(const clr_step_args* a, clr_stream_t stream) { return clr::step_fwd_b_impl(a, stream, nullptr); }

// `defer` != NULL: leave the per-CTA partials unsummed and describe them (the caller's finalize kernel sums them).
static int clr::step_fwd_b_impl(const clr_step_args* a, clr_stream_t stream, clr::PendingPack* defer) {
    int rc = clr::check_args(a);
    if (rc != CLR_OK) return rc;
    const clr::StepWs w = clr::carve(a);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int HW = a->H * a->W, R = 2 * a->K, C = a->C, K = a->K;
    const float* sums_s = a->packed1;
    const float* sums_t = a->packed1 + (size_t)R * (C + 1);
    rc = clr_align_finalize(sums_s, sums_t, K, C, a->stored_s, a->stored_t, a->first_s, a->first_t, a->decay,
                            a->w_intra, a->w_inter, a->P_s, a->P_t, nullptr, nullptr, a->g_s, a->g_t,
                            a->use_disc ? a->disc_vec : nullptr, a->use_disc ? a->disc_beta : nullptr, a->losses, stream);
    if (rc != CLR_OK) return rc;
    int n_hinge = 0, n_cons = 0;
    if (a->use_cons) {
        rc = clr::cons_fwd_partials(a->oT, a->oT_aug, a->masks, a->B_t, K, a->Hi, a->Wi, a->H, a->W,
                                    a->cons_threshold, w.cons, &n_cons, st);
        if (rc != CLR_OK) return rc;
    }
    int hinge_stride = 1 + K;
    const float* hinge_src = w.hinge;
    if (a->use_disc) {
        // one read of xs: dot products + active-set sums on the same shared-memory tile
        rc = CLR_ERR_UNSUPPORTED;
        if (clr::tunables().disc_impl != 1) {
            float* partial = reinterpret_cast<float*>(w.rows);
            float* hinge = partial + (size_t)320 * K * (C + 1);   // layout of clr_disc_fused_ws_bytes: [320][K][C+1] | [320]
            n_hinge = 320;
            rc = clr::disc_fused_impl(a->xs, a->ys, a->B_s, C, HW, K, a->disc_vec, a->disc_beta, a->margin,
                                      a->disc_coef, nullptr, partial, hinge, &n_hinge, st);
            if (rc == CLR_OK) {
                clr::launch_partial_reduce(partial, n_hinge, K, C, a->packed2, st);
                hinge_src = hinge;
                hinge_stride = 1;
            }
        }
        if (rc == CLR_ERR_UNSUPPORTED) {
            // two-pass form: per-pixel dots (read 1), then pooling of xs with the coefficient planes (read 2)
            rc = clr::disc_fwd_impl(a->xs, a->ys, a->B_s, C, HW, K, a->disc_vec, a->disc_beta, a->margin,
                                    a->disc_coef, nullptr, w.hinge, clr_disc_partials_cap(), &n_hinge, st);
            if (rc != CLR_OK) return rc;
            rc = clr::pool_fwd_impl(a->xs, a->disc_coef, CLR_W_EXPLICIT, a->B_s, a->packed2, nullptr, nullptr, 0, 0, nullptr,
                                    C, HW, K, w.rows, w.rows_bytes, st);
        }
        if (rc != CLR_OK) return rc;
    }
    if (defer) {
        *defer = clr::PendingPack{a->use_disc ? hinge_src : nullptr, n_hinge, hinge_stride, a->use_cons ? w.cons : nullptr, n_cons};
        return CLR_OK;
    }
    clr::launch_step_pack(a->use_disc ? hinge_src : nullptr, n_hinge, hinge_stride, a->use_cons ? w.cons : nullptr, n_cons,
                          a->packed2 + (size_t)K * (C + 1), st);
    return clr::launch_status();
}

int clr_step_fwd_c(const clr_step_args* a, clr_stream_t stream) {
    int rc = clr::check_args(a);
    if (rc != CLR_OK) return rc;
    const float ema = a->first_s ? 1.0f : (float)a->decay;
    return clr_disc_finalize(a->packed2, a->P_s, a->K, a->C, a->npx_global, a->w_disc, ema, a->grad_scale,
                             a->g_s, a->xtab, a->w_intra, a->w_inter, a->w_aug, a->aug_weight,
                             a->use_disc, a->use_cons, a->losses, stream);
}

// Single-GPU forward: no exchange points, so each partial reduce is folded into the finalize that consumes it
// (clr_finish.cuh), and the two finish bodies ride as the first CTAs of the streaming launch that follows them:
//   mc_stats -> retrify_weights -> pooling (xt + xs) -> [pool finish | consistency] -> discriminative
//            -> [disc finish | gradient of xt] -> gradient of xs                       (7 launches; was 10)
// `defer` != NULL: do not launch the disc finish; describe it instead (clr_step_run co-schedules it with the backward).
namespace clr {
static void bwd_doms(const clr_step_args* a, clr_bwd_dom (&d)[2]);

// receive-buffer layout (64-bit words): exchange 1 (packed1) [2][world][n1] | exchange 2 (packed2 + tail) [2][world][n2]
static size_t xchg_words1(int K, int C) { return (size_t)2 * 2 * K * (C + 1); }
static size_t xchg_words2(int K, int C) { return (size_t)K * (C + 1) + 4; }
static PeerXchg make_xchg(const clr_step_args* a, int which) {
    PeerXchg x{};
    x.world = a->world > 1 ? a->world : 1;
    x.rank = a->rank;
    x.seq = a->seq;
    x.err = a->losses + 7;
    x.pull = tunables().xchg_pull;
    const size_t n1 = xchg_words1(a->K, a->C), n2 = xchg_words2(a->K, a->C);
    x.n = (int)(which == 1 ? n1 : n2);
    const size_t off = which == 1 ? 0 : 2 * (size_t)x.world * n1;
    for (int q = 0; q < x.world && q < CLR_MAX_WORLD; ++q)
        x.rx[q] = a->world > 1 ? static_cast<unsigned long long*>(a->peer_rx[q]) + off : nullptr;
    return x;
}

// Schedule 2 of the single-call fused step (clr3: retrify target + discriminative term, one-read kernel available):
//
//   pool(xs) -> finish/source  ||  mc_stats -> retrify_weights -> pool(xt) -> [finish/target + align | consistency]  ||  disc(xs)
//            -> [disc finish | gradient of xt | gradient of xs (gated)]
//
// The source features do not depend on the MC statistics, and the discriminative pass depends on the SOURCE prototypes
// only.  So the source map is pooled first, the source half of the finish (EMA'd source prototypes, the discriminative
// term's vectors) runs as 33 CTAs of its own launch while the whole machine streams the MC logits (mc_stats skips its
// griddepcontrol.wait), and the discriminative kernel no longer waits for the latency chain that follows the target
// pooling -- that chain (target prototypes, alignment loss, all prototype gradients) now has the consistency AND the
// discriminative pass to hide behind.  Sharded, both halves carry their part of the in-kernel exchange: every cross-GPU
// rendezvous of the step then sits next to >= 20 us of independent streaming, so rank skew is absorbed instead of added.
// Same arithmetic as schedule 1 (bit-identical results: tests/test_gpu_step.py).  Measured (profiles/r02_schedule2.md): on ONE
// GPU it is 9 us SLOWER than schedule 1 (0.1855 vs 0.1762 ms) -- two pooling launches cost 51 instead of 45 us (each pays its
// own ramp and one-item tail), one more kernel boundary, and the discriminative CTAs still cannot become resident before the
// consistency CTAs leave.  Sharded, its cost is almost flat in the world size (+3.5 us at 8 GPUs against +14.8 us for schedule 1):
// 2 GPUs 0.1854 vs 0.1837 ms, 4 GPUs 0.1875 vs 0.1867 ms, 8 GPUs 0.1888 vs 0.1907 ms (two repetitions each, interleaved).  So
// "sched" = 0 (auto) picks schedule 2 -- with the split disc finish -- from 8 ranks on, schedule 1 below; 1 / 2 force one.
static bool use_schedule2(const clr_step_args* a) {
    const Tunables& t = tunables();
    const bool possible = a->use_retrify && a->use_disc && t.disc_impl != 1 && !t.finish_off && !t.hfuse_off && !t.flag_dep_off && !t.mc_fuse;
    return possible && (t.sched == 2 || (t.sched == 0 && a->world >= 8));
}

static int step_fwd_v2(const clr_step_args* a, cudaStream_t st, DiscFinishParams* defer, int* deferred) {
    const StepWs w = carve(a);
    clr_stream_t stream = st;
    const int HW = a->H * a->W, R = 2 * a->K, C = a->C, K = a->K;
    const int nfin = pool_finish_ctas(C);
    float* sums_s = a->packed1;
    float* sums_t = a->packed1 + (size_t)R * (C + 1);
    double* fin_s = w.fin;
    double* fin_t = w.fin + (size_t)nfin * (2 + CLR_MAX_K);
    unsigned int* counter = step_counters(w, C);
    const int early = tunables().fin_early_off ? 0 : 1;
    int rc;
    // 1. source pooling (first kernel of the step: zeroes the counters)
    PoolLayout lay_s{}, lay_t{};
    const size_t pb_s = pool_partial_bytes(a->B_s, C, HW, R);
    if (a->ev_pool_begin) cudaEventRecord(static_cast<cudaEvent_t>(a->ev_pool_begin), st);
    rc = pool_fwd_impl(a->xs, a->ys, CLR_W_COMPLEMENT, a->B_s, sums_s, nullptr, nullptr, 0, 0, nullptr,
                       C, HW, R, w.pool, pb_s, st, 0, 0, &lay_s, counter);
    if (a->ev_pool_end) cudaEventRecord(static_cast<cudaEvent_t>(a->ev_pool_end), st);
    if (rc != CLR_OK) return rc;
    // 2. source half of the finish
    PoolFinishParams pf{};
    pf.K = K; pf.C = C; pf.d = (float)a->decay; pf.omd = (float)(1.0 - a->decay);
    pf.w_intra = a->w_intra; pf.w_inter = a->w_inter;
    pf.first[0] = a->first_s; pf.first[1] = a->first_t;
    pf.stored[0] = a->stored_s; pf.stored[1] = a->stored_t; pf.P[0] = a->P_s; pf.P[1] = a->P_t;
    pf.g[0] = a->g_s; pf.g[1] = a->g_t; pf.sums[0] = sums_s; pf.sums[1] = sums_t;
    pf.losses = a->losses; pf.x = make_xchg(a, 1);
    PoolFinishParams ps = pf;
    ps.mode = 1;
    ps.partial[0] = lay_s.partial[0]; ps.slots[0] = lay_s.slots[0];
    ps.disc_vec = a->disc_vec; ps.disc_beta = a->disc_beta;
    ps.loss_partial = fin_s; ps.counter = counter + 4;
    ps.done_fin = counter + 5; ps.done_all = counter + 6; ps.early_signal = early;      // [5]: vectors out (early), [6]: CTA finished
    rc = pool_finish_launch(ps, st);
    if (rc != CLR_OK) return rc;
    // 3. MC statistics (not waiting for 2) + retrify weights
    rc = mc_stats_impl(a->preds, a->T, a->B_t, K, a->Hi, a->Wi, a->std_map, a->pred_mean, st, true, a->H);
    if (rc != CLR_OK) return rc;
    rc = clr_retrify_weights(a->oT_before, a->pred_mean, a->std_map, a->preds, a->T, a->B_t, K, a->H, a->W, a->Hi, a->Wi,
                             a->pseudo_thr, a->std_thr, a->wt_retrify, a->masks, nullptr, nullptr, stream);
    if (rc != CLR_OK) return rc;
    // 4. target pooling
    rc = pool_fwd_impl(a->xt, a->wt_retrify, CLR_W_EXPLICIT, a->B_t, sums_t, nullptr, nullptr, 0, 0, nullptr,
                       C, HW, R, w.pool + pb_s, w.pool_bytes - pb_s, st, 0, 0, &lay_t, nullptr, nullptr, TR_POOL_T);
    if (rc != CLR_OK) return rc;
    // 5. [target half + alignment | consistency]
    PoolFinishParams pt = pf;
    pt.mode = 2;
    pt.partial[1] = lay_t.partial[0]; pt.slots[1] = lay_t.slots[0];
    pt.loss_partial = fin_t; pt.counter = counter;
    pt.done_fin = nullptr; pt.done_all = counter + 2; pt.early_signal = 0;
    // waits for the source half's CTAs to have FINISHED ([6], not just published their vectors): completion of this launch then
    // implies completion of that one, and the discriminative kernel's exit check makes both visible to the backward launch
    pt.wait_src = counter + 6; pt.wait_src_n = (unsigned int)nfin; pt.wait_err = a->losses + 7;
    int n_cons = 0;
    unsigned int producers = (unsigned int)nfin;
    if (a->use_cons) {
        rc = cons_fwd_partials(a->oT, a->oT_aug, a->masks, a->B_t, K, a->Hi, a->Wi, a->H, a->W,
                               a->cons_threshold, w.cons, &n_cons, st, &pt);
        if (rc != CLR_OK) return rc;
        producers += (unsigned int)n_cons;
    } else {
        rc = pool_finish_launch(pt, st);
        if (rc != CLR_OK) return rc;
    }
    // 6. discriminative pass: depends on the source half only (flag), checks the [5]-launch's completion before it exits
    const float ema = a->first_s ? 1.0f : (float)a->decay;
    const double* cons = a->use_cons ? w.cons : nullptr;
    float* partial = reinterpret_cast<float*>(w.rows);
    float* hinge = partial + (size_t)320 * K * (C + 1);   // layout of clr_disc_fused_ws_bytes: [320][K][C+1] | [320]
    int n_hinge = 320;
    DiscFlagDep dep{counter + 5, counter + 2, (unsigned int)nfin, producers, a->losses + 7, early ? fin_s : nullptr};
    rc = disc_fused_impl(a->xs, a->ys, a->B_s, C, HW, K, a->disc_vec, a->disc_beta, a->margin,
                         a->disc_coef, nullptr, partial, hinge, &n_hinge, st, &dep);
    if (rc == CLR_OK) {
        DiscFinishParams df{};
        df.partial = partial; df.slots = n_hinge; df.packed2 = a->packed2; df.P_s = a->P_s; df.g_s = a->g_s;
        df.xtab = a->xtab; df.losses = a->losses; df.K = K; df.C = C; df.npx = a->npx_global;
        df.coef = (float)(2.0 / ((double)C * a->npx_global));
        df.w_disc = a->w_disc; df.ema_factor = ema; df.gscale = a->grad_scale; df.w_intra = a->w_intra;
        df.w_inter = a->w_inter; df.w_aug = a->w_aug; df.aug_weight = a->aug_weight; df.use_cons = a->use_cons;
        df.ps = PackSrc{hinge, n_hinge, 1, cons, n_cons};
        df.x = make_xchg(a, 2);
        if (defer) { *defer = df; *deferred = 1; return CLR_OK; }
        return disc_finish_launch(df, st);
    }
    if (rc != CLR_ERR_UNSUPPORTED) return rc;
    // geometry the one-read kernel cannot take (ragged planes): the two-pass form (ordinary kernel boundaries from here on)
    if (a->world > 1) return CLR_ERR_UNSUPPORTED;
    n_hinge = 0;
    rc = disc_fwd_impl(a->xs, a->ys, a->B_s, C, HW, K, a->disc_vec, a->disc_beta, a->margin,
                       a->disc_coef, nullptr, w.hinge, clr_disc_partials_cap(), &n_hinge, st);
    if (rc != CLR_OK) return rc;
    rc = pool_fwd_impl(a->xs, a->disc_coef, CLR_W_EXPLICIT, a->B_s, a->packed2, nullptr, nullptr, 0, 0, nullptr,
                       C, HW, K, w.rows, w.rows_bytes, st);
    if (rc != CLR_OK) return rc;
    return disc_finalize_impl(a->packed2, a->P_s, K, C, a->npx_global, a->w_disc, ema, a->grad_scale,
                              a->g_s, a->xtab, a->w_intra, a->w_inter, a->w_aug, a->aug_weight,
                              a->use_disc, a->use_cons, a->losses, w.hinge, n_hinge, 1 + K, cons, n_cons, st);
}

static int step_fwd_core(const clr_step_args* a, cudaStream_t st, DiscFinishParams* defer, int* deferred) {
    if (use_schedule2(a)) return step_fwd_v2(a, st, defer, deferred);
    const StepWs w = carve(a);
    clr_stream_t stream = st;
    const int HW = a->H * a->W, R = 2 * a->K, C = a->C, K = a->K;
    int rc;
    if (a->use_retrify) {
        // "mc_fuse" = 1: one pass (MC statistics + bilinear taps + pseudo-labels + masks + weight planes, pred_mean not
        // materialised).  Measured on B200: 38.9 us against 31.5 + 3.9 us for the two kernels (the per-CTA barrier and
        // epilogue drain the memory pipeline of a kernel that is co-limited by MUFU issue), so the default stays two kernels.
        rc = !tunables().mc_fuse ? CLR_ERR_UNSUPPORTED
                                   : mc_retrify_fused(a->preds, a->oT_before, a->T, a->B_t, K, a->H, a->W, a->Hi, a->Wi,
                                                      a->pseudo_thr, a->std_thr, a->std_map, nullptr, a->wt_retrify, a->masks, st);
        if (rc == CLR_ERR_UNSUPPORTED) {
            rc = mc_stats_impl(a->preds, a->T, a->B_t, K, a->Hi, a->Wi, a->std_map, a->pred_mean, st, false, a->H);
            if (rc != CLR_OK) return rc;
            rc = clr_retrify_weights(a->oT_before, a->pred_mean, a->std_map, a->preds, a->T, a->B_t, K, a->H, a->W, a->Hi, a->Wi,
                                     a->pseudo_thr, a->std_thr, a->wt_retrify, a->masks, nullptr, nullptr, stream);
        }
        if (rc != CLR_OK) return rc;
    }
    float* sums_s = a->packed1;
    float* sums_t = a->packed1 + (size_t)R * (C + 1);
    unsigned int* counter = step_counters(w, C);
    PoolLayout lay{};
    if (a->ev_pool_begin) cudaEventRecord(static_cast<cudaEvent_t>(a->ev_pool_begin), st);
    rc = pool_fwd_impl(a->xt, target_weights(a), target_fmt(a), a->B_t, sums_t,
                       a->xs, a->ys, CLR_W_COMPLEMENT, a->B_s, sums_s,
                       C, HW, R, w.pool, w.pool_bytes, st, 0, 0, &lay, counter);
    if (a->ev_pool_end) cudaEventRecord(static_cast<cudaEvent_t>(a->ev_pool_end), st);
    if (rc != CLR_OK) return rc;
    PoolFinishParams pf{};
    pf.partial[0] = lay.partial[1]; pf.partial[1] = lay.partial[0];       // pooling ran (target, source)
    pf.slots[0] = lay.slots[1]; pf.slots[1] = lay.slots[0];
    pf.sums[0] = sums_s; pf.sums[1] = sums_t;
    pf.stored[0] = a->stored_s; pf.stored[1] = a->stored_t; pf.P[0] = a->P_s; pf.P[1] = a->P_t;
    pf.g[0] = a->g_s; pf.g[1] = a->g_t; pf.first[0] = a->first_s; pf.first[1] = a->first_t;
    pf.K = K; pf.C = C; pf.d = (float)a->decay; pf.omd = (float)(1.0 - a->decay);
    pf.w_intra = a->w_intra; pf.w_inter = a->w_inter;
    pf.disc_vec = a->use_disc ? a->disc_vec : nullptr; pf.disc_beta = a->use_disc ? a->disc_beta : nullptr;
    pf.losses = a->losses; pf.loss_partial = w.fin; pf.counter = counter;
    pf.x = make_xchg(a, 1);
    // completion counters right behind the last-CTA counter (all three zeroed by the pooling kernel): the discriminative
    // kernel waits for the finish CTAs only, not for the grid (and its end-of-grid flush) they ride in
    const bool flag_dep = a->use_disc && tunables().disc_impl != 1 && !tunables().flag_dep_off;
    if (flag_dep) { pf.done_fin = counter + 1; pf.done_all = counter + 2; pf.early_signal = tunables().fin_early_off ? 0 : 1; }
    const bool align_only = !a->use_disc && !a->use_cons;      // the shipped trainer's step: totals come from the finish itself
    pf.write_total = align_only ? 1 : 0;
    int n_cons = 0;
    if (a->use_cons) {
        rc = cons_fwd_partials(a->oT, a->oT_aug, a->masks, a->B_t, K, a->Hi, a->Wi, a->H, a->W,
                               a->cons_threshold, w.cons, &n_cons, st, tunables().hfuse_off ? nullptr : &pf);
        if (rc != CLR_OK) return rc;
    }
    unsigned int producers = (unsigned int)pool_finish_ctas(C);      // CTAs that signal done_all
    if (!a->use_cons || tunables().hfuse_off) {
        rc = pool_finish_launch(pf, st);     // (stream order: after the consistency launch is fine, they are independent)
        if (rc != CLR_OK) return rc;
    } else {
        producers += (unsigned int)n_cons;
    }
    const float ema = a->first_s ? 1.0f : (float)a->decay;
    const double* cons = a->use_cons ? w.cons : nullptr;
    if (a->use_disc && tunables().disc_impl != 1) {
        float* partial = reinterpret_cast<float*>(w.rows);
        float* hinge = partial + (size_t)320 * K * (C + 1);   // layout of clr_disc_fused_ws_bytes: [320][K][C+1] | [320]
        int n_hinge = 320;
        // (a separate consistency launch sits between the finish launch and this kernel when hfuse_off: plain wait then)
        DiscFlagDep dep{counter + 1, counter + 2, (unsigned int)pool_finish_ctas(C), producers, a->losses + 7, tunables().fin_early_off ? nullptr : w.fin};
        const bool use_dep = flag_dep && !(a->use_cons && tunables().hfuse_off);
        rc = disc_fused_impl(a->xs, a->ys, a->B_s, C, HW, K, a->disc_vec, a->disc_beta, a->margin,
                             a->disc_coef, nullptr, partial, hinge, &n_hinge, st, use_dep ? &dep : nullptr);
        if (rc == CLR_OK) {
            DiscFinishParams df{};
            df.partial = partial; df.slots = n_hinge; df.packed2 = a->packed2; df.P_s = a->P_s; df.g_s = a->g_s;
            df.xtab = a->xtab; df.losses = a->losses; df.K = K; df.C = C; df.npx = a->npx_global;
            df.coef = (float)(2.0 / ((double)C * a->npx_global));
            df.w_disc = a->w_disc; df.ema_factor = ema; df.gscale = a->grad_scale; df.w_intra = a->w_intra;
            df.w_inter = a->w_inter; df.w_aug = a->w_aug; df.aug_weight = a->aug_weight; df.use_cons = a->use_cons;
            df.ps = PackSrc{hinge, n_hinge, 1, cons, n_cons};
            df.x = make_xchg(a, 2);
            if (defer && !tunables().hfuse_off) { *defer = df; *deferred = 1; return CLR_OK; }
            return disc_finish_launch(df, st);
        }
        if (rc != CLR_ERR_UNSUPPORTED) return rc;
    }
    // the in-kernel exchange of the discriminative / consistency numerators lives in disc_finish_body (one-read path)
    if (a->world > 1 && (a->use_disc || a->use_cons)) return CLR_ERR_UNSUPPORTED;
    if (align_only) return CLR_OK;
    int n_hinge = 0;
    if (a->use_disc) {
        // two-pass form: per-pixel dots (read 1), then pooling of xs with the coefficient planes (read 2)
        rc = disc_fwd_impl(a->xs, a->ys, a->B_s, C, HW, K, a->disc_vec, a->disc_beta, a->margin,
                           a->disc_coef, nullptr, w.hinge, clr_disc_partials_cap(), &n_hinge, st);
        if (rc != CLR_OK) return rc;
        rc = pool_fwd_impl(a->xs, a->disc_coef, CLR_W_EXPLICIT, a->B_s, a->packed2, nullptr, nullptr, 0, 0, nullptr,
                           C, HW, K, w.rows, w.rows_bytes, st);
        if (rc != CLR_OK) return rc;
    }
    return disc_finalize_impl(a->packed2, a->P_s, K, C, a->npx_global, a->w_disc, ema, a->grad_scale,
                              a->g_s, a->xtab, a->w_intra, a->w_inter, a->w_aug, a->aug_weight,
                              a->use_disc, a->use_cons, a->losses, a->use_disc ? w.hinge : nullptr, n_hinge, 1 + K,
                              cons, n_cons, st);
}

// the separate reduce / finalize kernels (the sharded path's kernels) on one GPU, for A/B runs ("finish_off" = 1)
static int step_fwd_unmerged(const clr_step_args* a, clr_stream_t stream) {
    int rc = clr_step_fwd_a(a, stream);
    if (rc != CLR_OK) return rc;
    PendingPack pk{};
    rc = step_fwd_b_impl(a, stream, &pk);
    if (rc != CLR_OK) return rc;
    const float ema0 = a->first_s ? 1.0f : (float)a->decay;
    return disc_finalize_impl(a->packed2, a->P_s, a->K, a->C, a->npx_global, a->w_disc, ema0, a->grad_scale,
                              a->g_s, a->xtab, a->w_intra, a->w_inter, a->w_aug, a->aug_weight,
                              a->use_disc, a->use_cons, a->losses, pk.hinge, pk.n_hinge, pk.hinge_stride,
                              pk.cons, pk.n_cons, static_cast<cudaStream_t>(stream));
}
}  // namespace clr

int clr_step_fwd(const clr_step_args* a, clr_stream_t stream) {
    int rc = clr::check_args(a);
    if (rc != CLR_OK) return rc;
    if (a->world > 1 && clr::tunables().finish_off) return CLR_ERR_UNSUPPORTED;
    if (clr::tunables().finish_off) return clr::step_fwd_unmerged(a, stream);
    return clr::step_fwd_core(a, static_cast<cudaStream_t>(stream), nullptr, nullptr);
}

namespace clr {
static void bwd_doms(const clr_step_args* a, clr_bwd_dom (&d)[2]) {
    const int R = 2 * a->K, C = a->C, K = a->K;
    const float* sums_s = a->packed1;
    const float* sums_t = a->packed1 + (size_t)R * (C + 1);
    d[0] = clr_bwd_dom{a->ys, a->g_s, sums_s, a->use_disc ? a->disc_coef : nullptr, a->use_disc ? a->xtab : nullptr,
                       a->gxs, a->gup, a->grad_scale, CLR_W_COMPLEMENT, a->B_s, a->use_disc ? K : 0};
    d[1] = clr_bwd_dom{target_weights(a), a->g_t, sums_t, nullptr, nullptr, a->gxt, a->gup, a->grad_scale,
                       target_fmt(a), a->B_t, 0};
}
}  // namespace clr

int clr_step_run(const clr_step_args* a, clr_stream_t stream) {
    int rc = clr::check_args(a);
    if (rc != CLR_OK) return rc;
    if (!a->gxs || !a->gxt) return CLR_ERR_BAD_ARG;
    if (clr::tunables().finish_off) {
        if (a->world > 1) return CLR_ERR_UNSUPPORTED;
        rc = clr::step_fwd_unmerged(a, stream);
        return rc != CLR_OK ? rc : clr_step_bwd(a, stream);
    }
    clr::DiscFinishParams df{};
    int deferred = 0;
    rc = clr::step_fwd_core(a, static_cast<cudaStream_t>(stream), &df, &deferred);
    if (rc != CLR_OK) return rc;
    if (!deferred) return clr_step_bwd(a, stream);
    // [disc finish | gradient of xt] in one launch, then the gradient of xs (needs the finish's table)
    clr_bwd_dom dd[2];
    clr::bwd_doms(a, dd);
    cudaStream_t s0 = static_cast<cudaStream_t>(stream);
    if (a->ev_bwd_begin) cudaEventRecord(static_cast<cudaEvent_t>(a->ev_bwd_begin), s0);
    // own launch for the disc finish: measured slower with schedule 1 (8 GPUs 0.1916 vs 0.1910 ms), faster with schedule 2
    // (2 GPUs 0.1854 vs 0.1931 ms) -- "dfin_split" = 0 (auto) follows the schedule, 1 = always, 2 = never
    const int split = clr::tunables().dfin_split;
    if (!clr::tunables().bwd_merge_off && (split == 1 || (split == 0 && clr::use_schedule2(a)))) {
        // [disc finish] as its own small launch (late trigger, bumps the gate), then [gradient of xt | gated gradient of xs]
        // WITHOUT griddepcontrol.wait: the finish grid's slow end-of-grid flush (it wrote peer memory) overlaps the gradient
        // write instead of extending the grid the next step has to wait for
        const clr::StepWs w = clr::carve(a);
        unsigned int* counter = clr::step_counters(w, a->C);
        df.gate_signal = counter + 3;
        rc = clr::disc_finish_launch(df, s0);
        if (rc == CLR_OK) rc = clr::pool_bwd_gated(&dd[1], &dd[0], a->C, a->H * a->W, a->K, counter + 3,
                                                   (unsigned int)clr::disc_finish_ctas(a->C), a->losses + 7, s0);
    } else if (!clr::tunables().bwd_merge_off) {
        // ONE launch: [disc finish | gradient of xt | gradient of xs]; the source CTAs are dispatched after the ~1000 target
        // CTAs and wait on the 4th counter word (zeroed by this step's pooling kernel, bumped by every finish CTA)
        const clr::StepWs w = clr::carve(a);
        unsigned int* counter = clr::step_counters(w, a->C);
        rc = clr::pool_bwd_merged(&dd[1], &dd[0], a->C, a->H * a->W, a->K, &df, counter + 3, a->losses + 7, s0);
    } else {
        rc = clr::pool_bwd_one(&dd[1], a->C, a->H * a->W, a->K, &df, false, s0);
        if (rc == CLR_OK) rc = clr::pool_bwd_one(&dd[0], a->C, a->H * a->W, a->K, nullptr, true, s0);
    }
    if (a->ev_bwd_end) cudaEventRecord(static_cast<cudaEvent_t>(a->ev_bwd_end), s0);
    if (rc != CLR_OK) return rc;
    if (a->use_cons && a->w_aug != 0.f && a->g_oT_aug) {
        const float* stats = a->packed2 + (size_t)a->K * (a->C + 1);
        rc = clr_cons_bwd(a->oT, a->oT_aug, a->masks, a->B_t, a->K, a->Hi, a->Wi, a->H, a->W, a->cons_threshold,
                          a->aug_weight, stats + 1, a->gup, a->grad_scale * a->w_aug, a->g_oT_aug, stream);
    }
    return rc;
}

int clr_step_bwd(const clr_step_args* a, clr_stream_t stream) {
    int rc = clr::check_args(a);
    if (rc != CLR_OK) return rc;
    if (!a->gxs || !a->gxt) return CLR_ERR_BAD_ARG;
    const int HW = a->H * a->W, C = a->C, K = a->K;
    clr_bwd_dom d[2];
    clr::bwd_doms(a, d);
    if (a->ev_bwd_begin) cudaEventRecord(static_cast<cudaEvent_t>(a->ev_bwd_begin), static_cast<cudaStream_t>(stream));
    rc = clr_pool_bwd_multi(d, 2, C, HW, K, stream);
    if (a->ev_bwd_end) cudaEventRecord(static_cast<cudaEvent_t>(a->ev_bwd_end), static_cast<cudaStream_t>(stream));
    if (rc != CLR_OK) return rc;
    if (a->use_cons && a->w_aug != 0.f && a->g_oT_aug) {
        // stats layout expected by clr_cons_bwd: [num, den, ...] = tail[1..2] of packed2
        const float* stats = a->packed2 + (size_t)K * (C + 1);   // {hinge num, cons num, cons den}: den at [2]
        rc = clr_cons_bwd(a->oT, a->oT_aug, a->masks, a->B_t, K, a->Hi, a->Wi, a->H, a->W, a->cons_threshold,
                          a->aug_weight, stats + 1, a->gup, a->grad_scale * a->w_aug, a->g_oT_aug, stream);
    }
    return rc;
}

int clr_step_schedule(const clr_step_args* a) {
    if (!a) return 0;
    return clr::use_schedule2(a) ? 2 : 1;
}

size_t clr_step_xchg_bytes(int world, int K, int C) {
    if (world < 1 || world > CLR_MAX_WORLD || K < 1 || K > CLR_MAX_K || C < 1) return 0;
    return sizeof(unsigned long long) * 2 * (size_t)world * (clr::xchg_words1(K, C) + clr::xchg_words2(K, C));
}

int clr_peer_alloc(size_t bytes, void** ptr) {
    if (!ptr || bytes == 0) return CLR_ERR_BAD_ARG;
    void* p = nullptr;
    CLR_RETURN_IF_CUDA(cudaMalloc(&p, bytes));
    CLR_RETURN_IF_CUDA(cudaMemset(p, 0, bytes));
    CLR_RETURN_IF_CUDA(cudaDeviceSynchronize());
    *ptr = p;
    return CLR_OK;
}
int clr_peer_free(void* ptr) {
    if (!ptr) return CLR_ERR_BAD_ARG;
    CLR_RETURN_IF_CUDA(cudaFree(ptr));
    return CLR_OK;
}
int clr_peer_export(void* ptr, unsigned char handle[64]) {
    if (!ptr || !handle) return CLR_ERR_BAD_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    CLR_RETURN_IF_CUDA(cudaIpcGetMemHandle(&h, ptr));
    memcpy(handle, &h, 64);
    return CLR_OK;
}
int clr_peer_open(const unsigned char handle[64], void** ptr) {
    if (!ptr || !handle) return CLR_ERR_BAD_ARG;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    void* p = nullptr;
    CLR_RETURN_IF_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *ptr = p;
    return CLR_OK;
}
int clr_peer_close(void* ptr) {
    if (!ptr) return CLR_ERR_BAD_ARG;
    CLR_RETURN_IF_CUDA(cudaIpcCloseMemHandle(ptr));
    return CLR_OK;
}

}  // extern "C"

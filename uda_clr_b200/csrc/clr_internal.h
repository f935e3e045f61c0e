// Cross-translation-unit entry points used by the fused step orchestration (step.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include "clr_b200.h"

namespace clr {

struct PoolLayout;
int pool_fwd_impl(const float* feat0, const float* w0, int fmt0, int B0, float* sums0,
                  const float* feat1, const float* w1, int fmt1, int B1, float* sums1,
                  int C, int HW, int R, void* ws, size_t ws_bytes, cudaStream_t st, int keep0 = 0, int keep1 = 0,
                  struct PoolLayout* skip_reduce_layout = nullptr, unsigned int* counter_reset = nullptr, float* mu0 = nullptr,
                  int trace_id = 2 /*TR_POOL*/);
size_t pool_partial_bytes(int B, int C, int HW, int R);
// sums[r][c] = sum_slot partial[slot][r][c] (fp64, fixed order) for a [slots][R][C+1] partial buffer
void launch_partial_reduce(const float* partial, int slots, int R, int C, float* sums, cudaStream_t st);

// One-read discriminative forward (disc_fused.cu); CLR_ERR_UNSUPPORTED -> use the two-pass form.
struct DiscFlagDep {   // flag dependency on the preceding [finish | consistency] launch (clr_common.cuh), or NULL
    const unsigned int* wait_fin; const unsigned int* wait_all;
    unsigned int wait_fin_n, wait_all_n;
    float* err;
    const double* beta_partial;   // the finish CTAs' loss partials ([wait_fin_n][2 + CLR_MAX_K]): the kernel sums beta itself
};
int disc_fused_impl(const float* xs, const float* ys, int B, int C, int HW, int K,
                    const float* disc_vec, const float* disc_beta, float margin,
                    float* coef, float* delta, float* partial, float* hinge, int* nparts, cudaStream_t st,
                    const DiscFlagDep* dep = nullptr);

int disc_fwd_impl(const float* xs, const float* ys, int B, int C, int HW, int K,
                  const float* disc_vec, const float* disc_beta, float margin,
                  float* coef, float* delta, float* partials, int partials_cap, int* nparts, cudaStream_t st);

struct PoolFinishParams;
int cons_fwd_partials(const float* oT, const float* oT_aug, const float* masks, int B, int K, int Hi, int Wi,
                      int H, int W, float threshold, double* partial, int* nblocks, cudaStream_t st,
                      const PoolFinishParams* fused_finish = nullptr);

void launch_step_pack(const float* hinge_partials, int n_hinge, int hinge_stride,
                      const double* cons_partials, int n_cons, float* tail, cudaStream_t st);

// clr_disc_finalize with the per-CTA partial sums folded in (no separate pack launch; single-GPU path).
int disc_finalize_impl(float* packed2, const float* P_s, int K, int C, double npx, float w_disc,
                       float ema_factor, float gscale, float* g_s, float* xtab,
                       float w_intra, float w_inter, float w_aug, float aug_weight, int use_disc, int use_cons,
                       float* losses, const float* hinge, int n_hinge, int hinge_stride,
                       const double* cons, int n_cons, cudaStream_t stream);

// Single-GPU fused step: merged reduce + finalize bodies (clr_finish.cuh), launched stand-alone (finalize.cu) or riding
// with the consistency pass (cons.cu) / the target-gradient write (pool_bwd.cu).
struct PoolFinishParams;
struct DiscFinishParams;
int pool_finish_launch(const PoolFinishParams& p, cudaStream_t st);
int disc_finish_launch(const DiscFinishParams& p, cudaStream_t st);
int pool_bwd_one(const clr_bwd_dom* dom, int C, int HW, int K, const DiscFinishParams* f, bool source, cudaStream_t st);
int pool_bwd_merged(const clr_bwd_dom* first, const clr_bwd_dom* gated, int C, int HW, int K, const DiscFinishParams* f,
                    unsigned int* gate, float* gate_err, cudaStream_t st);
int pool_bwd_gated(const clr_bwd_dom* first, const clr_bwd_dom* gated, int C, int HW, int K, unsigned int* gate, unsigned int gate_n,
                   float* gate_err, cudaStream_t st);

// clr_mc_stats with the option to skip griddepcontrol.wait (fused step, schedule 2: see the kernel)
int mc_stats_impl(const float* preds, int T, int B, int K, int Hi, int Wi, float* std_map, float* pred_mean, cudaStream_t st,
                  bool nowait, int tap_H = 0);
int retrify_weights_impl(const float* oT_before, const float* pred_mean, const float* std_map, const float* preds, int T,
                         int B, int K, int H, int W, int Hi, int Wi, float pseudo_thr, float std_thr,
                         float* weights, float* masks, float* pseudo_out, float* small_out, cudaStream_t stream, bool nowait);
// MC statistics + retrify weights in one pass (mc_stats.cu); CLR_ERR_UNSUPPORTED -> run the two kernels.
int mc_retrify_fused(const float* preds, const float* oT_before, int T, int B, int K, int H, int W, int Hi, int Wi,
                     float pseudo_thr, float std_thr, float* std_map, float* pred_mean /*nullable*/, float* weights,
                     float* masks, cudaStream_t st);

// Where pool_fwd_impl left its per-(b,chunk) partials (for callers that reduce them themselves).
struct PoolLayout { const float* partial[2]; int slots[2]; };

}  // namespace clr

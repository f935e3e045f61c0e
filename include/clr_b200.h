/*
 * clr_b200.h -- C ABI of the B200-native CLR hot path (libclr_b200.so).
 *
 * The reference (fengweie/UDA_CLR) has no FFI: its CLR path is plain Python over ATen
 * (utils/Utils.py:86-311, train_process/Trainer_prototype_full.py:328-449).  This header is the
 * boundary a maintainer binds instead (ctypes stub in INTEGRATION.md); each entry point cites the
 * reference lines it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - all tensors are fp32, contiguous, NCHW; "HW" is H*W, the contiguous pixel axis;
 *   - prototype rows ("R = 2K rows") are ordered obj_0..obj_{K-1}, bck_0..bck_{K-1}, i.e. the
 *     reference's return order (c0_obj, c1_obj, c0_bck, c1_bck) for K = 2 (utils/Utils.py:131);
 *   - "packed sums" are [R][C+1] floats: columns 0..C-1 hold S_r[c] = sum_{b,p} x[b,c,p] w_r[b,p],
 *     column C holds N_r = sum_{b,p} w_r[b,p].  This is the buffer that is all-reduced across GPUs;
 *   - stream is a cudaStream_t passed as void*; calls only enqueue work (no sync, no allocation) and are
 *     re-entrant from any thread (autograd runs backward on its own thread).  Global state is limited to
 *     the benchmark knobs of clr_set_tunable, the launch counter and the profiling trace pointer
 *     (atomics; change knobs only while no other thread is launching);
 *   - the fused step (clr_step_*) uses device-side waits instead of kernel boundaries in three places
 *     (flag dependency of the discriminative kernel, gated source-gradient CTAs, in-kernel exchange).  Every
 *     such wait has a ~2 s time-out that sets losses[7] = 1 and turns the step's losses / gradients into
 *     NaN (the EMA state is left untouched) -- callers poll losses[7] or rely on their NaN check.  The gated
 *     CTAs assume that the CTAs of one grid are dispatched in block-index order (true on every CUDA GPU to
 *     date, not a documented guarantee); clr_set_tunable("bwd_merge_off", 1) selects the gate-free form;
 *   - return value: 0 on success, a negative clr_status otherwise; nothing throws across the ABI.
 *     CUDA launch failures return CLR_ERR_CUDA_BASE - (int)cudaError_t.
 */
#ifndef CLR_B200_H
#define CLR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CLR_B200_VERSION 200 /* major*100 + minor; 2.0: clr_retrify_weights takes the MC logits (exact masks), new entry points */
#define CLR_MAX_K 8          /* classes per call */
#define CLR_MAX_WORLD 8      /* ranks of one NVLink domain that can share the fused step's in-kernel exchange */

typedef void* clr_stream_t;

enum clr_status {
    CLR_OK = 0,
    CLR_ERR_BAD_ARG = -1,     /* null pointer, non-positive size, K out of range */
    CLR_ERR_ALIGN = -2,       /* pointer not 4-byte aligned (16-byte alignment is detected, not required) */
    CLR_ERR_WORKSPACE = -3,   /* workspace too small */
    CLR_ERR_UNSUPPORTED = -4, /* combination not built */
    CLR_ERR_CUDA_BASE = -1000 /* CLR_ERR_CUDA_BASE - cudaError_t */
};

enum clr_weight_fmt {
    /* w is [B,K,HW]: w_obj,k = w_k, w_bck,k = 1 - w_k   (gen_prototype, utils/Utils.py:109-112) */
    CLR_W_COMPLEMENT = 0,
    /* w is [B,2K,HW]: rows obj_0.., bck_0..              (gen_prototype_retrify, utils/Utils.py:207-223) */
    CLR_W_EXPLICIT = 1
};

int clr_version(void);
const char* clr_status_string(int status);
/* Number of SMs / L2 bytes of the current device (grid sizing is derived from it; exposed for the bench). */
int clr_device_info(int* sm_count, int* l2_bytes);
/* Benchmark / debugging knobs (process-wide): "pool_impl" 0 auto / 1 LDG kernel / 2 TMA ring, "pool_stages" TMA ring
 * depth, "disc_impl" 0 fused / 1 two-pass, "mc_precise" 1 = clr_mc_stats / clr_mc_retrify evaluate the whole maps in
 * ATen's exact order (slow; parity tests).
 * Defaults select the fastest path. */
int clr_set_tunable(const char* name, int value);
/* Number of CUDA kernels this library has launched in this process so far (bench: "gpu_launches"). */
unsigned long long clr_launch_count(void);
/* Device-side kernel timeline (profiling aid; tools/timeline.py).  While enabled every kernel of the library stamps
 * %globaltimer into its slot: { earliest CTA start, earliest return from griddepcontrol.wait, latest CTA exit, CTAs }.
 * clr_trace_read synchronises the device, copies [clr_trace_slots()][4] uint64 to the host and resets the slots. */
int clr_trace_enable(int on);
int clr_trace_slots(void);
const char* clr_trace_name(int slot);
int clr_trace_read(unsigned long long* out_host);
/* cudaEvent_t helpers for clr_step_args.ev_* (the library records them around its own launches). */
int clr_event_create(void** ev);
int clr_event_destroy(void* ev);
int clr_event_elapsed_us(void* begin, void* end, float* us /*valid once both events completed*/);

/* ------------------------------------------------------------------------------------------------
 * Masked / confidence-weighted class-wise pooling  (replaces utils/Utils.py:114-126 -- the four
 * materialised products and eight reductions of gen_prototype -- and :212-223 of the retrify variant).
 * One read of feat.  Deterministic: per-CTA partials in `ws`, combined in fp64 in a fixed order.
 * ---------------------------------------------------------------------------------------------- */
size_t clr_pool_ws_bytes(int B, int C, int HW, int K);
int clr_pool_fwd(const float* feat /*[B,C,HW]*/, const float* w, int fmt, int B, int C, int HW, int K,
                 void* ws, size_t ws_bytes, float* sums /*[2K][C+1] out*/, clr_stream_t stream);
/* clr_pool_fwd + clr_proto_finalize in the same two launches (the reduce launch also writes mu [2K][C] = S_r / N_r): the
 * single-process form of gen_prototype's forward (utils/Utils.py:114-130).  Sharded callers all-reduce sums and call
 * clr_proto_finalize instead. */
int clr_pool_fwd_mu(const float* feat, const float* w, int fmt, int B, int C, int HW, int K,
                    void* ws, size_t ws_bytes, float* sums /*[2K][C+1] out*/, float* mu /*[2K][C] out*/, clr_stream_t stream);
/* Same kernel for R arbitrary explicit weight rows [B,R,HW] (1 <= R <= 16): sums is [R][C+1].
 * Used for the discriminative term's active-set sums and the per-sample (bmm-style) pooling. */
size_t clr_pool_rows_ws_bytes(int B, int C, int HW, int R);
int clr_pool_rows_fwd(const float* feat, const float* rows, int B, int C, int HW, int R,
                      void* ws, size_t ws_bytes, float* sums /*[R][C+1] out*/, clr_stream_t stream);
/* bmm-style per-sample pooling (Trainer_prototype.py:364-383, cal_prototype.py:156-175):
 *   proto[r][c] = mean_b( sum_p rows[b,r,p] x[b,c,p] / (sum_p rows[b,r,p] + 1) )
 * = clr_pool_rows_fwd_ps (per-sample packed sums [B][R][C+1]; ws >= R*(C+1)*4 + clr_pool_rows_ws_bytes)
 *   + clr_bmm_finalize (n_add = 1).  clr_pool_bwd_ps is the adjoint w.r.t. the features. */
int clr_pool_rows_fwd_ps(const float* feat, const float* rows, int B, int C, int HW, int R,
                         void* ws, size_t ws_bytes, float* sums_b /*[B][R][C+1] out*/, clr_stream_t stream);
int clr_bmm_finalize(const float* sums_b, int B, int R, int C, float n_add, float* out /*[R][C]*/, clr_stream_t stream);
int clr_pool_bwd_ps(const float* rows, int B, int C, int HW, int R, const float* g /*[R][C]*/, const float* sums_b,
                    float n_add, float scale, float* grad /*[B,C,HW] out*/, clr_stream_t stream);
/* Two domains (source, target) in ONE launch: the fused step's forward. sums0/sums1 as above. */
int clr_pool_fwd2(const float* feat0, const float* w0, int fmt0, int B0,
                  const float* feat1, const float* w1, int fmt1, int B1,
                  int C, int HW, int K, void* ws, size_t ws_bytes,
                  float* sums0, float* sums1, clr_stream_t stream);

/* mu_r = S_r / N_r, 0/0 -> NaN like the reference (utils/Utils.py:127-130). sums may be all-reduced. */
int clr_proto_finalize(const float* sums /*[R][C+1]*/, int R, int C, float* mu /*[R][C] out*/,
                       clr_stream_t stream);

/* Adjoint of pooling w.r.t. feat: grad[b,c,p] = scale * sum_r (g[r][c]/N_r) w_r[b,p]
 * (+ sum_k xtab[k][c] * xcoef[b,k,p] when xcoef != NULL: the discriminative term's direct gradient).
 * One write of grad, no read of feat.  Replaces autograd's Div/Sum/Mul backward chain of
 * utils/Utils.py:114-130. */
int clr_pool_bwd(const float* w, int fmt, int B, int C, int HW, int K,
                 const float* g /*[2K][C] dL/dmu*/, const float* sums /*[2K][C+1]*/, float scale,
                 const float* xcoef /*[B,Kx,HW] or NULL*/, const float* xtab /*[Kx][C] or NULL*/, int Kx,
                 float* grad /*[B,C,HW] out*/, clr_stream_t stream);
/* Up to two domains in ONE launch (the fused step's backward).  `scale_dev` (nullable) is a device scalar
 * multiplied into the result -- the upstream gradient of the step total, so no host sync is needed. */
typedef struct clr_bwd_dom {
    const float* w;       /* weight planes of this domain */
    const float* g;       /* [2K][C] dL/dmu */
    const float* sums;    /* [2K][C+1] (global) packed sums */
    const float* xcoef;   /* [B,Kx,HW] or NULL */
    const float* xtab;    /* [Kx][C] or NULL */
    float* grad;          /* [B,C,HW] out */
    const float* scale_dev;
    float scale;
    int fmt, B, Kx;
} clr_bwd_dom;
int clr_pool_bwd_multi(const clr_bwd_dom* doms, int ndom, int C, int HW, int K, clr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Per-pixel channel contractions: dots[b,q,p] = sum_c V[q][c] * feat[b,c,p]  (+ sumsq[b,p] = sum_c x^2).
 * One read of feat.  Building block of: the adjoint w.r.t. soft predictions, the pixel<->prototype
 * distance / cosine weight (Trainer_prototype.py:98-116, utils/Utils.py:86-88) and the discriminative
 * hinge (Trainer_prototype_mt bytecode L454-474).
 * ---------------------------------------------------------------------------------------------- */
int clr_pixel_dots(const float* feat, int B, int C, int HW, const float* V /*[Q][C]*/, int Q,
                   float* dots /*[B,Q,HW] out*/, float* sumsq /*[B,HW] out or NULL*/, clr_stream_t stream);

/* dL/dpred for soft predictions: complement fmt -> [B,K,HW] (= d w_obj - d w_bck); explicit -> [B,2K,HW]. */
size_t clr_pool_bwd_w_ws_bytes(int C, int K, int fmt);
int clr_pool_bwd_w(const float* feat, int fmt, int B, int C, int HW, int K,
                   const float* g /*[2K][C]*/, const float* sums /*[2K][C+1]*/, float scale,
                   void* ws, size_t ws_bytes, float* grad_w, clr_stream_t stream);

/* Pixel <-> prototype maps (Trainer_prototype.py:98-116, utils/Utils.py:86-88). */
int clr_proto_distance(const float* feat, int B, int C, int HW, const float* protos /*[Q][C]*/, int Q,
                       float* dist /*[B,Q,HW] out: ||proto_q - feat[b,:,p]||_2*/, clr_stream_t stream);
int clr_proto_cosine(const float* feat, int B, int C, int HW, const float* proto /*[C]*/, float* ws4 /*1 float*/,
                     float* out /*[B,1,HW]*/, clr_stream_t stream);
/* (x - min x) / (max x - min x) in place over n floats; ws >= 512 floats. */
int clr_minmax_normalize(float* x, size_t n, float* ws, clr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Prototype-guided discriminative hinge, forward pass over the source features
 * (Trainer_prototype_mt.cpython-38.pyc L454-474).  delta_k(p) = d_obj,k - d_bck,k is affine in x, so the
 * pass is K dot products per pixel with disc_vec[k] = P_obj,k - P_bck,k (from clr_align_finalize):
 *   coef[b,k,p]  = y_k [delta_k + m > 0] - (1 - y_k) [m - delta_k > 0]       (d(npx*loss)/d delta)
 *   partials[i]  = per-CTA { sum of hinge terms, sum coef_0, .. }             (row stride 1+K)
 * The active-set sums A_k[c] = sum coef*x follow from clr_pool_rows_fwd(xs, coef, R=K).
 * ---------------------------------------------------------------------------------------------- */
int clr_disc_partials_cap(void);
int clr_disc_fwd(const float* xs, const float* ys, int B, int C, int HW, int K,
                 const float* disc_vec /*[K][C]*/, const float* disc_beta /*[K]*/, float margin,
                 float* coef /*[B,K,HW] out*/, float* delta /*[B,K,HW] out or NULL*/,
                 float* partials /*[cap][1+K]*/, int partials_cap, int* nparts /*host out*/, clr_stream_t stream);

/* One-read form of the same pass: dot products AND the active-set sums A_k[c] = sum_p coef_k(p) x[c,p] from a
 * shared-memory tile (TMA bulk ring).  packed2 = [K][C+1] (col C = sum of coefficients) followed by
 * { hinge numerator, 0, 0, 0 }.  Returns CLR_ERR_UNSUPPORTED for ragged / very wide inputs: use the two-pass
 * form above then. */
size_t clr_disc_fused_ws_bytes(int C, int K);
int clr_disc_fused_fwd(const float* xs, const float* ys, int B, int C, int HW, int K,
                       const float* disc_vec, const float* disc_beta, float margin,
                       float* coef /*[B,K,HW] out*/, float* delta /*or NULL*/, void* ws, size_t ws_bytes,
                       float* packed2 /*[K][C+1] + 4 out*/, clr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * MC-dropout statistics + retrify weights (utils/Utils.py:161-223).
 * ---------------------------------------------------------------------------------------------- */
int clr_mc_stats(const float* preds /*[T*B,K,Hi,Wi] logits*/, int T, int B, int K, int Hi, int Wi,
                 float* std_map /*[B,K,Hi,Wi] out*/, float* pred_mean /*[B,K,Hi,Wi] out*/, clr_stream_t stream);
/* clr_mc_stats streams preds once with approximate sigmoids (std error <= ~3e-7).  The uncertainty mask
 * `std_small < std_thr` is an INTEGER output (bit-exact against eager torch on the same device): pass the MC logits
 * again as `preds` (+ T) and every pixel whose down-sampled std lies within 1e-5 of the threshold is re-evaluated from
 * them in ATen's exact order (sigmoid, two-accumulator Welford of torch.std's CUDA reduction, bilinear taps);
 * preds = NULL skips that guard band (knife-edge pixels may then differ from torch). */
int clr_retrify_weights(const float* oT_before /*[B,K,H,W]*/, const float* pred_mean, const float* std_map,
                        const float* preds /*[T*B,K,Hi,Wi] or NULL*/, int T,
                        int B, int K, int H, int W, int Hi, int Wi, float pseudo_thr /*0.75*/, float std_thr /*0.04*/,
                        float* weights /*[B,2K,H,W] out*/, float* masks /*[B,K,H,W] out, {0,2}*/,
                        float* pseudo_out /*[B,K,H,W] or NULL*/, float* small_out /*[2][B,K,H,W] or NULL*/,
                        clr_stream_t stream);

/* Both of the above in ONE pass over preds (utils/Utils.py:161-223 end to end): each CTA owns the image rows between
 * two consecutive bilinear source rows, so the taps of its feature row never leave shared memory.  pred_mean may be
 * NULL (the full-resolution mean is only ever down-sampled, :170).  Returns CLR_ERR_UNSUPPORTED when the geometry does
 * not allow it ((Hi-1) < 2 (H-1), Wi % 4 != 0, misaligned maps): call the two functions above then. */
int clr_mc_retrify(const float* preds, const float* oT_before, int T, int B, int K, int H, int W, int Hi, int Wi,
                   float pseudo_thr, float std_thr, float* std_map, float* pred_mean /*nullable*/,
                   float* weights /*[B,2K,H,W] out*/, float* masks /*[B,K,H,W] out*/, clr_stream_t stream);

/* MC statistics without the staging buffer of the trainer's loop (Trainer_prototype_full.py:359-368):
 * hand every MC forward's logits ([passes*B,K,Hi,Wi], pass-major like preds) to clr_mc_accumulate (first = 1 on the first
 * call of a step), then clr_mc_finalize writes std_map / pred_mean for clr_retrify_weights (preds = NULL: the raw logits
 * are gone, so the knife-edge guard of the mask is not available on this path).  state: clr_mc_state_floats() floats. */
size_t clr_mc_state_floats(int B, int K, int Hi, int Wi);
int clr_mc_accumulate(const float* logits, int passes, int B, int K, int Hi, int Wi, int first, float* state,
                      clr_stream_t stream);
int clr_mc_finalize(const float* state, int T, int B, int K, int Hi, int Wi, float* std_map, float* pred_mean,
                    clr_stream_t stream);

/* F.interpolate(target_map, size=(H, W), mode='nearest') of the hard labels (Trainer_prototype_full.py:329-330):
 * src [planes,Hi,Wi] -> dst [planes,H,W], ATen's source index min(floor(dst * (float)in / out), in - 1). */
int clr_label_downsample(const float* src, int planes, int Hi, int Wi, int H, int W, float* dst, clr_stream_t stream);

/* Stored-prototype EMA of Trainer.update_objective_SingleVector (Trainer_prototype.py:117-123), R vectors per launch:
 * out[r] = stored[r] * (1 - rate) + rate * v[r] unless sum_c v[r][c] == 0 (then out[r] = stored[r]); the zero test runs
 * on the device (the reference pays one .item() sync per vector).  out may alias stored. */
int clr_ema_rows(const float* v /*[R][C]*/, const float* stored /*[R][C]*/, int R, int C, float rate, float* out,
                 clr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Augmented-consistency masked BCE (Trainer_prototype_mt.cpython-38.pyc L502-561).
 * stats = { sum(m*l), sum(m), loss, 0 }.
 * ---------------------------------------------------------------------------------------------- */
size_t clr_cons_ws_bytes(void);
int clr_cons_fwd(const float* oT, const float* oT_aug, const float* masks, int B, int K, int Hi, int Wi, int H, int W,
                 float threshold, float aug_weight, void* ws, size_t ws_bytes, float* stats /*[4] out*/,
                 clr_stream_t stream);
int clr_cons_bwd(const float* oT, const float* oT_aug, const float* masks, int B, int K, int Hi, int Wi, int H, int W,
                 float threshold, float aug_weight, const float* stats, const float* gscale_dev, float gscale,
                 float* grad_oT_aug /*[B,K,Hi,Wi] out*/, clr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Image-resolution elementwise glue around the CLR block.
 *   seg loss : BCELoss(sigmoid(oS), target_map) + MSELoss(sigmoid(boundaryS), target_boundary), both means
 *              (Trainer_prototype_full.py:292-294).  out = { bce, mse, bce + mse, 0 }.  n2 may be 0 (BCE only).
 *              bwd writes d/d oS and d/d boundaryS of (gscale * *gup_dev) * (bce + mse); gup_dev may be NULL (= 1).
 *   entropy  : uncertainty_map = -sigmoid(o) * log(sigmoid(o) + smooth)  (:452, :481, :500) and its adjoint.
 * ---------------------------------------------------------------------------------------------- */
size_t clr_seg_loss_ws_bytes(void);
int clr_seg_loss_fwd(const float* oS, const float* target_map, size_t n1, const float* boundaryS,
                     const float* target_boundary, size_t n2, void* ws, size_t ws_bytes, float* out /*[4]*/,
                     clr_stream_t stream);
int clr_seg_loss_bwd(const float* oS, const float* target_map, size_t n1, const float* boundaryS,
                     const float* target_boundary, size_t n2, const float* gup_dev, float gscale,
                     float* g_oS, float* g_boundaryS, clr_stream_t stream);
int clr_entropy_fwd(const float* o, size_t n, float smooth, float* out, clr_stream_t stream);
int clr_entropy_bwd(const float* o, const float* gout, size_t n, float smooth, float* gin, clr_stream_t stream);
/* Validation counts (utils/metrics.py:118-168): counts[k][2*gt + pred] with pred = sigmoid(logit) > thr (exact fp32
 * decision), gt = target != 0; [K][4] unsigned 64-bit, zeroed by the call.  Dice / pixel accuracy / IoU follow. */
int clr_seg_counts(const float* logits /*[B,K,HW]*/, const float* target /*[B,K,HW]*/, int B, int K, size_t HW, float thr,
                   unsigned long long* counts /*[K][4] out*/, clr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * O(K*C) glue (Trainer_prototype_full.py:335-355, 378-398, 428-449): prototypes, EMA, alignment /
 * separation losses and their gradients; then the discriminative term's prototype gradients + totals.
 * losses = { intra, inter, disc, aug, total, 0, 0, 0 }.
 * ---------------------------------------------------------------------------------------------- */
int clr_align_finalize(const float* sums_s, const float* sums_t, int K, int C,
                       float* stored_s /*[2K][C] in/out*/, float* stored_t, int first_s, int first_t, double decay,
                       float w_intra, float w_inter, float* P_s /*[2K][C] out*/, float* P_t,
                       float* cur_s /*[2K][C] out or NULL*/, float* cur_t,
                       float* g_s /*[2K][C] out: dL/d cur_s*/, float* g_t,
                       float* disc_vec /*[K][C] out or NULL*/, float* disc_beta /*[K] out or NULL*/,
                       float* losses /*[8]*/, clr_stream_t stream);
int clr_disc_finalize(float* packed2 /*[K][C+1] | hinge num | cons num | cons den | 0*/, const float* P_s,
                      int K, int C, double npx, float w_disc, float ema_factor, float gscale,
                      float* g_s /*in/out*/, float* xtab /*[K][C] out*/,
                      float w_intra, float w_inter, float w_aug, float aug_weight, int use_disc, int use_cons,
                      float* losses, clr_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * The fused CLR step: A1 (hard source) + A1/A2 (target) + A4 + A5 [+ A9] [+ A10], forward and backward,
 * as three forward phases with the two possible cross-GPU exchange points between them, and one
 * backward.  Single GPU: call a, b, c back to back (clr_step_fwd does).  Everything is enqueued on
 * `stream`; no host synchronisation anywhere.
 *
 *   phase a: [mc_stats + retrify_weights] + pooling of both domains -> packed1 = [sums_s | sums_t]
 *            --- all-reduce(packed1) when sharded ---
 *   phase b: align_finalize; [consistency fwd]; [discriminative dots + active-set pooling] -> packed2
 *            --- all-reduce(packed2) when sharded ---
 *   phase c: disc_finalize (+ totals)
 *   bwd    : gradient write for both feature maps in one launch [+ consistency backward]
 * ---------------------------------------------------------------------------------------------- */
typedef struct clr_step_args {
    /* geometry */
    int B_s, B_t, C, H, W, K;
    int Hi, Wi, T;                 /* image-resolution geometry of preds / oT (0 when unused) */
    /* switches */
    int use_retrify, use_disc, use_cons;
    int wt_fmt;                    /* format of wt when !use_retrify (CLR_W_COMPLEMENT: sigmoid(oT_before) planes) */
    int first_s, first_t;          /* 1 on the first step (EMA copies), then 0 */
    /* hyper-parameters */
    double decay, npx_global;      /* npx_global = global B_s*H*W (discriminative mean) */
    float w_intra, w_inter, w_disc, w_aug, margin, aug_weight, cons_threshold, pseudo_thr, std_thr, grad_scale;
    /* inputs */
    const float* xs; const float* ys;     /* [B_s,C,H,W], [B_s,K,H,W] hard labels */
    const float* xt; const float* wt;     /* [B_t,C,H,W]; target weights (ignored when use_retrify) */
    const float* oT_before;               /* [B_t,K,H,W] (retrify) */
    const float* preds;                   /* [T*B_t,K,Hi,Wi] (retrify) */
    const float* oT; const float* oT_aug; /* [B_t,K,Hi,Wi] (consistency) */
    const float* gup;                     /* device scalar dL/dtotal for the backward, or NULL (= 1) */
    /* state, in/out */
    float* stored_s; float* stored_t;     /* [2K][C] */
    /* outputs */
    float* packed1;                       /* [2][2K][C+1] */
    float* packed2;                       /* [K][C+1] + 4 */
    float* P_s; float* P_t;               /* [2K][C] EMA'd prototypes */
    float* g_s; float* g_t;               /* [2K][C] dL/d(current prototypes) */
    float* losses;                        /* [8] */
    float* std_map; float* pred_mean;     /* [B_t,K,Hi,Wi] (retrify); pred_mean is scratch of the step: only the bilinear source
                                           * rows the down-sample reads are written (power-of-two image sizes), the rest is untouched */
    float* wt_retrify; float* masks;      /* [B_t,2K,H,W], [B_t,K,H,W] (retrify) */
    float* disc_coef;                     /* [B_s,K,H,W] (disc) */
    float* disc_vec; float* disc_beta; float* xtab; /* [K][C], [K], [K][C] (disc) */
    float* gxs; float* gxt;               /* [B,C,H,W] backward outputs */
    float* g_oT_aug;                      /* [B_t,K,Hi,Wi] backward output (consistency, when w_aug != 0) */
    /* workspace */
    void* ws; size_t ws_bytes;
    /* optional cudaEvent_t handles recorded on `stream` around the dominant kernel (the two-domain
     * pooling launch) so a harness can time it inside a live step; NULL = not recorded */
    void* ev_pool_begin; void* ev_pool_end;
    void* ev_bwd_begin; void* ev_bwd_end;
    void* reserved_ptr[3];                /* (a second-stream schedule lived here; measured slower, removed) */
    /* Sharded step with the exchange INSIDE the kernels (clr_step_run only; world <= 1: unused).  peer_rx[q] is rank q's
     * receive buffer (clr_step_xchg_bytes bytes, zeroed once, see clr_peer_*) mapped into THIS process; peer_rx[rank] is
     * the local one.  `seq` must be the same on every rank and change by +1 per clr_step_run call (start at 1).  The
     * finish stages push their packed sums to every peer as (value, seq) 64-bit words over NVLink and sum the G
     * contributions in rank order from their own buffer -- one NVLink traversal, no fences, no extra launches, every rank
     * obtains bit-identical sums.  losses[7] is set to 1 if a peer did not answer within ~2 s. */
    int world, rank;
    unsigned int seq; int reserved0;
    void* peer_rx[CLR_MAX_WORLD];
} clr_step_args;

/* ------------------------------------------------------------------------------------------------
 * TransNorm: the domain-split batch normalisation with the adaptive channel weight that the
 * reference's DeepLab uses with --use_TN.  Replaces `_BatchNorm.forward` of networks/sync_batchnorm/batchnorm.py
 * (:439-493 training, :494-521 eval; class BatchNorm2d :523, selected at networks/deeplabv3.py:17-23).
 * x, y, gy, gx: [B, C, HW] fp32 contiguous (HW = 1 for 2-D inputs); source = samples [0, B/2), target = the rest.
 * weight / bias [C] may be NULL (affine = False).  save [5][C] = { mean_s, mean_t, rstd_s, rstd_t, alpha } is
 * written by the forward and read by the backward.  ws: clr_tn_ws_bytes(C) bytes.
 *   clr_tn_fwd : y = ((x - mean_d) * rstd_d * weight + bias) * (1 + alpha); running estimates (may be NULL) are
 *                updated in place: r = (1 - momentum) r + momentum stat (unbiased variance), like F.batch_norm.
 *                3 launches: one read of x (statistics), O(C), one read + one write.
 *   clr_tn_bwd : gx, gweight [C], gbias [C] (either may be NULL); alpha is a constant (detached, :493).  eval_mode = 1:
 *                adjoint of clr_tn_eval (the statistics are constants too).
 *   clr_tn_eval: statistics = the running estimates; every sample is normalised with the TARGET ones (:497-509),
 *                alpha from both (:510-514).  B >= 1.
 * Returns CLR_ERR_UNSUPPORTED for C > 8192 or B > 65535. */
size_t clr_tn_ws_bytes(int C);
int clr_tn_fwd(const float* x, int B, int C, int HW, const float* weight, const float* bias,
               float* running_mean_s, float* running_var_s, float* running_mean_t, float* running_var_t,
               float momentum, float eps, void* ws, size_t ws_bytes, float* y, float* save, clr_stream_t stream);
int clr_tn_eval(const float* x, int B, int C, int HW, const float* weight, const float* bias,
                const float* running_mean_s, const float* running_var_s, const float* running_mean_t,
                const float* running_var_t, float eps, void* ws, size_t ws_bytes, float* y, float* save,
                clr_stream_t stream);
int clr_tn_bwd(const float* x, const float* gy, int B, int C, int HW, const float* weight, const float* save,
               int eval_mode, void* ws, size_t ws_bytes, float* gx, float* gweight, float* gbias, clr_stream_t stream);

/* Peer-visible device memory for the in-kernel exchange: plain cudaMalloc (zero-filled) + CUDA IPC handles, so that a
 * host in any language can wire the ranks of one node together (exchange the 64-byte handles over its own channel). */
int clr_peer_alloc(size_t bytes, void** ptr);
int clr_peer_free(void* ptr);
int clr_peer_export(void* ptr, unsigned char handle[64]);
int clr_peer_open(const unsigned char handle[64], void** ptr);
int clr_peer_close(void* ptr);
size_t clr_step_xchg_bytes(int world, int K, int C);

size_t clr_step_ws_bytes(const clr_step_args* a);
/* Which launch schedule clr_step_fwd / clr_step_run use for these arguments: 2 = source pooled first, finish halves hidden
 * behind the MC statistics / the discriminative pass (retrify + discriminative term), 1 = both maps pooled in one launch.
 * The optional ev_pool_* events bracket the two-domain pooling launch (1) or the source pooling launch (2). */
int clr_step_schedule(const clr_step_args* a);
int clr_step_fwd_a(const clr_step_args* a, clr_stream_t stream);
int clr_step_fwd_b(const clr_step_args* a, clr_stream_t stream);
int clr_step_fwd_c(const clr_step_args* a, clr_stream_t stream);
int clr_step_fwd(const clr_step_args* a, clr_stream_t stream);
int clr_step_bwd(const clr_step_args* a, clr_stream_t stream);
/* Forward AND backward in one call, for the common case that the step total enters the training loss with a known
 * (device-side, `gup`, default 1) coefficient: 6 launches; the two finish stages ride as the first CTAs of the
 * consistency pass / the target-gradient write.  Results are identical to clr_step_fwd + clr_step_bwd. */
int clr_step_run(const clr_step_args* a, clr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CLR_B200_H */

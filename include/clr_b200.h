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
 *   - stream is a cudaStream_t passed as void*; calls only enqueue work (no sync, no allocation,
 *     no global state, re-entrant from any thread -- autograd runs backward on its own thread);
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

#define CLR_B200_VERSION 100 /* major*100 + minor */
#define CLR_MAX_K 8          /* classes per call */

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

/* ------------------------------------------------------------------------------------------------
 * Masked / confidence-weighted class-wise pooling  (replaces utils/Utils.py:114-126 -- the four
 * materialised products and eight reductions of gen_prototype -- and :212-223 of the retrify variant).
 * One read of feat.  Deterministic: per-CTA partials in `ws`, combined in fp64 in a fixed order.
 * ---------------------------------------------------------------------------------------------- */
size_t clr_pool_ws_bytes(int B, int C, int HW, int K);
int clr_pool_fwd(const float* feat /*[B,C,HW]*/, const float* w, int fmt, int B, int C, int HW, int K,
                 void* ws, size_t ws_bytes, float* sums /*[2K][C+1] out*/, clr_stream_t stream);
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
int clr_pool_bwd2(const float* w0, int fmt0, int B0, const float* g0, const float* sums0, float scale0,
                  const float* xcoef0, const float* xtab0, int Kx0, float* grad0,
                  const float* w1, int fmt1, int B1, const float* g1, const float* sums1, float scale1,
                  float* grad1, int C, int HW, int K, clr_stream_t stream);

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

#ifdef __cplusplus
}
#endif
#endif /* CLR_B200_H */

// "Finish" stages of the single-GPU fused step: per-CTA partial reduce + the O(K*C) finalize arithmetic in ONE body
// each (no exchange point sits between reduce and finalize when nothing is sharded).  They are latency chains, not
// bandwidth, so the fused step never gives them a launch of their own: the bodies are co-scheduled as the first few
// CTAs of the streaming kernel that follows them in the step and does not depend on them --
//   pool_finish_body  rides with the consistency pass      (cons.cu:     pool_finish_cons_kernel)
//   disc_finish_body  rides with the target-gradient write (pool_bwd.cu: bwd_finish_kernel)
// -- which takes both chains (and two kernel boundaries) off the step's critical path.  finalize.cu also launches them
// stand-alone.  Both bodies are written for exactly kThreads (256) threads and any K <= CLR_MAX_K.
//
// Arithmetic = align_finalize_kernel / disc_finalize_kernel (finalize.cu), i.e. Trainer_prototype_full.py:335-355,
// 378-398, 428-449 and the Trainer_prototype_mt bytecode L454-474.
#pragma once
#include "clr_common.cuh"

namespace clr {

// Block-wide sums of NV doubles at once (one barrier): result valid in THREAD 0 only.
template <int NV>
__device__ __forceinline__ void block_sum_n(double (&v)[NV], double* sh /*[NV][32]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) sh[i * 32 + warp] = v[i];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = warp_sum(lane < nw ? sh[i * 32 + lane] : 0.0);
    }
}

// packed2 layout: [K][C+1] active-set sums (col C = n_k) | loss numerator | cons num | cons den | pad
struct PackSrc {   // per-CTA partials still to be summed (single-GPU path: no exchange between pack and finalize)
    const float* hinge; int n_hinge, hinge_stride;
    const double* cons; int n_cons;
};

// Sum of col[(sl + i*step) * stride] over the slots in [.., s_end): rounds of U predicated loads, all in flight at once
// (a plain remainder loop would serialise one L2 round trip per slot); fp64 accumulation in slot order.
template <int U>
__device__ __forceinline__ double strided_slot_sum(const float* __restrict__ col, int sl, int s_end, size_t stride, int step) {
    double s = 0.0;
    for (; sl < s_end; sl += U * step) {
        float v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = (sl + u * step < s_end) ? __ldcg(col + (size_t)(sl + u * step) * stride) : 0.f;
#pragma unroll
        for (int u = 0; u < U; ++u) s += (double)v[u];
    }
    return s;
}

// Two strided sums whose loads are issued TOGETHER (one L2 round trip instead of two dependent ones): a over
// col_a[(sa + i*step_a) * stride], b over col_b[(sb + i*step_b) * stride], slots below s_end; fp64, slot order.
template <int UA, int UB>
__device__ __forceinline__ void strided_slot_sum2(const float* __restrict__ col_a, int sa, int step_a, bool use_a,
                                                  const float* __restrict__ col_b, int sb, int step_b,
                                                  int s_end, size_t stride, double& out_a, double& out_b) {
    double a = 0.0, b = 0.0;
    while ((use_a && sa < s_end) || sb < s_end) {
        float va[UA], vb[UB];
#pragma unroll
        for (int u = 0; u < UA; ++u) va[u] = (use_a && sa + u * step_a < s_end) ? __ldcg(col_a + (size_t)(sa + u * step_a) * stride) : 0.f;
#pragma unroll
        for (int u = 0; u < UB; ++u) vb[u] = (sb + u * step_b < s_end) ? __ldcg(col_b + (size_t)(sb + u * step_b) * stride) : 0.f;
#pragma unroll
        for (int u = 0; u < UA; ++u) a += (double)va[u];
#pragma unroll
        for (int u = 0; u < UB; ++u) b += (double)vb[u];
        sa += UA * step_a;
        sb += UB * step_b;
    }
    out_a = a;
    out_b = b;
}

// ---------------------------------------------------------------------------------------------------------------
// In-kernel exchange of the packed sums between the ranks of one NVLink domain ("LL" style: value and sequence number
// travel in ONE 64-bit store, so a reader that sees the expected sequence number has the value -- no fences, no
// flags, one NVLink traversal).  Every rank owns a receive buffer [2 parities][world sources][n words]; a sender
// writes its word into slot [parity][its rank][idx] of EVERY rank's buffer (its own included), a receiver sums the
// `world` words of one index in rank order, so all ranks compute bit-identical totals.  Two parities: a rank can run
// at most one exchange ahead of the slowest peer (it needs that peer's next contribution to go further).
// ---------------------------------------------------------------------------------------------------------------
struct PeerXchg {
    unsigned long long* rx[CLR_MAX_WORLD];   // rank q's receive area for THIS exchange (peer-mapped); rx[rank] is local
    int world, rank;
    unsigned int seq;
    int n;                                   // words per source
    int pull;                                // 0: senders store into every peer's buffer, readers poll locally (push)
                                             // 1: senders store locally, readers poll the peers' buffers over NVLink (pull)
    float* err;                              // set to 1 on timeout (losses[7])
};
__device__ __forceinline__ void xchg_push(const PeerXchg& x, int idx, float v) {
    const unsigned long long w = ((unsigned long long)x.seq << 32) | (unsigned long long)__float_as_uint(v);
    const size_t off = ((size_t)(x.seq & 1u) * x.world + x.rank) * x.n + idx;
    if (x.pull) { *reinterpret_cast<volatile unsigned long long*>(x.rx[x.rank] + off) = w; return; }
    for (int q = 0; q < x.world; ++q) *reinterpret_cast<volatile unsigned long long*>(x.rx[q] + off) = w;
}
// A peer that does not answer within ~2 s: losses[7] is set, the caller's `*timed_out` (CTA-shared) is raised and the
// sum is NaN, so prototypes, losses and gradients of the step are poisoned LOUDLY (the trainer's NaN check fires,
// Trainer_prototype_full.py:298-299) instead of continuing with stale words; the finish bodies skip the EMA write-back.
// NI indices at once: all NI * world words are requested together (one round trip instead of NI * world dependent ones --
// measured at 8 ranks, polling the words one after the other cost 5.6 us per exchange on the step's critical path), then
// only the late ones are re-polled.  Sums in rank order, so every rank computes the same bits.
template <int NI>
__device__ __forceinline__ void xchg_pull_sums(const PeerXchg& x, const int (&idx)[NI], const bool (&use)[NI], float (&out)[NI],
                                               int* timed_out = nullptr) {
    unsigned long long w[NI][CLR_MAX_WORLD];
    const size_t base = (size_t)(x.seq & 1u) * x.world;
    auto addr = [&](int i, int q) {   // slot [parity][source q] of rank q's own buffer (pull) or of the local buffer (push)
        return reinterpret_cast<const volatile unsigned long long*>(x.rx[x.pull ? q : x.rank] + (base + q) * x.n + idx[i]);
    };
#pragma unroll
    for (int i = 0; i < NI; ++i)
#pragma unroll
        for (int q = 0; q < CLR_MAX_WORLD; ++q) {
            w[i][q] = 0ull;
            if (use[i] && q < x.world) w[i][q] = *addr(i, q);
        }
    auto late = [&](int i, int q) { return use[i] && q < x.world && (unsigned int)(w[i][q] >> 32) != x.seq; };
    bool all = true;
#pragma unroll
    for (int i = 0; i < NI; ++i)
#pragma unroll
        for (int q = 0; q < CLR_MAX_WORLD; ++q) all = all && !late(i, q);
    bool bad = false;
    if (!all) {
        const long long t0 = clock64();
        do {
            all = true;
#pragma unroll
            for (int i = 0; i < NI; ++i)
#pragma unroll
                for (int q = 0; q < CLR_MAX_WORLD; ++q)
                    if (late(i, q)) { w[i][q] = *addr(i, q); all = all && !late(i, q); }
            if (!all && clock64() - t0 > 4000000000LL) { if (x.err) *x.err = 1.f; bad = true; break; }     // ~2 s: a peer is gone
        } while (!all);
    }
    if (bad && timed_out) *timed_out = 1;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < CLR_MAX_WORLD; ++q)
            if (q < x.world) s += (double)__uint_as_float((unsigned int)(w[i][q] & 0xffffffffull));
        out[i] = bad ? __int_as_float(0x7fc00000) : (float)s;
    }
}
__device__ __forceinline__ float xchg_pull_sum(const PeerXchg& x, int idx, int* timed_out = nullptr) {
    const int i1[1] = {idx};
    const bool u1[1] = {true};
    float o1[1];
    xchg_pull_sums<1>(x, i1, u1, o1, timed_out);
    return o1[0];
}

// ---------------------------------------------------------------------------------------------------------------
// pool finish: CTA = 8 channels.  Warp w takes the (domain d, row r) pairs w, w+8, ..: it reduces the column block
// [c0, c0+8) of partial[d][slot][r][.] over the slots with 4 slot-lanes per channel (fp64, fixed order) plus the pair's
// weight-sum column; then K*8 threads do the align_finalize arithmetic for the CTA's channels; the loss terms are
// combined across CTAs by the last CTA to finish (per-CTA fp64 partials summed in a fixed order -> deterministic).
// `counter` is zeroed by the pooling kernel of the same step (stream order), so no initialisation contract leaks out.
// ---------------------------------------------------------------------------------------------------------------
struct PoolFinishParams {
    const float* partial[2];   // [slots][R][C+1]   (0 = source, 1 = target)
    float* sums[2];            // packed sums out
    float* stored[2];
    float* P[2];
    float* g[2];
    int slots[2];
    int first[2];
    int K, C;
    float d, omd, w_intra, w_inter;
    float* disc_vec;
    float* disc_beta;
    float* losses;
    double* loss_partial;      // [ctas][2 + CLR_MAX_K]
    unsigned int* counter;
    PeerXchg x;                // world > 1: sum the packed sums over the ranks before the finalize arithmetic
    unsigned int* done_fin;    // optional completion counters (cta_signal): finish CTAs only / every CTA of the launch
    unsigned int* done_all;
    int early_signal;          // 1: done_fin is bumped before the last-CTA combine (the consumer sums beta itself), 0: at the CTA's end
    int write_total;           // alignment-only step (no discriminative / consistency term): the last CTA also writes the totals
    // Split form (fused step, schedule 2): the SOURCE half runs right after the source pooling (mode 1: sums, EMA'd source
    // prototypes, separation loss, the discriminative term's vectors) hidden behind the MC statistics; the TARGET half
    // (mode 2: target sums / prototypes, alignment loss, all gradients) runs after the target pooling, reads the source
    // prototypes the first half left in P[0] and waits (already satisfied in practice) on that half's completion counter.
    int mode;                  // 0 = both domains in one body, 1 = source half, 2 = target half
    const unsigned int* wait_src; unsigned int wait_src_n; float* wait_err;     // mode 2
};
static inline int pool_finish_ctas(int C) { return (C + 7) / 8; }

__device__ __forceinline__ void pool_finish_body(const PoolFinishParams& p, const int cta, const int ncta) {
    constexpr int NL = 2 + CLR_MAX_K;
    const int K = p.K, R = 2 * K, C = p.C, n = R * (C + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ch = lane >> 2, sl0 = lane & 3;
    const int c = cta * 8 + ch;
    const bool do_s = p.mode != 2, do_t = p.mode != 1;            // which domains THIS body reduces / finalises
    const int d_lo = do_s ? 0 : 1, d_hi = do_t ? 2 : 1;
    __shared__ float S[2][2 * CLR_MAX_K][8];
    __shared__ float Nn[2][2 * CLR_MAX_K];
    __shared__ double lp[8 * CLR_MAX_K][NL];
    __shared__ bool is_last;
    __shared__ int xchg_bad;       // a peer timed out (world > 1): skip the EMA write-back, everything downstream is NaN
    if (tid == 0) xchg_bad = 0;
    // the stored (EMA) prototypes of this thread's (class, channel) entries: issued before the slot sums, which they do not
    // depend on, so the chain below has one L2 round trip less
    float st_pre[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
    if (tid < K * 8 && cta * 8 + (tid & 7) < C) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const size_t e = (size_t)((tid >> 3) + h * K) * C + cta * 8 + (tid & 7);
            if (do_s && !p.first[0]) st_pre[0][h] = __ldcg(p.stored[0] + e);
            if (do_t && !p.first[1]) st_pre[1][h] = __ldcg(p.stored[1] + e);
        }
    }
    for (int pair = d_lo * R + warp; pair < d_hi * R; pair += kWarps) {
        const int d = pair / R, r = pair - d * R;
        const float* part = p.partial[d];
        const int slots = p.slots[d];
        double s, nn;              // 64 slots: ONE round of loads for the channel column and the weight-sum column
        strided_slot_sum2<16, 2>(part + (size_t)r * (C + 1) + (c < C ? c : 0), sl0, 4, c < C,
                                 part + (size_t)r * (C + 1) + C, lane, 32, slots, (size_t)n, s, nn);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        nn = warp_sum(nn);
        if (p.x.world > 1) {          // this rank's contribution goes to every rank (word index = position in packed1)
            if (sl0 == 0 && c < C) xchg_push(p.x, (d * R + r) * (C + 1) + c, (float)s);
            if (lane == 0 && cta == 0) xchg_push(p.x, (d * R + r) * (C + 1) + C, (float)nn);
        } else {
            if (sl0 == 0 && c < C) { S[d][r][ch] = (float)s; p.sums[d][(size_t)r * (C + 1) + c] = (float)s; }
            if (lane == 0) { Nn[d][r] = (float)nn; if (cta == 0) p.sums[d][(size_t)r * (C + 1) + C] = (float)nn; }
        }
    }
    if (p.x.world > 1) {
        __syncthreads();           // xchg_bad = 0 is visible before anybody can raise it
        // global sums: items [0, nd*R*8) are (pair, channel) entries, the next nd*R the weight-sum columns (nd = domains of
        // this body); looped, because 2R*8 + 2R exceeds the 256 threads of the CTA for K = 8 (R = 16)
        const int npairs = (d_hi - d_lo) * R;
        for (int it = tid; it < npairs * 8 + npairs; it += kThreads) {
            if (it < npairs * 8) {
                const int pair = d_lo * R + (it >> 3), j = it & 7, d = pair / R, r = pair - d * R, cc = cta * 8 + j;
                if (cc < C) {
                    const float v = xchg_pull_sum(p.x, (d * R + r) * (C + 1) + cc, &xchg_bad);
                    S[d][r][j] = v;
                    p.sums[d][(size_t)r * (C + 1) + cc] = v;
                }
            } else {
                const int pair = d_lo * R + (it - npairs * 8), d = pair / R, r = pair - d * R;
                const float v = xchg_pull_sum(p.x, (d * R + r) * (C + 1) + C, &xchg_bad);
                Nn[d][r] = v;
                if (cta == 0) p.sums[d][(size_t)r * (C + 1) + C] = v;
            }
        }
    }
    if (p.mode == 2 && p.wait_src) {
        // the source half's prototypes (P[0]) are read below through coherent loads; its completion counter was bumped tens
        // of microseconds ago in practice (it ran behind the MC statistics) -- this wait is the formal ordering.  A miss
        // (~2 s) sets losses[7] and poisons this half like a lost peer does.
        if (tid == 0 && !spin_until_at_least(p.wait_src, p.wait_src_n)) { if (p.wait_err) *p.wait_err = 1.f; xchg_bad = 1; }
    }
    __syncthreads();
    const bool keep_state = xchg_bad == 0;
    if (tid < K * 8) {
        const int k = tid >> 3, j = tid & 7, cc = cta * 8 + j;
        double acc[NL];
#pragma unroll
        for (int i = 0; i < NL; ++i) acc[i] = 0.0;
        if (cc < C) {
            const float dd = p.d, omd = p.omd, invC = 1.0f / (float)C;
            const float ds = p.first[0] ? 1.f : dd, dt = p.first[1] ? 1.f : dd;
            float ps[2], pt[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int rr = k + h * K;
                const size_t e = (size_t)rr * C + cc;
                if (do_s) {
                    const float cs = S[0][rr][j] / Nn[0][rr];                     // utils/Utils.py:127-130
                    ps[h] = p.first[0] ? cs : __fadd_rn(__fmul_rn(omd, st_pre[0][h]), __fmul_rn(dd, cs));
                    p.P[0][e] = ps[h];
                    if (keep_state) p.stored[0][e] = ps[h];                       // .detach() copy (Trainer_prototype_full.py:341-344)
                } else {
                    ps[h] = keep_state ? __ldcg(p.P[0] + e) : __int_as_float(0x7fc00000);   // written by the source half
                }
                if (do_t) {
                    const float ct = S[1][rr][j] / Nn[1][rr];
                    pt[h] = p.first[1] ? ct : __fadd_rn(__fmul_rn(omd, st_pre[1][h]), __fmul_rn(dd, ct));
                    p.P[1][e] = pt[h];
                    if (keep_state) p.stored[1][e] = pt[h];                       // (:384-387)
                    const double df = (double)ps[h] - (double)pt[h];
                    acc[0] += df * df;
                }
            }
            const float dob = ps[0] - ps[1];
            if (do_s) acc[1] += (double)dob * dob;
            if (do_t) {
                const float gsep = ds * p.w_inter * 2.0f * dob * invC;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const size_t e = (size_t)(k + h * K) * C + cc;
                    const float gi = p.w_intra * 2.0f * (ps[h] - pt[h]) * invC;
                    p.g[0][e] = ds * gi + (h == 0 ? gsep : -gsep);
                    p.g[1][e] = -dt * gi;
                }
            }
            if (do_s) {
                if (p.disc_vec) p.disc_vec[(size_t)k * C + cc] = dob;
#pragma unroll
                for (int kk = 0; kk < CLR_MAX_K; ++kk)
                    if (kk == k) acc[2 + kk] = (double)ps[0] * ps[0] - (double)ps[1] * ps[1];
            }
        }
#pragma unroll
        for (int i = 0; i < NL; ++i) lp[tid][i] = acc[i];
    }
    __syncthreads();
    if (tid < NL) {
        double t = 0.0;
        for (int e = 0; e < K * 8; ++e) t += lp[e][tid];
        __stcg(p.loss_partial + (size_t)cta * NL + tid, t);
        __threadfence();
    }
    __syncthreads();
    // Flag dependency (clr_common.cuh): everything the discriminative kernel needs from THIS CTA is out now -- its slice of
    // disc_vec and its loss partials (the kernel sums the per-class column itself, in the order used below) -- so the
    // consumer is released here, before the last-CTA combine, which only serves the logged losses.  done_all (grid-completion
    // transitivity) is still bumped at the very end of the CTA by the kernel wrapper.
    if (p.done_fin && p.early_signal) cta_signal(p.done_fin, nullptr);
    if (tid == 0) {
        const unsigned prev = atomicAdd(p.counter, 1u);
        is_last = (prev == (unsigned)ncta - 1u);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        // warp i sums value i over the CTAs: lanes take CTAs lane, lane+32, .. (loads in flight together), fixed order.
        // value 0 = alignment (needs both domains: written by the body that has the target), 1 = separation and 2.. = the
        // discriminative offsets (source only)
        for (int i = warp; i < 2 + K; i += kWarps) {
            if ((i == 0 && !do_t) || (i >= 1 && !do_s)) continue;            // warp-uniform
            double t = 0.0;
            for (int b = lane; b < ncta; b += 32) t += __ldcg(p.loss_partial + (size_t)b * NL + i);
            t = warp_sum(t);
            if (lane == 0) {
                const float val = (float)(t * (1.0 / (double)C));
                if (i < 2) p.losses[i] = val;
                else if (p.disc_beta) p.disc_beta[i - 2] = val;
            }
        }
        if (p.write_total) {
            __syncthreads();       // (is_last is CTA-uniform) losses[0..1] written above by lanes 0 of warps 0 / 1
            if (tid == 0) {
                p.losses[2] = 0.f; p.losses[3] = 0.f;
                p.losses[4] = p.w_intra * p.losses[0] + p.w_inter * p.losses[1];
                p.losses[5] = 0.f; p.losses[6] = 0.f;
                if (p.x.world <= 1) p.losses[7] = 0.f;
            }
        }
        if (tid == 0) *p.counter = 0u;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// disc finish: CTA = 8 channels.  Warp w takes the (class k, slot quarter q) pairs w, w+8, ..: it reduces the per-CTA
// partials of the fused discriminative kernel ([slots][K][C+1]) for the CTA's channels; then K*8 threads apply the
// disc_finalize arithmetic.  One extra (last) CTA owns no channels: it folds the hinge / consistency per-CTA partials
// into the loss tail and writes the step totals.  No cross-CTA step.
// ---------------------------------------------------------------------------------------------------------------
struct DiscFinishParams {
    const float* partial; int slots;
    float* packed2; const float* P_s; float* g_s; float* xtab; float* losses;
    int K, C;
    double npx;
    float coef;       // 2 / (C * npx)
    float w_disc, ema_factor, gscale, w_intra, w_inter, w_aug, aug_weight;
    int use_cons;
    PackSrc ps;
    PeerXchg x;       // world > 1: sum the active-set sums and the loss numerators over the ranks
    unsigned int* gate_signal;   // stand-alone launch in front of a gated backward launch: bumped once per CTA when it is done
};
static inline int disc_finish_ctas(int C) { return (C + 7) / 8 + 1; }

__device__ __forceinline__ void disc_finish_body(const DiscFinishParams& p, const int cta, const int ncta) {
    const int K = p.K, C = p.C, n = K * (C + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ch = lane >> 2, sl0 = lane & 3;
    const int c = cta * 8 + ch;
    __shared__ double Sq[CLR_MAX_K][4][8];
    __shared__ double Nq[CLR_MAX_K][4];
    __shared__ double shp[3 * 32];
    const bool loss_cta = cta == ncta - 1;
    const int per = (p.slots + 3) / 4;
    if (!loss_cta) {
        for (int pair = warp; pair < 4 * K; pair += kWarps) {
            const int k = pair >> 2, q = pair & 3;
            const int s_begin = q * per, s_end = (s_begin + per) < p.slots ? (s_begin + per) : p.slots;
            double s, nn;          // 296 slots / 4 quarters: one round for both columns
            strided_slot_sum2<20, 3>(p.partial + (size_t)k * (C + 1) + (c < C ? c : 0), s_begin + sl0, 4, c < C,
                                     p.partial + (size_t)k * (C + 1) + C, s_begin + lane, 32, s_end, (size_t)n, s, nn);
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            nn = warp_sum(nn);
            if (sl0 == 0) Sq[k][q][ch] = s;
            if (lane == 0) Nq[k][q] = nn;
        }
        __syncthreads();
        if (tid < K * 8) {
            const int kk = tid >> 3, j = tid & 7, cc = cta * 8 + j;
            float nk = (float)(((Nq[kk][0] + Nq[kk][1]) + Nq[kk][2]) + Nq[kk][3]);
            float A = (float)(((Sq[kk][0][j] + Sq[kk][1][j]) + Sq[kk][2][j]) + Sq[kk][3][j]);
            if (p.x.world > 1) {
                if (cc < C) xchg_push(p.x, kk * (C + 1) + cc, A);
                if (cta == 0 && j == 0) xchg_push(p.x, kk * (C + 1) + C, nk);
                const int i2[2] = {kk * (C + 1) + (cc < C ? cc : 0), kk * (C + 1) + C};
                const bool u2[2] = {cc < C, true};
                float o2[2];
                xchg_pull_sums<2>(p.x, i2, u2, o2);
                if (cc < C) A = o2[0];
                nk = o2[1];
            }
            if (cc < C) {
                p.packed2[(size_t)kk * (C + 1) + cc] = A;
                const float po = p.P_s[(size_t)kk * C + cc], pb = p.P_s[(size_t)(K + kk) * C + cc];
                p.g_s[(size_t)kk * C + cc] += p.ema_factor * p.w_disc * p.coef * (nk * po - A);
                p.g_s[(size_t)(K + kk) * C + cc] -= p.ema_factor * p.w_disc * p.coef * (nk * pb - A);
                p.xtab[(size_t)kk * C + cc] = -p.gscale * p.w_disc * p.coef * (po - pb);
            }
            if (cta == 0 && j == 0) p.packed2[(size_t)kk * (C + 1) + C] = nk;
        }
        return;
    }
    // ---- the loss CTA ----
    double v[3] = {0.0, 0.0, 0.0};
    if (p.ps.hinge) v[0] = strided_slot_sum<4>(p.ps.hinge, tid, p.ps.n_hinge, (size_t)p.ps.hinge_stride, kThreads);
    if (p.ps.cons) {
        const double2* c2 = reinterpret_cast<const double2*>(p.ps.cons);
        for (int i = tid; i < p.ps.n_cons; i += 4 * kThreads) {
            double2 t[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) t[u] = (i + u * kThreads < p.ps.n_cons) ? __ldcg(c2 + i + u * kThreads) : make_double2(0.0, 0.0);
#pragma unroll
            for (int u = 0; u < 4; ++u) { v[1] += t[u].x; v[2] += t[u].y; }
        }
    }
    block_sum_n<3>(v, shp);
    if (tid == 0) {
        float* tail = p.packed2 + (size_t)K * (C + 1);
        float t[3] = {(float)v[0], (float)v[1], (float)v[2]};
        if (p.x.world > 1) {
            for (int i = 0; i < 3; ++i) xchg_push(p.x, K * (C + 1) + i, t[i]);
            const int i3[3] = {K * (C + 1), K * (C + 1) + 1, K * (C + 1) + 2};
            const bool u3[3] = {true, true, true};
            xchg_pull_sums<3>(p.x, i3, u3, t);
        }
        tail[0] = t[0]; tail[1] = t[1]; tail[2] = t[2]; tail[3] = 0.f;
        const float disc = (float)((double)t[0] / p.npx);
        const float aug = p.use_cons ? (float)((double)t[1] / (double)t[2] * (double)p.aug_weight) : 0.f;
        p.losses[2] = disc;
        p.losses[3] = aug;
        p.losses[4] = p.w_intra * p.losses[0] + p.w_inter * p.losses[1] + p.w_disc * disc + p.w_aug * aug;
        p.losses[5] = 0.f; p.losses[6] = 0.f;
        if (p.x.world <= 1) p.losses[7] = 0.f;     // sharded: [7] is the exchange-timeout flag (host zeroes it once)
    }
}

}  // namespace clr

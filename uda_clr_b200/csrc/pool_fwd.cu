// clr_pool_fwd: class-wise weighted pooling  S[r][c] = sum_{b,p} x[b,c,p] * w_r[b,p],  N[r] = sum w_r.
//
// Replaces the 4 materialised [B,C,H,W] products + 8 reductions of utils/Utils.py:114-126 (and
// :212-223 for the retrify weights) with ONE read of the feature map.
//
// Bound: HBM.  Algorithmic bytes per domain = 4*B*C*HW (features) + 4*B*WP*HW (weight planes).
//
// Decomposition (both kernels)
//   item   = (domain, sample b, pixel chunk of PX pixels, group of CG channels)  -> 4*PX*CG bytes
//   grid   = persistent: each CTA owns a contiguous range of items, so it walks the channel groups
//            of one (b, chunk) before moving on and re-stages that chunk's R weight rows in shared
//            memory only when (b, chunk) changes.
//   thread = owns VEC*REPS fixed pixels of the chunk, keeps CG*R fp32 accumulators; per item the warp
//            does one transposing butterfly (31 shuffles) and the 8 warps combine through shared memory.
//   output = per-(b,chunk) partials [R][C+1] (column C = weight sums), combined across (b,chunk)
//            in fp64 and in a fixed order by pool_reduce_kernel -> bit-stable run to run, no atomics.
//
// Two data paths for the feature rows
//   pool_fwd_tma_kernel : a producer warp streams each item's CG channel rows (PX*4 contiguous bytes
//            each) into a shared-memory ring with 1-D bulk async copies (cp.async.bulk + mbarrier
//            complete_tx, SASS UBLKCP) so that several 32-64 KB stages are in flight per SM regardless
//            of what the 8 consumer warps are doing (FMA, butterfly, barrier).  Needs HW % 4 == 0 and
//            16-byte aligned bases.
//   pool_fwd_ldg_kernel : 128-bit non-allocating loads straight to registers; also the scalar
//            fallback (VEC = 1) for ragged planes / 4-byte-aligned bases.
#include "clr_common.cuh"
#include "clr_internal.h"

namespace clr {

struct PoolDom {
    const float* feat;
    const float* w;
    float* partial;   // [B*nChunk][R][C+1]
    float* sums;      // [R][C+1]
    float* mu;        // optional [R][C]: the reduce kernel also writes S_r[c] / N_r (single-GPU drop-in: no finalize launch)
    int B;
    int fmt;
    int items;        // B*nChunk*nGroup
    int slots;        // B*nChunk
    int keep_l2;      // 1: this feature map is read again soon (evict-last L2 policy), 0: streaming (evict-first)
};

struct PoolParams {
    PoolDom dom[2];
    int ndom;
    int C, HW;
    int nChunk, nGroup;
    int total;
    int stages;       // TMA ring depth
    int reduce_trace_id;
    int trace_id;                  // TR_POOL, or TR_POOL_T for the target pooling launch of schedule 2
    int skip_reduce;               // 1: the caller reduces the partials itself (pool_finish_kernel)
    unsigned int* counter_reset;   // optional: 8 words zeroed by CTA 0 (the finish stages' last-CTA / completion counters, the gate)
};

// Register blocking: a thread keeps CG x R accumulators -- 32 of them up to R = 8, 64 beyond (R = 16 would otherwise
// leave 2 channels = 8 KB per item, and the per-item butterfly + barriers, not HBM, would bound the kernel).
constexpr int pool_nacc(int R) { return R > 8 ? 64 : 32; }
constexpr int pool_cg(int R) { return (pool_nacc(R) / R) < 16 ? (pool_nacc(R) / R) : 16; }   // channels per item
constexpr int pool_reps(int R) { return (R >= 3 && R <= 8) ? 2 : 1; }    // VEC-wide pixel groups per thread

struct ItemCoord { int d, slot, grp, b, chunk; };

__device__ __forceinline__ ItemCoord decode_item(const PoolParams& p, int it) {
    ItemCoord c;
    c.d = (p.ndom > 1 && it >= p.dom[0].items) ? 1 : 0;
    const int local = it - (c.d ? p.dom[0].items : 0);
    c.slot = local / p.nGroup;
    c.grp = local - c.slot * p.nGroup;
    c.b = c.slot / p.nChunk;
    c.chunk = c.slot - c.b * p.nChunk;
    return c;
}

// Shared-memory layout of the staged weight rows.  Plain: [r][PX].  PAIR (even R >= 8, LDG kernel): rows 2q and 2q+1 are
// interleaved pixel by pixel, [q][PX][2], so that one 128-bit load yields the (w_2q, w_2q+1) pairs of two pixels -- the
// operands of the packed FFMA2 (two fp32 FMAs per issue slot on sm_100) that halves the FMA instruction count of the
// issue-bound K >= 3 pooling; every accumulator still sees the same products in the same order (bit-identical results).
// The pairs of pixels {0,1} and {2,3} of a thread's 4-pixel group live in two planes, [q][2][PX/4][4], so that consecutive
// threads read consecutive 16-byte words (stride 32 bytes would be a 2-way bank conflict on every 128-bit load).
// Measured in the live step (B200, config 1 with K classes; "pool_pair" = 2 selects the scalar loop): R = 16 (K = 8) pooling
// 114.3 -> 108.9 us, R = 8 (K = 4) 69.5 -> 63.0 us, R = 6 (K = 3) 51.2 -> 53.9 us (slower: few FMAs per staged pair) -> R >= 8.
constexpr bool pool_pair(int R) { return R % 2 == 0 && R >= 8; }
template <int PX, bool PAIR>
__device__ __forceinline__ int wsm_idx(int r, int px) {
    if (!PAIR) return r * PX + px;
    const int t = px >> 2, v = px & 3;
    return ((((r >> 1) * 2 + (v >> 1)) * (PX / 4) + t) << 2) + ((v & 1) << 1) + (r & 1);
}
template <int VEC, int PX, bool PAIR>
__device__ __forceinline__ void wsm_store(float* wsm, int r, int px, const Pack<VEC>& w) {
    if constexpr (PAIR) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) wsm[wsm_idx<PX, true>(r, px + v)] = w.v[v];
    } else {
        st_keep<VEC>(wsm + r * PX + px, w);
    }
}

// Stage the R weight rows of (b, chunk) in shared memory (complement rows are materialised here:
// w_bck = 1 - w_obj, utils/Utils.py:111-112).  Pixels past the plane are staged as 0 for every row.
template <int R, int VEC, int REPS, int NT = kThreads, bool PAIR = false>
__device__ __forceinline__ void stage_weights(const PoolDom& D, int b, int px0, int HW, float* wsm, int tid) {
    constexpr int PX = NT * VEC * REPS, K = R / 2;
    const int WP = (D.fmt == CLR_W_COMPLEMENT) ? K : R;
    const float* wb = D.w + (size_t)b * WP * HW;
#pragma unroll
    for (int rep = 0; rep < REPS; ++rep) {
        const int off = (rep * NT + tid) * VEC;
        const bool ok = px0 + off < HW;   // VEC-granular: HW % VEC == 0 on the VEC = 4 paths
        if (D.fmt == CLR_W_COMPLEMENT) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                Pack<VEC> wo, wc;
#pragma unroll
                for (int v = 0; v < VEC; ++v) { wo.v[v] = 0.f; wc.v[v] = 0.f; }
                if (ok) {
                    wo = ld_keep<VEC>(wb + (size_t)k * HW + px0 + off);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) wc.v[v] = 1.0f - wo.v[v];
                }
                wsm_store<VEC, PX, PAIR>(wsm, k, off, wo);
                wsm_store<VEC, PX, PAIR>(wsm, K + k, off, wc);
            }
        } else {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                Pack<VEC> wr;
#pragma unroll
                for (int v = 0; v < VEC; ++v) wr.v[v] = 0.f;
                if (ok) wr = ld_keep<VEC>(wb + (size_t)r * HW + px0 + off);
                wsm_store<VEC, PX, PAIR>(wsm, r, off, wr);
            }
        }
    }
}

template <int VEC>
__device__ __forceinline__ Pack<VEC> lds_pack(const float* p) {
    Pack<VEC> r;
    if constexpr (VEC == 4) {
        const float4 t = *reinterpret_cast<const float4*>(p);
        r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    } else {
        r.v[0] = *p;
    }
    return r;
}

// CTA-level combine of the per-thread accumulators of one item and store of the partial row slices.
// sync() is the barrier over the 256 compute threads.
template <int R, int CG, int VEC, int REPS, int NT = kThreads, bool PAIR = false, typename Sync>
__device__ __forceinline__ void reduce_and_store(float (&acc)[pool_nacc(R)], const float* wsm, float* red, int& parity,
                                                 float* out, int C, int c0, bool owns_counts,
                                                 int tid, Sync sync) {
    constexpr int PX = NT * VEC * REPS, NB = pool_nacc(R) / 32, kWarps = NT / 32;   // (shadows the 256-thread constant)
    static_assert(NT / 32 >= NB, "one warp per 32-accumulator block in the cross-warp combine");
    const int lane = tid & 31, warp = tid >> 5;
    float nsum[R];
    if (owns_counts) {   // the item that owns channel group 0 also sums this chunk's weights
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float s = 0.f;
#pragma unroll
            for (int rep = 0; rep < REPS; ++rep) {
                const int off = (rep * NT + tid) * VEC;
#pragma unroll
                for (int v = 0; v < VEC; ++v) s += wsm[wsm_idx<PX, PAIR>(r, off + v)];
            }
            nsum[r] = s;
        }
    }
    // accumulator i = j*R + r; block b of 32 goes through one transposing butterfly: lane l ends up with entry 32b + l
    float* redp = red + parity * (NB * kWarps * 32);
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        float blk[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) blk[i] = acc[32 * b + i];
        redp[(b * kWarps + warp) * 32 + lane] = warp_sum_transpose32(blk, lane);
    }
    sync();
    if (warp < NB) {
        float s = 0.f;
#pragma unroll
        for (int wq = 0; wq < kWarps; ++wq) s += redp[(warp * kWarps + wq) * 32 + lane];
        const int i = 32 * warp + lane, j = i / R, r = i - j * R;
        if (j < CG && c0 + j < C) out[(size_t)r * (C + 1) + c0 + j] = s;
    }
    parity ^= 1;
    if (owns_counts) {   // CTA-uniform
#pragma unroll
        for (int r = 0; r < R; ++r) nsum[r] = warp_sum(nsum[r]);
        float* redn = red + parity * (NB * kWarps * 32);
        if (lane == 0) {
#pragma unroll
            for (int r = 0; r < R; ++r) redn[warp * 32 + r] = nsum[r];
        }
        sync();
        if (warp == 0 && lane < R) {
            float s = 0.f;
#pragma unroll
            for (int wq = 0; wq < kWarps; ++wq) s += redn[wq * 32 + lane];
            out[(size_t)lane * (C + 1) + C] = s;
        }
        parity ^= 1;
    }
}

// ------------------------------------------------------------------------------------------------
// LDG path (also the scalar fallback)
// ------------------------------------------------------------------------------------------------
// NT = threads per CTA.  The chunk (PX pixels, fixed per R so that the partial layout does not depend on NT) is spread
// over NT threads: NT = 128 gives every thread twice the pixels per item, i.e. twice the FMAs per accumulator between two
// transposing butterflies -- for R > 8 (64 accumulators, 4 channels x 1024 pixels per item) the butterfly's FSEL / SHFL /
// FADD were as many instructions as the FMAs themselves (ncu source page, K = 8: 33.5 M FFMA vs 33.3 M).
// `bid` of `nblk` CTAs walk the items.  (Round 2 shared this body's launch with the MC statistics -- [pool(xs) | mc_stats] as one
// persistent grid -- and lost: at this kernel's 126 registers the MC CTAs run at 2 per SM and take 58 us; profiles/r02_schedule2.md.)
template <int R, int VEC, int NT, bool PAIR_ON = true>
__device__ __forceinline__ void pool_ldg_body(const PoolParams& p, const int bid, const int nblk) {
    if (p.counter_reset && bid == 0 && threadIdx.x < 8) p.counter_reset[threadIdx.x] = 0u;   // last-CTA counters + completion counters + gate
    constexpr int CG = pool_cg(R), REPS = pool_reps(R) * (kThreads / NT), PX = NT * VEC * REPS;
    constexpr bool PAIR = PAIR_ON && pool_pair(R) && VEC == 4;
    static_assert(PX == kThreads * VEC * pool_reps(R), "chunk size is independent of the CTA size");
    extern __shared__ __align__(16) float smem[];
    float* wsm = smem;                 // [R][PX]
    float* red = smem + R * PX;        // [2][NB][kWarps][32]
    const int tid = threadIdx.x;
    int begin, end;
    partition(p.total, nblk, bid, begin, end);
    int cur_key = -1, parity = 0;
    auto sync = [] { __syncthreads(); };
    // (Letting every CTA walk its share of domain 0 before its share of domain 1 -- so that the source map is what the grid
    // read last and the discriminative pass finds more of it in L2 -- was measured in round 2: the re-read's DRAM bytes
    // did not move (109.6 vs 109.9 MB of 135.3) and the step got 1.7 us slower; profiles/r02_l2_harvest.md.)
    for (int it = begin; it < end; ++it) {
        const ItemCoord ic = decode_item(p, it);
        const PoolDom& D = p.dom[ic.d];
        const int px0 = ic.chunk * PX;
        const int key = ic.d * 0x40000000 + ic.slot;
        if (key != cur_key) {
            __syncthreads();
            cur_key = key;
            stage_weights<R, VEC, REPS, NT, PAIR>(D, ic.b, px0, p.HW, wsm, tid);
            __syncthreads();
        }
        float acc[pool_nacc(R)];
#pragma unroll
        for (int i = 0; i < pool_nacc(R); ++i) acc[i] = 0.f;
        const int c0 = ic.grp * CG;
        const float* xb = D.feat + ((size_t)ic.b * p.C + c0) * p.HW + px0;
#pragma unroll
        for (int rep = 0; rep < REPS; ++rep) {
            const int off = (rep * NT + tid) * VEC;
            const bool ok = px0 + off < p.HW;
            Pack<VEC> x[CG];
#pragma unroll
            for (int j = 0; j < CG; ++j) {
#pragma unroll
                for (int v = 0; v < VEC; ++v) x[j].v[v] = 0.f;
                if (ok && c0 + j < p.C) x[j] = ld_stream<VEC>(xb + (size_t)j * p.HW + off);
            }
            // weight row outermost: one row vector live at a time (R = 16 rows would not fit in registers otherwise)
            if constexpr (PAIR) {
                // packed form: (acc[j][2q], acc[j][2q+1]) += (x, x) * (w_2q, w_2q+1), pixel by pixel, in the scalar order
#pragma unroll
                for (int q = 0; q < R / 2; ++q) {
                    float2 wp[VEC];
#pragma unroll
                    for (int h = 0; h < VEC / 2; ++h) {
                        const float4 t = *reinterpret_cast<const float4*>(wsm + ((size_t)(q * 2 + h) * (PX / 4) + (off >> 2)) * 4);
                        wp[2 * h] = make_float2(t.x, t.y);
                        wp[2 * h + 1] = make_float2(t.z, t.w);
                    }
#pragma unroll
                    for (int j = 0; j < CG; ++j) {
                        float2 a = make_float2(acc[j * R + 2 * q], acc[j * R + 2 * q + 1]);
#pragma unroll
                        for (int v = 0; v < VEC; ++v) a = __ffma2_rn(make_float2(x[j].v[v], x[j].v[v]), wp[v], a);
                        acc[j * R + 2 * q] = a.x;
                        acc[j * R + 2 * q + 1] = a.y;
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const Pack<VEC> w = lds_pack<VEC>(wsm + r * PX + off);
#pragma unroll
                    for (int j = 0; j < CG; ++j)
#pragma unroll
                        for (int v = 0; v < VEC; ++v) acc[j * R + r] = fmaf(x[j].v[v], w.v[v], acc[j * R + r]);
                }
            }
        }
        float* out = D.partial + (size_t)ic.slot * R * (p.C + 1);
        reduce_and_store<R, CG, VEC, REPS, NT, PAIR>(acc, wsm, red, parity, out, p.C, c0, ic.grp == 0, tid, sync);
    }
}

template <int R, int VEC, int NT, bool PAIR_ON = true>
__global__ void __launch_bounds__(NT, 2) pool_fwd_ldg_kernel(const PoolParams p) {
    kernel_begin(p.trace_id);
    pool_ldg_body<R, VEC, NT, PAIR_ON>(p, blockIdx.x, gridDim.x);
    trace_exit(p.trace_id);
}

// ------------------------------------------------------------------------------------------------
// TMA path: warp 8 = producer (bulk async copies into a ring of `stages` buffers), warps 0-7 = compute
// ------------------------------------------------------------------------------------------------
constexpr int kPoolTmaThreads = kThreads + 32;
constexpr int kMaxStages = 8;

template <int R>
struct PoolTmaSmem {
    static constexpr int CG = pool_cg(R), REPS = pool_reps(R), PX = kThreads * 4 * REPS;
    static constexpr size_t stage_bytes = sizeof(float) * CG * PX;
    static constexpr size_t w_bytes = sizeof(float) * R * PX;
    static constexpr size_t red_bytes = sizeof(float) * 2 * (pool_nacc(R) / 32) * kWarps * 32;
    static constexpr size_t bar_bytes = sizeof(uint64_t) * 2 * kMaxStages;
    static size_t total(int stages) { return stages * stage_bytes + w_bytes + red_bytes + bar_bytes; }
};

template <int R>
__global__ void __launch_bounds__(kPoolTmaThreads, 1) pool_fwd_tma_kernel(const PoolParams p) {
    kernel_begin(p.trace_id);
    if (p.counter_reset && blockIdx.x == 0 && threadIdx.x < 8) p.counter_reset[threadIdx.x] = 0u;   // last-CTA counters + completion counters + gate
    using SM = PoolTmaSmem<R>;
    constexpr int CG = SM::CG, REPS = SM::REPS, PX = SM::PX, VEC = 4;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* xs = reinterpret_cast<float*>(smem_raw);                                    // [stages][CG][PX]
    float* wsm = reinterpret_cast<float*>(smem_raw + p.stages * SM::stage_bytes);      // [R][PX]
    float* red = wsm + R * PX;                                                         // [2][NB][kWarps][32]
    uint64_t* full = reinterpret_cast<uint64_t*>(red + 2 * (pool_nacc(R) / 32) * kWarps * 32);   // [kMaxStages]
    uint64_t* empty = full + kMaxStages;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kWarps); }
        fence_mbar_init();
    }
    __syncthreads();

    int begin, end;
    partition(p.total, gridDim.x, blockIdx.x, begin, end);

    if (warp == kWarps) {
        // ---------------- producer ----------------
        if (lane == 0) {
            const uint64_t pol_stream = policy_evict_first(), pol_keep = policy_evict_last();
            int stage = 0;
            uint32_t phase = 0;
            for (int it = begin; it < end; ++it) {
                const ItemCoord ic = decode_item(p, it);
                const PoolDom& D = p.dom[ic.d];
                const int px0 = ic.chunk * PX, c0 = ic.grp * CG;
                const int rows = (p.C - c0) < CG ? (p.C - c0) : CG;
                const int npx = (p.HW - px0) < PX ? (p.HW - px0) : PX;
                mbar_wait(&empty[stage], phase ^ 1u);
                mbar_arrive_expect_tx(&full[stage], (uint32_t)(rows * npx * sizeof(float)));
                const float* src = D.feat + ((size_t)ic.b * p.C + c0) * p.HW + px0;
                float* dst = xs + (size_t)stage * CG * PX;
                const uint64_t pol = D.keep_l2 ? pol_keep : pol_stream;
                for (int j = 0; j < rows; ++j)
                    bulk_g2s(dst + j * PX, src + (size_t)j * p.HW, (uint32_t)(npx * sizeof(float)), &full[stage], pol);
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
        }
        return;
    }

    // ---------------- consumers (256 threads) ----------------
    auto sync = [] { named_bar_sync(1, kThreads); };
    int cur_key = -1, parity = 0, stage = 0;
    uint32_t phase = 0;
    for (int it = begin; it < end; ++it) {
        const ItemCoord ic = decode_item(p, it);
        const PoolDom& D = p.dom[ic.d];
        const int px0 = ic.chunk * PX;
        const int key = ic.d * 0x40000000 + ic.slot;
        if (key != cur_key) {
            sync();
            cur_key = key;
            stage_weights<R, VEC, REPS>(D, ic.b, px0, p.HW, wsm, tid);
            sync();
        }
        float acc[pool_nacc(R)];
#pragma unroll
        for (int i = 0; i < pool_nacc(R); ++i) acc[i] = 0.f;
        const int c0 = ic.grp * CG;
        const float* xst = xs + (size_t)stage * CG * PX;
        mbar_wait(&full[stage], phase);
#pragma unroll
        for (int rep = 0; rep < REPS; ++rep) {
            const int off = (rep * kThreads + tid) * VEC;
            const bool ok = px0 + off < p.HW;   // bytes past the copied span are stale: never read them
            Pack<VEC> x[CG];
#pragma unroll
            for (int j = 0; j < CG; ++j) {
#pragma unroll
                for (int v = 0; v < VEC; ++v) x[j].v[v] = 0.f;
                if (ok && c0 + j < p.C) x[j] = lds_pack<VEC>(xst + j * PX + off);
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const Pack<VEC> w = lds_pack<VEC>(wsm + r * PX + off);
#pragma unroll
                for (int j = 0; j < CG; ++j)
#pragma unroll
                    for (int v = 0; v < VEC; ++v) acc[j * R + r] = fmaf(x[j].v[v], w.v[v], acc[j * R + r]);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);   // this warp is done with the stage: producer may refill
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        float* out = D.partial + (size_t)ic.slot * R * (p.C + 1);
        reduce_and_store<R, CG, VEC, REPS>(acc, wsm, red, parity, out, p.C, c0, ic.grp == 0, tid, sync);
    }
    trace_exit(p.trace_id);
}

// sums[r][c] = sum over (b,chunk) slots of partial[slot][r][c], fp64, fixed order.
__global__ void __launch_bounds__(256) pool_reduce_kernel(const PoolParams p, int R) {
    kernel_begin(p.reduce_trace_id);
    const int d = blockIdx.y;
    const PoolDom& D = p.dom[d];
    const int n = R * (p.C + 1);
    // blockDim = (32 columns, 8 slot-lanes)
    const int col = blockIdx.x * 32 + threadIdx.x;
    __shared__ double sh[8][33];
    double s = 0.0;
    if (col < n) {
        // 8 independent loads per round (the slots of one column are n floats apart)
        int sl = threadIdx.y;
        for (; sl + 56 < D.slots; sl += 64) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = D.partial[(size_t)(sl + 8 * u) * n + col];
#pragma unroll
            for (int u = 0; u < 8; ++u) s += (double)v[u];
        }
        for (; sl < D.slots; sl += 8) s += (double)D.partial[(size_t)sl * n + col];
    }
    sh[threadIdx.y][threadIdx.x] = s;
    // mu requested: every thread also sums its row's weight-sum column (same slot split, same order as the thread that owns
    // that column), so the prototype S_r[c] / N_r leaves this launch too -- from the fp32-rounded sums, like clr_proto_finalize
    __shared__ double shn[8][33];
    const int row = col < n ? col / (p.C + 1) : 0;
    if (D.mu) {
        double sn = 0.0;
        if (col < n) {
            const size_t ncol = (size_t)row * (p.C + 1) + p.C;
            int sl = threadIdx.y;
            for (; sl + 56 < D.slots; sl += 64) {
                float v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = D.partial[(size_t)(sl + 8 * u) * n + ncol];
#pragma unroll
                for (int u = 0; u < 8; ++u) sn += (double)v[u];
            }
            for (; sl < D.slots; sl += 8) sn += (double)D.partial[(size_t)sl * n + ncol];
        }
        shn[threadIdx.y][threadIdx.x] = sn;
    }
    __syncthreads();
    if (threadIdx.y == 0 && col < n) {
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += sh[i][threadIdx.x];
        D.sums[col] = (float)t;
        const int c = col - row * (p.C + 1);
        if (D.mu && c < p.C) {
            double tn = 0.0;
#pragma unroll
            for (int i = 0; i < 8; ++i) tn += shn[i][threadIdx.x];
            D.mu[(size_t)row * p.C + c] = (float)t / (float)tn;      // 0/0 -> NaN, as the reference (utils/Utils.py:127-130)
        }
    }
    trace_exit(p.reduce_trace_id);
}

// Per-sample variant: sums_b[b][r][c] = sum over the chunks of sample b only (bmm-style pooling keeps samples apart).
__global__ void __launch_bounds__(256) pool_reduce_ps_kernel(const float* __restrict__ partial, int nChunk, int R, int C,
                                                             float* __restrict__ sums_b) {
    kernel_begin(TR_OTHER);
    const int b = blockIdx.y, n = R * (C + 1);
    const int col = blockIdx.x * 256 + threadIdx.x;
    if (col >= n) return;
    double s = 0.0;
    for (int ch = 0; ch < nChunk; ++ch) s += (double)partial[((size_t)b * nChunk + ch) * n + col];
    sums_b[(size_t)b * n + col] = (float)s;
}

// out[r][c] = mean_b S_b[r][c] / (N_b[r] + n_add)      (Trainer_prototype.py:366-368: bmm / (sum + 1), mean over batch)
__global__ void bmm_finalize_kernel(const float* __restrict__ sums_b, int B, int R, int C, float n_add, float* __restrict__ out) {
    kernel_begin(TR_OTHER);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R * C) return;
    const int r = i / C, c = i - r * C;
    float acc = 0.f;
    for (int b = 0; b < B; ++b) {
        const float* sb = sums_b + ((size_t)b * R + r) * (C + 1);
        acc += sb[c] / (sb[C] + n_add);
    }
    out[i] = acc / (float)B;
}

__global__ void proto_finalize_kernel(const float* __restrict__ sums, int R, int C, float* __restrict__ mu) {
    kernel_begin(TR_OTHER);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R * C) return;
    const int r = i / C, c = i - r * C;
    mu[i] = sums[(size_t)r * (C + 1) + c] / sums[(size_t)r * (C + 1) + C];   // 0/0 -> NaN, as the reference
}

void launch_partial_reduce(const float* partial, int slots, int R, int C, float* sums, cudaStream_t st) {
    PoolParams p{};
    p.ndom = 1; p.C = C;
    p.dom[0].partial = const_cast<float*>(partial);
    p.dom[0].sums = sums;
    p.dom[0].slots = slots;
    p.reduce_trace_id = TR_DISC_REDUCE;
    const int n = R * (C + 1);
    clr::launch_k(pool_reduce_kernel, dim3((n + 31) / 32, 1), dim3(32, 8), 0, st, p, R);
}

static void launch_reduce(const PoolParams& p, int R, cudaStream_t st) {
    const int n = R * (p.C + 1);
    clr::launch_k(pool_reduce_kernel, dim3((n + 31) / 32, p.ndom), dim3(32, 8), 0, st, p, R);
}

template <int R, int VEC, int NT>
static int launch_ldg_nt(const PoolParams& p, cudaStream_t st) {
    constexpr int PX = kThreads * VEC * pool_reps(R);
    constexpr size_t smem = sizeof(float) * (R * PX + 2 * (pool_nacc(R) / 32) * (NT / 32) * 32);
    auto kern = pool_fwd_ldg_kernel<R, VEC, NT>;
    if constexpr (pool_pair(R) && VEC == 4) {
        if (tunables().pool_pair == 2) kern = pool_fwd_ldg_kernel<R, VEC, NT, false>;      // A/B: the scalar-FMA inner loop
    }
    int occ = 0;
    { const int rc = kernel_occupancy(reinterpret_cast<const void*>(kern), NT, smem, &occ); if (rc != CLR_OK) return rc; }
    if (occ < 1) occ = 1;
    int grid = device_facts().sms * occ;
    if (grid > p.total) grid = p.total;
    clr::launch_k(kern, grid, NT, smem, st, p);
    if (!p.skip_reduce) launch_reduce(p, R, st);
    return launch_status();
}

// CTA size: 128 threads (2 CTAs / SM, ~254 registers: every load of an item in flight at once) where the per-item
// butterfly would otherwise rival the FMAs (R > 8: K >= 5), 256 otherwise.  Measured in the live step (bench.py --K):
// K = 8 pooling 129 -> 114 us; K = 4 unchanged (69.7 / 68.5 us); K = 2 (HBM-bound) 45.4 -> 49.7 us, hence the switch
// point.  Compiling the 128-thread form for 3 CTAs / SM (168 registers, spills) lost everywhere.  "pool_threads" tunable:
// 128 / 256 force one of them for A/B runs.
template <int R, int VEC>
static int launch_ldg(const PoolParams& p, cudaStream_t st) {
    const int want = tunables().pool_threads;
    const bool small = want == 128 || (want != 256 && R > 8);
    if constexpr (VEC == 4) {
        if (small) return launch_ldg_nt<R, VEC, 128>(p, st);
    }
    return launch_ldg_nt<R, VEC, kThreads>(p, st);
}

template <int R>
static int launch_tma(PoolParams p, cudaStream_t st) {
    using SM = PoolTmaSmem<R>;
    const size_t budget = (size_t)device_facts().max_smem_optin;
    int stages = tunables().pool_stages > 0 ? tunables().pool_stages : 4;
    if (stages > kMaxStages) stages = kMaxStages;
    while (stages > 1 && SM::total(stages) > budget) --stages;
    if (SM::total(stages) > budget) return CLR_ERR_UNSUPPORTED;
    p.stages = stages;
    const size_t smem = SM::total(stages);
    auto kern = pool_fwd_tma_kernel<R>;
    CLR_RETURN_IF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CLR_RETURN_IF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kPoolTmaThreads, smem));
    if (occ < 1) occ = 1;
    int grid = device_facts().sms * occ;
    if (grid > p.total) grid = p.total;
    clr::launch_k(kern, grid, kPoolTmaThreads, smem, st, p);
    if (!p.skip_reduce) launch_reduce(p, R, st);
    return launch_status();
}

static int dispatch(int R, bool vec4, bool tma, const PoolParams& p, cudaStream_t st) {
    switch (R) {
#define CLR_CASE(r) case r: return tma ? launch_tma<r>(p, st) : (vec4 ? launch_ldg<r, 4>(p, st) : launch_ldg<r, 1>(p, st));
        CLR_CASE(1) CLR_CASE(2) CLR_CASE(3) CLR_CASE(4) CLR_CASE(5) CLR_CASE(6) CLR_CASE(7) CLR_CASE(8)
        CLR_CASE(9) CLR_CASE(10) CLR_CASE(11) CLR_CASE(12) CLR_CASE(13) CLR_CASE(14) CLR_CASE(15) CLR_CASE(16)
#undef CLR_CASE
    }
    return CLR_ERR_UNSUPPORTED;
}

static int chunk_px(int R, int vec) { return kThreads * vec * pool_reps(R); }

// Workspace: partials for one domain, sized for the scalar (smallest-chunk) path.
static size_t partial_floats(int B, int C, int HW, int R) {
    const int px = chunk_px(R, 1);
    const size_t slots = (size_t)B * ((HW + px - 1) / px);
    return slots * R * (C + 1);
}

size_t pool_partial_bytes(int B, int C, int HW, int R) { return sizeof(float) * partial_floats(B, C, HW, R); }

// R = number of output rows (2K for the prototype formats; any 1..16 for explicit rows).
int pool_fwd_impl(const float* feat0, const float* w0, int fmt0, int B0, float* sums0,
                  const float* feat1, const float* w1, int fmt1, int B1, float* sums1,
                  int C, int HW, int R, void* ws, size_t ws_bytes, cudaStream_t st, int keep0, int keep1,
                  PoolLayout* skip_reduce_layout, unsigned int* counter_reset, float* mu0, int trace_id) {
    const int ndom = feat1 ? 2 : 1;
    CLR_CHECK_ARG(feat0 && w0 && sums0 && ws && B0 > 0 && C > 0 && HW > 0 && R >= 1 && R <= 2 * CLR_MAX_K);
    CLR_CHECK_ARG(fmt0 == CLR_W_COMPLEMENT || fmt0 == CLR_W_EXPLICIT);
    CLR_CHECK_ARG(fmt0 != CLR_W_COMPLEMENT || R % 2 == 0);
    if (ndom == 2) {
        CLR_CHECK_ARG(w1 && sums1 && B1 > 0 && (fmt1 == CLR_W_COMPLEMENT || fmt1 == CLR_W_EXPLICIT));
        CLR_CHECK_ARG(fmt1 != CLR_W_COMPLEMENT || R % 2 == 0);
    }
    if (!aligned4(feat0) || !aligned4(w0) || (feat1 && (!aligned4(feat1) || !aligned4(w1)))) return CLR_ERR_ALIGN;
    const bool vec4 = (HW % 4 == 0) && aligned16(feat0) && aligned16(w0) && (!feat1 || (aligned16(feat1) && aligned16(w1)));
    const bool tma = vec4 && tunables().pool_impl == 2;
    const int vec = vec4 ? 4 : 1;
    const int px = chunk_px(R, vec);
    const int CG = pool_cg(R);

    PoolParams p{};
    p.ndom = ndom; p.C = C; p.HW = HW;
    p.nChunk = (HW + px - 1) / px;
    p.nGroup = (C + CG - 1) / CG;
    const size_t need = sizeof(float) * (partial_floats(B0, C, HW, R) + (ndom == 2 ? partial_floats(B1, C, HW, R) : 0));
    if (ws_bytes < need) return CLR_ERR_WORKSPACE;
    float* wsf = static_cast<float*>(ws);
    const long long items0 = (long long)B0 * p.nChunk * p.nGroup;
    const long long items1 = ndom == 2 ? (long long)B1 * p.nChunk * p.nGroup : 0;
    if (items0 + items1 > 0x3fffffff) return CLR_ERR_UNSUPPORTED;
    p.dom[0] = PoolDom{feat0, w0, wsf, sums0, mu0, B0, fmt0, (int)items0, B0 * p.nChunk, keep0};
    if (ndom == 2)
        p.dom[1] = PoolDom{feat1, w1, wsf + partial_floats(B0, C, HW, R), sums1, nullptr, B1, fmt1, (int)items1, B1 * p.nChunk, keep1};
    p.total = (int)(items0 + items1);
    p.reduce_trace_id = TR_POOL_REDUCE;
    p.trace_id = trace_id;
    p.counter_reset = counter_reset;
    if (skip_reduce_layout) {
        p.skip_reduce = 1;
        skip_reduce_layout->partial[0] = p.dom[0].partial; skip_reduce_layout->slots[0] = p.dom[0].slots;
        skip_reduce_layout->partial[1] = ndom == 2 ? p.dom[1].partial : nullptr;
        skip_reduce_layout->slots[1] = ndom == 2 ? p.dom[1].slots : 0;
    }
    return dispatch(R, vec4, tma, p, st);
}

}  // namespace clr

extern "C" {

size_t clr_pool_ws_bytes(int B, int C, int HW, int K) {
    if (B <= 0 || C <= 0 || HW <= 0 || K < 1 || K > CLR_MAX_K) return 0;
    return sizeof(float) * clr::partial_floats(B, C, HW, 2 * K);
}

size_t clr_pool_rows_ws_bytes(int B, int C, int HW, int R) {
    if (B <= 0 || C <= 0 || HW <= 0 || R < 1 || R > 2 * CLR_MAX_K) return 0;
    return sizeof(float) * clr::partial_floats(B, C, HW, R);
}

int clr_pool_rows_fwd(const float* feat, const float* rows, int B, int C, int HW, int R,
                      void* ws, size_t ws_bytes, float* sums, clr_stream_t stream) {
    return clr::pool_fwd_impl(feat, rows, CLR_W_EXPLICIT, B, sums, nullptr, nullptr, 0, 0, nullptr, C, HW, R, ws, ws_bytes,
                              static_cast<cudaStream_t>(stream));
}

int clr_pool_rows_fwd_ps(const float* feat, const float* rows, int B, int C, int HW, int R,
                         void* ws, size_t ws_bytes, float* sums_b, clr_stream_t stream) {
    if (!sums_b || !ws || ws_bytes < sizeof(float) * ((size_t)R * (C + 1)) + clr_pool_rows_ws_bytes(B, C, HW, R)) return CLR_ERR_WORKSPACE;
    // run the ordinary pooling (its batch-wide sums land in the head of the workspace and are ignored), then
    // re-reduce the per-(b,chunk) partials sample by sample
    float* scratch = static_cast<float*>(ws);
    char* partial = static_cast<char*>(ws) + sizeof(float) * (size_t)R * (C + 1);
    const size_t pbytes = ws_bytes - sizeof(float) * (size_t)R * (C + 1);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = clr::pool_fwd_impl(feat, rows, CLR_W_EXPLICIT, B, scratch, nullptr, nullptr, 0, 0, nullptr, C, HW, R, partial, pbytes, st);
    if (rc != CLR_OK) return rc;
    const bool vec4 = (HW % 4 == 0) && clr::aligned16(feat) && clr::aligned16(rows);
    const int px = clr::chunk_px(R, vec4 ? 4 : 1);
    const int nChunk = (HW + px - 1) / px;
    const int n = R * (C + 1);
    clr::launch_k(clr::pool_reduce_ps_kernel, dim3((n + 255) / 256, B), 256, 0, st,
                  reinterpret_cast<const float*>(partial), nChunk, R, C, sums_b);
    return clr::launch_status();
}

int clr_bmm_finalize(const float* sums_b, int B, int R, int C, float n_add, float* out, clr_stream_t stream) {
    if (!sums_b || !out || B < 1 || R < 1 || C < 1) return CLR_ERR_BAD_ARG;
    const int n = R * C;
    clr::launch_k(clr::bmm_finalize_kernel, (n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream), sums_b, B, R, C, n_add, out);
    return clr::launch_status();
}

int clr_pool_fwd(const float* feat, const float* w, int fmt, int B, int C, int HW, int K,
                 void* ws, size_t ws_bytes, float* sums, clr_stream_t stream) {
    if (K < 1 || K > CLR_MAX_K) return CLR_ERR_BAD_ARG;
    return clr::pool_fwd_impl(feat, w, fmt, B, sums, nullptr, nullptr, 0, 0, nullptr, C, HW, 2 * K, ws, ws_bytes,
                              static_cast<cudaStream_t>(stream));
}

int clr_pool_fwd_mu(const float* feat, const float* w, int fmt, int B, int C, int HW, int K,
                    void* ws, size_t ws_bytes, float* sums, float* mu, clr_stream_t stream) {
    if (K < 1 || K > CLR_MAX_K || !mu) return CLR_ERR_BAD_ARG;
    return clr::pool_fwd_impl(feat, w, fmt, B, sums, nullptr, nullptr, 0, 0, nullptr, C, HW, 2 * K, ws, ws_bytes,
                              static_cast<cudaStream_t>(stream), 0, 0, nullptr, nullptr, mu);
}

int clr_pool_fwd2(const float* feat0, const float* w0, int fmt0, int B0,
                  const float* feat1, const float* w1, int fmt1, int B1,
                  int C, int HW, int K, void* ws, size_t ws_bytes,
                  float* sums0, float* sums1, clr_stream_t stream) {
    if (!feat1 || K < 1 || K > CLR_MAX_K) return CLR_ERR_BAD_ARG;
    return clr::pool_fwd_impl(feat0, w0, fmt0, B0, sums0, feat1, w1, fmt1, B1, sums1, C, HW, 2 * K, ws, ws_bytes,
                              static_cast<cudaStream_t>(stream));
}

int clr_proto_finalize(const float* sums, int R, int C, float* mu, clr_stream_t stream) {
    if (!sums || !mu || R <= 0 || C <= 0) return CLR_ERR_BAD_ARG;
    const int n = R * C;
    clr::launch_k(clr::proto_finalize_kernel, (n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream), sums, R, C, mu);
    return clr::launch_status();
}

}  // extern "C"

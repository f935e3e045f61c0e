// Prototype-guided discriminative hinge, forward, in ONE read of the source features
// (Trainer_prototype_mt.cpython-38.pyc L454-474).
//
// Per pixel the hinge argument delta_k = d_obj,k - d_bck,k is affine in x (a dot product with
// D_k = P_obj,k - P_bck,k over the channel axis); the loss's gradient w.r.t. the prototypes needs the
// active-set sums A_k[c] = sum_p coef_k(p) x[c,p] (a reduction over the pixel axis with coefficients that
// are only known once the pixel's dot products are complete).  Both reductions run over the same
// [C x TP] tile while it sits in shared memory, so the feature map is read once instead of twice
// (clr_disc_fwd + clr_pool_rows_fwd).
//
// Bound: HBM.  Algorithmic bytes = 4*B*C*HW (features) + 2*4*B*K*HW (labels in, coefficients out).
//
//   tile     = all C channels x TP (=64 or 32) consecutive pixels of one sample: C rows of 4*TP bytes
//   fetch    = all 256 threads issue 16-byte cp.async (LDGSTS) chunks of the tile STAGES-1 ahead into a ring of
//              shared-memory stages (commit / wait groups).  Bulk-TMA row copies were measured slower here:
//              a row is only 4*TP = 128-256 bytes, far below the size at which cp.async.bulk is efficient.
//   compute  = the same 256 threads.  phase 1: thread (pixel quad g, channel slice s) accumulates K dot products
//              over its slice, slices are combined through shared memory, the epilogue evaluates the
//              hinge and writes coef (global + shared).  phase 2: thread = channel c walks its row and
//              accumulates K coefficient-weighted sums in registers ACROSS ALL TILES of the CTA.
//              Rows are padded by 4 floats so both access patterns are bank-conflict free.
//   output   = per-CTA partials [K][C+1] (col C = sum of coefficients) + per-CTA hinge sum; combined in
//              fp64, fixed order, by the pooling reduce kernel -> deterministic.
#include <cuda.h>   // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, no libcuda link)
#include "clr_common.cuh"
#include "clr_internal.h"

namespace clr {

// Per-phase cycle accounting of CTA 0 (build with -DCLR_PHASE_PROFILE; read through the trace slots dbg0..dbg4)
#ifdef CLR_PHASE_PROFILE
#define PH_DECL long long ph_t = clock64(), ph_acc[5] = {0, 0, 0, 0, 0};
#define PH_MARK(i) { const long long now_ = clock64(); ph_acc[i] += now_ - ph_t; ph_t = now_; }
#define PH_FLUSH if (g_trace_dev && blockIdx.x == 0 && threadIdx.x == 0) { for (int i_ = 0; i_ < 5; ++i_) { g_trace_dev[TR_DBG0 + i_].t_first = 0; g_trace_dev[TR_DBG0 + i_].t_ready = 0; g_trace_dev[TR_DBG0 + i_].t_last = (unsigned long long)ph_acc[i_]; g_trace_dev[TR_DBG0 + i_].n_cta = 1; } }
#else
#define PH_DECL
#define PH_MARK(i)
#define PH_FLUSH
#endif

struct DiscParams {
    const float* xs;
    const float* ys;        // [B,K,HW]
    const float* V;         // [K][C]  D_k
    const float* beta;      // [K]
    float* coef;            // [B,K,HW]
    float* delta;           // [B,K,HW] or null
    float* partial;         // [grid][K][C+1]
    float* hinge;           // [grid]
    float alpha, margin;
    int B, C, HW, tilesPerSample, total, stages;
    int reverse;            // 1: walk the tiles in DESCENDING address order (L2 harvest of what the pooling pass read last)
    int rows_box, nbox;     // TMA path: the [C x 32] tile is fetched as nbox boxes of rows_box channel rows (rows_box % 8 == 0)
    // flag dependency (clr_common.cuh): instead of griddepcontrol.wait on the whole producer grid, wait until `wait_fin`
    // reaches wait_fin_n before the first read of V / beta, and until `wait_all` reaches wait_all_n before exiting
    const unsigned int* wait_fin; const unsigned int* wait_all;
    unsigned int wait_fin_n, wait_all_n;
    float* err;             // set to 1 if a wait times out
    // with the flag dependency the per-class offsets are summed here from the finish CTAs' loss partials
    // ([beta_nparts][2 + CLR_MAX_K] doubles, column 2 + k), in pool_finish_body's own order -> the same float
    const double* beta_partial; int beta_nparts;
};

constexpr int kDiscPad = 4;
constexpr int kDiscMaxParts = 320;   // >= resident CTAs (2 per SM): one partial per CTA
constexpr int kDiscCols = 64;      // phase 2: 64 channel columns x 4 pixel quarters = 256 threads
constexpr int kDiscMaxCPT = 12;    // channels per thread in phase 2 (C <= 768)
constexpr int kDiscMaxAcc = 48;    // (C/64)*K register accumulators at most

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// One box of a rank-3 tensor map (pixels, channels, samples) -> shared memory; completion on `bar` (complete_tx bytes).
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tmap, int x, int y, int z, uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;"
        ::"r"(smem_u32(dst)), "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)), "l"(policy) : "memory");
}

// TMA = false: tile rows fetched with 16-byte cp.async (LDGSTS) by all threads into padded rows.
// TMA = true : the whole [C x 32] tile is ONE cp.async.bulk.tensor box per <= 256 channel rows (rank-3 tensor map
//              (pixel, channel, sample), 128-byte swizzle instead of padding, zero fill past ragged edges), issued by one
//              thread; completion through an mbarrier per stage.  The LDGSTS form is bound by the SM's load/store
//              pipe (measured: ~2300 cycles per tile spent behind the next tile's 2048 LDGSTS, none waiting for
//              data), the tensor form leaves that pipe to the two compute phases.
template <int K, int TP, int STAGES, int NT, bool TMA, int CPT, int MB = ((TP == 32 && STAGES <= 4) ? 2 : 1)>
__global__ void __launch_bounds__(NT, MB) disc_fused_kernel(const DiscParams p, const __grid_constant__ CUtensorMap tmap) {
    trace_enter(TR_DISC);
    pdl_trigger();
    if (!p.wait_fin) pdl_wait();      // else: the feature tiles are inputs -- start fetching, wait for the flag below
    trace_ready(TR_DISC);
    static_assert(!TMA || TP == 32, "the tensor-map path uses 128-byte rows");
    constexpr int RS = TMA ? TP : TP + kDiscPad; // row stride (floats): dense + swizzled (TMA) or padded (cp.async)
    constexpr int NG = TP / 4;                   // pixel quads (16-byte chunks) per row
    constexpr int NS = NT / NG;                  // channel slices in phase 1
    constexpr int NW = NT / 32;                  // warps
    constexpr int PS = NT / kDiscCols;           // pixel splits in phase 2
    constexpr int NE = (K * TP + NT - 1) / NT;   // epilogue iterations per thread
    extern __shared__ __align__(1024) unsigned char smem_raw[];              // 128B swizzle: stages 1024-byte aligned
    const int rows_total = TMA ? p.rows_box * p.nbox : p.C;
    const size_t stage_floats = (size_t)rows_total * RS;
    float* tiles = reinterpret_cast<float*>(smem_raw);                       // [STAGES][rows][RS]
    float* red = tiles + (size_t)STAGES * stage_floats;                      // [NW][K][TP]
    float* cfs = red + (size_t)NW * K * TP;                              // [K][TP]
    float* wred = cfs + (size_t)K * TP;                                      // [1 + NE][NW]
    uint64_t* full = reinterpret_cast<uint64_t*>(wred + (1 + NE) * NW + ((1 + NE) * NW & 1));   // [STAGES] (TMA), 8-byte aligned
    float* Vs = reinterpret_cast<float*>(full + STAGES);                     // [K][C] contraction vectors + [K] offsets
    float* betas = Vs + (size_t)K * p.C;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (TMA) {
        if (tid == 0) {
            for (int st = 0; st < STAGES; ++st) mbar_init(&full[st], 1);
            fence_mbar_init();
        }
        __syncthreads();
    }
    // physical float offset of 16-byte chunk `q` in row `c`: XOR swizzle (TMA) or identity (padded rows)
    auto chunk_off = [&](int c, int q) -> int { return TMA ? 4 * (q ^ (c & 7)) : 4 * q; };

    // Tiles are dealt round-robin (CTA k: tiles k, k+grid, ..): at any moment the resident CTAs fetch ADJACENT
    // 128-byte segments of the same channel rows, which keeps DRAM pages open across CTAs -- a contiguous range
    // per CTA would have every CTA touching its own page of every row.
    const int begin = blockIdx.x, end = p.total, step = gridDim.x;

    // Asynchronous tile fetch by all 256 threads: 16-byte cp.async (LDGSTS) chunks; thread (q = tid % NG chunk
    // column, r0 = tid / NG row) copies rows r0, r0 + NS, ..: pointer increments only.  Columns past a ragged plane
    // end are zero-filled.
    const int fq = tid % NG, fr0 = tid / NG;
    uint64_t pol_stream = 0;
    if (TMA && tid == 0) pol_stream = policy_evict_first();
    auto issue = [&](int b, int tile, int stage) {
        if (TMA) {
            if (tid == 0 && b < p.B) {
                mbar_arrive_expect_tx(&full[stage], (uint32_t)(stage_floats * sizeof(float)));
                for (int j = 0; j < p.nbox; ++j)
                    tma_load_3d(tiles + (size_t)stage * stage_floats + (size_t)j * p.rows_box * RS, &tmap, tile * TP, j * p.rows_box, b,
                                &full[stage], pol_stream);
            }
            return;
        }
        if (b < p.B) {
            const int px0 = tile * TP;
            const int nchunk = ((p.HW - px0) < TP ? (p.HW - px0) : TP) / 4;
            const float* src = p.xs + ((size_t)b * p.C + fr0) * p.HW + px0 + 4 * fq;
            float* dst = tiles + (size_t)stage * stage_floats + (size_t)fr0 * RS + 4 * fq;
            const size_t sstep = (size_t)NS * p.HW;
            if (fq < nchunk) {
                for (int c = fr0; c < p.C; c += NS, src += sstep, dst += NS * RS) cp_async16(dst, src);
            } else {
                for (int c = fr0; c < p.C; c += NS, dst += NS * RS) *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        cp_async_commit();
    };
    // (b, tile) of the item `step` further on
    auto advance = [&](int& b, int& tile) {
        if (p.reverse) {
            tile -= step;
            while (tile < 0) { tile += p.tilesPerSample; --b; }
        } else {
            tile += step;
            while (tile >= p.tilesPerSample) { tile -= p.tilesPerSample; ++b; }
        }
    };

    const int g = tid % NG, s = tid / NG;
    float A[CPT][K];             // CPT = channels per thread in phase 2 (5: C <= 320, 12: C <= 768)
#pragma unroll
    for (int i = 0; i < CPT; ++i)
#pragma unroll
        for (int k = 0; k < K; ++k) A[i][k] = 0.f;
    float hinge_sum = 0.f;               // epilogue threads: running sums of their (k, pixel) entries
    float ncf[NE];
#pragma unroll
    for (int i = 0; i < NE; ++i) ncf[i] = 0.f;

    const int first = p.reverse ? (p.total - 1 - begin) : begin;             // (negative only when this CTA has no item)
    int b = first >= 0 ? first / p.tilesPerSample : 0, tile = first >= 0 ? first - b * p.tilesPerSample : 0;   // current item
    int fb = b, ftile = tile;                                               // next item to fetch
    int fetched = begin;
#pragma unroll
    for (int j = 0; j < STAGES - 1; ++j) {
        issue(fetched < end ? fb : p.B, ftile, j);
        advance(fb, ftile); fetched += step;
    }
    // the contraction vectors come from the finish CTAs of the preceding launch: wait for THEM (not for the whole grid),
    // then stage V / beta in shared memory through coherent loads (they were written while this kernel was resident,
    // so the non-coherent __ldg path must not be used for them)
    // (a missed dependency -- ~2 s -- sets losses[7] and poisons the contraction vectors with NaN: the step's losses and
    // gradients turn NaN instead of being computed from stale vectors)
    __shared__ int dep_bad;
    bool stale = false;
    if (p.wait_fin) {
        if (tid == 0) {
            dep_bad = spin_until_at_least(p.wait_fin, p.wait_fin_n) ? 0 : 1;
            if (dep_bad && p.err) *p.err = 1.f;
        }
        __syncthreads();
        stale = dep_bad != 0;
    }
    for (int i = tid; i < K * p.C; i += NT) Vs[i] = stale ? __int_as_float(0x7fc00000) : __ldcg(p.V + i);
    if (p.beta_partial) {
        if (warp < K) {      // K <= 8 = warps of the 256-thread CTA
            double t = 0.0;
            for (int bb = lane; bb < p.beta_nparts; bb += 32) t += __ldcg(p.beta_partial + (size_t)bb * (2 + CLR_MAX_K) + 2 + warp);
            t = warp_sum(t);
            if (lane == 0) betas[warp] = (float)(t * (1.0 / (double)p.C));
        }
    } else if (tid < K) {
        betas[tid] = __ldcg(p.beta + tid);
    }
    __syncthreads();
    int stage = 0;
    uint32_t phase = 0;
    PH_DECL
    for (int it = begin; it < end; it += step, advance(b, tile)) {
        const int px0 = tile * TP;
        const int npx = (p.HW - px0) < TP ? (p.HW - px0) : TP;
        // labels of this thread's epilogue entries: issued now, consumed after phase 1
        float yv[NE];
#pragma unroll
        for (int ei = 0; ei < NE; ++ei) {
            const int e = tid + ei * NT;
            const int k = e / TP, j = e - k * TP;
            yv[ei] = (e < K * TP && j < npx) ? __ldg(p.ys + ((size_t)b * K + k) * p.HW + px0 + j) : 0.f;
        }
        PH_MARK(4)
        if (TMA) mbar_wait(&full[stage], phase);     // the tile's boxes have landed
        else cp_async_wait<STAGES - 2>();            // this thread's copies of tile `it` have landed
        __syncthreads();                 // ... everyone's have, and everyone is done with the previous tile
        PH_MARK(0)
        {
            int nst = stage + STAGES - 1;
            if (nst >= STAGES) nst -= STAGES;
            issue(fetched < end ? fb : p.B, ftile, nst);  // refill the stage the previous tile occupied
            advance(fb, ftile); fetched += step;
        }
        const float* xt = tiles + (size_t)stage * stage_floats;
        // ---- phase 1: K dot products over the channel axis ----------------------------------------
        float acc[K][4];
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int v = 0; v < 4; ++v) acc[k][v] = 0.f;
#pragma unroll 4
        for (int c = s; c < p.C; c += NS) {
            const float4 x = *reinterpret_cast<const float4*>(xt + (size_t)c * RS + chunk_off(c, g));
            float vv[K];   // D_k[c] from the shared-memory copy
#pragma unroll
            for (int k = 0; k < K; ++k) vv[k] = Vs[(size_t)k * p.C + c];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const float vk = vv[k];
                acc[k][0] = fmaf(x.x, vk, acc[k][0]);
                acc[k][1] = fmaf(x.y, vk, acc[k][1]);
                acc[k][2] = fmaf(x.z, vk, acc[k][2]);
                acc[k][3] = fmaf(x.w, vk, acc[k][3]);
            }
        }
        // the 32/NG slices of a warp share their pixel quads: combine them in registers first
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int v = 0; v < 4; ++v) {
#pragma unroll
                for (int o = NG; o < 32; o <<= 1) acc[k][v] += __shfl_xor_sync(0xffffffffu, acc[k][v], o);
            }
        if (lane < NG) {
#pragma unroll
            for (int k = 0; k < K; ++k)
                *reinterpret_cast<float4*>(red + ((size_t)warp * K + k) * TP + 4 * g) =
                    make_float4(acc[k][0], acc[k][1], acc[k][2], acc[k][3]);
        }
        __syncthreads();
        PH_MARK(1)
        // ---- epilogue: thread e -> (k, pixel j) ------------------------------------------------------
#pragma unroll
        for (int ei = 0; ei < NE; ++ei) {
            const int e = tid + ei * NT;
            if (e >= K * TP) break;
            const int k = e / TP, j = e - k * TP;
            float dot = 0.f;
#pragma unroll
            for (int q = 0; q < NW; ++q) dot += red[((size_t)q * K + k) * TP + j];
            float cf = 0.f;
            if (j < npx) {
                const size_t o = ((size_t)b * K + k) * p.HW + px0 + j;
                const float delta = fmaf(p.alpha, dot, betas[k]);
                const float ho = delta + p.margin, hb = p.margin - delta;
                hinge_sum += yv[ei] * fmaxf(ho, 0.f) + (1.f - yv[ei]) * fmaxf(hb, 0.f);
                cf = (ho > 0.f ? yv[ei] : 0.f) - (hb > 0.f ? (1.f - yv[ei]) : 0.f);
                p.coef[o] = cf;
                if (p.delta) p.delta[o] = delta;
            }
            cfs[e] = cf;
            ncf[ei] += cf;
        }
        __syncthreads();
        PH_MARK(2)
        // ---- phase 2: coefficient-weighted sums over the pixel axis.  Thread = (channel column cp, pixel quarter h):
        //      it walks channels cp, cp+64, .. over its quarter of the tile's pixels, so the coefficient loads
        //      are shared by all of the thread's channels (12 LDS.128 per 64 FMA at C = 256, K = 2)
        {
            const int cp = tid % kDiscCols, h = tid / kDiscCols;
            constexpr int JQ = NG / PS;            // pixel quads per split
            static_assert(JQ >= 1, "tile too narrow for this CTA size");
            float4 cf[JQ][K];
#pragma unroll
            for (int jq = 0; jq < JQ; ++jq)
#pragma unroll
                for (int k = 0; k < K; ++k) cf[jq][k] = *reinterpret_cast<const float4*>(cfs + k * TP + 4 * (h * JQ + jq));
#pragma unroll
            for (int i = 0; i < CPT; ++i) {
                const int c = cp + i * kDiscCols;
                if (c < p.C) {
                    const float* row = xt + (size_t)c * RS;
#pragma unroll
                    for (int jq = 0; jq < JQ; ++jq) {
                        const float4 x = *reinterpret_cast<const float4*>(row + chunk_off(c, h * JQ + jq));
#pragma unroll
                        for (int k = 0; k < K; ++k) {
                            A[i][k] = fmaf(x.x, cf[jq][k].x, A[i][k]);
                            A[i][k] = fmaf(x.y, cf[jq][k].y, A[i][k]);
                            A[i][k] = fmaf(x.z, cf[jq][k].z, A[i][k]);
                            A[i][k] = fmaf(x.w, cf[jq][k].w, A[i][k]);
                        }
                    }
                }
            }
        }
        PH_MARK(3)
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        // `red` is next written after the next tile's first barrier, `cfs` after its second: both after every
        // thread finished this tile's phase 2, so no trailing barrier is needed
    }
    PH_FLUSH
    if (!TMA) cp_async_wait<0>();
    // ---- CTA partials: the 4 pixel quarters of every (k, c) are combined through shared memory (tiles are free now)
    __syncthreads();
    {
        float* comb = tiles;                     // [PS][K][C]
        const int cp = tid % kDiscCols, h = tid / kDiscCols;
#pragma unroll
        for (int i = 0; i < CPT; ++i) {
            const int c = cp + i * kDiscCols;
            if (c < p.C) {
#pragma unroll
                for (int k = 0; k < K; ++k) comb[((size_t)h * K + k) * p.C + c] = A[i][k];
            }
        }
    }
    __syncthreads();
    float* out = p.partial + (size_t)blockIdx.x * K * (p.C + 1);
    for (int i = tid; i < K * p.C; i += NT) {
        const int k = i / p.C, c = i - k * p.C;
        const float* comb = tiles;
        float t = 0.f;
#pragma unroll
        for (int h = 0; h < PS; ++h) t += comb[(size_t)h * K * p.C + i];
        out[(size_t)k * (p.C + 1) + c] = t;
    }
    const float hs = warp_sum(hinge_sum);
    if (lane == 0) wred[warp] = hs;
#pragma unroll
    for (int ei = 0; ei < NE; ++ei) {
        const float t = warp_sum(ncf[ei]);
        if (lane == 0) wred[(1 + ei) * NW + warp] = t;
    }
    __syncthreads();
    if (tid == 0) {
        float t = 0.f;
        for (int w = 0; w < NW; ++w) t += wred[w];
        p.hinge[blockIdx.x] = t;
    }
    if (tid < K) {
        // entry e = ei*256 + warp*32 + lane belongs to class e / TP; TP is a multiple of 32, so a warp is one class
        float t = 0.f;
        for (int ei = 0; ei < NE; ++ei)
            for (int w = 0; w < NW; ++w)
                if ((ei * NT + w * 32) / TP == tid) t += wred[(1 + ei) * NW + w];
        out[(size_t)tid * (p.C + 1) + p.C] = t;
    }
    // completing this grid must imply that the WHOLE producer launch is complete (its streaming CTAs too): the next
    // kernel's griddepcontrol.wait only covers this grid
    if (p.wait_all && tid == 0 && !spin_until_at_least(p.wait_all, p.wait_all_n) && p.err) *p.err = 1.f;
    trace_exit(TR_DISC);
}

template <int K, int TP, int NT>
static size_t disc_smem(int C, int stages) {
    constexpr int RS = TP + kDiscPad, NW = NT / 32, NE = (K * TP + NT - 1) / NT;
    size_t fl = (size_t)stages * C * RS + (size_t)NW * K * TP + (size_t)K * TP + (1 + NE) * NW;
    const size_t comb = (size_t)(NT / kDiscCols) * K * C;     // end-of-kernel combine buffer aliases the tile ring
    if (comb > (size_t)stages * C * RS) fl += comb - (size_t)stages * C * RS;
    return (fl + (size_t)K * C + K + 2) * sizeof(float) + 16 + sizeof(uint64_t) * 8;
}

// tensor-map path: dense 128-byte rows (rows_total >= C, a multiple of 8), one mbarrier per stage
template <int K, int NT>
static size_t disc_smem_tma(int C, int rows_total, int stages) {
    constexpr int TP = 32, NW = NT / 32, NE = (K * TP + NT - 1) / NT;
    size_t fl = (size_t)stages * rows_total * TP + (size_t)NW * K * TP + (size_t)K * TP + (1 + NE) * NW + 2;
    const size_t comb = (size_t)(NT / kDiscCols) * K * C;
    if (comb > (size_t)stages * rows_total * TP) fl += comb - (size_t)stages * rows_total * TP;
    return (fl + (size_t)K * C + K + 2) * sizeof(float) + sizeof(uint64_t) * stages;
}

static int finish_launch_geometry(DiscParams& p, int TP, int occ, int* nparts) {
    p.tilesPerSample = (p.HW + TP - 1) / TP;
    const long long total = (long long)p.B * p.tilesPerSample;
    if (total > 0x3fffffff) return -1;
    p.total = (int)total;
    p.reverse = tunables().disc_reverse;
    int grid = device_facts().sms * occ;
    if (grid > p.total) grid = p.total;
    if (grid > *nparts) grid = *nparts;
    *nparts = grid;
    return grid;
}

constexpr int kDiscSmallCPT = 5;   // C <= 320 (the reference's 256 / 305 channels): 5*K accumulators instead of 12*K

template <int K, int TP, int STAGES, int NT, int CPT>
static int launch_disc_c(DiscParams& p, int* nparts, cudaStream_t st) {
    const size_t smem = disc_smem<K, TP, NT>(p.C, STAGES);
    auto kern = disc_fused_kernel<K, TP, STAGES, NT, false, CPT>;
    int occ = 0;
    { const int rc = kernel_occupancy(reinterpret_cast<const void*>(kern), NT, smem, &occ); if (rc != CLR_OK) return rc; }
    if (occ < 1) return CLR_ERR_UNSUPPORTED;
    const int grid = finish_launch_geometry(p, TP, occ, nparts);
    if (grid < 1) return CLR_ERR_UNSUPPORTED;
    CUtensorMap unused{};
    clr::launch_k(kern, grid, NT, smem, st, p, unused);
    return launch_status();
}
template <int K, int TP, int STAGES, int NT>
static int launch_disc_s(DiscParams& p, int* nparts, cudaStream_t st) {
    return p.C <= kDiscSmallCPT * kDiscCols ? launch_disc_c<K, TP, STAGES, NT, kDiscSmallCPT>(p, nparts, st)
                                            : launch_disc_c<K, TP, STAGES, NT, kDiscMaxCPT>(p, nparts, st);
}

// ---- tensor map for xs viewed as (pixel HW, channel C, sample B), box = (32 pixels, rows_box channels, 1 sample) -----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            f = nullptr;
        }
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}

static bool make_feature_tmap(const float* xs, int B, int C, int HW, int rows_box, CUtensorMap* out) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    const cuuint64_t gdim[3] = {(cuuint64_t)HW, (cuuint64_t)C, (cuuint64_t)B};
    const cuuint64_t gstr[2] = {(cuuint64_t)HW * sizeof(float), (cuuint64_t)C * HW * sizeof(float)};
    const cuuint32_t box[3] = {32u, (cuuint32_t)rows_box, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    return enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(xs), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int K, int STAGES, int NT, int CPT, int MB = 2>
static int launch_disc_tma_c(DiscParams& p, const CUtensorMap& tmap, int* nparts, cudaStream_t st) {
    const size_t smem = disc_smem_tma<K, NT>(p.C, p.rows_box * p.nbox, STAGES);
    auto kern = disc_fused_kernel<K, 32, STAGES, NT, true, CPT, MB>;
    int occ = 0;
    { const int rc = kernel_occupancy(reinterpret_cast<const void*>(kern), NT, smem, &occ); if (rc != CLR_OK) return rc; }
    if (occ < 1) return CLR_ERR_UNSUPPORTED;
    const int grid = finish_launch_geometry(p, 32, occ, nparts);
    if (grid < 1) return CLR_ERR_UNSUPPORTED;
    clr::launch_k(kern, grid, NT, smem, st, p, tmap);
    return launch_status();
}
template <int K, int STAGES, int NT>
static int launch_disc_tma_s(DiscParams& p, const CUtensorMap& tmap, int* nparts, cudaStream_t st) {
    return p.C <= kDiscSmallCPT * kDiscCols ? launch_disc_tma_c<K, STAGES, NT, kDiscSmallCPT>(p, tmap, nparts, st)
                                            : launch_disc_tma_c<K, STAGES, NT, kDiscMaxCPT>(p, tmap, nparts, st);
}

// Tensor-map path: two CTAs per SM with a 3-stage (else 2-stage) ring; 512-thread CTAs on request when K <= 2.
template <int K>
static int launch_disc_tma(DiscParams& p, int* nparts, cudaStream_t st) {
    if (p.HW < 32 || p.HW % 4 != 0 || !aligned16(p.xs)) return CLR_ERR_UNSUPPORTED;
    p.nbox = (p.C + 255) / 256;
    p.rows_box = (((p.C + p.nbox - 1) / p.nbox) + 7) & ~7;
    CUtensorMap tmap;
    if (!make_feature_tmap(p.xs, p.B, p.C, p.HW, p.rows_box, &tmap)) return CLR_ERR_UNSUPPORTED;
    const size_t budget = (size_t)device_facts().max_smem_optin;
    const size_t per_cta = (budget - 2048) / 2;
    const int rt = p.rows_box * p.nbox;
    constexpr bool wide = (K <= 2);
    if constexpr (wide) {
        if (tunables().disc_threads == 512) {
            if (disc_smem_tma<K, 512>(p.C, rt, 3) <= per_cta) return launch_disc_tma_s<K, 3, 512>(p, tmap, nparts, st);
            if (disc_smem_tma<K, 512>(p.C, rt, 2) <= per_cta) return launch_disc_tma_s<K, 2, 512>(p, tmap, nparts, st);
        }
        // "disc_ctas" = 3: THREE resident CTAs per SM on a 2-stage ring (<= 85 registers), C <= 320: the kernel is bound by
        // its per-tile latency chain, not by DRAM (profiles/r02_l2_harvest.md) -- a third CTA to interleave
        if (tunables().disc_ctas == 3 && p.C <= kDiscSmallCPT * kDiscCols && disc_smem_tma<K, 256>(p.C, rt, 2) <= (budget - 3072) / 3)
            return launch_disc_tma_c<K, 2, 256, kDiscSmallCPT, 3>(p, tmap, nparts, st);
    }
    if (disc_smem_tma<K, 256>(p.C, rt, 3) <= per_cta) return launch_disc_tma_s<K, 3, 256>(p, tmap, nparts, st);
    if (disc_smem_tma<K, 256>(p.C, rt, 2) <= per_cta) return launch_disc_tma_s<K, 2, 256>(p, tmap, nparts, st);
    if (disc_smem_tma<K, 256>(p.C, rt, 2) <= budget) return launch_disc_tma_s<K, 2, 256>(p, tmap, nparts, st);
    return CLR_ERR_UNSUPPORTED;
}

template <int K, int TP>
static int launch_disc(DiscParams& p, int* nparts, cudaStream_t st) {
    const size_t budget = (size_t)device_facts().max_smem_optin;
    // TP = 32 aims at two resident CTAs per SM (half the budget each)
    const size_t per_cta = (TP == 32) ? (budget - 2048) / 2 : budget;
    // 512-thread CTAs (32 warps per SM at two CTAs) when the accumulators fit the 64-register budget
    constexpr bool wide = (K <= 2);
    if constexpr (wide) {
        if (tunables().disc_threads == 512) {
            if (disc_smem<K, TP, 512>(p.C, 3) <= per_cta) return launch_disc_s<K, TP, 3, 512>(p, nparts, st);
            if (disc_smem<K, TP, 512>(p.C, 2) <= budget) return launch_disc_s<K, TP, 2, 512>(p, nparts, st);
        }
    }
    if (disc_smem<K, TP, 256>(p.C, 4) <= per_cta) return launch_disc_s<K, TP, 4, 256>(p, nparts, st);
    if (disc_smem<K, TP, 256>(p.C, 3) <= per_cta) return launch_disc_s<K, TP, 3, 256>(p, nparts, st);
    if (disc_smem<K, TP, 256>(p.C, 2) <= budget) return launch_disc_s<K, TP, 2, 256>(p, nparts, st);
    return CLR_ERR_UNSUPPORTED;
}

static int dispatch_disc_tma(int K, DiscParams& p, int* nparts, cudaStream_t st) {
    switch (K) {
        case 1: return launch_disc_tma<1>(p, nparts, st);
        case 2: return launch_disc_tma<2>(p, nparts, st);
        case 3: return launch_disc_tma<3>(p, nparts, st);
        case 4: return launch_disc_tma<4>(p, nparts, st);
        case 5: return launch_disc_tma<5>(p, nparts, st);
        case 6: return launch_disc_tma<6>(p, nparts, st);
        case 7: return launch_disc_tma<7>(p, nparts, st);
        case 8: return launch_disc_tma<8>(p, nparts, st);
    }
    return CLR_ERR_UNSUPPORTED;
}

template <int TP>
static int dispatch_disc_k(int K, DiscParams& p, int* nparts, cudaStream_t st) {
    switch (K) {
        case 1: return launch_disc<1, TP>(p, nparts, st);
        case 2: return launch_disc<2, TP>(p, nparts, st);
        case 3: return launch_disc<3, TP>(p, nparts, st);
        case 4: return launch_disc<4, TP>(p, nparts, st);
        case 5: return launch_disc<5, TP>(p, nparts, st);
        case 6: return launch_disc<6, TP>(p, nparts, st);
        case 7: return launch_disc<7, TP>(p, nparts, st);
        case 8: return launch_disc<8, TP>(p, nparts, st);
    }
    return CLR_ERR_UNSUPPORTED;
}

// Returns CLR_ERR_UNSUPPORTED when the tile does not fit (very wide C) or the planes are not 16-byte
// friendly; the caller then falls back to the two-pass form (clr_disc_fwd + clr_pool_rows_fwd).
int disc_fused_impl(const float* xs, const float* ys, int B, int C, int HW, int K,
                    const float* disc_vec, const float* disc_beta, float margin,
                    float* coef, float* delta, float* partial, float* hinge, int* nparts, cudaStream_t st,
                    const DiscFlagDep* dep) {
    CLR_CHECK_ARG(xs && ys && disc_vec && disc_beta && coef && partial && hinge && nparts && *nparts > 0);
    CLR_CHECK_ARG(B > 0 && C > 0 && HW > 0 && K >= 1 && K <= CLR_MAX_K);
    if (HW % 4 != 0 || !aligned16(xs) || C > kDiscMaxCPT * kDiscCols) return CLR_ERR_UNSUPPORTED;
    if (((C + kDiscCols - 1) / kDiscCols) * K > kDiscMaxAcc) return CLR_ERR_UNSUPPORTED;
    DiscParams p{};
    p.xs = xs; p.ys = ys; p.V = disc_vec; p.beta = disc_beta; p.coef = coef; p.delta = delta;
    p.partial = partial; p.hinge = hinge; p.alpha = -2.0f / (float)C; p.margin = margin;
    p.B = B; p.C = C; p.HW = HW;
    if (dep) {
        p.wait_fin = dep->wait_fin; p.wait_all = dep->wait_all; p.wait_fin_n = dep->wait_fin_n; p.wait_all_n = dep->wait_all_n; p.err = dep->err;
        p.beta_partial = dep->beta_partial; p.beta_nparts = (int)dep->wait_fin_n;
    }
    // "disc_tile" tunable: 0 auto (32-pixel tiles, two CTAs per SM), 64 = 64-pixel tiles, one CTA per SM
    int n = *nparts;
    int rc = CLR_ERR_UNSUPPORTED;
    // "disc_impl": 0 auto = tensor-map TMA tiles, 2 = cp.async (LDGSTS) tiles, (1 = two-pass form, handled by the caller)
    if (tunables().disc_impl != 2 && tunables().disc_tile != 64) rc = dispatch_disc_tma(K, p, &n, st);
    if (rc == CLR_OK) { *nparts = n; return rc; }
    if (rc != CLR_ERR_UNSUPPORTED) return rc;
    n = *nparts;
    if (tunables().disc_tile == 64) rc = dispatch_disc_k<64>(K, p, &n, st);
    if (rc == CLR_ERR_UNSUPPORTED) { n = *nparts; rc = dispatch_disc_k<32>(K, p, &n, st); }
    if (rc == CLR_ERR_UNSUPPORTED && tunables().disc_tile != 64) { n = *nparts; rc = dispatch_disc_k<64>(K, p, &n, st); }
    if (rc == CLR_OK) *nparts = n;
    return rc;
}

}  // namespace clr

extern "C" {

size_t clr_disc_fused_ws_bytes(int C, int K) {
    if (C < 1 || K < 1 || K > CLR_MAX_K) return 0;
    return sizeof(float) * ((size_t)clr::kDiscMaxParts * K * (C + 1) + clr::kDiscMaxParts);
}

int clr_disc_fused_fwd(const float* xs, const float* ys, int B, int C, int HW, int K,
                       const float* disc_vec, const float* disc_beta, float margin,
                       float* coef, float* delta, void* ws, size_t ws_bytes, float* packed2, clr_stream_t stream) {
    if (!ws || !packed2) return CLR_ERR_BAD_ARG;
    if (ws_bytes < clr_disc_fused_ws_bytes(C, K)) return CLR_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* partial = static_cast<float*>(ws);
    float* hinge = partial + (size_t)clr::kDiscMaxParts * K * (C + 1);
    int nparts = clr::kDiscMaxParts;
    int rc = clr::disc_fused_impl(xs, ys, B, C, HW, K, disc_vec, disc_beta, margin, coef, delta, partial, hinge, &nparts, st);
    if (rc != CLR_OK) return rc;
    clr::launch_partial_reduce(partial, nparts, K, C, packed2, st);
    clr::launch_step_pack(hinge, nparts, 1, nullptr, 0, packed2 + (size_t)K * (C + 1), st);
    return clr::launch_status();
}

}  // extern "C"

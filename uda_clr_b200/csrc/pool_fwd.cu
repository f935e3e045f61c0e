// clr_pool_fwd: class-wise weighted pooling  S[r][c] = sum_{b,p} x[b,c,p] * w_r[b,p],  N[r] = sum w_r.
//
// Replaces the 4 materialised [B,C,H,W] products + 8 reductions of utils/Utils.py:114-126 (and
// :212-223 for the retrify weights) with ONE read of the feature map.
//
// Bound: HBM.  Algorithmic bytes per domain = 4*B*C*HW (features) + 4*B*WP*HW (weight planes).
//
// Layout / decomposition
//   item   = (domain, sample b, pixel chunk of PX pixels, group of CG channels)  -> 4*PX*CG bytes
//   grid   = persistent: (#SM x resident CTAs), each CTA owns a contiguous range of items, so a
//            CTA walks the channel groups of one (b, chunk) before moving on and re-stages the
//            chunk's weight planes in shared memory only when (b, chunk) changes.
//   thread = owns VEC*REPS fixed pixels of the chunk; per item it streams CG channel rows with
//            128-bit non-allocating loads (CG*REPS independent loads in flight), keeps CG*2K fp32
//            accumulators, then the warp does one transposing butterfly (31 shuffles) and the
//            8 warps combine through shared memory.
//   output = per-(b,chunk) partials [R][C+1] (column C = weight sums), combined across (b,chunk)
//            in fp64 and in a fixed order by pool_reduce_kernel -> bit-stable run to run.
#include "clr_common.cuh"

namespace clr {

struct PoolDom {
    const float* feat;
    const float* w;
    float* partial;   // [B*nChunk][R][C+1]
    float* sums;      // [R][C+1]
    int B;
    int fmt;
    int items;        // B*nChunk*nGroup
    int slots;        // B*nChunk
};

struct PoolParams {
    PoolDom dom[2];
    int ndom;
    int C, HW;
    int nChunk, nGroup;
    int total;
};

template <int K, int VEC>
struct PoolCfg {
    static constexpr int R = 2 * K;
    static constexpr int CG = 32 / R;                       // channels per item (accumulators <= 32)
    static constexpr int REPS = (K <= 2) ? 2 : 1;           // pixels per thread = VEC*REPS
    static constexpr int PX = kThreads * VEC * REPS;        // pixels per chunk
    static constexpr int SMEM_W = 2 * K * PX;               // explicit planes worst case (floats)
    static constexpr int SMEM_RED = 2 * kWarps * 32;        // double-buffered cross-warp scratch
    static constexpr size_t SMEM_BYTES = sizeof(float) * (SMEM_W + SMEM_RED);
};

template <int K, int VEC>
__global__ void __launch_bounds__(kThreads, 2) pool_fwd_kernel(const PoolParams p) {
    using Cfg = PoolCfg<K, VEC>;
    constexpr int R = Cfg::R, CG = Cfg::CG, REPS = Cfg::REPS, PX = Cfg::PX;
    extern __shared__ __align__(16) float smem[];
    float* wsm = smem;                         // [R][PX]  (obj rows then bck rows, already complemented)
    float* red = smem + Cfg::SMEM_W;           // [2][kWarps][32]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int begin, end;
    partition(p.total, gridDim.x, blockIdx.x, begin, end);

    int cur_slot_key = -1;
    int parity = 0;
    for (int it = begin; it < end; ++it) {
        const int d = (p.ndom > 1 && it >= p.dom[0].items) ? 1 : 0;
        const PoolDom& D = p.dom[d];
        const int local = it - (d ? p.dom[0].items : 0);
        const int slot = local / p.nGroup, grp = local - slot * p.nGroup;
        const int b = slot / p.nChunk, chunk = slot - b * p.nChunk;
        const int px0 = chunk * PX;
        const int slot_key = d * 0x40000000 + slot;

        if (slot_key != cur_slot_key) {
            // ---- stage this (b, chunk)'s 2K weight rows in shared memory -------------------------
            __syncthreads();   // previous item's readers of wsm are done
            cur_slot_key = slot_key;
            const int WP = (D.fmt == CLR_W_COMPLEMENT) ? K : R;
            const float* wb = D.w + (size_t)b * WP * p.HW;
#pragma unroll
            for (int rep = 0; rep < REPS; ++rep) {
                const int off = (rep * kThreads + tid) * VEC;
                const bool ok = px0 + off < p.HW;   // VEC-granular: HW % VEC == 0 on the VEC=4 path
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    Pack<VEC> wo, wb_;
                    if (ok) {
                        wo = ld_keep<VEC>(wb + (size_t)k * p.HW + px0 + off);
                        if (D.fmt == CLR_W_COMPLEMENT) {
#pragma unroll
                            for (int v = 0; v < VEC; ++v) wb_.v[v] = 1.0f - wo.v[v];   // utils/Utils.py:111-112
                        } else {
                            wb_ = ld_keep<VEC>(wb + (size_t)(K + k) * p.HW + px0 + off);
                        }
                    } else {
#pragma unroll
                        for (int v = 0; v < VEC; ++v) { wo.v[v] = 0.f; wb_.v[v] = 0.f; }
                    }
                    st_keep<VEC>(wsm + k * PX + off, wo);
                    st_keep<VEC>(wsm + (K + k) * PX + off, wb_);
                }
            }
            __syncthreads();
        }

        // ---- accumulate CG channels x R rows over this thread's pixels ------------------------------
        float acc[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = 0.f;
        const int c0 = grp * CG;
        const float* xb = D.feat + ((size_t)b * p.C + c0) * p.HW + px0;
#pragma unroll
        for (int rep = 0; rep < REPS; ++rep) {
            const int off = (rep * kThreads + tid) * VEC;
            const bool ok = px0 + off < p.HW;
            Pack<VEC> x[CG];
#pragma unroll
            for (int j = 0; j < CG; ++j) {
                if (ok && c0 + j < p.C) x[j] = ld_stream<VEC>(xb + (size_t)j * p.HW + off);
                else {
#pragma unroll
                    for (int v = 0; v < VEC; ++v) x[j].v[v] = 0.f;
                }
            }
            Pack<VEC> w[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if constexpr (VEC == 4) {
                    const float4 t = *reinterpret_cast<const float4*>(wsm + r * PX + off);
                    w[r].v[0] = t.x; w[r].v[1] = t.y; w[r].v[2] = t.z; w[r].v[3] = t.w;
                } else {
                    w[r].v[0] = wsm[r * PX + off];
                }
            }
#pragma unroll
            for (int j = 0; j < CG; ++j)
#pragma unroll
                for (int r = 0; r < R; ++r)
#pragma unroll
                    for (int v = 0; v < VEC; ++v)
                        acc[j * R + r] = fmaf(x[j].v[v], w[r].v[v], acc[j * R + r]);
        }
        // the item that owns channel group 0 also sums this chunk's weights (padding is staged as 0)
        float nsum[R];
        if (grp == 0) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float s = 0.f;
#pragma unroll
                for (int rep = 0; rep < REPS; ++rep) {
                    const int off = (rep * kThreads + tid) * VEC;
#pragma unroll
                    for (int v = 0; v < VEC; ++v) s += wsm[r * PX + off + v];
                }
                nsum[r] = s;
            }
        }

        // ---- CTA reduction: butterfly inside the warp, shared memory across warps -------------------
        const float tot = warp_sum_transpose32(acc, lane);
        float* redp = red + parity * (kWarps * 32);
        redp[warp * 32 + lane] = tot;
        __syncthreads();
        float* out = D.partial + (size_t)slot * R * (p.C + 1);
        if (warp == 0) {
            float s = 0.f;
#pragma unroll
            for (int wq = 0; wq < kWarps; ++wq) s += redp[wq * 32 + lane];
            const int j = lane / R, r = lane - j * R;
            if (j < CG && c0 + j < p.C) out[(size_t)r * (p.C + 1) + c0 + j] = s;
        }
        parity ^= 1;

        if (grp == 0) {   // CTA-uniform branch
#pragma unroll
            for (int r = 0; r < R; ++r) nsum[r] = warp_sum(nsum[r]);
            float* redn = red + parity * (kWarps * 32);
            if (lane == 0) {
#pragma unroll
                for (int r = 0; r < R; ++r) redn[warp * 32 + r] = nsum[r];
            }
            __syncthreads();
            if (warp == 0 && lane < R) {
                float s = 0.f;
#pragma unroll
                for (int wq = 0; wq < kWarps; ++wq) s += redn[wq * 32 + lane];
                out[(size_t)lane * (p.C + 1) + p.C] = s;
            }
            parity ^= 1;
        }
    }
}

// sums[r][c] = sum over (b,chunk) slots of partial[slot][r][c], fp64, fixed order.
__global__ void __launch_bounds__(256) pool_reduce_kernel(const PoolParams p, int R) {
    const int d = blockIdx.y;
    const PoolDom& D = p.dom[d];
    const int n = R * (p.C + 1);
    // blockDim = (32 columns, 8 slot-lanes)
    const int col = blockIdx.x * 32 + threadIdx.x;
    __shared__ double sh[8][33];
    double s = 0.0;
    if (col < n)
        for (int sl = threadIdx.y; sl < D.slots; sl += 8) s += (double)D.partial[(size_t)sl * n + col];
    sh[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && col < n) {
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += sh[i][threadIdx.x];
        D.sums[col] = (float)t;
    }
}

__global__ void proto_finalize_kernel(const float* __restrict__ sums, int R, int C, float* __restrict__ mu) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R * C) return;
    const int r = i / C, c = i - r * C;
    mu[i] = sums[(size_t)r * (C + 1) + c] / sums[(size_t)r * (C + 1) + C];   // 0/0 -> NaN, as the reference
}

template <int K, int VEC>
static int launch_pool(const PoolParams& p, cudaStream_t st) {
    using Cfg = PoolCfg<K, VEC>;
    static int occ_cache = 0;   // benign race
    auto kern = pool_fwd_kernel<K, VEC>;
    if (occ_cache == 0) {
        CLR_RETURN_IF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES));
        int occ = 0;
        CLR_RETURN_IF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, Cfg::SMEM_BYTES));
        occ_cache = occ > 0 ? occ : 1;
    }
    int grid = device_facts().sms * occ_cache;
    if (grid > p.total) grid = p.total;
    kern<<<grid, kThreads, Cfg::SMEM_BYTES, st>>>(p);
    const int n = 2 * K * (p.C + 1);
    pool_reduce_kernel<<<dim3((n + 31) / 32, p.ndom), dim3(32, 8), 0, st>>>(p, 2 * K);
    return launch_status();
}

template <int VEC>
static int dispatch_k(int K, const PoolParams& p, cudaStream_t st) {
    switch (K) {
        case 1: return launch_pool<1, VEC>(p, st);
        case 2: return launch_pool<2, VEC>(p, st);
        case 3: return launch_pool<3, VEC>(p, st);
        case 4: return launch_pool<4, VEC>(p, st);
        case 5: return launch_pool<5, VEC>(p, st);
        case 6: return launch_pool<6, VEC>(p, st);
        case 7: return launch_pool<7, VEC>(p, st);
        case 8: return launch_pool<8, VEC>(p, st);
    }
    return CLR_ERR_UNSUPPORTED;
}

static int chunk_px(int K, int vec) { return kThreads * vec * (K <= 2 ? 2 : 1); }

// Workspace: partials for up to two domains, sized for the scalar (smallest-chunk) path.
static size_t partial_floats(int B, int C, int HW, int K) {
    const int px = chunk_px(K, 1);
    const size_t slots = (size_t)B * ((HW + px - 1) / px);
    return slots * 2 * K * (C + 1);
}

int pool_fwd_impl(const float* feat0, const float* w0, int fmt0, int B0, float* sums0,
                  const float* feat1, const float* w1, int fmt1, int B1, float* sums1,
                  int C, int HW, int K, void* ws, size_t ws_bytes, cudaStream_t st) {
    const int ndom = feat1 ? 2 : 1;
    CLR_CHECK_ARG(feat0 && w0 && sums0 && ws && B0 > 0 && C > 0 && HW > 0 && K >= 1 && K <= CLR_MAX_K);
    CLR_CHECK_ARG(fmt0 == CLR_W_COMPLEMENT || fmt0 == CLR_W_EXPLICIT);
    if (ndom == 2) CLR_CHECK_ARG(w1 && sums1 && B1 > 0 && (fmt1 == CLR_W_COMPLEMENT || fmt1 == CLR_W_EXPLICIT));
    if (!aligned4(feat0) || !aligned4(w0) || (feat1 && (!aligned4(feat1) || !aligned4(w1)))) return CLR_ERR_ALIGN;
    // the partition index is an int; the per-element offsets are size_t
    bool vec4 = (HW % 4 == 0) && aligned16(feat0) && aligned16(w0) && (!feat1 || (aligned16(feat1) && aligned16(w1)));
    const int vec = vec4 ? 4 : 1;
    const int px = chunk_px(K, vec);
    const int CG = 32 / (2 * K);

    PoolParams p{};
    p.ndom = ndom; p.C = C; p.HW = HW;
    p.nChunk = (HW + px - 1) / px;
    p.nGroup = (C + CG - 1) / CG;
    const size_t need = sizeof(float) * (partial_floats(B0, C, HW, K) + (ndom == 2 ? partial_floats(B1, C, HW, K) : 0));
    if (ws_bytes < need) return CLR_ERR_WORKSPACE;
    float* wsf = static_cast<float*>(ws);
    const long long items0 = (long long)B0 * p.nChunk * p.nGroup;
    const long long items1 = ndom == 2 ? (long long)B1 * p.nChunk * p.nGroup : 0;
    if (items0 + items1 > 0x3fffffff) return CLR_ERR_UNSUPPORTED;
    p.dom[0] = PoolDom{feat0, w0, wsf, sums0, B0, fmt0, (int)items0, B0 * p.nChunk};
    if (ndom == 2)
        p.dom[1] = PoolDom{feat1, w1, wsf + partial_floats(B0, C, HW, K), sums1, B1, fmt1, (int)items1, B1 * p.nChunk};
    p.total = (int)(items0 + items1);
    return vec4 ? dispatch_k<4>(K, p, st) : dispatch_k<1>(K, p, st);
}

}  // namespace clr

extern "C" {

size_t clr_pool_ws_bytes(int B, int C, int HW, int K) {
    if (B <= 0 || C <= 0 || HW <= 0 || K < 1 || K > CLR_MAX_K) return 0;
    return sizeof(float) * clr::partial_floats(B, C, HW, K);
}

int clr_pool_fwd(const float* feat, const float* w, int fmt, int B, int C, int HW, int K,
                 void* ws, size_t ws_bytes, float* sums, clr_stream_t stream) {
    return clr::pool_fwd_impl(feat, w, fmt, B, sums, nullptr, nullptr, 0, 0, nullptr, C, HW, K, ws, ws_bytes,
                              static_cast<cudaStream_t>(stream));
}

int clr_pool_fwd2(const float* feat0, const float* w0, int fmt0, int B0,
                  const float* feat1, const float* w1, int fmt1, int B1,
                  int C, int HW, int K, void* ws, size_t ws_bytes,
                  float* sums0, float* sums1, clr_stream_t stream) {
    if (!feat1) return CLR_ERR_BAD_ARG;
    return clr::pool_fwd_impl(feat0, w0, fmt0, B0, sums0, feat1, w1, fmt1, B1, sums1, C, HW, K, ws, ws_bytes,
                              static_cast<cudaStream_t>(stream));
}

int clr_proto_finalize(const float* sums, int R, int C, float* mu, clr_stream_t stream) {
    if (!sums || !mu || R <= 0 || C <= 0) return CLR_ERR_BAD_ARG;
    const int n = R * C;
    clr::proto_finalize_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(sums, R, C, mu);
    return clr::launch_status();
}

}  // extern "C"

// Shared device/host helpers for the CLR kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "clr_b200.h"

namespace clr {

constexpr int kThreads = 256;           // CTA size of every streaming kernel
constexpr int kWarps = kThreads / 32;

#define CLR_CHECK_ARG(cond) do { if (!(cond)) return CLR_ERR_BAD_ARG; } while (0)
#define CLR_RETURN_IF_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return CLR_ERR_CUDA_BASE - (int)_e; } while (0)

static inline int launch_status() {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) { cudaGetLastError(); return CLR_ERR_CUDA_BASE - (int)e; }
    return CLR_OK;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline bool aligned4(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 3u) == 0; }

// Per-device facts, cached per device id (read-mostly; benign race: every writer stores the same values).
struct DeviceFacts { int sms; int l2; int max_smem_optin; };
const DeviceFacts& device_facts();

// Process-wide kernel-selection knobs for benchmarking (clr_set_tunable); defaults pick the fastest path.  This is the
// library's only mutable global state besides the launch counter and the profiling trace pointer: plain ints that a
// launch reads once; changing one while another thread is launching is not synchronised (bench / test use only).
struct Tunables {
    int pool_impl;     // 0 = auto (128-bit LDG kernel, 2 CTAs/SM: fastest in the live step), 1 = force LDG, 2 = force the TMA ring
    int pool_stages;   // TMA ring depth (0 = auto)
    int pool_pair;     // LDG pooling kernel, even R >= 8: 0 = packed FFMA2 inner loop, 2 = scalar FMAs (A/B)
    int pool_threads;  // LDG pooling kernel CTA size: 0 = auto (128 for R > 8, else 256), 128 / 256 = forced
    int disc_impl;     // 0 = auto (one-read fused kernel, tensor-map TMA tiles), 1 = two-pass form, 2 = one-read kernel with cp.async tiles
    int disc_threads;  // 0 = auto (256-thread CTAs: fastest in the live step), 512 = 512-thread CTAs when K <= 2
    int disc_ctas;     // 3 = fused discriminative kernel with three CTAs per SM on a 2-stage ring (K <= 2, C <= 320)
    int disc_tile;     // 0 = auto, 64 = force 64-pixel tiles in the fused discriminative kernel
    int pdl_off;       // 1 = do not use programmatic dependent launch
    int mc_precise;    // 1 = clr_mc_stats / clr_mc_retrify evaluate std / mean exactly like ATen's CUDA reductions (slow; tests), 0 = streaming
    int finish_off;    // 1 = single-GPU step uses the separate reduce / finalize kernels instead of the merged finish kernels
    int hfuse_off;     // 1 = finish bodies get launches of their own instead of riding with cons / the target-gradient write
    int fin_early_off; // 1 = the pooling finish releases the discriminative kernel only at its very end (after the last-CTA combine)
    int bwd_merge_off; // 1 = clr_step_run writes the two gradient maps with two launches ([finish | xt], then xs) instead of one
    int mc_all_rows;   // 1 = the fused step's MC statistics write the full-resolution mean map on every row (A/B; default: bilinear source rows only)
    int mc_generic;    // 1 = clr_mc_stats does not use the T == 8 specialisation (A/B runs)
    int mc_fuse;       // 1 = the fused step uses the one-pass mc_retrify kernel instead of mc_stats + retrify_weights
    int mc_split;      // 1 = one-pass mc_retrify kernel splits image rows into column blocks (more, smaller CTAs)
    int flag_dep_off;  // 1 = the discriminative kernel waits for the whole [finish | consistency] grid (griddepcontrol.wait)
                       //     instead of the finish CTAs' completion counter
    int disc_reverse;  // 1 = the one-read discriminative kernel walks its tiles in descending address order (round 2: the re-read of
                       //     xs then misses DRAM for 102 instead of 110 of 135 MB, the kernel is not faster -- it is not DRAM-bound)
    int sched;         // fused step launch schedule (step.cu): 0 = auto (schedule 2 from 8 ranks on, else 1), 1, 2
    int dfin_split;    // clr_step_run: disc finish as its own launch in front of a no-wait gated backward: 0 = auto (with schedule 2), 1, 2 = never
    int xchg_pull;     // in-kernel exchange: 1 = readers poll the peers' buffers (no remote stores), 0 = senders push
    void* trace_buf;   // device TraceRec[kTraceSlots] or NULL (clr_trace_set): device-side timeline of the kernels
};
Tunables& tunables();

// Programmatic dependent launch: every kernel of the library starts with pdl_wait() and is launched through
// launch_k(), which sets cudaLaunchAttributeProgrammaticStreamSerialization so that the NEXT kernel's launch
// overlaps this one's tail; griddepcontrol.wait then blocks until the predecessor grid has completed and its
// writes are visible.  ("pdl" tunable = 0 turns the attribute off; the wait is then a no-op.)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Process-wide count of kernel launches issued by this library (clr_launch_count); relaxed atomic.
void count_launch();

// ---- device-side timeline (profiling aid; inert unless clr_trace_set installed a buffer) ---------------------
// Every kernel stamps %globaltimer into its slot: t_first = earliest CTA start (before griddepcontrol.wait),
// t_ready = earliest return from griddepcontrol.wait (predecessor complete), t_last = latest CTA exit.  This is the
// only way to see the live step's overlap / gaps without a timeline profiler (no nsys in the image); ncu
// serialises kernels and event records break programmatic dependent launch.
struct TraceRec { unsigned long long t_first, t_ready, t_last, n_cta; };
enum TraceId { TR_MC_STATS = 0, TR_RETRIFY, TR_POOL, TR_POOL_REDUCE, TR_ALIGN, TR_CONS, TR_DISC, TR_DISC_REDUCE,
               TR_DISC_FIN, TR_BWD_T, TR_BWD_S, TR_BWD_BOTH, TR_CONS_BWD, TR_PACK, TR_OTHER, TR_POOL_T /*schedule 2: target pooling*/, TR_DBG0 = 16, TR_DBG1, TR_DBG2, TR_DBG3, TR_DBG4, TR_DBG5,
               TR_DBG6, TR_FIN_S /*schedule 2: source half of the pooling finish*/, kTraceSlots = 24 };
static __device__ TraceRec* g_trace_dev = nullptr;     // one copy per translation unit, installed by launch_k
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ bool trace_leader() { return threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0; }
__device__ __forceinline__ void trace_enter(int id) {
    TraceRec* t = g_trace_dev;
    if (t && trace_leader()) { atomicMin(&t[id].t_first, global_ns()); atomicAdd(&t[id].n_cta, 1ull); }
}
__device__ __forceinline__ void trace_ready(int id) {
    TraceRec* t = g_trace_dev;
    if (t && trace_leader()) atomicMin(&t[id].t_ready, global_ns());
}
__device__ __forceinline__ void trace_exit(int id) {
    TraceRec* t = g_trace_dev;
    if (t && trace_leader()) atomicMax(&t[id].t_last, global_ns());
}
// phase marks for ad-hoc instrumentation of one kernel (slots TR_DBG*): first / last time any CTA passed the mark
__device__ __forceinline__ void trace_mark(int id) {
    TraceRec* t = g_trace_dev;
    if (t && trace_leader()) { const unsigned long long now = global_ns(); atomicMin(&t[id].t_first, now); atomicMin(&t[id].t_ready, now);
                               atomicMax(&t[id].t_last, now); atomicAdd(&t[id].n_cta, 1ull); }
}
// ---- flag dependencies (finer than grid completion) ---------------------------------------------------------------
// A kernel whose consumer needs only what a FEW of its CTAs produce (the finish CTAs riding in front of a streaming
// launch) publishes a counter: every thread fences, the CTA synchronises, one thread increments.  The consumer skips
// griddepcontrol.wait, spins on the counter with one thread, and from then on reads the produced data through
// coherent loads only.  This removes the grid-completion latency -- notably the system-scope flush a grid pays at
// its end once it has touched peer memory (the in-kernel exchange) -- from the step's critical path.
__device__ __forceinline__ void cta_signal(unsigned int* counter_a, unsigned int* counter_b) {
    __threadfence();
    __syncthreads();
    if (trace_leader()) {
        if (counter_a) atomicAdd(counter_a, 1u);
        if (counter_b) atomicAdd(counter_b, 1u);
    }
}
// one thread: returns false on timeout (~2 s)
__device__ __forceinline__ bool spin_until_at_least(const unsigned int* counter, unsigned int expected) {
    const volatile unsigned int* c = counter;
    if (*c < expected) {
        const long long t0 = clock64();
        while (*c < expected)
            if (clock64() - t0 > 4000000000LL) return false;
    }
    __threadfence();
    return true;
}

// kernel prologue: stamp; let the NEXT kernel of the stream start launching right away (its CTAs become resident as
// ours retire and park in their own griddepcontrol.wait, so launch latency and CTA ramp-up leave the critical path --
// the wait still blocks until this whole grid has completed and flushed, so ordering is unchanged); wait for the
// predecessor grid; stamp
__device__ __forceinline__ void kernel_begin(int id) { trace_enter(id); pdl_trigger(); pdl_wait(); trace_ready(id); }
// The same with the trigger AFTER the wait, for the producers of a flag dependency: their consumer skips
// griddepcontrol.wait and trusts completion counters that the kernel BEFORE the producer zeroes, so it must not be
// launched before that kernel has completed.  (With the early trigger a chain of small grids can become resident all at
// once -- e.g. a CUDA-graph replay of a small step -- and the consumer would read the previous step's counters, still at
// their final values, before the reset.)  Triggering after the wait makes "reset -> producer started -> consumer
// launched" a happens-before chain; the consumer still overlaps the producer's whole body.
__device__ __forceinline__ void kernel_begin_late_trigger(int id) { trace_enter(id); pdl_wait(); pdl_trigger(); trace_ready(id); }

// ---- streaming loads / stores --------------------------------------------------------------------
// Feature maps are touched exactly once per pass: read through the non-coherent path without
// allocating in L1; gradients are written with an evict-first hint.
template <int VEC> struct Pack { float v[VEC]; };

template <int VEC>
__device__ __forceinline__ Pack<VEC> ld_stream(const float* p);
template <>
__device__ __forceinline__ Pack<4> ld_stream<4>(const float* p) {
    Pack<4> r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]) : "l"(p));
    return r;
}
template <>
__device__ __forceinline__ Pack<1> ld_stream<1>(const float* p) {
    Pack<1> r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r.v[0]) : "l"(p));
    return r;
}

// The same with an explicit L2 policy (createpolicy.*): e.g. evict-first for a stream that must not displace lines a later
// kernel of the step re-reads from L2.
template <int VEC>
__device__ __forceinline__ Pack<VEC> ld_stream_hint(const float* p, uint64_t pol);
template <>
__device__ __forceinline__ Pack<4> ld_stream_hint<4>(const float* p, uint64_t pol) {
    Pack<4> r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]) : "l"(p), "l"(pol));
    return r;
}
template <>
__device__ __forceinline__ Pack<1> ld_stream_hint<1>(const float* p, uint64_t pol) {
    Pack<1> r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(r.v[0]) : "l"(p), "l"(pol));
    return r;
}

// Small, re-read planes (labels / weights): ordinary read-only loads so they stay in L1/L2.
template <int VEC>
__device__ __forceinline__ Pack<VEC> ld_keep(const float* p);
template <>
__device__ __forceinline__ Pack<4> ld_keep<4>(const float* p) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    Pack<4> r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; return r;
}
template <>
__device__ __forceinline__ Pack<1> ld_keep<1>(const float* p) {
    Pack<1> r; r.v[0] = __ldg(p); return r;
}

template <int VEC>
__device__ __forceinline__ void st_stream(float* p, const Pack<VEC>& r);
template <>
__device__ __forceinline__ void st_stream<4>(float* p, const Pack<4>& r) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(r.v[0], r.v[1], r.v[2], r.v[3]));
}
template <>
__device__ __forceinline__ void st_stream<1>(float* p, const Pack<1>& r) { __stcs(p, r.v[0]); }

template <int VEC>
__device__ __forceinline__ void st_keep(float* p, const Pack<VEC>& r);
template <>
__device__ __forceinline__ void st_keep<4>(float* p, const Pack<4>& r) {
    *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
}
template <>
__device__ __forceinline__ void st_keep<1>(float* p, const Pack<1>& r) { *p = r.v[0]; }

// ---- warp reductions -----------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Transposing butterfly: every lane holds 32 partial values v[0..31]; afterwards lane i holds
// sum over lanes of v[i] (returned).  31 shuffles instead of 32*5.
__device__ __forceinline__ float warp_sum_transpose32(float (&v)[32], int lane) {
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool upper = (lane & half) != 0;
#pragma unroll
        for (int j = 0; j < half; ++j) {
            const float send = upper ? v[j] : v[j + half];
            const float keep = upper ? v[j + half] : v[j];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
    return v[0];
}

// ---- mbarrier + bulk async copy (TMA engine, 1-D form: cp.async.bulk -> SASS UBLKCP) -------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// make mbarrier initialisation visible to the async proxy before the first bulk copy targets it
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy; completion is signalled on `bar` as `bytes` of transaction count.
// dst/src 16-byte aligned, bytes a multiple of 16.  Streaming data: L2 evict-first policy.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// named barrier among a subset of the CTA's warps (id 1..15; 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace clr

#include <atomic>
#include <utility>
namespace clr {

template <typename... KArgs, typename... Args>
static inline void launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = tunables().pdl_off ? 0 : 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    {   // install / remove the trace buffer pointer in THIS translation unit's copy of g_trace_dev (stream-ordered).
        // Launches may come from several host threads (autograd runs backward on its own): the bookkeeping word is atomic,
        // and two threads that race merely install the same pointer twice.
        static std::atomic<void*> installed{nullptr};
        void* want = tunables().trace_buf;
        if (want != installed.load(std::memory_order_acquire)) {
            cudaMemcpyToSymbolAsync(g_trace_dev, &want, sizeof(want), 0, cudaMemcpyHostToDevice, st);
            installed.store(want, std::memory_order_release);
        }
    }
    count_launch();
    cudaLaunchKernelEx(&cfg, kern, KArgs(std::forward<Args>(args))...);
}

// cudaFuncSetAttribute(max dynamic smem) + cudaOccupancyMaxActiveBlocksPerMultiprocessor are constants of (kernel, device,
// block size, dynamic smem): asked once per thread and cached, instead of two driver round trips in front of every launch
// (the drop-in ops and the fused step are enqueued from the host every training step).
static inline int kernel_occupancy(const void* kern, int threads, size_t smem, int* occ) {
    struct Entry { const void* kern; int dev, threads; size_t smem; int occ; };
    constexpr int kCap = 32;
    static thread_local Entry cache[kCap];
    static thread_local int used = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    for (int i = 0; i < used; ++i)
        if (cache[i].kern == kern && cache[i].dev == dev && cache[i].threads == threads && cache[i].smem == smem) { *occ = cache[i].occ; return CLR_OK; }
    if (smem > 48 * 1024) CLR_RETURN_IF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int o = 0;
    CLR_RETURN_IF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, threads, smem));
    if (used < kCap) cache[used++] = Entry{kern, dev, threads, smem, o};
    *occ = o;
    return CLR_OK;
}

// Static contiguous partition of `total` items over `parts` workers.
__device__ __forceinline__ void partition(int total, int parts, int idx, int& begin, int& end) {
    const int per = total / parts, rem = total % parts;
    begin = idx * per + (idx < rem ? idx : rem);
    end = begin + per + (idx < rem ? 1 : 0);
}

}  // namespace clr

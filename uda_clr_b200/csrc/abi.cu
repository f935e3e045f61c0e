// Library-level entry points of libclr_b200.so: version, status strings, device facts.
#include "clr_common.cuh"

namespace clr {

const DeviceFacts& device_facts() {
    static DeviceFacts facts[64];
    static bool have[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!have[dev]) {
        DeviceFacts f{148, 0, 48 * 1024};
        cudaDeviceGetAttribute(&f.sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&f.l2, cudaDevAttrL2CacheSize, dev);
        cudaDeviceGetAttribute(&f.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        if (f.sms <= 0) f.sms = 148;
        facts[dev] = f;
        have[dev] = true;
    }
    return facts[dev];
}

static unsigned long long g_launches = 0;
void count_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }

Tunables& tunables() {
    static Tunables t{};
    return t;
}

}  // namespace clr

#include <string.h>

extern "C" {

int clr_set_tunable(const char* name, int value) {
    if (!name) return CLR_ERR_BAD_ARG;
    clr::Tunables& t = clr::tunables();
    if (!strcmp(name, "pool_impl")) t.pool_impl = value;
    else if (!strcmp(name, "pool_stages")) t.pool_stages = value;
    else if (!strcmp(name, "pool_threads")) t.pool_threads = value;
    else if (!strcmp(name, "pool_pair")) t.pool_pair = value;
    else if (!strcmp(name, "mc_precise")) t.mc_precise = value;
    else if (!strcmp(name, "disc_impl")) t.disc_impl = value;
    else if (!strcmp(name, "disc_tile")) t.disc_tile = value;
    else if (!strcmp(name, "disc_ctas")) t.disc_ctas = value;
    else if (!strcmp(name, "disc_threads")) t.disc_threads = value;
    else if (!strcmp(name, "pdl_off")) t.pdl_off = value;
    else if (!strcmp(name, "finish_off")) t.finish_off = value;
    else if (!strcmp(name, "hfuse_off")) t.hfuse_off = value;
    else if (!strcmp(name, "xchg_pull")) t.xchg_pull = value;
    else if (!strcmp(name, "flag_dep_off")) t.flag_dep_off = value;
    else if (!strcmp(name, "mc_fuse")) t.mc_fuse = value;
    else if (!strcmp(name, "mc_generic")) t.mc_generic = value;
    else if (!strcmp(name, "mc_all_rows")) t.mc_all_rows = value;
    else if (!strcmp(name, "bwd_merge_off")) t.bwd_merge_off = value;
    else if (!strcmp(name, "fin_early_off")) t.fin_early_off = value;
    else if (!strcmp(name, "mc_split")) t.mc_split = value;
    else if (!strcmp(name, "disc_reverse")) t.disc_reverse = value;
    else if (!strcmp(name, "sched")) t.sched = value;
    else if (!strcmp(name, "dfin_split")) t.dfin_split = value;

    else return CLR_ERR_BAD_ARG;
    return CLR_OK;
}

/* Device-side kernel timeline (clr_common.cuh TraceRec).  The buffer is allocated once per process and never freed,
 * so a stale pointer in a translation unit that has not launched since tracing was switched off stays harmless. */
static clr::TraceRec* g_trace_alloc = nullptr;
static int trace_reset() {
    clr::TraceRec init[clr::kTraceSlots];
    for (int i = 0; i < clr::kTraceSlots; ++i) init[i] = clr::TraceRec{~0ull, ~0ull, 0ull, 0ull};
    CLR_RETURN_IF_CUDA(cudaMemcpy(g_trace_alloc, init, sizeof(init), cudaMemcpyHostToDevice));
    return CLR_OK;
}
int clr_trace_enable(int on) {
    if (!on) { clr::tunables().trace_buf = nullptr; return CLR_OK; }
    if (!g_trace_alloc) CLR_RETURN_IF_CUDA(cudaMalloc(&g_trace_alloc, sizeof(clr::TraceRec) * clr::kTraceSlots));
    const int rc = trace_reset();
    if (rc != CLR_OK) return rc;
    clr::tunables().trace_buf = g_trace_alloc;
    return CLR_OK;
}
int clr_trace_slots(void) { return clr::kTraceSlots; }
const char* clr_trace_name(int slot) {
    static const char* names[clr::kTraceSlots] = {"mc_stats", "retrify_weights", "pool_fwd", "pool_reduce", "align_finalize",
        "cons_fwd", "disc_fused", "disc_reduce", "disc_finalize", "pool_bwd_target", "pool_bwd_source", "pool_bwd_both",
        "cons_bwd", "step_pack", "other", "pool_fwd_target", "dbg0", "dbg1", "dbg2", "dbg3", "dbg4", "dbg5", "dbg6", "finish_source"};
    return (slot >= 0 && slot < clr::kTraceSlots) ? names[slot] : "";
}
int clr_trace_read(unsigned long long* out_host) {
    if (!out_host || !g_trace_alloc) return CLR_ERR_BAD_ARG;
    CLR_RETURN_IF_CUDA(cudaDeviceSynchronize());
    CLR_RETURN_IF_CUDA(cudaMemcpy(out_host, g_trace_alloc, sizeof(clr::TraceRec) * clr::kTraceSlots, cudaMemcpyDeviceToHost));
    return trace_reset();
}

int clr_version(void) { return CLR_B200_VERSION; }

unsigned long long clr_launch_count(void) { return __atomic_load_n(&clr::g_launches, __ATOMIC_RELAXED); }

const char* clr_status_string(int status) {
    switch (status) {
        case CLR_OK: return "ok";
        case CLR_ERR_BAD_ARG: return "bad argument (null pointer, non-positive size or K out of range)";
        case CLR_ERR_ALIGN: return "pointer not aligned to 4 bytes";
        case CLR_ERR_WORKSPACE: return "workspace too small";
        case CLR_ERR_UNSUPPORTED: return "unsupported size or combination";
        default: break;
    }
    if (status <= CLR_ERR_CUDA_BASE) return cudaGetErrorString((cudaError_t)(CLR_ERR_CUDA_BASE - status));
    return "unknown status";
}

/* Timing events for harnesses that want the library to bracket its own launches (clr_step_args.ev_*). */
int clr_event_create(void** ev) {
    if (!ev) return CLR_ERR_BAD_ARG;
    cudaEvent_t e;
    CLR_RETURN_IF_CUDA(cudaEventCreate(&e));
    *ev = e;
    return CLR_OK;
}
int clr_event_destroy(void* ev) {
    if (!ev) return CLR_ERR_BAD_ARG;
    CLR_RETURN_IF_CUDA(cudaEventDestroy(static_cast<cudaEvent_t>(ev)));
    return CLR_OK;
}
int clr_event_elapsed_us(void* begin, void* end, float* us) {
    if (!begin || !end || !us) return CLR_ERR_BAD_ARG;
    float ms = 0.f;
    CLR_RETURN_IF_CUDA(cudaEventElapsedTime(&ms, static_cast<cudaEvent_t>(begin), static_cast<cudaEvent_t>(end)));
    *us = ms * 1000.f;
    return CLR_OK;
}

int clr_device_info(int* sm_count, int* l2_bytes) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); return CLR_ERR_UNSUPPORTED; }
    const clr::DeviceFacts& f = clr::device_facts();
    if (sm_count) *sm_count = f.sms;
    if (l2_bytes) *l2_bytes = f.l2;
    return CLR_OK;
}

}  // extern "C"

// Stand-alone HBM bandwidth probe for sizing expectations (not part of the library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/bwprobe tools/bwprobe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ float4 ldnc(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

template <int U>
__global__ void __launch_bounds__(256) read_ldg(const float4* __restrict__ src, size_t n4, float* __restrict__ out) {
    // contiguous chunk per CTA, threads stride inside
    const size_t per = (n4 + gridDim.x - 1) / gridDim.x;
    const size_t b = (size_t)blockIdx.x * per, e = min(b + per, n4);
    float acc = 0.f;
    size_t i = b + threadIdx.x;
    for (; i + (size_t)(U - 1) * 256 < e; i += (size_t)U * 256) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = ldnc(src + i + (size_t)u * 256);
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    for (; i < e; i += 256) { float4 v = ldnc(src + i); acc += v.x + v.y + v.z + v.w; }
    if (acc == 123.456f) out[blockIdx.x * 256 + threadIdx.x] = acc;
}

__global__ void __launch_bounds__(256) write_st(float4* __restrict__ dst, size_t n4) {
    const size_t per = (n4 + gridDim.x - 1) / gridDim.x;
    const size_t b = (size_t)blockIdx.x * per, e = min(b + per, n4);
    const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
    for (size_t i = b + threadIdx.x; i < e; i += 256) __stcs(dst + i, v);
}

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// TMA bulk ring: one producer thread, consumers only wait/arrive (pure fetch ceiling) and read one float4 per stage
template <int STAGE_BYTES>
__global__ void __launch_bounds__(288, 1) read_tma(const char* __restrict__ src, size_t bytes, int stages, float* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)stages * STAGE_BYTES);
    uint64_t* empty = full + 8;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < stages; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&full[s])), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&empty[s])), "r"(8));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t nchunks = bytes / STAGE_BYTES;
    const size_t per = nchunks / gridDim.x, rem = nchunks % gridDim.x;
    const size_t b = blockIdx.x * per + min((size_t)blockIdx.x, rem);
    const size_t e = b + per + (blockIdx.x < rem ? 1 : 0);
    if (warp == 8) {
        if (lane == 0) {
            int st = 0; uint32_t ph = 0;
            for (size_t c = b; c < e; ++c) {
                asm volatile("{\n.reg .pred p;\nW1:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D1;\nbra W1;\nD1:\n}\n" ::"r"(s32(&empty[st])), "r"(ph ^ 1u) : "memory");
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[st])), "r"((uint32_t)STAGE_BYTES) : "memory");
                // 8 rows of STAGE_BYTES/8 to mimic the pooling kernel's issue pattern
                for (int j = 0; j < 8; ++j)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(s32(smem + (size_t)st * STAGE_BYTES + j * (STAGE_BYTES / 8))), "l"(src + c * STAGE_BYTES + j * (STAGE_BYTES / 8)),
                                   "r"((uint32_t)(STAGE_BYTES / 8)), "r"(s32(&full[st])) : "memory");
                if (++st == stages) { st = 0; ph ^= 1u; }
            }
        }
        return;
    }
    int st = 0; uint32_t ph = 0;
    float acc = 0.f;
    for (size_t c = b; c < e; ++c) {
        asm volatile("{\n.reg .pred p;\nW2:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D2;\nbra W2;\nD2:\n}\n" ::"r"(s32(&full[st])), "r"(ph) : "memory");
        acc += reinterpret_cast<const float*>(smem + (size_t)st * STAGE_BYTES)[tid];
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[st])) : "memory");
        if (++st == stages) { st = 0; ph ^= 1u; }
    }
    if (acc == 123.456f) out[blockIdx.x * 256 + tid] = acc;
}


// ---- access-pattern probes for the one-read discriminative kernel: tile = [C rows x TP pixels], row stride HW floats ----
__device__ __forceinline__ void cpa16(void* d, const void* g) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(d)), "l"(g) : "memory");
}
// cp.async ring of STAGES tiles, NT threads, round-robin tiles; "compute" = one shared-memory read per thread per tile
template <int TP, int STAGES, int NT>
__global__ void __launch_bounds__(NT) read_tile_cpasync(const float* __restrict__ src, int B, int C, int HW, float* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* tiles = reinterpret_cast<float*>(smem);
    constexpr int RS = TP + 4, NG = TP / 4, NS = NT / NG;
    const int tps = HW / TP, total = B * tps;
    const int tid = threadIdx.x, fq = tid % NG, fr0 = tid / NG;
    const size_t stage_floats = (size_t)C * RS;
    auto issue = [&](int t, int stage) {
        if (t < total) {
            const int b = t / tps, tile = t - b * tps;
            const float* g = src + ((size_t)b * C + fr0) * HW + tile * TP + 4 * fq;
            float* d = tiles + stage * stage_floats + (size_t)fr0 * RS + 4 * fq;
            for (int c = fr0; c < C; c += NS, g += (size_t)NS * HW, d += NS * RS) cpa16(d, g);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int fetched = blockIdx.x;
    for (int j = 0; j < STAGES - 1; ++j) { issue(fetched, j); fetched += gridDim.x; }
    int stage = 0;
    float acc = 0.f;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
        asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 2) : "memory");
        __syncthreads();
        int nst = stage + STAGES - 1; if (nst >= STAGES) nst -= STAGES;
        issue(fetched, nst); fetched += gridDim.x;
        acc += tiles[stage * stage_floats + (tid % C) * RS + (tid / C) % TP];
        if (++stage == STAGES) stage = 0;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (acc == 123.456f) out[blockIdx.x * NT + tid] = acc;
}
// tile straight to registers: warp w holds channels w, w+NW, ..; lane = pixel quad (TP = 128 -> 512-byte warp rows);
// next tile's loads are issued before the current one is consumed (register double buffer of depth CPT)
template <int CPT, int NT>
__global__ void __launch_bounds__(NT) read_tile_regs(const float* __restrict__ src, int B, int C, int HW, float* out) {
    constexpr int TP = 128, NW = NT / 32;
    const int tps = HW / TP, total = B * tps;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc = 0.f;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const int b = t / tps, tile = t - b * tps;
        const float4* g = reinterpret_cast<const float4*>(src + ((size_t)b * C + warp) * HW + tile * TP) + lane;
        float4 v[CPT];
#pragma unroll
        for (int i = 0; i < CPT; ++i) v[i] = ldnc(g + (size_t)i * NW * (HW / 4));
#pragma unroll
        for (int i = 0; i < CPT; ++i) acc += v[i].x + v[i].y + v[i].z + v[i].w;
    }
    if (acc == 123.456f) out[blockIdx.x * NT + threadIdx.x] = acc;
}


// ---- TMA tensor-map tile fetch: one cp.async.bulk.tensor.3d per [rows x 32 px] box, 128B swizzle, mbarrier ring ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) { printf("cuTensorMapEncodeTiled unavailable\n"); exit(1); }
    return (EncodeTiledFn)fn;
}
template <int STAGES, int NT>
__global__ void __launch_bounds__(NT) read_tile_tma(const __grid_constant__ CUtensorMap tmap, int B, int C, int HW, int rows_box, int nbox, float* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int rt = rows_box * nbox;
    const size_t stage_bytes = (size_t)rt * 128;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * stage_bytes);
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&full[s])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int tps = (HW + 31) / 32, total = B * tps;
    auto issue = [&](int t, int stage) {
        if (t < total && tid == 0) {
            const int b = t / tps, tile = t - b * tps;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[stage])), "r"((uint32_t)stage_bytes) : "memory");
            for (int j = 0; j < nbox; ++j)
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                             ::"r"(s32(smem + stage * stage_bytes + (size_t)j * rows_box * 128)), "l"(&tmap), "r"(tile * 32), "r"(j * rows_box), "r"(b),
                               "r"(s32(&full[stage])) : "memory");
        }
    };
    int fetched = blockIdx.x;
    for (int j = 0; j < STAGES - 1; ++j) { issue(fetched, j); fetched += gridDim.x; }
    int stage = 0; uint32_t phase = 0;
    float acc = 0.f;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
        asm volatile("{\n.reg .pred p;\nW3:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D3;\nbra W3;\nD3:\n}\n" ::"r"(s32(&full[stage])), "r"(phase) : "memory");
        __syncthreads();
        int nst = stage + STAGES - 1; if (nst >= STAGES) nst -= STAGES;
        issue(fetched, nst); fetched += gridDim.x;
        acc += reinterpret_cast<const float*>(smem + stage * stage_bytes)[(tid % C) * 32 + (tid / C) % 32];
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
    }
    if (acc == 123.456f) out[blockIdx.x * NT + tid] = acc;
}

template <typename F>
float time_us(F f, int iters = 20) {
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int i = 0; i < 3; ++i) f(i);
    CK(cudaDeviceSynchronize());
    std::vector<float> ts;
    for (int i = 0; i < iters; ++i) {
        CK(cudaEventRecord(a)); f(i); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b)); ts.push_back(ms * 1000.f);
    }
    std::sort(ts.begin(), ts.end());
    return ts[ts.size() / 2];
}

int main(int argc, char** argv) {
    const bool tiles_only = argc > 1 && argv[1][0] == 't';
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const size_t sizes[] = {(size_t)128 << 20, (size_t)256 << 20, (size_t)1024 << 20};
    const int NB = 3;
    float* out; CK(cudaMalloc(&out, 4 << 20));
    for (size_t bytes : sizes) {
        if (tiles_only) break;
        char* buf[NB];
        for (int i = 0; i < NB; ++i) { CK(cudaMalloc(&buf[i], bytes)); CK(cudaMemset(buf[i], 0, bytes)); }
        const size_t n4 = bytes / 16;
        printf("== %zu MiB\n", bytes >> 20);
        for (int mult : {2, 4, 8, 16}) {
            float t = time_us([&](int i) { read_ldg<8><<<sms * mult, 256>>>((const float4*)buf[i % NB], n4, out); });
            printf("read_ldg U=8 grid=%dxSM  %8.2f us  %7.1f GB/s\n", mult, t, bytes / t / 1e3);
        }
        {
            float t = time_us([&](int i) { read_ldg<16><<<sms * 4, 256>>>((const float4*)buf[i % NB], n4, out); });
            printf("read_ldg U=16 grid=4xSM %8.2f us  %7.1f GB/s\n", t, bytes / t / 1e3);
        }
        for (int stages : {2, 3}) {
            const size_t smem = (size_t)stages * 65536 + 128;
            CK(cudaFuncSetAttribute(read_tma<65536>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            float t = time_us([&](int i) { read_tma<65536><<<sms, 288, smem>>>(buf[i % NB], bytes, stages, out); });
            printf("read_tma 64KB x%d        %8.2f us  %7.1f GB/s\n", stages, t, bytes / t / 1e3);
        }
        for (int stages : {2, 4, 6}) {
            const size_t smem = (size_t)stages * 32768 + 128;
            CK(cudaFuncSetAttribute(read_tma<32768>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            float t = time_us([&](int i) { read_tma<32768><<<sms, 288, smem>>>(buf[i % NB], bytes, stages, out); });
            printf("read_tma 32KB x%d        %8.2f us  %7.1f GB/s\n", stages, t, bytes / t / 1e3);
        }
        for (int mult : {4, 8, 16}) {
            float t = time_us([&](int i) { write_st<<<sms * mult, 256>>>((float4*)buf[i % NB], n4); });
            printf("write_st grid=%dxSM      %8.2f us  %7.1f GB/s\n", mult, t, bytes / t / 1e3);
        }
        {
            float t = time_us([&](int i) { CK(cudaMemcpyAsync(buf[(i + 1) % NB], buf[i % NB], bytes, cudaMemcpyDeviceToDevice)); });
            printf("memcpy d2d               %8.2f us  %7.1f GB/s (R+W)\n", t, 2.0 * bytes / t / 1e3);
        }
        for (int i = 0; i < NB; ++i) CK(cudaFree(buf[i]));
    }

    {   // tile-pattern probes on the config-1 feature map [8,256,128*128]
        const int B = 8, C = 256, HW = 16384;
        const size_t bytes = (size_t)B * C * HW * 4;
        float* buf[NB];
        for (int i = 0; i < NB; ++i) { CK(cudaMalloc(&buf[i], bytes)); CK(cudaMemset(buf[i], 0, bytes)); }
        printf("== tile patterns, %zu MiB\n", bytes >> 20);
#define TILE_PROBE(TP, ST, NT, MULT) { \
            const size_t smem = (size_t)ST * C * (TP + 4) * 4; \
            CK(cudaFuncSetAttribute(read_tile_cpasync<TP, ST, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            float t = time_us([&](int i) { read_tile_cpasync<TP, ST, NT><<<sms * MULT, NT, smem>>>(buf[i % NB], B, C, HW, out); }); \
            CK(cudaGetLastError()); \
            printf("tile cp.async TP=%d stages=%d NT=%d ctas/SM=%d  %8.2f us  %7.1f GB/s\n", TP, ST, NT, MULT, t, bytes / t / 1e3); }
        TILE_PROBE(32, 3, 512, 2)
        TILE_PROBE(32, 3, 256, 2)
        TILE_PROBE(32, 2, 512, 2)
        TILE_PROBE(64, 3, 512, 1)
        TILE_PROBE(64, 2, 512, 1)
        TILE_PROBE(64, 2, 1024, 1)
        TILE_PROBE(16, 3, 256, 4)
#define REG_PROBE(CPT, NT, MULT) { \
            float t = time_us([&](int i) { read_tile_regs<CPT, NT><<<sms * MULT, NT>>>(buf[i % NB], B, C, HW, out); }); \
            CK(cudaGetLastError()); \
            printf("tile regs TP=128 CPT=%d NT=%d ctas/SM=%d  %8.2f us  %7.1f GB/s\n", CPT, NT, MULT, t, bytes / t / 1e3); }
        REG_PROBE(16, 512, 1)
        REG_PROBE(16, 512, 2)
        REG_PROBE(32, 256, 2)
        REG_PROBE(32, 256, 4)
        REG_PROBE(8, 1024, 1)
        REG_PROBE(8, 1024, 2)

        {
            EncodeTiledFn enc = encode_tiled();
            const int rows_box = 256, nbox = 1;
#define TMA_PROBE(ST, NT, MULT) { \
            CUtensorMap maps[NB]; \
            for (int i = 0; i < NB; ++i) { \
                cuuint64_t gdim[3] = {(cuuint64_t)HW, (cuuint64_t)C, (cuuint64_t)B}; \
                cuuint64_t gstr[2] = {(cuuint64_t)HW * 4, (cuuint64_t)C * HW * 4}; \
                cuuint32_t box[3] = {32, (cuuint32_t)rows_box, 1}; \
                cuuint32_t estr[3] = {1, 1, 1}; \
                CUresult r = enc(&maps[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, buf[i], gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, \
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); \
                if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); } \
            } \
            const size_t smem = (size_t)ST * rows_box * nbox * 128 + 64 + 1024; \
            CK(cudaFuncSetAttribute(read_tile_tma<ST, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            float t = time_us([&](int i) { read_tile_tma<ST, NT><<<sms * MULT, NT, smem>>>(maps[i % NB], B, C, HW, rows_box, nbox, out); }); \
            CK(cudaGetLastError()); \
            printf("tile TMA tensor TP=32 stages=%d NT=%d ctas/SM=%d  %8.2f us  %7.1f GB/s\n", ST, NT, MULT, t, bytes / t / 1e3); }
            TMA_PROBE(3, 256, 2)
            TMA_PROBE(2, 256, 2)
            TMA_PROBE(2, 256, 3)
            TMA_PROBE(6, 256, 1)
            TMA_PROBE(3, 512, 2)
        }
        for (int i = 0; i < NB; ++i) CK(cudaFree(buf[i]));
    }
    return 0;
}

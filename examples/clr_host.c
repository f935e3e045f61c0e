/* Minimal C host of the CLR C ABI (include/clr_b200.h): class-wise pooling of a feature map, prototypes, and the
 * write-only adjoint -- what utils/Utils.py:108-131 (gen_prototype) does forward and autograd does backward.
 *
 *   gcc -std=c99 -I include -I /usr/local/cuda/include examples/clr_host.c -o /tmp/clr_host \
 *       -L uda_clr_b200/lib -lclr_b200 -L /usr/local/cuda/lib64 -lcudart -lm -Wl,-rpath,$PWD/uda_clr_b200/lib
 *   /tmp/clr_host            (needs a CUDA device; prints "clr_host ok")
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <cuda_runtime_api.h>
#include "clr_b200.h"

#define CHECK_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA: %s\n", cudaGetErrorString(e_)); return 2; } } while (0)
#define CHECK_CLR(x) do { int s_ = (x); if (s_ != CLR_OK) { fprintf(stderr, "clr: %s (%d)\n", clr_status_string(s_), s_); return 3; } } while (0)

int main(void) {
    const int B = 2, C = 24, H = 16, W = 16, K = 2, HW = H * W, R = 2 * K;
    const size_t nx = (size_t)B * C * HW, ny = (size_t)B * K * HW;
    float* hx = (float*)malloc(nx * sizeof(float));
    float* hy = (float*)malloc(ny * sizeof(float));
    float* hg = (float*)malloc((size_t)R * C * sizeof(float));
    unsigned s = 12345u;
    for (size_t i = 0; i < nx; ++i) { s = s * 1664525u + 1013904223u; hx[i] = (float)((s >> 8) & 0xffff) / 32768.0f - 1.0f; }
    for (size_t i = 0; i < ny; ++i) { s = s * 1664525u + 1013904223u; hy[i] = ((s >> 16) & 3u) == 0u ? 1.0f : 0.0f; }   /* hard labels */
    for (int i = 0; i < R * C; ++i) hg[i] = 0.01f * (float)(i % 7 - 3);

    float *dx, *dy, *dsums, *dmu, *dg, *dgrad;
    void* dws;
    const size_t ws_bytes = clr_pool_ws_bytes(B, C, HW, K);
    CHECK_CUDA(cudaMalloc((void**)&dx, nx * sizeof(float)));
    CHECK_CUDA(cudaMalloc((void**)&dy, ny * sizeof(float)));
    CHECK_CUDA(cudaMalloc((void**)&dsums, (size_t)R * (C + 1) * sizeof(float)));
    CHECK_CUDA(cudaMalloc((void**)&dmu, (size_t)R * C * sizeof(float)));
    CHECK_CUDA(cudaMalloc((void**)&dg, (size_t)R * C * sizeof(float)));
    CHECK_CUDA(cudaMalloc((void**)&dgrad, nx * sizeof(float)));
    CHECK_CUDA(cudaMalloc(&dws, ws_bytes));
    CHECK_CUDA(cudaMemcpy(dx, hx, nx * sizeof(float), cudaMemcpyHostToDevice));
    CHECK_CUDA(cudaMemcpy(dy, hy, ny * sizeof(float), cudaMemcpyHostToDevice));
    CHECK_CUDA(cudaMemcpy(dg, hg, (size_t)R * C * sizeof(float), cudaMemcpyHostToDevice));

    cudaStream_t st;
    CHECK_CUDA(cudaStreamCreate(&st));
    CHECK_CLR(clr_pool_fwd(dx, dy, CLR_W_COMPLEMENT, B, C, HW, K, dws, ws_bytes, dsums, st));
    CHECK_CLR(clr_proto_finalize(dsums, R, C, dmu, st));
    CHECK_CLR(clr_pool_bwd(dy, CLR_W_COMPLEMENT, B, C, HW, K, dg, dsums, 1.0f, NULL, NULL, 0, dgrad, st));
    CHECK_CUDA(cudaStreamSynchronize(st));

    float* mu = (float*)malloc((size_t)R * C * sizeof(float));
    float* sums = (float*)malloc((size_t)R * (C + 1) * sizeof(float));
    float* grad = (float*)malloc(nx * sizeof(float));
    CHECK_CUDA(cudaMemcpy(mu, dmu, (size_t)R * C * sizeof(float), cudaMemcpyDeviceToHost));
    CHECK_CUDA(cudaMemcpy(sums, dsums, (size_t)R * (C + 1) * sizeof(float), cudaMemcpyDeviceToHost));
    CHECK_CUDA(cudaMemcpy(grad, dgrad, nx * sizeof(float), cudaMemcpyDeviceToHost));

    /* check against the closed form in double on the host */
    double worst_mu = 0.0, worst_g = 0.0, ref_mu = 0.0, ref_g = 0.0;
    for (int r = 0; r < R; ++r) {
        const int k = r % K, bck = r >= K;
        double N = 0.0;
        for (int b = 0; b < B; ++b) for (int p = 0; p < HW; ++p) { const double w = hy[((size_t)b * K + k) * HW + p]; N += bck ? 1.0 - w : w; }
        if (fabs(N - (double)sums[(size_t)r * (C + 1) + C]) != 0.0) { fprintf(stderr, "count mismatch row %d\n", r); return 4; }
        for (int c = 0; c < C; ++c) {
            double S = 0.0;
            for (int b = 0; b < B; ++b) for (int p = 0; p < HW; ++p) {
                const double w = hy[((size_t)b * K + k) * HW + p];
                S += (bck ? 1.0 - w : w) * hx[((size_t)b * C + c) * HW + p];
            }
            const double m = S / N, d = fabs(m - (double)mu[(size_t)r * C + c]);
            if (d > worst_mu) worst_mu = d;
            if (fabs(m) > ref_mu) ref_mu = fabs(m);
        }
    }
    for (int b = 0; b < B; ++b) for (int c = 0; c < C; ++c) for (int p = 0; p < HW; ++p) {
        double gsum = 0.0;
        for (int r = 0; r < R; ++r) {
            const int k = r % K, bck = r >= K;
            const double w = hy[((size_t)b * K + k) * HW + p];
            gsum += (double)hg[(size_t)r * C + c] / (double)sums[(size_t)r * (C + 1) + C] * (bck ? 1.0 - w : w);
        }
        const double d = fabs(gsum - (double)grad[((size_t)b * C + c) * HW + p]);
        if (d > worst_g) worst_g = d;
        if (fabs(gsum) > ref_g) ref_g = fabs(gsum);
    }
    printf("prototype relerr %.2e, gradient relerr %.2e, kernels launched %llu\n", worst_mu / ref_mu, worst_g / ref_g, clr_launch_count());
    if (worst_mu / ref_mu > 1e-5 || worst_g / ref_g > 1e-4) return 5;
    printf("clr_host ok\n");
    return 0;
}

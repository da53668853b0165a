// skb_peak.cu -- FP32 / FP64 FMA-pipe micro-benchmark: the denominator of the roofline of the compute-bound modes
// (closest row, validity bytes), which SURVEY.md 8(d) asks the builder to measure because MEASURED_PEAKS.json holds only
// the HBM and bf16 tensor figures.  Measurement aid, not part of the hot path.
//
// Every thread runs CHAINS independent dependent-FMA chains out of registers; the grid fills every SM with 2048 resident
// threads.  flop = 2 * threads * CHAINS * width * iters.  `packed` uses fma.rn.f32x2 (two fp32 FMAs per instruction, the
// form the step kernel's wrench contraction uses).
#include <cuda_runtime.h>

#include <string>

#include "skb_internal.h"

namespace skb {

constexpr int PEAK_CHAINS = 8;
constexpr int PEAK_UNROLL = 16;

__global__ void __launch_bounds__(1024) fma_peak_f32(float *out, int iters, float a, float b) {
    float x[PEAK_CHAINS];
#pragma unroll
    for (int i = 0; i < PEAK_CHAINS; ++i) x[i] = (float)(threadIdx.x + i) * 1e-3f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < PEAK_UNROLL; ++u)
#pragma unroll
            for (int i = 0; i < PEAK_CHAINS; ++i) x[i] = __fmaf_rn(x[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PEAK_CHAINS; ++i) s += x[i];
    if (s == 123.456f) out[0] = s;                    // keeps the chains alive without a store on the timed path
}

__global__ void __launch_bounds__(1024) fma_peak_f32x2(float *out, int iters, float a, float b) {
    unsigned long long x[PEAK_CHAINS];
    unsigned long long aa, bb;
    asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
#pragma unroll
    for (int i = 0; i < PEAK_CHAINS; ++i) {
        const float v = (float)(threadIdx.x + i) * 1e-3f;
        asm("mov.b64 %0, {%1, %1};" : "=l"(x[i]) : "f"(v));
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < PEAK_UNROLL; ++u)
#pragma unroll
            for (int i = 0; i < PEAK_CHAINS; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(aa), "l"(bb));
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PEAK_CHAINS; ++i) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i]));
        s += lo + hi;
    }
    if (s == 123.456f) out[0] = s;
}

__global__ void __launch_bounds__(1024) fma_peak_f64(float *out, int iters, double a, double b) {
    double x[PEAK_CHAINS];
#pragma unroll
    for (int i = 0; i < PEAK_CHAINS; ++i) x[i] = (double)(threadIdx.x + i) * 1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < PEAK_UNROLL; ++u)
#pragma unroll
            for (int i = 0; i < PEAK_CHAINS; ++i) x[i] = __fma_rn(x[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < PEAK_CHAINS; ++i) s += x[i];
    if (s == 123.456) out[0] = (float)s;
}

}  // namespace skb

using namespace skb;

#define CUDA_OK(expr)                                                                          \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return set_error(std::string(#expr) + ": " + cudaGetErrorString(e__));             \
    } while (0)

extern "C" int skb_fma_peak(int32_t kind, int32_t iters, double *tflops_out) {
    if (!tflops_out || iters < 1 || kind < 0 || kind > 2) return set_error("fma_peak: bad arguments (kind 0 f32, 1 f32x2, 2 f64)");
    int dev = 0, sms = 0;
    CUDA_OK(cudaGetDevice(&dev));
    CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    float *out = nullptr;
    CUDA_OK(cudaMalloc(&out, 16));
    cudaEvent_t e0, e1;
    CUDA_OK(cudaEventCreate(&e0));
    CUDA_OK(cudaEventCreate(&e1));
    const int grid = sms * 2, block = 1024;
    auto launch = [&](int n) {
        if (kind == 0) fma_peak_f32<<<grid, block>>>(out, n, 0.999f, 1e-3f);
        else if (kind == 1) fma_peak_f32x2<<<grid, block>>>(out, n, 0.999f, 1e-3f);
        else fma_peak_f64<<<grid, block>>>(out, n, 0.999, 1e-3);
    };
    launch(iters / 4 + 1);                                     // warm-up (clocks, instruction cache)
    CUDA_OK(cudaEventRecord(e0));
    launch(iters);
    CUDA_OK(cudaEventRecord(e1));
    CUDA_OK(cudaEventSynchronize(e1));
    CUDA_OK(cudaGetLastError());
    float ms = 0.f;
    CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    const double width = kind == 1 ? 2.0 : 1.0;
    const double flop = 2.0 * grid * (double)block * PEAK_CHAINS * PEAK_UNROLL * width * (double)iters;
    *tflops_out = flop / (ms * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    ++g_launch_count;
    return 0;
}

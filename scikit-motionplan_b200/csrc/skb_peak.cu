// skb_peak.cu -- FP32 / FP64 FMA-pipe micro-benchmark: the denominator of the roofline of the compute-bound modes
// (closest row, validity bytes), which SURVEY.md 8(d) asks the builder to measure because MEASURED_PEAKS.json holds only
// the HBM and bf16 tensor figures.  Measurement aid, not part of the hot path.
//
// Every thread runs CHAINS independent dependent-FMA chains out of registers; the grid fills every SM with 2048 resident
// threads.  flop = 2 * threads * CHAINS * width * iters.  `packed` uses fma.rn.f32x2 (two fp32 FMAs per instruction, the
// form the step kernel's wrench contraction uses).
#include <cuda_runtime.h>

#include <string>

#include "skb_device.cuh"
#include "skb_internal.h"

namespace skb {

// ---- write-stream micro-benchmark: what a kernel that only WRITES can reach on this GPU.  MEASURED_PEAKS.json's HBM figure
// is a copy (read + write); the collision kernels write 650 bytes for every byte they read, so the ceiling of their store
// path is measured separately: kind 0 = 16-byte streaming stores from registers (st.global.cs.v4), kind 1 = the step kernel's own
// path -- a block assembled in shared memory leaves as one cp.async.bulk (TMA 1-D) store with the L2 evict-first policy,
// `block_bytes` per instruction, one staging block per warp, the next block written after the previous one has been read.
__global__ void __launch_bounds__(256) store_peak_stg(float4 *out, long long n16) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const float v = (float)threadIdx.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) __stcs(out + i, make_float4(v, v, v, v));
}
__global__ void __launch_bounds__(256) store_peak_bulk(unsigned char *out, long long n_blocks, int block_bytes) {
    extern __shared__ __align__(128) unsigned char peak_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    float4 *stage = reinterpret_cast<float4 *>(peak_smem + (size_t)warp * block_bytes);
    const uint64_t pol = l2_evict_first_policy();
    const long long step = (long long)gridDim.x * n_warps;
    for (long long b = (long long)blockIdx.x * n_warps + warp; b < n_blocks; b += step) {
        if (lane == 0) bulk_wait_read<0>();
        __syncwarp();
        const float v = (float)b;
        for (int i = lane; i < block_bytes / 16; i += 32) stage[i] = make_float4(v, v, v, v);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
            bulk_store_hint(out + b * block_bytes, stage, (uint32_t)block_bytes, pol);
            bulk_commit();
        }
    }
    bulk_wait_read<0>();
}

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

// kind 0: streaming 16-byte stores; kind 1: shared memory -> TMA bulk stores of `block_bytes` (multiple of 16, <= 8192) each
extern "C" int skb_store_peak(int32_t kind, int64_t bytes, int32_t block_bytes, double *gbs_out) {
    if (!gbs_out || bytes < (1 << 20) || kind < 0 || kind > 1) return set_error("store_peak: bad arguments (kind 0 stg, 1 bulk)");
    if (kind == 1 && (block_bytes < 16 || block_bytes > 8192 || block_bytes % 16)) return set_error("store_peak: block_bytes must be a multiple of 16 in [16, 8192]");
    int dev = 0, sms = 0;
    CUDA_OK(cudaGetDevice(&dev));
    CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    unsigned char *out = nullptr;
    CUDA_OK(cudaMalloc(&out, (size_t)bytes));
    cudaEvent_t e0, e1;
    CUDA_OK(cudaEventCreate(&e0));
    CUDA_OK(cudaEventCreate(&e1));
    const long long n_blocks = kind == 1 ? bytes / block_bytes : 0;
    const int warps = 8;
    const size_t smem = kind == 1 ? (size_t)warps * block_bytes : 0;
    int per_sm = 8;
    if (kind == 1) {
        CUDA_OK(cudaFuncSetAttribute(store_peak_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, store_peak_bulk, warps * 32, smem));
        per_sm = per_sm < 1 ? 1 : (per_sm > 3 ? 3 : per_sm);                     // the step kernel keeps 3 CTAs of 8 warps per SM
    }
    auto launch = [&]() {
        if (kind == 0) store_peak_stg<<<sms * 8, 256>>>(reinterpret_cast<float4 *>(out), bytes / 16);
        else store_peak_bulk<<<sms * per_sm, warps * 32, smem>>>(out, n_blocks, block_bytes);
    };
    launch();                                                  // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        CUDA_OK(cudaEventRecord(e0));
        launch();
        CUDA_OK(cudaEventRecord(e1));
        CUDA_OK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
        best = ms < best ? ms : best;
    }
    CUDA_OK(cudaGetLastError());
    const double written = kind == 1 ? (double)n_blocks * block_bytes : (double)(bytes / 16) * 16.0;
    *gbs_out = written / (best * 1e-3) / 1e9;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    g_launch_count += 6;
    return 0;
}

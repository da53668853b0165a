// skb_edges.cu -- batched motion-step (edge) discretisation and the per-edge AND, the device half of
// `is_valid_motion_step` for many edges at once (BASELINE config 5 "RRT edge validity").
//
//     interpolate_fractions(motion_step_box, q1, q2, include_q1)      skmp/solver/motion_step_box.py:8-41
//     is_valid_motion_step: every sample q1 + (q2 - q1) * frac valid  skmp/solver/motion_step_box.py:44-57
//
// All arithmetic on the fractions is fp64 and sequential exactly like the reference's Python loop
// (`travel_rate += step_ratio`), so sample lists are bit-identical to the reference's:
//   edge_count_kernel   thread per edge: active axis (first arg-max of |diff| / box), step ratio, number of samples
//   edge_sample_kernel  warp per edge: sample k of the edge -> q1 + (q2 - q1) * frac_k, written as fp32 or fp64 rows
//   segment_all_kernel  thread per edge: AND of the validity bytes of its samples (an edge without samples is valid)
#include <cuda_runtime.h>

#include <string>

#include "skb_internal.h"

namespace skb {

struct EdgePlan { double step; int n; int kind; };   // kind 0: no sample, 1: only [1.0], 2: walk

__device__ __forceinline__ EdgePlan edge_plan(const double *q1, const double *q2, const double *box, int D) {
    int active = 0;
    double best = -1.0;
    for (int d = 0; d < D; ++d) {
        const double s = fabs(q2[d] - q1[d]) / box[d];
        if (s > best) { best = s; active = d; }           // first maximum wins (np.argmax)
    }
    const double da = fabs(q2[active] - q1[active]);
    EdgePlan p;
    p.step = 0.0; p.n = 0; p.kind = 0;
    if (da < 1e-6) return p;
    p.step = box[active] / da;
    if (p.step > 1.0) { p.kind = 1; return p; }
    p.kind = 2;
    double travel = 0.0;
    while (travel + p.step < 1.0) { travel += p.step; ++p.n; }
    return p;
}

// number of samples of a walked edge and whether a final 1.0 is appended
__device__ __forceinline__ int walk_count(const EdgePlan &p, int include_q1, bool &append_one) {
    double last = 0.0;
    for (int k = 0; k < p.n; ++k) last += p.step;
    const int listed = (include_q1 ? 1 : 0) + p.n;
    append_one = listed == 0 || fabs(last - 1.0) > 1e-6;
    return listed + (append_one ? 1 : 0);
}

__global__ void edge_count_kernel(const double *q1, const double *q2, const double *box, long long E, int D, int include_q1,
                                  long long *counts) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (long long)gridDim.x * blockDim.x) {
        const EdgePlan p = edge_plan(q1 + e * D, q2 + e * D, box, D);
        long long c = 0;
        if (p.kind == 1) c = 1;
        else if (p.kind == 2) { bool ap; c = walk_count(p, include_q1, ap); }
        counts[e] = c;
    }
}

template <typename T>
__global__ void edge_sample_kernel(const double *q1, const double *q2, const double *box, long long E, int D, int include_q1,
                                   const long long *offsets, T *samples) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long e = warp; e < E; e += n_warps) {
        const double *a = q1 + e * D, *b = q2 + e * D;
        const EdgePlan p = edge_plan(a, b, box, D);
        const long long o0 = offsets[e];
        const int cnt = (int)(offsets[e + 1] - o0);
        for (int i = lane; i < cnt * D; i += 32) {
            const int k = i / D, d = i - k * D;
            double frac = 1.0;
            if (p.kind == 2) {
                const int steps = include_q1 ? k : k + 1;             // number of sequential additions behind sample k
                if (steps <= p.n) { frac = 0.0; for (int s = 0; s < steps; ++s) frac += p.step; }
            }
            samples[(o0 + k) * D + d] = (T)__dadd_rn(a[d], __dmul_rn(b[d] - a[d], frac));   // no FMA contraction: numpy rounds twice
        }
    }
}

__global__ void segment_all_kernel(const uint8_t *valid, const long long *offsets, long long E, uint8_t *out) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (long long)gridDim.x * blockDim.x) {
        uint8_t ok = 1;
        for (long long i = offsets[e]; i < offsets[e + 1]; ++i) ok &= valid[i] != 0;
        out[e] = ok;
    }
}

static int grid_for(long long n, int threads) {
    long long g = (n + threads - 1) / threads;
    return (int)(g < 1 ? 1 : (g > 148 * 32 ? 148 * 32 : g));
}

}  // namespace skb

using namespace skb;

#define CHECK_LAUNCH(what)                                                                       \
    do {                                                                                         \
        cudaError_t e__ = cudaGetLastError();                                                    \
        if (e__ != cudaSuccess) return set_error(std::string(what ": ") + cudaGetErrorString(e__)); \
        ++g_launch_count;                                                                        \
    } while (0)

extern "C" {

int skb_edge_counts(const double *q1_dev, const double *q2_dev, const double *box_dev, int64_t E, int32_t D, int32_t include_q1,
                    int64_t *counts_dev, void *stream) {
    if (E < 0 || D < 1) return set_error("edge_counts: bad sizes");
    if (E == 0) return 0;
    if (!q1_dev || !q2_dev || !box_dev || !counts_dev) return set_error("edge_counts: bad arguments");
    edge_count_kernel<<<grid_for(E, 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        q1_dev, q2_dev, box_dev, E, D, include_q1, reinterpret_cast<long long *>(counts_dev));
    CHECK_LAUNCH("edge_counts");
    return 0;
}

int skb_edge_samples(const double *q1_dev, const double *q2_dev, const double *box_dev, int64_t E, int32_t D, int32_t include_q1,
                     const int64_t *offsets_dev, int32_t out_dtype, void *samples_dev, void *stream) {
    if (E < 0 || D < 1) return set_error("edge_samples: bad sizes");
    if (E == 0) return 0;
    if (!q1_dev || !q2_dev || !box_dev || !offsets_dev || !samples_dev) return set_error("edge_samples: bad arguments");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int grid = grid_for(E * 32, 256);
    if (out_dtype == SKB_F32)
        edge_sample_kernel<float><<<grid, 256, 0, st>>>(q1_dev, q2_dev, box_dev, E, D, include_q1,
                                                        reinterpret_cast<const long long *>(offsets_dev), reinterpret_cast<float *>(samples_dev));
    else if (out_dtype == SKB_F64)
        edge_sample_kernel<double><<<grid, 256, 0, st>>>(q1_dev, q2_dev, box_dev, E, D, include_q1,
                                                         reinterpret_cast<const long long *>(offsets_dev), reinterpret_cast<double *>(samples_dev));
    else return set_error("bad dtype");
    CHECK_LAUNCH("edge_samples");
    return 0;
}

int skb_segment_all(const uint8_t *valid_dev, const int64_t *offsets_dev, int64_t E, uint8_t *out_dev, void *stream) {
    if (E < 0) return set_error("segment_all: bad sizes");
    if (E == 0) return 0;
    if (!offsets_dev || !out_dev) return set_error("segment_all: bad arguments");
    segment_all_kernel<<<grid_for(E, 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        valid_dev, reinterpret_cast<const long long *>(offsets_dev), E, out_dev);
    CHECK_LAUNCH("segment_all");
    return 0;
}

}  // extern "C"

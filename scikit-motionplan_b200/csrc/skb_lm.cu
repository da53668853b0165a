// skb_lm.cu -- one damped Gauss-Newton (Levenberg-Marquardt) step for a batch of seeds, the linear-algebra half of the
// batched multi-seed IK / projection that replaces the reference's sequential SLSQP restarts
// (skmp/satisfy.py:45-143, solver/ompl_solver.py:146-154, solver/myrrt_solver.py:152-160).
//
//     q_new[b] = clamp( q[b] + dq[b], lb, ub ),   (J_b^T J_b + lam_b (I + diag(J_b^T J_b))) dq[b] = -J_b^T r_b
//
// One warp per seed.  The residual rows of seed b (equality rows and hinge rows of the violated inequalities, built by
// skmp_b200/satisfy.py from the constraint kernels' outputs) are streamed through shared memory one row at a time while
// the lanes accumulate the lower triangle of J^T J and J^T r; the D x D system (D <= 64) is then factorised in place by
// a warp-cooperative Cholesky and solved by forward / backward substitution -- no library call, no host round trip.
#include <cuda_runtime.h>

#include <algorithm>
#include <string>

#include "skb_internal.h"

namespace skb {

template <typename T>
__global__ void __launch_bounds__(128) lm_step_kernel(const T *__restrict__ J, const T *__restrict__ r, const T *__restrict__ lam,
                                                      const T *__restrict__ q, const T *__restrict__ lb, const T *__restrict__ ub,
                                                      T *__restrict__ q_new, int B, int M, int D, int warp_elems) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const int b = blockIdx.x * n_warps + warp;
    if (b >= B) return;
    T *A = reinterpret_cast<T *>(smem_raw) + (size_t)warp * warp_elems;     // [D][DP] lower triangle used
    const int DP = D + 1;                                                    // odd-ish pitch against bank conflicts
    T *g = A + D * DP;                                                       // J^T r, then the solution
    T *row = g + D;                                                          // one residual row of J (D) + its r (1)
    const int n_tri = D * (D + 1) / 2;
    for (int e = lane; e < D * DP + D; e += 32) A[e] = T(0);
    __syncwarp();
    const T *Jb = J + (size_t)b * M * D;
    const T *rb = r + (size_t)b * M;
    for (int m = 0; m < M; ++m) {
        const T rm = rb[m];
        if (rm == T(0)) {                      // inactive hinge row: contributes nothing (uniform test, rb[m] is per seed)
            bool any = false;
            for (int d = lane; d < D; d += 32) any |= Jb[(size_t)m * D + d] != T(0);
            if (!__any_sync(0xffffffffu, any)) continue;
        }
        for (int d = lane; d < D; d += 32) row[d] = Jb[(size_t)m * D + d];
        __syncwarp();
        for (int e = lane; e < n_tri; e += 32) {
            // e -> (i, j) with i >= j: i = floor((sqrt(8e + 1) - 1) / 2)
            int i = (int)((sqrtf(8.f * (float)e + 1.f) - 1.f) * 0.5f);
            while (i * (i + 1) / 2 > e) --i;
            while ((i + 1) * (i + 2) / 2 <= e) ++i;
            const int j = e - i * (i + 1) / 2;
            A[i * DP + j] += row[i] * row[j];
        }
        for (int d = lane; d < D; d += 32) g[d] += row[d] * rm;
        __syncwarp();
    }
    // damping (Marquardt scaling) and right-hand side
    const T l = lam[b];
    for (int d = lane; d < D; d += 32) {
        A[d * DP + d] += l * (T(1) + A[d * DP + d]);
        g[d] = -g[d];
    }
    __syncwarp();
    // in-place Cholesky A = L L^T (lower), right-looking, one column at a time
    for (int k = 0; k < D; ++k) {
        const T dkk = sqrt(max(A[k * DP + k], sizeof(T) == 8 ? T(1e-200) : T(1e-30)));
        __syncwarp();
        for (int i = k + lane; i < D; i += 32) A[i * DP + k] = (i == k) ? dkk : A[i * DP + k] / dkk;
        __syncwarp();
        const int rem = D - k - 1;
        for (int e = lane; e < rem * (rem + 1) / 2; e += 32) {
            int i = (int)((sqrtf(8.f * (float)e + 1.f) - 1.f) * 0.5f);
            while (i * (i + 1) / 2 > e) --i;
            while ((i + 1) * (i + 2) / 2 <= e) ++i;
            const int j = e - i * (i + 1) / 2;
            const int ii = k + 1 + i, jj = k + 1 + j;
            A[ii * DP + jj] -= A[ii * DP + k] * A[jj * DP + k];
        }
        __syncwarp();
    }
    // forward substitution L y = g, then backward L^T x = y (columns handled by the lanes, one pivot at a time)
    for (int k = 0; k < D; ++k) {
        const T yk = g[k] / A[k * DP + k];
        __syncwarp();
        if (lane == 0) g[k] = yk;
        for (int i = k + 1 + lane; i < D; i += 32) g[i] -= A[i * DP + k] * yk;
        __syncwarp();
    }
    for (int k = D - 1; k >= 0; --k) {
        const T xk = g[k] / A[k * DP + k];
        __syncwarp();
        if (lane == 0) g[k] = xk;
        for (int i = lane; i < k; i += 32) g[i] -= A[k * DP + i] * xk;
        __syncwarp();
    }
    for (int d = lane; d < D; d += 32) {
        T v = q[(size_t)b * D + d] + g[d];
        v = v < lb[d] ? lb[d] : (v > ub[d] ? ub[d] : v);
        q_new[(size_t)b * D + d] = v;
    }
}

template <typename T>
static int launch_lm_t(const void *J, const void *r, const void *lam, const void *q, const void *lb, const void *ub, void *q_new,
                       int B, int M, int D, cudaStream_t stream) {
    const int warp_elems = ((D * (D + 1) + D + D + 1) + 3) / 4 * 4;
    const int warps = 4;
    const size_t smem = (size_t)warps * warp_elems * sizeof(T);
    auto kern = lm_step_kernel<T>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(std::string("lm_step: ") + cudaGetErrorString(e));
    const int grid = (B + warps - 1) / warps;
    kern<<<grid, warps * 32, smem, stream>>>(reinterpret_cast<const T *>(J), reinterpret_cast<const T *>(r), reinterpret_cast<const T *>(lam),
                                             reinterpret_cast<const T *>(q), reinterpret_cast<const T *>(lb), reinterpret_cast<const T *>(ub),
                                             reinterpret_cast<T *>(q_new), B, M, D, warp_elems);
    e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(std::string("lm_step: ") + cudaGetErrorString(e));
    ++g_launch_count;
    return 0;
}

}  // namespace skb

using namespace skb;

extern "C" int skb_lm_step(int32_t dtype, const void *J_dev, const void *r_dev, const void *lam_dev, const void *q_dev,
                           const void *lb_dev, const void *ub_dev, void *q_new_dev, int64_t B, int32_t M, int32_t D, void *stream) {
    if (B < 0 || M < 0 || D < 1 || D > 64) return set_error("lm_step: bad sizes (1 <= D <= 64)");
    if (B == 0) return 0;
    if (!J_dev || !r_dev || !lam_dev || !q_dev || !lb_dev || !ub_dev || !q_new_dev) return set_error("lm_step: bad arguments");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == SKB_F32) return launch_lm_t<float>(J_dev, r_dev, lam_dev, q_dev, lb_dev, ub_dev, q_new_dev, (int)B, M, D, st);
    if (dtype == SKB_F64) return launch_lm_t<double>(J_dev, r_dev, lam_dev, q_dev, lb_dev, ub_dev, q_new_dev, (int)B, M, D, st);
    return set_error("bad dtype");
}

// skb_com.cu -- centre-of-mass forward kinematics + Jacobian, and the stability constraint built on it.
//
//     tinyfk.solve_com_fk(qs, joint_ids, action_link_ids, action_forces, base_type, with_jacobian)
//                                                                    (call site skmp/constraint.py:901-908)
//     COMStabilityConst._evaluate: value = -box.sdf(com), J = -grad . J_com    (skmp/constraint.py:899-928)
//
// The plan's features are the weighted points of the robot (one per massive link at its inertial origin, weight =
// mass; one per "action" link, weight = force / g), sorted by the DFS preorder of the node they hang on, so the points
// below any joint form a contiguous range.  One warp owns a tile of G configurations:
//   phase 1  frames + twists, three lanes per (configuration, node)                 -> fk_phase (skb_device.cuh)
//   phase 2  per configuration: lane per point, (w p, w) -> warp inclusive scan -> prefix table in shared memory;
//            com = prefix[n] / W;  lane per column d:  J_com[:, d] = (omega_d x S_d + M_d v_d) / W  with the subtree
//            first moment S_d and mass M_d read as a difference of two prefix entries (no per-joint reduction).
//   what = SKB_COM_POINT      out_vals (N,3), out_jac (N,3,D)
//   what = SKB_COM_STABILITY  out_vals (N,1) = -sdf(com), out_jac (N,1,D) = -grad sdf(com) . J_com (analytic gradient)
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <string>

#include "skb_device.cuh"
#include "skb_internal.h"

namespace skb {

template <typename T>
struct ComParams {
    const int32_t *I;
    const T *F;
    const T *q;
    long long N;
    long long n_tiles;
    const DevPrim<T> *prims;
    int n_prims;
    int what;
    T *out_vals;
    T *out_jac;
    int i_total, f_total;
    int off_F, off_prims, off_warp, warp_elems;
    int off_lv, n_nodes;
    int woff_q, woff_sc, woff_frames, woff_tw, woff_prefix;
    int fstride, tstride;
    int G, logG, slots3;
    int D, n_points, n_levels;
};

template <typename T>
__global__ void __launch_bounds__(256) com_kernel(const ComParams<T> prm) {
    using V4 = Vec4<T>;
    using Rl = Real<T>;
    extern __shared__ __align__(16) unsigned char smem[];
    int32_t *sI = reinterpret_cast<int32_t *>(smem);
    T *sF = reinterpret_cast<T *>(smem + prm.off_F);
    DevPrim<T> *sP = reinterpret_cast<DevPrim<T> *>(smem + prm.off_prims);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n_warps = blockDim.x >> 5;
    for (int i = tid; i < prm.i_total; i += blockDim.x) sI[i] = prm.I[i];
    for (int i = tid; i < prm.f_total; i += blockDim.x) sF[i] = prm.F[i];
    {
        const int words = prm.n_prims * (int)(sizeof(DevPrim<T>) / 4);
        const int32_t *src = reinterpret_cast<const int32_t *>(prm.prims);
        int32_t *dst = reinterpret_cast<int32_t *>(sP);
        for (int i = tid; i < words; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();

    const int D = prm.D, n_points = prm.n_points, n_levels = prm.n_levels;
    const int fstride = prm.fstride, tstride = prm.tstride, G = prm.G, logG = prm.logG;
    const int32_t *levelI = sI + sI[C_LEVEL_I];
    int4 *lvrec = reinterpret_cast<int4 *>(smem + prm.off_lv);
    build_level_records(lvrec, sI + sI[C_NODE_I], sI + sI[C_LEVELNODE_I], prm.n_nodes);
    __syncthreads();
    const int32_t *ptNode = sI + sI[C_PTNODE_I];
    const int2 *colRange = reinterpret_cast<const int2 *>(sI + sI[C_COLRANGE_I]);
    const V4 *nodeF = reinterpret_cast<const V4 *>(sF + sI[C_NODE_F]);
    const V4 *ptA = reinterpret_cast<const V4 *>(sF + sI[C_PT_F]);

    T *wbase = reinterpret_cast<T *>(smem + prm.off_warp) + (size_t)warp * prm.warp_elems;
    T *wq = wbase + prm.woff_q;
    T *wsc = wbase + prm.woff_sc;
    T *wframes = wbase + prm.woff_frames;
    T *wtw = wbase + prm.woff_tw;
    V4 *wprefix = reinterpret_cast<V4 *>(wbase + prm.woff_prefix);

    for (long long tile = (long long)blockIdx.x * n_warps + warp; tile < prm.n_tiles; tile += (long long)gridDim.x * n_warps) {
        const long long n0 = tile * G;
        const int g_live = (int)min((long long)G, prm.N - n0);
        __syncwarp();
        for (int i = lane; i < g_live * D; i += 32) {
            const T th = prm.q[n0 * D + i];
            T sn, cs;
            Rl::sincos(th, &sn, &cs);
            wq[i] = th;
            wsc[2 * i] = sn; wsc[2 * i + 1] = cs;
        }
        __syncwarp();
        fk_phase<T, true, false>(lane, G, logG, prm.slots3, g_live, D, n_levels, lvrec, levelI, nodeF, wq, wsc,
                                 wframes, fstride, wtw, tstride);

        for (int c = 0; c < g_live; ++c) {
            const long long n = n0 + c;
            const V4 *fr = reinterpret_cast<const V4 *>(wframes + c * fstride);
            const T *tw = wtw + c * tstride;
            // ---- weighted points -> inclusive prefix sums of (w p, w) in preorder
            V4 carry; carry.x = carry.y = carry.z = carry.w = T(0);
            if (lane == 0) wprefix[0] = carry;
            for (int r0 = 0; r0 < n_points; r0 += 32) {
                const int i = r0 + lane;
                const bool has = i < n_points;
                const int node = ptNode[has ? i : 0];
                V4 A = ptA[has ? i : 0];
                if (!has) A.w = T(0);
                const V4 f0 = fr[3 * node], f1 = fr[3 * node + 1], f2 = fr[3 * node + 2];
                V4 v;
                v.x = A.w * (f0.w + f0.x * A.x + f0.y * A.y + f0.z * A.z);
                v.y = A.w * (f1.w + f1.x * A.x + f1.y * A.y + f1.z * A.z);
                v.z = A.w * (f2.w + f2.x * A.x + f2.y * A.y + f2.z * A.z);
                v.w = A.w;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const T tx = __shfl_up_sync(0xffffffffu, v.x, o), ty = __shfl_up_sync(0xffffffffu, v.y, o);
                    const T tz = __shfl_up_sync(0xffffffffu, v.z, o), tw_ = __shfl_up_sync(0xffffffffu, v.w, o);
                    if (lane >= o) { v.x += tx; v.y += ty; v.z += tz; v.w += tw_; }
                }
                v.x += carry.x; v.y += carry.y; v.z += carry.z; v.w += carry.w;
                if (has) wprefix[i + 1] = v;
                carry.x = __shfl_sync(0xffffffffu, v.x, 31); carry.y = __shfl_sync(0xffffffffu, v.y, 31);
                carry.z = __shfl_sync(0xffffffffu, v.z, 31); carry.w = __shfl_sync(0xffffffffu, v.w, 31);
            }
            __syncwarp();
            const T invW = T(1) / carry.w;
            const T cx = carry.x * invW, cy = carry.y * invW, cz = carry.z * invW;
            T gx = T(0), gy = T(0), gz = T(0);
            if (prm.what == SKB_COM_STABILITY) {
                const T d = sdf_union<T>(sP, prm.n_prims, cx, cy, cz, gx, gy, gz);
                if (lane == 0) prm.out_vals[n] = -d;
            } else if (lane < 3) {
                prm.out_vals[n * 3 + lane] = lane == 0 ? cx : (lane == 1 ? cy : cz);
            }
            if (prm.out_jac != nullptr) {
                for (int col = lane; col < D; col += 32) {
                    const int2 rg = colRange[col];
                    const V4 a = wprefix[rg.x], b = wprefix[rg.y];
                    const T sx = b.x - a.x, sy = b.y - a.y, sz = b.z - a.z, m = b.w - a.w;
                    const V4 w = *reinterpret_cast<const V4 *>(tw + col * 8);
                    const V4 v = *reinterpret_cast<const V4 *>(tw + col * 8 + 4);
                    const bool any = rg.y > rg.x;       // a column whose joint moves no weighted point is structurally zero
                    const T jx = any ? (w.y * sz - w.z * sy + m * v.x) * invW : T(0);
                    const T jy = any ? (w.z * sx - w.x * sz + m * v.y) * invW : T(0);
                    const T jz = any ? (w.x * sy - w.y * sx + m * v.z) * invW : T(0);
                    if (prm.what == SKB_COM_STABILITY) prm.out_jac[n * D + col] = -(gx * jx + gy * jy + gz * jz);
                    else {
                        T *g = prm.out_jac + n * 3 * D + col;
                        g[0] = jx; g[D] = jy; g[2 * D] = jz;
                    }
                }
            }
            __syncwarp();
        }
    }
}

#define CUDA_OK(expr)                                                                          \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return set_error(std::string(#expr) + ": " + cudaGetErrorString(e__));             \
    } while (0)

static int align_up(int x, int a) { return (x + a - 1) / a * a; }
static int odd_quads(int words) { int q = (words + 3) / 4; if (q % 2 == 0) q += 1; return q * 4; }

template <typename T>
static int launch_com_t(const skb_plan *p, const skb_com_program *cp, const void *q, int64_t N, const void *dev_prims, int n_prims,
                        int what, void *out_vals, void *out_jac, cudaStream_t stream) {
    ComParams<T> prm;
    memset(&prm, 0, sizeof(prm));
    prm.I = reinterpret_cast<const int32_t *>(cp->dI);
    prm.F = reinterpret_cast<const T *>(sizeof(T) == 4 ? cp->dF32 : cp->dF64);
    prm.q = reinterpret_cast<const T *>(q);
    prm.N = N;
    prm.prims = reinterpret_cast<const DevPrim<T> *>(dev_prims);
    prm.n_prims = n_prims;
    prm.what = what;
    prm.out_vals = reinterpret_cast<T *>(out_vals);
    prm.out_jac = reinterpret_cast<T *>(out_jac);
    prm.i_total = (int)cp->I.size();
    prm.f_total = (int)cp->F.size();
    const int D = p->D;
    prm.D = D; prm.n_points = cp->n_points; prm.n_levels = cp->n_levels;
    prm.off_F = align_up(prm.i_total * 4, 16);
    prm.off_prims = align_up(prm.off_F + prm.f_total * (int)sizeof(T), 16);
    prm.off_lv = align_up(prm.off_prims + std::max(1, n_prims) * (int)sizeof(DevPrim<T>), 16);
    prm.n_nodes = cp->n_nodes;
    prm.off_warp = align_up(prm.off_lv + cp->n_nodes * 16, 128);
    const int al = 16 / (int)sizeof(T);
    prm.fstride = odd_quads(cp->n_nodes * 12);     // elements of T per configuration
    prm.tstride = odd_quads(D * 8);
    const size_t smem_limit = 227 * 1024;
    auto layout = [&](int g, int warps) {
        int e = 0;
        prm.woff_q = e;      e += align_up(g * D, al);
        prm.woff_sc = e;     e += align_up(2 * g * D, al);
        prm.woff_frames = e; e += g * prm.fstride;
        prm.woff_tw = e;     e += g * prm.tstride;
        prm.woff_prefix = e; e += 4 * (cp->n_points + 1);
        prm.warp_elems = align_up(e, al);
        return (size_t)prm.off_warp + (size_t)warps * prm.warp_elems * sizeof(T);
    };
    static const int cand[][2] = {{8, 8}, {4, 8}, {2, 8}, {1, 8}, {1, 4}, {1, 2}, {1, 1}};
    int G = 0, warps = 0;
    for (int ctas = 2; ctas >= 1 && G == 0; --ctas)
        for (auto &c : cand)
            if ((layout(c[0], c[1]) + 1024) * ctas <= smem_limit) { G = c[0]; warps = c[1]; break; }
    if (G == 0) return set_error("com kernel: shared-memory footprint exceeds 227 KB");
    const size_t smem = layout(G, warps);
    prm.G = G;
    prm.logG = 0;
    while ((1 << prm.logG) < G) ++prm.logG;
    prm.slots3 = 32 / (3 * G);
    prm.n_tiles = (N + G - 1) / G;
    auto kern = com_kernel<T>;
    CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0, sms = 148, dev = 0;
    CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, warps * 32, smem));
    if (per_sm < 1) return set_error("com kernel: zero occupancy");
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long want = (prm.n_tiles + warps - 1) / warps;
    const int grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)sms * per_sm));
    kern<<<grid, warps * 32, smem, stream>>>(prm);
    CUDA_OK(cudaGetLastError());
    ++g_launch_count;
    return 0;
}

int launch_com(const skb_plan *p, const skb_com_program *cp, int dtype, const void *q, int64_t N, const void *dev_prims,
               int n_prims, int what, void *out_vals, void *out_jac, void *stream_) {
    if (N <= 0) return 0;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (dtype == SKB_F32) return launch_com_t<float>(p, cp, q, N, dev_prims, n_prims, what, out_vals, out_jac, stream);
    if (dtype == SKB_F64) return launch_com_t<double>(p, cp, q, N, dev_prims, n_prims, what, out_vals, out_jac, stream);
    return set_error("bad dtype");
}

}  // namespace skb

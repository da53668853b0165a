// skb_pose.cu -- pose features (position + roll-pitch-yaw or quaternion) and their Jacobians on the warp-cooperative
// forward kinematics:
//
//     EndEffectorKinematicsMap.map with RotationType.RPY / XYZW          (skmp/kinematics.py:43-63,153-162)
//     PoseConstraint._evaluate = flatten(map(qs)) - target               (skmp/constraint.py:451-472)
//
// One warp owns a tile of G configurations.
//   phase 1   frames + twists, three lanes per (configuration, node)                     -> fk_phase (skb_device.cuh)
//   phase 1q  (XYZW) feature quaternions: the PRODUCT of the factors Q_pre * Qz(theta / 2) along the feature's chain like the
//             reference's, not a conversion of the rotation matrix (sign!); four lanes per (configuration, feature)
//   phase 1c  lane per (configuration, feature): position, rpy angles / quaternion -> values out, and the few scalars the
//             rotational Jacobian rows need into a small shared record
//   phase 2   lane per (configuration, column) of the tile, flattened: the lane keeps ITS joint twist (omega, v) in
//             registers and walks the features; row k of feature f is one value per lane, so every store instruction
//             writes consecutive columns of one output row (coalesced), structural zeros included.
//                 position rows   omega x p + v
//                 RPY rows        [ (cy wx + sy wy) / cp,  cy wy - sy wx,  wz + sp (cy wx + sy wy) / cp ]
//                 XYZW rows       1/2 Xi(q) omega
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>

#include "skb_device.cuh"
#include "skb_internal.h"

namespace skb {

template <typename T>
struct PoseParams {
    const int32_t *I;
    const T *F;
    const T *q;
    long long N;
    long long n_tiles;
    T *out_pts;
    T *out_jac;
    int i_total, f_total;
    int off_F, off_warp, warp_elems;
    int off_lv, n_nodes;
    int woff_q, woff_sc, woff_sch, woff_frames, woff_tw, woff_quat, woff_feat;
    int fstride, tstride, qstride, featstride;
    int G, logG, slots3;
    int D, n_feat, n_levels, rot_type, tdim;
    unsigned div_magic;                  // ceil(2^32 / D): i / D == __umulhi(i, div_magic) for i < 2^16
};

template <typename T>
__device__ __forceinline__ void qmul(const T a[4], const T b[4], T o[4]) {    // (x, y, z, w)
    o[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
    o[1] = a[3] * b[1] - a[0] * b[2] + a[1] * b[3] + a[2] * b[0];
    o[2] = a[3] * b[2] + a[0] * b[1] - a[1] * b[0] + a[2] * b[3];
    o[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
}

template <typename T, int ROT>
__global__ void __launch_bounds__(256) pose_kernel(const PoseParams<T> prm) {
    using V4 = Vec4<T>;
    using Rl = Real<T>;
    constexpr int TD = ROT == SKB_ROT_RPY ? 6 : 7;
    extern __shared__ __align__(16) unsigned char smem[];
    int32_t *sI = reinterpret_cast<int32_t *>(smem);
    T *sF = reinterpret_cast<T *>(smem + prm.off_F);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n_warps = blockDim.x >> 5;
    for (int i = tid; i < prm.i_total; i += blockDim.x) sI[i] = prm.I[i];
    for (int i = tid; i < prm.f_total; i += blockDim.x) sF[i] = prm.F[i];
    __syncthreads();

    const int D = prm.D, F = prm.n_feat, n_levels = prm.n_levels, G = prm.G, logG = prm.logG;
    const int fstride = prm.fstride, tstride = prm.tstride, qstride = prm.qstride, featstride = prm.featstride;
    const int32_t *nodeI = sI + sI[E_NODE_I];
    const int32_t *levelI = sI + sI[E_LEVEL_I];
    const int32_t *levelNodes = sI + sI[E_LEVELNODE_I];
    int4 *lvrec = reinterpret_cast<int4 *>(smem + prm.off_lv);
    build_level_records(lvrec, nodeI, levelNodes, prm.n_nodes);
    __syncthreads();
    const int4 *featI = reinterpret_cast<const int4 *>(sI + sI[E_FEAT_I]);
    const V4 *nodeF = reinterpret_cast<const V4 *>(sF + sI[E_NODE_F]);
    const int32_t *chainI = sI + sI[E_CHAIN_I];
    const V4 *chainQ = reinterpret_cast<const V4 *>(sF + sI[E_CHAINQ_F]);
    const int chain_len = sI[E_CHAIN_LEN];
    const T *featF = sF + sI[E_FEAT_F];                 // 20 per feature: off(3) pad | R(9) pad(3) | quat(4)

    T *wbase = reinterpret_cast<T *>(smem + prm.off_warp) + (size_t)warp * prm.warp_elems;
    T *wq = wbase + prm.woff_q;
    T *wsc = wbase + prm.woff_sc;
    T *wsch = wbase + prm.woff_sch;
    T *wframes = wbase + prm.woff_frames;
    T *wtw = wbase + prm.woff_tw;
    T *wquat = wbase + prm.woff_quat;
    T *wfeat = wbase + prm.woff_feat;
    const long long FTD = (long long)F * TD * D;

    for (long long tile = (long long)blockIdx.x * n_warps + warp; tile < prm.n_tiles; tile += (long long)gridDim.x * n_warps) {
        const long long n0 = tile * G;
        const int g_live = (int)min((long long)G, prm.N - n0);
        __syncwarp();
        for (int i = lane; i < g_live * D; i += 32) {
            const T th = prm.q[n0 * D + i];
            T sn, cs;
            wq[i] = th;
            if (ROT == SKB_ROT_XYZW) {
                // one sincos per joint value: the quaternion factors need the half angle, the frames get
                // sin = 2 sh ch, cos = (ch - sh)(ch + sh) from it (each within 2 ulp of the direct evaluation)
                T sh, ch;
                Rl::sincos(T(0.5) * th, &sh, &ch);
                wsch[2 * i] = sh; wsch[2 * i + 1] = ch;
                sn = T(2) * sh * ch; cs = (ch - sh) * (ch + sh);
            } else Rl::sincos(th, &sn, &cs);
            wsc[2 * i] = sn; wsc[2 * i + 1] = cs;
        }
        __syncwarp();
        fk_phase<T, true, false>(lane, G, logG, prm.slots3, g_live, D, n_levels, lvrec, levelI, nodeF, wq, wsc,
                                 wframes, fstride, wtw, tstride);

        // ------------------------------------------------------------ phase 1q: feature quaternions (XYZW)
        // Four lanes per (configuration, feature): each multiplies a quarter of the chain's factors  Q_k = C_k * Qz(q_k / 2)
        // (constants merged on the host), two shuffle rounds combine the partial products in chain order (the quaternion
        // product is associative), lane 0 of the group appends the feature's own offset.
        if (ROT == SKB_ROT_XYZW) {
            const int L4 = chain_len >> 2;
            for (int i0 = 0; i0 < g_live * F * 4; i0 += 32) {
                const int i = i0 + lane;
                const bool act = i < g_live * F * 4;
                const int part = i & 3, cf = act ? i >> 2 : 0, c = cf / F, f = cf - c * F;
                T acc[4] = {T(0), T(0), T(0), T(1)};
                const int32_t *cc = chainI + f * chain_len + part * L4;
                const V4 *cq = chainQ + f * chain_len + part * L4;
                for (int k = 0; k < L4; ++k) {
                    const int col = cc[k];
                    const V4 C4 = cq[k];
                    T sh = T(0), ch = T(1);
                    if (col >= 0) { sh = wsch[2 * (c * D + col)]; ch = wsch[2 * (c * D + col) + 1]; }
                    // C * Qz(theta / 2) = (x c + y s, y c - x s, z c + w s, w c - z s)
                    const T b[4] = {C4.x * ch + C4.y * sh, C4.y * ch - C4.x * sh, C4.z * ch + C4.w * sh, C4.w * ch - C4.z * sh};
                    T o[4];
                    qmul<T>(acc, b, o);
                    acc[0] = o[0]; acc[1] = o[1]; acc[2] = o[2]; acc[3] = o[3];
                }
#pragma unroll
                for (int step = 1; step <= 2; step <<= 1) {                  // (P0 P1)(P2 P3)
                    T nb[4], o[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) nb[j] = __shfl_down_sync(0xffffffffu, acc[j], step);
                    qmul<T>(acc, nb, o);
                    if ((part & (2 * step - 1)) == 0) { acc[0] = o[0]; acc[1] = o[1]; acc[2] = o[2]; acc[3] = o[3]; }
                }
                if (act && part == 0) {
                    T qf[4];
                    qmul<T>(acc, featF + f * 20 + 16, qf);
                    V4 out; out.x = qf[0]; out.y = qf[1]; out.z = qf[2]; out.w = qf[3];
                    reinterpret_cast<V4 *>(wquat + c * qstride)[f] = out;
                }
            }
            __syncwarp();
        }

        // ------------------------------------------------------------ phase 1c: feature poses -> values + Jacobian scalars
        for (int i = lane; i < g_live * F; i += 32) {
            const int c = i / F, f = i - c * F;
            const int node = featI[f].x;
            const T *ff = featF + f * 20;
            const V4 *fr = reinterpret_cast<const V4 *>(wframes + c * fstride);
            const V4 f0 = fr[3 * node], f1 = fr[3 * node + 1], f2 = fr[3 * node + 2];
            const T px = f0.w + f0.x * ff[0] + f0.y * ff[1] + f0.z * ff[2];
            const T py = f1.w + f1.x * ff[0] + f1.y * ff[1] + f1.z * ff[2];
            const T pz = f2.w + f2.x * ff[0] + f2.y * ff[1] + f2.z * ff[2];
            T *vr = prm.out_pts + ((n0 + c) * F + f) * TD;
            T *fb = wfeat + (c * F + f) * featstride;
            vr[0] = px; vr[1] = py; vr[2] = pz;
            fb[0] = px; fb[1] = py; fb[2] = pz;
            if (ROT == SKB_ROT_RPY) {
                const T *Ro = ff + 4;      // feature rotation = R_node * R_off; URDF roll-pitch-yaw of it
                const T r00 = f0.x * Ro[0] + f0.y * Ro[3] + f0.z * Ro[6];
                const T r10 = f1.x * Ro[0] + f1.y * Ro[3] + f1.z * Ro[6];
                const T r20 = f2.x * Ro[0] + f2.y * Ro[3] + f2.z * Ro[6];
                const T r21 = f2.x * Ro[1] + f2.y * Ro[4] + f2.z * Ro[7];
                const T r22 = f2.x * Ro[2] + f2.y * Ro[5] + f2.z * Ro[8];
                const T roll = Rl::atan2(r21, r22);
                const T sarg = -r20;
                const T pitch = sarg <= T(-1) ? T(-1.5707963267948966) : (sarg >= T(1) ? T(1.5707963267948966) : Rl::asin(sarg));
                const T yaw = Rl::atan2(r10, r00);
                vr[3] = roll; vr[4] = pitch; vr[5] = yaw;
                T sp, cp, sy, cy;
                Rl::sincos(pitch, &sp, &cp);
                Rl::sincos(yaw, &sy, &cy);
                fb[3] = cy; fb[4] = sy; fb[5] = T(1) / cp; fb[6] = sp;
            } else {
                const V4 qf = reinterpret_cast<const V4 *>(wquat + c * qstride)[f];       // phase 1q
                vr[3] = qf.x; vr[4] = qf.y; vr[5] = qf.z; vr[6] = qf.w;
                fb[3] = qf.x; fb[4] = qf.y; fb[5] = qf.z; fb[6] = qf.w;
            }
        }
        __syncwarp();
        if (prm.out_jac == nullptr) continue;

        // ------------------------------------------------------------ phase 2: lane per (configuration, column)
        for (int i = lane; i < g_live * D; i += 32) {
            const int c = (int)__umulhi((unsigned)i, prm.div_magic), d = i - c * D;       // i / D, exact for i < 2^16
            const T *tw = wtw + c * tstride + d * 8;
            const V4 w = *reinterpret_cast<const V4 *>(tw), v = *reinterpret_cast<const V4 *>(tw + 4);
            T *row = prm.out_jac + ((n0 + c) * FTD + d);            // row 0 of feature 0; every further row is D elements on
            const V4 *fb4 = reinterpret_cast<const V4 *>(wfeat + c * F * featstride);
            for (int f = 0; f < F; ++f) {
                const int4 fi = featI[f];
                const unsigned long long am = ((unsigned long long)(unsigned)fi.z << 32) | (unsigned)fi.y;
                const bool mine = (am >> d) & 1ull;
                const V4 fa = fb4[2 * f], fb = fb4[2 * f + 1];       // {px py pz a0} {a1 a2 a3 .}
                const T px = fa.x, py = fa.y, pz = fa.z, a0 = fa.w, a1 = fb.x, a2 = fb.y, a3 = fb.z;
                const T wx = mine ? w.x : T(0), wy = mine ? w.y : T(0), wz = mine ? w.z : T(0);
                const T vx = mine ? v.x : T(0), vy = mine ? v.y : T(0), vz = mine ? v.z : T(0);
                *row = wy * pz - wz * py + vx; row += D;
                *row = wz * px - wx * pz + vy; row += D;
                *row = wx * py - wy * px + vz; row += D;
                if (ROT == SKB_ROT_RPY) {
                    const T rr = (a0 * wx + a1 * wy) * a2;          // (cy wx + sy wy) / cp
                    *row = rr; row += D;
                    *row = a0 * wy - a1 * wx; row += D;
                    *row = wz + a3 * rr; row += D;
                } else {
                    *row = T(0.5) * (a3 * wx + (wy * a2 - wz * a1)); row += D;
                    *row = T(0.5) * (a3 * wy + (wz * a0 - wx * a2)); row += D;
                    *row = T(0.5) * (a3 * wz + (wx * a1 - wy * a0)); row += D;
                    *row = T(-0.5) * (wx * a0 + wy * a1 + wz * a2); row += D;
                }
            }
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

template <typename T, int ROT>
static int launch_pose_inst(const skb_plan *p, const skb_pose_program *pp, const void *q, int64_t N, void *out_pts, void *out_jac,
                            cudaStream_t stream) {
    PoseParams<T> prm;
    memset(&prm, 0, sizeof(prm));
    prm.I = reinterpret_cast<const int32_t *>(pp->dI);
    prm.F = reinterpret_cast<const T *>(sizeof(T) == 4 ? pp->dF32 : pp->dF64);
    prm.q = reinterpret_cast<const T *>(q);
    prm.N = N;
    prm.out_pts = reinterpret_cast<T *>(out_pts);
    prm.out_jac = reinterpret_cast<T *>(out_jac);
    prm.i_total = (int)pp->I.size();
    prm.f_total = (int)pp->F.size();
    const int D = p->D, F = p->F;
    prm.div_magic = (unsigned)((0x100000000ull + (unsigned long long)D - 1) / (unsigned long long)D);
    prm.D = D; prm.n_feat = F; prm.n_levels = pp->n_levels; prm.rot_type = ROT; prm.tdim = ROT == SKB_ROT_RPY ? 6 : 7;
    prm.off_F = align_up(prm.i_total * 4, 16);
    prm.off_lv = align_up(prm.off_F + prm.f_total * (int)sizeof(T), 16);
    prm.n_nodes = pp->n_nodes;
    prm.off_warp = align_up(prm.off_lv + pp->n_nodes * 16, 128);
    const int al = 16 / (int)sizeof(T);
    prm.fstride = odd_quads(pp->n_nodes * 12);
    prm.tstride = odd_quads(D * 8);
    prm.qstride = ROT == SKB_ROT_XYZW ? odd_quads(F * 4) : 0;          // one quaternion per feature
    prm.featstride = 8;
    const size_t smem_limit = 227 * 1024;
    auto layout = [&](int g, int warps) {
        int e = 0;
        prm.woff_q = e;      e += align_up(g * D, al);
        prm.woff_sc = e;     e += align_up(2 * g * D, al);
        prm.woff_sch = e;    e += ROT == SKB_ROT_XYZW ? align_up(2 * g * D, al) : 0;
        prm.woff_frames = e; e += g * prm.fstride;
        prm.woff_tw = e;     e += g * prm.tstride;
        prm.woff_quat = e;   e += g * prm.qstride;
        prm.woff_feat = e;   e += align_up(g * F * prm.featstride, al);
        prm.warp_elems = align_up(e, al);
        return (size_t)prm.off_warp + (size_t)warps * prm.warp_elems * sizeof(T);
    };
    // (G, warps per CTA) by the same issue-slot score as the step kernel: resident warps^0.75 / slots per configuration
    auto kern = pose_kernel<T, ROT>;
    CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_limit));
    static const int cand[][2] = {{8, 8}, {8, 4}, {4, 8}, {4, 6}, {4, 4}, {2, 8}, {2, 6}, {2, 4}, {1, 8}, {1, 4}, {1, 2}, {1, 1}};
    auto fk_slots = [&](int g) {
        const int s3 = 32 / (3 * g);
        double it = 0;
        for (int lv = 0; lv < pp->n_levels; ++lv) it += (pp->I[pp->I[E_LEVEL_I] + 2 * lv + 1] + s3 - 1) / s3;
        return (ROT == SKB_ROT_XYZW ? 130.0 : 90.0) * it / g;
    };
    int G = 0, warps = 0, per_sm = 0;
    double best_score = -1.0;
    for (auto &c : cand) {
        const size_t sm = layout(c[0], c[1]);
        if (sm > smem_limit) continue;
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, c[1] * 32, sm) != cudaSuccess || occ < 1) continue;
        const double score = std::pow((double)std::min(c[1] * occ, 48), 0.75) / (fk_slots(c[0]) + 10.0 * F * (ROT == SKB_ROT_RPY ? 6 : 7) * ((D + 31) / 32));
        if (score > best_score * 1.0001) { best_score = score; G = c[0]; warps = c[1]; per_sm = occ; }
    }
    if (G == 0) return set_error("pose kernel: shared-memory footprint exceeds 227 KB");
    const size_t smem = layout(G, warps);
    prm.G = G;
    prm.logG = 0;
    while ((1 << prm.logG) < G) ++prm.logG;
    prm.slots3 = 32 / (3 * G);
    prm.n_tiles = (N + G - 1) / G;
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long want = (prm.n_tiles + warps - 1) / warps;
    const int grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)sms * per_sm));
    kern<<<grid, warps * 32, smem, stream>>>(prm);
    CUDA_OK(cudaGetLastError());
    ++g_launch_count;
    return 0;
}

int launch_pose(const skb_plan *p, const skb_pose_program *pp, int dtype, const void *q, int64_t N, void *out_pts, void *out_jac,
                void *stream_) {
    if (N <= 0) return 0;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    const bool rpy = p->rot_type == SKB_ROT_RPY;
    if (dtype == SKB_F32)
        return rpy ? launch_pose_inst<float, SKB_ROT_RPY>(p, pp, q, N, out_pts, out_jac, stream)
                   : launch_pose_inst<float, SKB_ROT_XYZW>(p, pp, q, N, out_pts, out_jac, stream);
    if (dtype == SKB_F64)
        return rpy ? launch_pose_inst<double, SKB_ROT_RPY>(p, pp, q, N, out_pts, out_jac, stream)
                   : launch_pose_inst<double, SKB_ROT_XYZW>(p, pp, q, N, out_pts, out_jac, stream);
    return set_error("bad dtype");
}

}  // namespace skb

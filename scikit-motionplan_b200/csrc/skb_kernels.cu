// skb_kernels.cu -- hand-written sm_100a kernels of the collision-constraint hot path.
//
//   tree_kernel<T, DEPTH, KIND, WITH_JAC>
//     KIND_MAP       K1/K4: KinematicsMapBase.map          (skmp/kinematics.py:43-63)
//     KIND_COLLFREE  K1+K2: CollFreeConst._evaluate        (skmp/constraint.py:287-374)
//   pairwise_kernel<T>  K3: PairWiseSelfCollFreeConst._evaluate (skmp/constraint.py:749-778)
//   sdf_eval_kernel<T>      sdf(points)                    (skmp/constraint.py:353)
//
// Execution model (tree_kernel): one warp owns a tile of 32 configurations, one lane per
// configuration.  The compiled robot program (nodes = controlled joints in DFS preorder, with all
// fixed links / non-controlled joints / base pose folded into constants and every joint axis
// rotated onto +z) is staged in shared memory once per CTA and interpreted in lock-step by all
// lanes, so control flow is warp-uniform.  Per lane the current link transform and a *stack* of
// spatial joint twists (omega, v) of the controlled ancestors live in registers.  For a sphere at
// p with SDF gradient g the Jacobian row is the twist/wrench pairing
//        J[d] = omega_d . (p x g) + v_d . g        (6 FMA per non-zero entry)
// and structurally-zero entries are never computed: rows are assembled in a per-warp shared
// staging buffer whose zero columns are maintained by a host-compiled zero-fill schedule, then
// each lane ships its configuration's contiguous run to HBM with one 1-D bulk async store
// (cp.async.bulk shared::cta -> global, SASS UBLKCP), so sphere centres and per-sphere rows never
// round-trip through HBM and the write stream is issued by the TMA unit, not by LSU instructions.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <string>

#include "skb_device.cuh"
#include "skb_internal.h"

namespace skb {

template <typename T>
struct TreeParams {
    const int32_t *I;
    const T *F;
    const T *q;
    long long N;
    long long n_tiles;
    const DevPrim<T> *prims;
    int n_prims;
    int mode;
    T margin;
    T *out_vals;
    T *out_jac;
    uint8_t *out_valid;
    int i_total, f_total;
    int off_F, off_prims, off_warp, warp_elems;      // bytes, bytes, bytes, elements of T per warp
    int woff_q, woff_stage, woff_v, woff_best;       // element offsets inside a warp region
    int vp, dp;                                      // padded strides of the value / best rows
};

template <typename T, int DEPTH>
struct Lane {
    T R[9];
    T t[3];
    T Q[4];
    T w[DEPTH][3];
    T v[DEPTH][3];
};

template <typename T>
__device__ __forceinline__ void quat_mul(const T *a, const T *b, T *o) {
    T x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
    T y = a[3] * b[1] - a[0] * b[2] + a[1] * b[3] + a[2] * b[0];
    T z = a[3] * b[2] + a[0] * b[1] - a[1] * b[0] + a[2] * b[3];
    T w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
    o[0] = x; o[1] = y; o[2] = z; o[3] = w;
}

template <typename T, int DEPTH, int KIND, bool WITH_JAC>
__global__ void __launch_bounds__(256) tree_kernel(const TreeParams<T> prm) {
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

    const int n_nodes = sI[H_N_NODES], D = sI[H_D], n_feat = sI[H_N_FEAT], rot_type = sI[H_ROT_TYPE];
    const int n_buf = sI[H_N_BUF], RUNP = sI[H_RUNP], fstride = sI[H_FEAT_STRIDE];
    const int node_i = sI[H_NODE_I], chunk_i = sI[H_CHUNK_I], feat_f = sI[H_FEAT_F];
    const int ROWS = (KIND == KIND_MAP) ? sI[H_ROWS] : 1;
    const long long ROW = (long long)n_feat * ROWS * D;     // jacobian elements per configuration
    const long long VROW = (long long)n_feat * ROWS;        // value elements per configuration
    const bool closest = (KIND == KIND_COLLFREE) && prm.mode == SKB_MODE_CLOSEST;

    T *wbase = reinterpret_cast<T *>(smem + prm.off_warp) + (size_t)warp * prm.warp_elems;
    T *wq = wbase + prm.woff_q;
    T *wstage = wbase + prm.woff_stage;
    T *wv = wbase + prm.woff_v + lane * prm.vp;
    T *wbest = wbase + prm.woff_best;

    if (WITH_JAC && !closest) {
        for (int i = lane; i < n_buf * 32 * RUNP; i += 32) wstage[i] = T(0);
    }
    __syncwarp();

    T slots[MAX_SLOTS][16];
    Lane<T, DEPTH> L;

    for (long long tile = (long long)blockIdx.x * n_warps + warp; tile < prm.n_tiles; tile += (long long)gridDim.x * n_warps) {
        const long long n0 = tile * 32, n = n0 + lane;
        const bool live = n < prm.N;
        const int rows_here = (int)min((long long)32, prm.N - n0);
        __syncwarp();
        for (int i = lane; i < rows_here * D; i += 32) wq[i] = prm.q[n0 * D + i];
        __syncwarp();
        const T *myq = wq + min(lane, rows_here - 1) * D;

        T vmin = Rl::inf();
        T best = Rl::inf();
        int best_idx = 0x7fffffff;
        if (closest && WITH_JAC)
            for (int c = 0; c < D; ++c) wbest[lane * prm.dp + c] = T(0);

        for (int ni = 0; ni < n_nodes; ++ni) {
            const int32_t *nd = sI + node_i + ni * NODE_I;
            const T *nf = sF + ni * NODE_F;
            const int type = nd[N_TYPE], src = nd[N_SRC], di = nd[N_DI];
            // ---- Y = parent * pre
            T Y[9], yt[3], yq[4];
            if (src == SRC_IDENTITY) {
#pragma unroll
                for (int k = 0; k < 9; ++k) Y[k] = nf[NF_R + k];
#pragma unroll
                for (int k = 0; k < 3; ++k) yt[k] = nf[NF_T + k];
#pragma unroll
                for (int k = 0; k < 4; ++k) yq[k] = nf[NF_Q + k];
            } else {
                if (src >= 0) {
#pragma unroll
                    for (int k = 0; k < 9; ++k) L.R[k] = slots[src][k];
#pragma unroll
                    for (int k = 0; k < 3; ++k) L.t[k] = slots[src][9 + k];
#pragma unroll
                    for (int k = 0; k < 4; ++k) L.Q[k] = slots[src][12 + k];
                }
#pragma unroll
                for (int r = 0; r < 3; ++r) {
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        Y[r * 3 + c] = L.R[r * 3 + 0] * nf[NF_R + c] + L.R[r * 3 + 1] * nf[NF_R + 3 + c] + L.R[r * 3 + 2] * nf[NF_R + 6 + c];
                    yt[r] = L.t[r] + L.R[r * 3 + 0] * nf[NF_T + 0] + L.R[r * 3 + 1] * nf[NF_T + 1] + L.R[r * 3 + 2] * nf[NF_T + 2];
                }
                if (KIND == KIND_MAP && rot_type == SKB_ROT_XYZW) quat_mul<T>(L.Q, nf + NF_Q, yq);
            }
            // ---- joint motion about / along local +z, and the joint's spatial twist
            T tw[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
            if (type == NODE_REVOLUTE) {
                const T th = myq[nd[N_COL]];
                T s, c;
                Rl::sincos(th, &s, &c);
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const T a = Y[r * 3 + 0], b = Y[r * 3 + 1];
                    L.R[r * 3 + 0] = c * a + s * b;
                    L.R[r * 3 + 1] = c * b - s * a;
                    L.R[r * 3 + 2] = Y[r * 3 + 2];
                    L.t[r] = yt[r];
                }
                if (WITH_JAC) {
                    const T ax = Y[2], ay = Y[5], az = Y[8];
                    tw[0] = ax; tw[1] = ay; tw[2] = az;
                    tw[3] = yt[1] * az - yt[2] * ay;     // v = o x a
                    tw[4] = yt[2] * ax - yt[0] * az;
                    tw[5] = yt[0] * ay - yt[1] * ax;
                }
                if (KIND == KIND_MAP && rot_type == SKB_ROT_XYZW) {
                    T sh, ch;
                    Rl::sincos(T(0.5) * th, &sh, &ch);
                    const T qz[4] = {T(0), T(0), sh, ch};
                    quat_mul<T>(yq, qz, L.Q);
                }
            } else {
#pragma unroll
                for (int k = 0; k < 9; ++k) L.R[k] = Y[k];
#pragma unroll
                for (int k = 0; k < 3; ++k) L.t[k] = yt[k];
#pragma unroll
                for (int k = 0; k < 4; ++k) L.Q[k] = yq[k];
                if (type == NODE_PRISMATIC) {
                    const T th = myq[nd[N_COL]];
                    L.t[0] += th * Y[2]; L.t[1] += th * Y[5]; L.t[2] += th * Y[8];
                    tw[3] = Y[2]; tw[4] = Y[5]; tw[5] = Y[8];
                }
            }
            if (WITH_JAC && di >= 0) {
#pragma unroll
                for (int i = 0; i < DEPTH; ++i)
                    if (i == di) {
                        L.w[i][0] = tw[0]; L.w[i][1] = tw[1]; L.w[i][2] = tw[2];
                        L.v[i][0] = tw[3]; L.v[i][1] = tw[4]; L.v[i][2] = tw[5];
                    }
            }
            const int save = nd[N_SAVE];
            if (save >= 0) {
#pragma unroll
                for (int k = 0; k < 9; ++k) slots[save][k] = L.R[k];
#pragma unroll
                for (int k = 0; k < 3; ++k) slots[save][9 + k] = L.t[k];
#pragma unroll
                for (int k = 0; k < 4; ++k) slots[save][12 + k] = L.Q[k];
            }

            const int ch_begin = nd[N_CHUNK_BEGIN], ch_end = nd[N_CHUNK_END];
            if (ch_begin == ch_end) continue;
            const int active = di + 1;
            int cols[DEPTH];
            if (WITH_JAC) {
#pragma unroll
                for (int i = 0; i < DEPTH; ++i) cols[i] = (i < active) ? sI[nd[N_ANC] + i] : 0;
            }

            for (int ci = ch_begin; ci < ch_end; ++ci) {
                const int32_t *ch = sI + chunk_i + ci * CHUNK_I;
                const int cnt = ch[C_COUNT], out0 = ch[C_OUT0], flags = ch[C_PAD];
                T *st = wstage + ((size_t)ch[C_BUF] * 32 + lane) * RUNP;
                if (WITH_JAC && !closest) {
                    // the staging buffer may still be being read by the bulk store issued n_buf chunks ago
                    if (n_buf == 1) bulk_wait_read<0>();
                    else if (n_buf == 2) bulk_wait_read<1>();
                    else if (n_buf == 3) bulk_wait_read<2>();
                    else bulk_wait_read<3>();
                    const int nz = ch[C_ZERO_COUNT];
                    for (int z = 0; z < nz; ++z) {
                        const int zc = sI[ch[C_ZERO_BEGIN] + z];
                        for (int s = 0; s < cnt * ROWS; ++s) st[s * D + zc] = T(0);
                    }
                }
                const T *fr = sF + feat_f + (size_t)ch[C_FEAT_BEGIN] * fstride;
                for (int fi = 0; fi < cnt; ++fi, fr += fstride) {
                    const T px = L.t[0] + L.R[0] * fr[0] + L.R[1] * fr[1] + L.R[2] * fr[2];
                    const T py = L.t[1] + L.R[3] * fr[0] + L.R[4] * fr[1] + L.R[5] * fr[2];
                    const T pz = L.t[2] + L.R[6] * fr[0] + L.R[7] * fr[1] + L.R[8] * fr[2];
                    if (KIND == KIND_COLLFREE) {
                        T gx, gy, gz;
                        const T d = sdf_union<T>(sP, prm.n_prims, px, py, pz, gx, gy, gz);
                        const T val = d - fr[3] - prm.margin;
                        vmin = (val < vmin || val != val) ? val : vmin;      // NaN sticks (np.min)
                        // wrench of the unit "force" g applied at p: (p x g, g)
                        const T mx = py * gz - pz * gy, my = pz * gx - px * gz, mz = px * gy - py * gx;
                        if (!closest) {
                            wv[fi] = val;
                            if (WITH_JAC) {
                                T *row = st + fi * D;
#pragma unroll
                                for (int i = 0; i < DEPTH; ++i)
                                    if (i < active)
                                        row[cols[i]] = L.w[i][0] * mx + L.w[i][1] * my + L.w[i][2] * mz +
                                                       L.v[i][0] * gx + L.v[i][1] * gy + L.v[i][2] * gz;
                            }
                        } else {
                            // argmin over spheres of sdf - radius, first index wins ties (np.argmin)
                            const int idx = out0 + fi;
                            const bool better = (val < best) || (val == best && idx < best_idx);
                            if (better) { best = val; best_idx = idx; }
                            if (WITH_JAC && __any_sync(0xffffffffu, better)) {
                                T *row = wbest + lane * prm.dp;
                                if (better)
                                    for (int c = 0; c < D; ++c) row[c] = T(0);
#pragma unroll
                                for (int i = 0; i < DEPTH; ++i)
                                    if (i < active) {
                                        const T j = L.w[i][0] * mx + L.w[i][1] * my + L.w[i][2] * mz +
                                                    L.v[i][0] * gx + L.v[i][1] * gy + L.v[i][2] * gz;
                                        if (better) row[cols[i]] = j;
                                    }
                            }
                        }
                    } else {   // KIND_MAP
                        T *vr = wv + fi * ROWS;
                        vr[0] = px; vr[1] = py; vr[2] = pz;
                        T *row = st + (size_t)fi * ROWS * D;
                        if (WITH_JAC) {
#pragma unroll
                            for (int i = 0; i < DEPTH; ++i)
                                if (i < active) {
                                    const int c = cols[i];
                                    row[c] = L.w[i][1] * pz - L.w[i][2] * py + L.v[i][0];
                                    row[D + c] = L.w[i][2] * px - L.w[i][0] * pz + L.v[i][1];
                                    row[2 * D + c] = L.w[i][0] * py - L.w[i][1] * px + L.v[i][2];
                                }
                        }
                        if (rot_type == SKB_ROT_RPY) {
                            // feature rotation = R * Roff; URDF roll-pitch-yaw of it
                            const T *Ro = fr + 4;
                            const T r00 = L.R[0] * Ro[0] + L.R[1] * Ro[3] + L.R[2] * Ro[6];
                            const T r10 = L.R[3] * Ro[0] + L.R[4] * Ro[3] + L.R[5] * Ro[6];
                            const T r20 = L.R[6] * Ro[0] + L.R[7] * Ro[3] + L.R[8] * Ro[6];
                            const T r21 = L.R[6] * Ro[1] + L.R[7] * Ro[4] + L.R[8] * Ro[7];
                            const T r22 = L.R[6] * Ro[2] + L.R[7] * Ro[5] + L.R[8] * Ro[8];
                            const T roll = Rl::atan2(r21, r22);
                            const T sarg = -r20;
                            const T pitch = sarg <= T(-1) ? T(-1.5707963267948966) : (sarg >= T(1) ? T(1.5707963267948966) : Rl::asin(sarg));
                            const T yaw = Rl::atan2(r10, r00);
                            vr[3] = roll; vr[4] = pitch; vr[5] = yaw;
                            if (WITH_JAC) {
                                T sp, cp, sy, cy;
                                Rl::sincos(pitch, &sp, &cp);
                                Rl::sincos(yaw, &sy, &cy);
                                const T icp = T(1) / cp;
#pragma unroll
                                for (int i = 0; i < DEPTH; ++i)
                                    if (i < active) {
                                        const int c = cols[i];
                                        const T rr = (cy * L.w[i][0] + sy * L.w[i][1]) * icp;
                                        row[3 * D + c] = rr;
                                        row[4 * D + c] = cy * L.w[i][1] - sy * L.w[i][0];
                                        row[5 * D + c] = L.w[i][2] + sp * rr;
                                    }
                            }
                        } else if (rot_type == SKB_ROT_XYZW) {
                            T qf[4];
                            quat_mul<T>(L.Q, fr + 13, qf);
                            vr[3] = qf[0]; vr[4] = qf[1]; vr[5] = qf[2]; vr[6] = qf[3];
                            if (WITH_JAC) {
#pragma unroll
                                for (int i = 0; i < DEPTH; ++i)
                                    if (i < active) {
                                        const int c = cols[i];
                                        const T wx = L.w[i][0], wy = L.w[i][1], wz = L.w[i][2];
                                        row[3 * D + c] = T(0.5) * (qf[3] * wx + (wy * qf[2] - wz * qf[1]));
                                        row[4 * D + c] = T(0.5) * (qf[3] * wy + (wz * qf[0] - wx * qf[2]));
                                        row[5 * D + c] = T(0.5) * (qf[3] * wz + (wx * qf[1] - wy * qf[0]));
                                        row[6 * D + c] = T(-0.5) * (wx * qf[0] + wy * qf[1] + wz * qf[2]);
                                    }
                            }
                        }
                    }
                }   // features of the chunk
                if (closest) continue;

                // ---- ship the chunk: jacobian run via bulk async store, values via vector stores
                if (WITH_JAC) {
                    const int run = cnt * ROWS * D;
                    T *gdst = prm.out_jac + n * ROW + (long long)out0 * ROWS * D;
                    if (flags & 1) {
                        fence_async_smem();
                        if (live) bulk_store(gdst, st, (uint32_t)(run * sizeof(T)));
                    } else {
                        __syncwarp();
                        const T *sb = wstage + (size_t)ch[C_BUF] * 32 * RUNP;
                        for (int c = 0; c < rows_here; ++c) {
                            T *g = prm.out_jac + (n0 + c) * ROW + (long long)out0 * ROWS * D;
                            for (int k = lane; k < run; k += 32) g[k] = sb[(size_t)c * RUNP + k];
                        }
                        __syncwarp();
                    }
                    bulk_commit();
                }
                if (prm.out_vals != nullptr && live) {
                    const int nv = cnt * ROWS;
                    T *g = prm.out_vals + n * VROW + (long long)out0 * ROWS;
                    if (flags & 2) {
                        constexpr int VEC = 16 / (int)sizeof(T);
                        for (int k = 0; k < nv; k += VEC)
                            *reinterpret_cast<int4 *>(g + k) = *reinterpret_cast<const int4 *>(wv + k);
                    } else {
                        for (int k = 0; k < nv; ++k) g[k] = wv[k];
                    }
                }
            }   // chunks
        }       // nodes

        if (KIND == KIND_COLLFREE) {
            if (closest) {
                if (prm.out_vals != nullptr && live) prm.out_vals[n] = best;
                if (WITH_JAC) {
                    __syncwarp();
                    T *g = prm.out_jac + n0 * D;
                    for (int i = lane; i < rows_here * D; i += 32) g[i] = wbest[(i / D) * prm.dp + (i % D)];
                    __syncwarp();
                }
            }
            if (prm.out_valid != nullptr && live) prm.out_valid[n] = vmin > T(0) ? 1 : 0;
        }
    }   // tiles
    if (WITH_JAC) bulk_wait_read<0>();   // shared memory must outlive the last bulk reads
}

// ------------------------------------------------------------------ sdf(points)
template <typename T>
__global__ void sdf_eval_kernel(const DevPrim<T> *prims, int n_prims, const T *pts, long long M, T *out_vals, T *out_grads) {
    extern __shared__ __align__(16) unsigned char smem[];
    DevPrim<T> *sP = reinterpret_cast<DevPrim<T> *>(smem);
    const int words = n_prims * (int)(sizeof(DevPrim<T>) / 4);
    for (int i = threadIdx.x; i < words; i += blockDim.x)
        reinterpret_cast<int32_t *>(sP)[i] = reinterpret_cast<const int32_t *>(prims)[i];
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (long long)gridDim.x * blockDim.x) {
        T gx, gy, gz;
        const T d = sdf_union<T>(sP, n_prims, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], gx, gy, gz);
        out_vals[i] = d;
        if (out_grads) { out_grads[3 * i] = gx; out_grads[3 * i + 1] = gy; out_grads[3 * i + 2] = gz; }
    }
}

// ------------------------------------------------------------------ pairwise self-collision (K3)
// One lane per configuration.  Every node is evaluated; sphere centres and node twists are kept
// per lane in shared memory ([word][lane] layout, bank-conflict free).  For a pair (i, j) with
// delta = xi - xj the value is |delta|^2 - (ri+rj)^2 and only joints on the tree path between the
// two spheres contribute:  d/dq_d = +-2 (omega_d . (x x delta) + v_d . delta).
template <typename T>
struct PairParams {
    const int32_t *I;
    const T *F;
    const T *q;
    long long N;
    long long n_tiles;
    int mode;
    T *out_vals;
    T *out_jac;
    uint8_t *out_valid;
    int i_total, f_total;
    int off_F, off_warp, warp_elems;
    int woff_q, woff_pos, woff_tw, woff_row;
    int dp;
};

template <typename T, bool WITH_JAC>
__global__ void __launch_bounds__(128) pairwise_kernel(const PairParams<T> prm) {
    using Rl = Real<T>;
    extern __shared__ __align__(16) unsigned char smem[];
    int32_t *sI = reinterpret_cast<int32_t *>(smem);
    T *sF = reinterpret_cast<T *>(smem + prm.off_F);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n_warps = blockDim.x >> 5;
    for (int i = tid; i < prm.i_total; i += blockDim.x) sI[i] = prm.I[i];
    for (int i = tid; i < prm.f_total; i += blockDim.x) sF[i] = prm.F[i];
    __syncthreads();
    const int n_nodes = sI[P_N_NODES], D = sI[P_D], P = sI[P_N_PAIRS];
    const int node_i = sI[P_NODE_I], featnode_i = sI[P_FEATNODE_I], pair_i = sI[P_PAIR_I], class_i = sI[P_CLASS_I];
    const int feat_f = sI[P_FEAT_F], pair_f = sI[P_PAIR_F];
    const bool closest = prm.mode == SKB_MODE_CLOSEST;

    T *wbase = reinterpret_cast<T *>(smem + prm.off_warp) + (size_t)warp * prm.warp_elems;
    T *wq = wbase + prm.woff_q;
    T *wpos = wbase + prm.woff_pos + lane;       // [(f*3+k)*32]
    T *wtw = wbase + prm.woff_tw + lane;         // [(node*6+k)*32]
    T *wrow = wbase + prm.woff_row + lane * prm.dp;
    T slots[MAX_SLOTS][12];

    for (long long tile = (long long)blockIdx.x * n_warps + warp; tile < prm.n_tiles; tile += (long long)gridDim.x * n_warps) {
        const long long n0 = tile * 32, n = n0 + lane;
        const bool live = n < prm.N;
        const int rows_here = (int)min((long long)32, prm.N - n0);
        __syncwarp();
        for (int i = lane; i < rows_here * D; i += 32) wq[i] = prm.q[n0 * D + i];
        __syncwarp();
        const T *myq = wq + min(lane, rows_here - 1) * D;

        T R[9], t[3];
        for (int ni = 0; ni < n_nodes; ++ni) {
            const int32_t *nd = sI + node_i + ni * NODE_I;
            const T *nf = sF + ni * NODE_F;
            const int type = nd[N_TYPE], src = nd[N_SRC];
            T Y[9], yt[3];
            if (src == SRC_IDENTITY) {
#pragma unroll
                for (int k = 0; k < 9; ++k) Y[k] = nf[NF_R + k];
#pragma unroll
                for (int k = 0; k < 3; ++k) yt[k] = nf[NF_T + k];
            } else {
                if (src >= 0) {
#pragma unroll
                    for (int k = 0; k < 9; ++k) R[k] = slots[src][k];
#pragma unroll
                    for (int k = 0; k < 3; ++k) t[k] = slots[src][9 + k];
                }
#pragma unroll
                for (int r = 0; r < 3; ++r) {
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        Y[r * 3 + c] = R[r * 3 + 0] * nf[NF_R + c] + R[r * 3 + 1] * nf[NF_R + 3 + c] + R[r * 3 + 2] * nf[NF_R + 6 + c];
                    yt[r] = t[r] + R[r * 3 + 0] * nf[NF_T + 0] + R[r * 3 + 1] * nf[NF_T + 1] + R[r * 3 + 2] * nf[NF_T + 2];
                }
            }
            T tw[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
            if (type == NODE_REVOLUTE) {
                T s, c;
                Rl::sincos(myq[nd[N_COL]], &s, &c);
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const T a = Y[r * 3 + 0], b = Y[r * 3 + 1];
                    R[r * 3 + 0] = c * a + s * b; R[r * 3 + 1] = c * b - s * a; R[r * 3 + 2] = Y[r * 3 + 2];
                    t[r] = yt[r];
                }
                const T ax = Y[2], ay = Y[5], az = Y[8];
                tw[0] = ax; tw[1] = ay; tw[2] = az;
                tw[3] = yt[1] * az - yt[2] * ay; tw[4] = yt[2] * ax - yt[0] * az; tw[5] = yt[0] * ay - yt[1] * ax;
            } else {
#pragma unroll
                for (int k = 0; k < 9; ++k) R[k] = Y[k];
#pragma unroll
                for (int k = 0; k < 3; ++k) t[k] = yt[k];
                if (type == NODE_PRISMATIC) {
                    const T th = myq[nd[N_COL]];
                    t[0] += th * Y[2]; t[1] += th * Y[5]; t[2] += th * Y[8];
                    tw[3] = Y[2]; tw[4] = Y[5]; tw[5] = Y[8];
                }
            }
            if (WITH_JAC) {
#pragma unroll
                for (int k = 0; k < 6; ++k) wtw[(ni * 6 + k) * 32] = tw[k];
            }
            const int save = nd[N_SAVE];
            if (save >= 0) {
#pragma unroll
                for (int k = 0; k < 9; ++k) slots[save][k] = R[k];
#pragma unroll
                for (int k = 0; k < 3; ++k) slots[save][9 + k] = t[k];
            }
            // features bound to this node (any output index)
            for (int fk = nd[N_CHUNK_BEGIN]; fk < nd[N_CHUNK_END]; ++fk) {
                const int f = sI[featnode_i + fk];
                const T *fr = sF + feat_f + f * 4;
                wpos[(f * 3 + 0) * 32] = t[0] + R[0] * fr[0] + R[1] * fr[1] + R[2] * fr[2];
                wpos[(f * 3 + 1) * 32] = t[1] + R[3] * fr[0] + R[4] * fr[1] + R[5] * fr[2];
                wpos[(f * 3 + 2) * 32] = t[2] + R[6] * fr[0] + R[7] * fr[1] + R[8] * fr[2];
            }
        }

        T best = Rl::inf();
        int best_pair = -1;
        T vmin = Rl::inf();
        for (int pi = 0; pi < P; ++pi) {
            const int32_t *pr = sI + pair_i + pi * 4;
            const int fi = pr[0], fj = pr[1];
            const T xi = wpos[(fi * 3 + 0) * 32], yi = wpos[(fi * 3 + 1) * 32], zi = wpos[(fi * 3 + 2) * 32];
            const T xj = wpos[(fj * 3 + 0) * 32], yj = wpos[(fj * 3 + 1) * 32], zj = wpos[(fj * 3 + 2) * 32];
            const T dx = xi - xj, dy = yi - yj, dz = zi - zj;
            const T val = dx * dx + dy * dy + dz * dz - sF[pair_f + pi];
            vmin = (val < vmin || val != val) ? val : vmin;      // NaN sticks (np.min)
            if (closest) {
                if (val < best) { best = val; best_pair = pi; }   // first index wins ties (np.argmin)
                continue;
            }
            if (live && prm.out_vals) prm.out_vals[n * P + pi] = val;
            if (WITH_JAC && live) {
                T *g = prm.out_jac + (n * P + pi) * D;
                for (int c = 0; c < D; ++c) g[c] = T(0);
                const int32_t *cl = sI + class_i + pr[2] * 2;
                const int32_t *tm = sI + cl[0];
                for (int k = 0; k < cl[1]; ++k, tm += 3) {
                    const T *w = wtw + tm[0] * 6 * 32;
                    const bool ui = tm[2] > 0;
                    const T ax = ui ? xi : xj, ay = ui ? yi : yj, az = ui ? zi : zj;
                    const T mx = ay * dz - az * dy, my = az * dx - ax * dz, mz = ax * dy - ay * dx;
                    const T j = w[0] * mx + w[32] * my + w[64] * mz + w[96] * dx + w[128] * dy + w[160] * dz;
                    g[tm[1]] = (ui ? T(2) : T(-2)) * j;
                }
            }
        }
        if (closest) {
            if (live && prm.out_vals) prm.out_vals[n] = best;
            if (WITH_JAC) {
                for (int c = 0; c < D; ++c) wrow[c] = T(0);
                if (best_pair >= 0) {
                    const int32_t *pr = sI + pair_i + best_pair * 4;
                    const int fi = pr[0], fj = pr[1];
                    const T xi = wpos[(fi * 3 + 0) * 32], yi = wpos[(fi * 3 + 1) * 32], zi = wpos[(fi * 3 + 2) * 32];
                    const T xj = wpos[(fj * 3 + 0) * 32], yj = wpos[(fj * 3 + 1) * 32], zj = wpos[(fj * 3 + 2) * 32];
                    const T dx = xi - xj, dy = yi - yj, dz = zi - zj;
                    const int32_t *cl = sI + class_i + pr[2] * 2;
                    const int32_t *tm = sI + cl[0];
                    for (int k = 0; k < cl[1]; ++k, tm += 3) {
                        const T *w = wtw + tm[0] * 6 * 32;
                        const bool ui = tm[2] > 0;
                        const T ax = ui ? xi : xj, ay = ui ? yi : yj, az = ui ? zi : zj;
                        const T mx = ay * dz - az * dy, my = az * dx - ax * dz, mz = ax * dy - ay * dx;
                        const T j = w[0] * mx + w[32] * my + w[64] * mz + w[96] * dx + w[128] * dy + w[160] * dz;
                        wrow[tm[1]] = (ui ? T(2) : T(-2)) * j;
                    }
                }
                __syncwarp();
                T *g = prm.out_jac + n0 * D;
                const T *rb = wbase + prm.woff_row;
                for (int i = lane; i < rows_here * D; i += 32) g[i] = rb[(i / D) * prm.dp + (i % D)];
                __syncwarp();
            }
        }
        if (prm.out_valid != nullptr && live) prm.out_valid[n] = vmin > T(0) ? 1 : 0;
    }
}

// ------------------------------------------------------------------ launchers
#define CUDA_OK(expr)                                                                          \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return set_error(std::string(#expr) + ": " + cudaGetErrorString(e__));             \
    } while (0)

static int align_up(int x, int a) { return (x + a - 1) / a * a; }

static int sm_count_cached() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

template <typename T, int DEPTH, int KIND, bool WITH_JAC>
static int launch_tree_inst(const skb_plan *p, const skb_schedule *sch, const void *q, int64_t N, const void *dev_prims,
                            int n_prims, double margin, int mode, void *out_vals, void *out_jac, uint8_t *out_valid,
                            cudaStream_t stream) {
    TreeParams<T> prm;
    prm.I = reinterpret_cast<const int32_t *>(sch->dI);
    prm.F = reinterpret_cast<const T *>(sizeof(T) == 4 ? sch->dF32 : sch->dF64);
    prm.q = reinterpret_cast<const T *>(q);
    prm.N = N;
    prm.n_tiles = (N + 31) / 32;
    prm.prims = reinterpret_cast<const DevPrim<T> *>(dev_prims);
    prm.n_prims = n_prims;
    prm.mode = mode;
    prm.margin = (T)margin;
    prm.out_vals = reinterpret_cast<T *>(out_vals);
    prm.out_jac = reinterpret_cast<T *>(out_jac);
    prm.out_valid = out_valid;
    prm.i_total = (int)sch->I.size();
    prm.f_total = (int)sch->F.size();
    const int D = p->D;
    const int rows = sch->rows;
    const bool closest = (KIND == KIND_COLLFREE) && mode == SKB_MODE_CLOSEST;
    prm.off_F = align_up(prm.i_total * 4, 16);
    prm.off_prims = align_up(prm.off_F + prm.f_total * (int)sizeof(T), 16);
    prm.off_warp = align_up(prm.off_prims + n_prims * (int)sizeof(DevPrim<T>), 128);
    // per-warp region (elements of T); every sub-region starts 16-byte aligned
    const int al = 16 / (int)sizeof(T);
    int e = 0;
    prm.woff_q = e;      e += align_up(32 * D, al);
    prm.woff_stage = e;  e += (WITH_JAC && !closest) ? sch->n_buf * 32 * sch->runp : 0;
    prm.vp = align_up(std::max(1, sch->chunk_cap * rows), al) + al;     // +al keeps lanes on distinct 16-byte banks
    if ((prm.vp / al) % 2 == 0) prm.vp += al;
    prm.woff_v = e;      e += 32 * prm.vp;
    prm.dp = D | 1;
    prm.woff_best = e;   e += closest ? align_up(32 * prm.dp, al) : 0;
    prm.warp_elems = align_up(e, al);
    int warps = p->tune_warps > 0 ? p->tune_warps : 4;
    const size_t warp_bytes = (size_t)prm.warp_elems * sizeof(T);
    const size_t smem_limit = 227 * 1024;
    while (warps > 1 && prm.off_warp + warps * warp_bytes > smem_limit) warps /= 2;
    const size_t smem = prm.off_warp + warps * warp_bytes;
    if (smem > smem_limit) return set_error("tree kernel: shared-memory footprint exceeds 227 KB (robot too large for one warp's staging)");
    auto kern = tree_kernel<T, DEPTH, KIND, WITH_JAC>;
    CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, warps * 32, smem));
    if (per_sm < 1) return set_error("tree kernel: zero occupancy");
    if (p->tune_ctas > 0) per_sm = std::min(per_sm, p->tune_ctas);
    const long long want = (prm.n_tiles + warps - 1) / warps;
    const int grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)sm_count_cached() * per_sm));
    kern<<<grid, warps * 32, smem, stream>>>(prm);
    CUDA_OK(cudaGetLastError());
    ++g_launch_count;
    return 0;
}

template <typename T, int KIND, bool WITH_JAC>
static int launch_tree_depth(const skb_plan *p, const skb_schedule *sch, const void *q, int64_t N, const void *dev_prims,
                             int n_prims, double margin, int mode, void *out_vals, void *out_jac, uint8_t *out_valid,
                             cudaStream_t stream) {
    const int depth = WITH_JAC ? p->max_active : 0;
    if (depth <= 8)
        return launch_tree_inst<T, 8, KIND, WITH_JAC>(p, sch, q, N, dev_prims, n_prims, margin, mode, out_vals, out_jac, out_valid, stream);
    if (depth <= 20)
        return launch_tree_inst<T, 20, KIND, WITH_JAC>(p, sch, q, N, dev_prims, n_prims, margin, mode, out_vals, out_jac, out_valid, stream);
    return launch_tree_inst<T, 48, KIND, WITH_JAC>(p, sch, q, N, dev_prims, n_prims, margin, mode, out_vals, out_jac, out_valid, stream);
}

template <typename T>
static int launch_tree_t(const skb_plan *p, const skb_schedule *sch, int kind, const void *q, int64_t N, const void *dev_prims,
                         int n_prims, double margin, int mode, void *out_vals, void *out_jac, uint8_t *out_valid,
                         cudaStream_t stream) {
    const bool wj = out_jac != nullptr;
    if (kind == KIND_MAP) {
        if (wj) return launch_tree_depth<T, KIND_MAP, true>(p, sch, q, N, dev_prims, n_prims, margin, mode, out_vals, out_jac, out_valid, stream);
        return launch_tree_depth<T, KIND_MAP, false>(p, sch, q, N, dev_prims, n_prims, margin, mode, out_vals, out_jac, out_valid, stream);
    }
    if (wj) return launch_tree_depth<T, KIND_COLLFREE, true>(p, sch, q, N, dev_prims, n_prims, margin, mode, out_vals, out_jac, out_valid, stream);
    return launch_tree_depth<T, KIND_COLLFREE, false>(p, sch, q, N, dev_prims, n_prims, margin, mode, out_vals, out_jac, out_valid, stream);
}

int launch_tree(const skb_plan *p, const skb_schedule *sch, int kind, int dtype, const void *q, int64_t N,
                const void *dev_prims, int n_prims, double margin, int mode, void *out_vals, void *out_jac,
                uint8_t *out_valid, void *stream) {
    if (N <= 0) return 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == SKB_F32) return launch_tree_t<float>(p, sch, kind, q, N, dev_prims, n_prims, margin, mode, out_vals, out_jac, out_valid, st);
    if (dtype == SKB_F64) return launch_tree_t<double>(p, sch, kind, q, N, dev_prims, n_prims, margin, mode, out_vals, out_jac, out_valid, st);
    return set_error("bad dtype");
}

template <typename T, bool WITH_JAC>
static int launch_pairwise_inst(const skb_plan *p, const skb_pair_program *pp, const void *q, int64_t N, int mode,
                                void *out_vals, void *out_jac, uint8_t *out_valid, cudaStream_t stream) {
    PairParams<T> prm;
    prm.I = reinterpret_cast<const int32_t *>(pp->dI);
    prm.F = reinterpret_cast<const T *>(sizeof(T) == 4 ? pp->dF32 : pp->dF64);
    prm.q = reinterpret_cast<const T *>(q);
    prm.N = N;
    prm.n_tiles = (N + 31) / 32;
    prm.mode = mode;
    prm.out_vals = reinterpret_cast<T *>(out_vals);
    prm.out_jac = reinterpret_cast<T *>(out_jac);
    prm.out_valid = out_valid;
    prm.i_total = (int)pp->I.size();
    prm.f_total = (int)pp->F.size();
    prm.off_F = align_up(prm.i_total * 4, 16);
    prm.off_warp = align_up(prm.off_F + prm.f_total * (int)sizeof(T), 128);
    const int D = p->D, al = 16 / (int)sizeof(T);
    int e = 0;
    prm.woff_q = e;   e += align_up(32 * D, al);
    prm.woff_pos = e; e += p->F * 3 * 32;
    prm.woff_tw = e;  e += WITH_JAC ? pp->n_nodes * 6 * 32 : 0;
    prm.dp = D | 1;
    prm.woff_row = e; e += align_up(32 * prm.dp, al);
    prm.warp_elems = align_up(e, al);
    int warps = 4;
    const size_t warp_bytes = (size_t)prm.warp_elems * sizeof(T);
    const size_t smem_limit = 227 * 1024;
    while (warps > 1 && prm.off_warp + warps * warp_bytes > smem_limit) warps /= 2;
    const size_t smem = prm.off_warp + warps * warp_bytes;
    if (smem > smem_limit) return set_error("pairwise kernel: shared-memory footprint exceeds 227 KB");
    auto kern = pairwise_kernel<T, WITH_JAC>;
    CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, warps * 32, smem));
    if (per_sm < 1) return set_error("pairwise kernel: zero occupancy");
    const long long want = (prm.n_tiles + warps - 1) / warps;
    const int grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)sm_count_cached() * per_sm));
    kern<<<grid, warps * 32, smem, stream>>>(prm);
    CUDA_OK(cudaGetLastError());
    ++g_launch_count;
    return 0;
}

int launch_pairwise(const skb_plan *p, const skb_pair_program *pp, int dtype, const void *q, int64_t N, int mode,
                    void *out_vals, void *out_jac, uint8_t *out_valid, void *stream) {
    if (N <= 0) return 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool wj = out_jac != nullptr;
    if (dtype == SKB_F32)
        return wj ? launch_pairwise_inst<float, true>(p, pp, q, N, mode, out_vals, out_jac, out_valid, st)
                  : launch_pairwise_inst<float, false>(p, pp, q, N, mode, out_vals, out_jac, out_valid, st);
    if (dtype == SKB_F64)
        return wj ? launch_pairwise_inst<double, true>(p, pp, q, N, mode, out_vals, out_jac, out_valid, st)
                  : launch_pairwise_inst<double, false>(p, pp, q, N, mode, out_vals, out_jac, out_valid, st);
    return set_error("bad dtype");
}

template <typename T>
static int launch_sdf_eval_t(const void *dev_prims, int n_prims, const void *pts, int64_t M, void *out_vals, void *out_grads,
                             cudaStream_t stream) {
    const size_t smem = (size_t)std::max(1, n_prims) * sizeof(DevPrim<T>);
    if (smem > 48 * 1024) return set_error("sdf_eval: too many primitives");
    const int threads = 256;
    const int grid = (int)std::max<long long>(1, std::min<long long>((M + threads - 1) / threads, (long long)sm_count_cached() * 8));
    sdf_eval_kernel<T><<<grid, threads, smem, stream>>>(reinterpret_cast<const DevPrim<T> *>(dev_prims), n_prims,
                                                        reinterpret_cast<const T *>(pts), M, reinterpret_cast<T *>(out_vals),
                                                        reinterpret_cast<T *>(out_grads));
    CUDA_OK(cudaGetLastError());
    ++g_launch_count;
    return 0;
}

int launch_sdf_eval(const void *dev_prims, int n_prims, int dtype, const void *pts, int64_t M, void *out_vals, void *out_grads,
                    void *stream) {
    if (M <= 0) return 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == SKB_F32) return launch_sdf_eval_t<float>(dev_prims, n_prims, pts, M, out_vals, out_grads, st);
    if (dtype == SKB_F64) return launch_sdf_eval_t<double>(dev_prims, n_prims, pts, M, out_vals, out_grads, st);
    return set_error("bad dtype");
}

}  // namespace skb

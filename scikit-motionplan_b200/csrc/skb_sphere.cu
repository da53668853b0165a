// skb_sphere.cu -- the fused FK + SDF + Jacobian kernel for point features (collision spheres).
//
//   sphere_kernel<T, KIND, WITH_JAC, VW>
//     KIND_COLLFREE  K1+K2: CollFreeConst._evaluate            (skmp/constraint.py:287-374)
//     KIND_MAP       K1:    KinematicsMapBase.map, IGNORE rot  (skmp/kinematics.py:43-63)
//
// Execution model: one warp owns a tile of G configurations (G = 4..32, chosen from the
// shared-memory budget) and runs two phases per tile without any block-level barrier.
//
//   phase 1  forward kinematics, level by level: lane = (configuration, node of the level).  A lane
//            composes  parent frame x constant pre-transform x Rz(q)/Tz(q)  (every joint axis is
//            folded onto +z by the plan compiler) and stores the node frame [R|t] plus the joint's
//            spatial twist (omega, v = o x omega) in the warp's shared-memory slab.
//   phase 2  one lane per sphere: a "round" is up to 32 consecutive output spheres of ONE
//            configuration.  The lane transforms its sphere centre, evaluates the world SDF (value
//            + analytic gradient; voxel grids through one 32-byte corner-pack gather), forms the
//            wrench (p x g, g) and contracts it with the twists of all D columns
//                 J[s][d] = omega_d . (p x g) + v_d . g     if joint d is an ancestor of s, else 0
//            (twists are warp-broadcast shared loads, so the contraction is warp-amortised).  Rows
//            are assembled in registers, staged as the dense [cnt][D] block that IS the output
//            layout, and shipped with coalesced 16-byte streaming stores (evict-first, so the
//            write stream does not evict the L2-resident grid).  Sphere centres and per-sphere rows
//            never round-trip through HBM.
//   closest / validity: warp-shuffle min / arg-min over the spheres of a configuration.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>

#include "skb_device.cuh"
#include "skb_internal.h"

namespace skb {

enum { MAX_ROUNDS = 32 };

#ifndef SKB_LB_THREADS
#define SKB_LB_THREADS 256      // max threads per CTA
#define SKB_LB_BLOCKS 3         // min CTAs per SM (register cap = 65536 / product)
#endif

template <typename T>
struct SphereParams {
    const int32_t *I;
    const T *F;
    const T *q;
    long long N;
    long long n_tiles;
    const DevPrim<T> *prims;
    int n_prims;
    T margin;
    T *out_vals;
    T *out_jac;
    uint8_t *out_valid;
    int i_total, f_total;
    int off_F, off_prims, off_warp, warp_elems;   // bytes, bytes, bytes, elements of T per warp
    int off_lv, n_nodes;                          // per-CTA level-record table (bytes), node count
    int woff_q, woff_sc, woff_frames, woff_tw, woff_stage; // element offsets inside a warp slab
    int fstride, tstride;                         // elements of T per configuration: frames, twists
    int G, logG, slots3;                          // configurations per tile (power of two <= 8), nodes per FK pass
    int D, S, n_rounds, n_levels;                 // copies of the program header (uniform operands)
    int copy_vec;                                 // 1: every round block is 16-byte aligned in out_jac
    int stage_elems;                              // elements of T per staging buffer (two per warp)
    // Warp-uniform data the hot loop reads every iteration lives in the kernel parameters (constant bank):
    // the compiler then keeps mask tests / branches on the uniform datapath and folds the primitive's
    // constants straight into FFMA operands.
    unsigned long long round_cm[MAX_ROUNDS];      // per round: columns any of its spheres depends on
    DevPrim<T> prim0;                             // copy of primitive 0 (single-primitive fast path)
};

template <typename T>
__device__ __forceinline__ void st_stream(T *p, T v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream4(float *p, const float4 &v) { __stcs(reinterpret_cast<float4 *>(p), v); }
__device__ __forceinline__ void st_stream4(double *p, const double2 &v) { __stcs(reinterpret_cast<double2 *>(p), v); }

// packed fp32x2 FMA (sm_100: SASS FFMA2): both halves in one issue slot
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b),
                       rc = *reinterpret_cast<unsigned long long *>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2 *>(&rd);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b), rd;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2 *>(&rd);
}

// How the kernel reaches the world SDF.
enum { SDF_UNION = 0,       // any union of primitives, table staged in shared memory
       SDF_SINGLE = 1,      // one primitive of any type, constants read from the kernel parameters
       SDF_GRID_FAST = 2 }; // one fp32 corner-packed voxel grid with identity rotation (the cfg3 hot path)

// fp32 cell-packed grid, identity rotation: same polynomial evaluation as sdf_prim_world -> sdf_grid_local,
// minus the type dispatch, rotation, value-type and 64-bit index generality.
__device__ __forceinline__ float sdf_grid_fast(const DevPrim<float> &pr, float px, float py, float pz, float &gx, float &gy, float &gz) {
    const float ux = ((px - pr.pos[0]) - pr.par[0]) * pr.par[3], uy = ((py - pr.pos[1]) - pr.par[1]) * pr.par[4],
                uz = ((pz - pr.pos[2]) - pr.par[2]) * pr.par[5];
    const float tx = float(pr.dims[0] - 1), ty = float(pr.dims[1] - 1), tz = float(pr.dims[2] - 1);
    gx = gy = gz = 0.f;
    if (!(ux >= 0.f && ux <= tx && uy >= 0.f && uy <= ty && uz >= 0.f && uz <= tz)) return pr.par[6];
    const int ix = min((int)floorf(ux), pr.dims[0] - 2), iy = min((int)floorf(uy), pr.dims[1] - 2), iz = min((int)floorf(uz), pr.dims[2] - 2);
    const float fx = ux - float(ix), fy = uy - float(iy), fz = uz - float(iz);
    const unsigned cell = ((unsigned)ix * (unsigned)(pr.dims[1] - 1) + (unsigned)iy) * (unsigned)(pr.dims[2] - 1) + (unsigned)iz;
    float a[8], hx, hy, hz;
    ldg_cell(reinterpret_cast<const float *>(pr.cells) + (size_t)cell * 8, a);
    const float val = trilinear_poly<float>(a, fx, fy, fz, hx, hy, hz);
    gx = hx * pr.par[3]; gy = hy * pr.par[4]; gz = hz * pr.par[5];
    return val;
}
template <typename T>
__device__ __forceinline__ T sdf_grid_fast(const DevPrim<T> &pr, T px, T py, T pz, T &gx, T &gy, T &gz) {
    return sdf_prim_world<T>(pr, px, py, pz, gx, gy, gz);     // only instantiated for float; keeps templates well-formed
}

// Twist storage per configuration.  Scalar layout: 8 elements per column [wx wy wz 0 | vx vy vz 0].
// Pair layout (fp32 collfree, even D): 12 floats per column pair (j, j+1), component-interleaved
// [wx_j wx_j1 wy_j wy_j1 | wz_j wz_j1 vx_j vx_j1 | vy_j vy_j1 vz_j vz_j1] so that three LDS.128 feed six FFMA2.
template <typename T, int KIND, int VW>
struct TwistLayout {
    static constexpr bool PAIR = (sizeof(T) == 4) && (KIND == KIND_COLLFREE) && (VW >= 2);
};

template <typename T, int KIND, bool WITH_JAC, bool CLOSEST, int SDFK, int VW, typename MaskT>
__global__ void __launch_bounds__(SKB_LB_THREADS, SKB_LB_BLOCKS) sphere_kernel(const __grid_constant__ SphereParams<T> prm) {
    using Rl = Real<T>;
    using V4 = Vec4<T>;
    constexpr int ROWS = (KIND == KIND_MAP) ? 3 : 1;
    constexpr bool PAIR = TwistLayout<T, KIND, VW>::PAIR && !CLOSEST;
    extern __shared__ __align__(16) unsigned char smem[];
    int32_t *sI = reinterpret_cast<int32_t *>(smem);
    T *sF = reinterpret_cast<T *>(smem + prm.off_F);
    DevPrim<T> *sP = reinterpret_cast<DevPrim<T> *>(smem + prm.off_prims);

    // the warp index goes through a shuffle so the compiler treats it (and every loop bound derived from it)
    // as warp-uniform: mask tests and branches of the hot loop then run on the uniform datapath
    const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), n_warps = blockDim.x >> 5;
    for (int i = tid; i < prm.i_total; i += blockDim.x) sI[i] = prm.I[i];
    for (int i = tid; i < prm.f_total; i += blockDim.x) sF[i] = prm.F[i];
    if (KIND == KIND_COLLFREE && SDFK == SDF_UNION) {
        const int words = prm.n_prims * (int)(sizeof(DevPrim<T>) / 4);
        const int32_t *src = reinterpret_cast<const int32_t *>(prm.prims);
        int32_t *dst = reinterpret_cast<int32_t *>(sP);
        for (int i = tid; i < words; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();

    const int n_levels = prm.n_levels, n_rounds = prm.n_rounds, D = prm.D, S = prm.S;
    const int fstride = prm.fstride, tstride = prm.tstride;
    const int32_t *levelI = sI + sI[S_LEVEL_I];
    int4 *lvrec = reinterpret_cast<int4 *>(smem + prm.off_lv);
    build_level_records(lvrec, sI + sI[S_NODE_I], sI + sI[S_LEVELNODE_I], prm.n_nodes);
    __syncthreads();
    const int4 *roundI = reinterpret_cast<const int4 *>(sI + sI[S_ROUND_I]);
    const int4 *tabB = reinterpret_cast<const int4 *>(sI + sI[S_TABB_I]);
    const V4 *nodeF = reinterpret_cast<const V4 *>(sF + sI[S_NODE_F]);
    const V4 *tabA = reinterpret_cast<const V4 *>(sF + sI[S_TABA_F]);
    const int G = prm.G, logG = prm.logG;

    T *wbase = reinterpret_cast<T *>(smem + prm.off_warp) + (size_t)warp * prm.warp_elems;
    T *wq = wbase + prm.woff_q;
    T *wsc = wbase + prm.woff_sc;
    T *wframes = wbase + prm.woff_frames;
    T *wtw = wbase + prm.woff_tw;
    T *wstage = wbase + prm.woff_stage;
    const int stage_elems = prm.stage_elems;
    int stage_sel = 0;
    const uint64_t l2_policy = l2_evict_first_policy();

    for (long long tile = (long long)blockIdx.x * n_warps + warp; tile < prm.n_tiles; tile += (long long)gridDim.x * n_warps) {
        const long long n0 = tile * G;
        const int g_live = (int)min((long long)G, prm.N - n0);
        __syncwarp();
        // q tile + one vectorised sin/cos pass over all G*D joint values (32 lanes busy), so the level walk below
        // -- where only G x (nodes of the level) lanes are busy -- just picks (sin, cos) up from shared memory
        for (int i = lane; i < g_live * D; i += 32) {
            const T th = prm.q[n0 * D + i];
            T sn, cs;
            Rl::sincos(th, &sn, &cs);
            wq[i] = th;
            wsc[2 * i] = sn; wsc[2 * i + 1] = cs;
        }
        __syncwarp();

        // ------------------------------------------------------------ phase 1: frames + twists (three lanes per node)
        fk_phase<T, WITH_JAC, PAIR>(lane, G, logG, prm.slots3, g_live, D, n_levels, lvrec, levelI, nodeF, wq, wsc,
                                    wframes, fstride, wtw, tstride);

        // ------------------------------------------------------------ phase 2: one lane per sphere
        // per-configuration reduction state (validity / closest modes only)
        T vmin = Rl::inf();
        T best = Rl::inf();
        int best_idx = 0x7fffffff;
        T bgx = T(0), bgy = T(0), bgz = T(0), bmx = T(0), bmy = T(0), bmz = T(0);
        unsigned long long bmask = 0;

        // one (configuration c, round r) step; A / B are the round's per-lane sphere constants
        auto step = [&](const int c, const int r, const int4 rd, const V4 &A, const int4 &B) {
            const long long n = n0 + c;
            const V4 *fr = reinterpret_cast<const V4 *>(wframes + c * fstride);
            const T *tw = wtw + c * tstride;
            {
                const bool has = B.w != 0;
                const V4 f0 = fr[3 * B.x], f1 = fr[3 * B.x + 1], f2 = fr[3 * B.x + 2];
                const T px = f0.w + f0.x * A.x + f0.y * A.y + f0.z * A.z;
                const T py = f1.w + f1.x * A.x + f1.y * A.y + f1.z * A.z;
                const T pz = f2.w + f2.x * A.x + f2.y * A.y + f2.z * A.z;
                T gx = T(0), gy = T(0), gz = T(0);

                if (KIND == KIND_COLLFREE) {
                    const T d = SDFK == SDF_GRID_FAST ? sdf_grid_fast(prm.prim0, px, py, pz, gx, gy, gz)
                              : SDFK == SDF_SINGLE    ? sdf_prim_world<T>(prm.prim0, px, py, pz, gx, gy, gz)
                                                      : sdf_union<T>(sP, prm.n_prims, px, py, pz, gx, gy, gz);
                    const T val = d - A.w - prm.margin;
                    if (has) vmin = (val < vmin || val != val) ? val : vmin;      // NaN sticks (np.min)
                    if (CLOSEST) {
                        // arg-min over spheres of sdf - radius, first index wins ties (np.argmin)
                        if (has && val < best) {
                            best = val; best_idx = rd.x + lane;
                            bmask = ((unsigned long long)(unsigned)B.z << 32) | (unsigned)B.y;
                            bgx = gx; bgy = gy; bgz = gz;
                            wrench_moment<T>(px, py, pz, gx, gy, gz, bmx, bmy, bmz);
                        }
                        return;
                    }
                    if (prm.out_vals != nullptr && has) st_stream(prm.out_vals + n * S + rd.x + lane, val);
                } else {
                    if (has) {
                        T *g = prm.out_vals + (n * S + rd.x + lane) * 3;
                        st_stream(g, px); st_stream(g + 1, py); st_stream(g + 2, pz);
                    }
                }
                if (!WITH_JAC || CLOSEST) return;

                // ---- Jacobian rows of this lane's sphere, all D columns, assembled VW at a time
                const unsigned long long cm = prm.round_cm[r];               // warp-uniform (constant bank)
                const MaskT am = sizeof(MaskT) == 8 ? (MaskT)(((unsigned long long)(unsigned)B.z << 32) | (unsigned)B.y) : (MaskT)(unsigned)B.y;
                T mx, my, mz;                                                 // wrench moment p x g
                wrench_moment<T>(px, py, pz, gx, gy, gz, mx, my, mz);
                T *row0 = wstage + stage_sel * stage_elems;
                T *row = row0 + lane * (ROWS * D);
                // the bulk store issued from this buffer two rounds ago must have finished reading it
                if (lane == 0) bulk_wait_read<1>();
                __syncwarp();
                if (PAIR) {
                    const float2 mx2 = make_float2((float)mx, (float)mx), my2 = make_float2((float)my, (float)my), mz2 = make_float2((float)mz, (float)mz);
                    const float2 gx2 = make_float2((float)gx, (float)gx), gy2 = make_float2((float)gy, (float)gy), gz2 = make_float2((float)gz, (float)gz);
                    for (int j = 0; j < D; j += VW) {
                        alignas(16) float v0[VW];
#pragma unroll
                        for (int u = 0; u < VW; u += 2) {
                            const int col = j + u;
                            float2 acc = make_float2(0.f, 0.f);
                            if ((cm >> col) & 3ull) {                   // uniform: someone in this round needs col or col + 1
                                const float4 *t4 = reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(tw) + (col >> 1) * 12);
                                const float4 a = t4[0], b = t4[1], cc = t4[2];
                                acc = mul2(make_float2(a.x, a.y), mx2);
                                acc = fma2(make_float2(a.z, a.w), my2, acc);
                                acc = fma2(make_float2(b.x, b.y), mz2, acc);
                                acc = fma2(make_float2(b.z, b.w), gx2, acc);
                                acc = fma2(make_float2(cc.x, cc.y), gy2, acc);
                                acc = fma2(make_float2(cc.z, cc.w), gz2, acc);
                                const MaskT sh = am >> col;
                                acc.x = (sh & 1) ? acc.x : 0.f;
                                acc.y = (sh & 2) ? acc.y : 0.f;
                            }
                            v0[u] = acc.x; v0[u + 1] = acc.y;
                        }
                        if (VW == 2) *reinterpret_cast<float2 *>(reinterpret_cast<float *>(row) + j) = *reinterpret_cast<float2 *>(v0);
                        else *reinterpret_cast<float4 *>(reinterpret_cast<float *>(row) + j) = *reinterpret_cast<float4 *>(v0);
                    }
                } else {
                    for (int j = 0; j < D; j += VW) {
                        alignas(16) T v0[VW], v1[VW], v2[VW];
#pragma unroll
                        for (int u = 0; u < VW; ++u) {
                            const int col = j + u;
                            v0[u] = T(0); v1[u] = T(0); v2[u] = T(0);
                            if ((cm >> col) & 1ull) {                   // uniform: someone in this round needs column col
                                const V4 w = *reinterpret_cast<const V4 *>(tw + col * 8);
                                const V4 v = *reinterpret_cast<const V4 *>(tw + col * 8 + 4);
                                const bool mine = (am >> col) & 1;
                                if (KIND == KIND_COLLFREE) {
                                    const T jv = dot6<T>(w.x, w.y, w.z, v.x, v.y, v.z, mx, my, mz, gx, gy, gz);
                                    v0[u] = mine ? jv : T(0);
                                } else {                                // d p / d q_col = omega x p + v
                                    v0[u] = mine ? (w.y * pz - w.z * py + v.x) : T(0);
                                    v1[u] = mine ? (w.z * px - w.x * pz + v.y) : T(0);
                                    v2[u] = mine ? (w.x * py - w.y * px + v.z) : T(0);
                                }
                            }
                        }
                        if (VW == 1) {
                            row[j] = v0[0];
                            if (ROWS == 3) { row[D + j] = v1[0]; row[2 * D + j] = v2[0]; }
                        } else if (VW * sizeof(T) == 8) {
                            *reinterpret_cast<float2 *>(row + j) = *reinterpret_cast<float2 *>(v0);
                            if (ROWS == 3) {
                                *reinterpret_cast<float2 *>(row + D + j) = *reinterpret_cast<float2 *>(v1);
                                *reinterpret_cast<float2 *>(row + 2 * D + j) = *reinterpret_cast<float2 *>(v2);
                            }
                        } else {
                            *reinterpret_cast<float4 *>(row + j) = *reinterpret_cast<float4 *>(v0);
                            if (ROWS == 3) {
                                *reinterpret_cast<float4 *>(row + D + j) = *reinterpret_cast<float4 *>(v1);
                                *reinterpret_cast<float4 *>(row + 2 * D + j) = *reinterpret_cast<float4 *>(v2);
                            }
                        }
                    }
                }
                // ---- ship the round: the dense [cnt][ROWS*D] block IS the output layout and is contiguous in HBM
                {
                    const int len = rd.y * ROWS * D;
                    T *g = prm.out_jac + (n * S + rd.x) * (long long)(ROWS * D);
                    if (prm.copy_vec) {
                        // one bulk async store (TMA, SASS UBLKCP) per round, issued by lane 0; no LSU traffic for the copy
                        fence_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            bulk_store_hint(g, row0, (uint32_t)(len * sizeof(T)), l2_policy);
                            bulk_commit();
                        }
                    } else {
                        __syncwarp();
#pragma unroll 4
                        for (int i = lane; i < len; i += 32) st_stream(g + i, row0[i]);
                    }
                }
                stage_sel ^= 1;
            }
        };   // step

        // per-configuration epilogue of the reducing modes
        auto finish = [&](const int c) {
            const long long n = n0 + c;
            const T *tw = wtw + c * tstride;
            if (KIND == KIND_COLLFREE) {
                if (prm.out_valid != nullptr) {
                    T m = vmin;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) { const T x = __shfl_xor_sync(0xffffffffu, m, o); m = (x < m || x != x) ? x : m; }
                    if (lane == 0) prm.out_valid[n] = m > T(0) ? 1 : 0;
                }
                if (CLOSEST) {
                    T m = best; int mi = best_idx;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const T x = __shfl_xor_sync(0xffffffffu, m, o);
                        const int xi = __shfl_xor_sync(0xffffffffu, mi, o);
                        if (x < m || (x == m && xi < mi)) { m = x; mi = xi; }
                    }
                    if (prm.out_vals != nullptr && lane == 0) prm.out_vals[n] = m;
                    if (WITH_JAC) {
                        const int src = __ffs(__ballot_sync(0xffffffffu, best_idx == mi && best == m)) - 1;
                        const T gx = __shfl_sync(0xffffffffu, bgx, src), gy = __shfl_sync(0xffffffffu, bgy, src), gz = __shfl_sync(0xffffffffu, bgz, src);
                        const T mx = __shfl_sync(0xffffffffu, bmx, src), my = __shfl_sync(0xffffffffu, bmy, src), mz = __shfl_sync(0xffffffffu, bmz, src);
                        const unsigned long long am = __shfl_sync(0xffffffffu, bmask, src);
                        for (int col = lane; col < D; col += 32) {     // one lane per column of the single row
                            T jv = T(0);
                            if ((am >> col) & 1ull) {
                                const V4 w = *reinterpret_cast<const V4 *>(tw + col * 8);
                                const V4 v = *reinterpret_cast<const V4 *>(tw + col * 8 + 4);
                                jv = dot6<T>(w.x, w.y, w.z, v.x, v.y, v.z, mx, my, mz, gx, gy, gz);
                            }
                            prm.out_jac[n * D + col] = jv;
                        }
                    }
                }
            }
            vmin = Rl::inf(); best = Rl::inf(); best_idx = 0x7fffffff;
        };

        if (!CLOSEST && prm.out_valid == nullptr) {
            // no per-configuration reduction: rounds outermost, so each round's sphere table is read once per tile
            for (int r = 0; r < n_rounds; ++r) {
                const int4 rd = roundI[r];
                const V4 A = tabA[r * 32 + lane];
                const int4 B = tabB[r * 32 + lane];
                for (int c = 0; c < g_live; ++c) step(c, r, rd, A, B);
            }
        } else {
            for (int c = 0; c < g_live; ++c) {
                for (int r = 0; r < n_rounds; ++r) step(c, r, roundI[r], tabA[r * 32 + lane], tabB[r * 32 + lane]);
                finish(c);
            }
        }
    }       // tiles
    if (WITH_JAC && !CLOSEST) bulk_wait_read<0>();   // shared memory must outlive the last bulk reads
}

// ------------------------------------------------------------------ cell-pack expansion of a voxel grid
// Per cell the eight coefficients of its trilinear polynomial (trilinear_poly, skb_device.cuh), formed in fp64 from the
// corner samples and rounded once to the grid's value type: 32 B (fp32) per cell, one LDG.256 per lookup.
template <typename V>
__global__ void expand_cells_kernel(const V *grid, int nx, int ny, int nz, V *cells) {
    const long long total = (long long)(nx - 1) * (ny - 1) * (nz - 1);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int iz = (int)(i % (nz - 1));
        const long long t = i / (nz - 1);
        const int iy = (int)(t % (ny - 1)), ix = (int)(t / (ny - 1));
        const long long sy = nz, sx = (long long)ny * nz, b = ix * sx + iy * sy + iz;
        const double c000 = grid[b], c001 = grid[b + 1], c010 = grid[b + sy], c011 = grid[b + sy + 1];
        const double c100 = grid[b + sx], c101 = grid[b + sx + 1], c110 = grid[b + sx + sy], c111 = grid[b + sx + sy + 1];
        V *o = cells + i * 8;
        o[0] = (V)c000;
        o[1] = (V)(c100 - c000); o[2] = (V)(c010 - c000); o[3] = (V)(c001 - c000);
        o[4] = (V)(c110 - c100 - c010 + c000);
        o[5] = (V)(c101 - c100 - c001 + c000);
        o[6] = (V)(c011 - c010 - c001 + c000);
        o[7] = (V)(c111 - c110 - c101 - c011 + c100 + c010 + c001 - c000);
    }
}

#define CUDA_OK(expr)                                                                          \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return set_error(std::string(#expr) + ": " + cudaGetErrorString(e__));             \
    } while (0)

int launch_expand_cells(const void *grid, int is_f64, const int32_t dims[3], void *cells) {
    const long long total = (long long)(dims[0] - 1) * (dims[1] - 1) * (dims[2] - 1);
    const int threads = 256;
    const int blocks = (int)std::min<long long>((total + threads - 1) / threads, 148 * 16);
    if (is_f64) expand_cells_kernel<double><<<blocks, threads>>>(reinterpret_cast<const double *>(grid), dims[0], dims[1], dims[2], reinterpret_cast<double *>(cells));
    else expand_cells_kernel<float><<<blocks, threads>>>(reinterpret_cast<const float *>(grid), dims[0], dims[1], dims[2], reinterpret_cast<float *>(cells));
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaDeviceSynchronize());
    return 0;
}

// ------------------------------------------------------------------ launcher
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

template <typename T, int KIND, bool WITH_JAC, bool CLOSEST, int SDFK, int VW, typename MaskT>
static int launch_sphere_inst(const skb_plan *p, const skb_sphere_program *sp, const void *q, int64_t N, const void *dev_prims,
                              const void *host_prims, int n_prims, double margin, void *out_vals, void *out_jac,
                              uint8_t *out_valid, cudaStream_t stream) {
    constexpr int ROWS = (KIND == KIND_MAP) ? 3 : 1;
    constexpr bool PAIR = TwistLayout<T, KIND, VW>::PAIR && !CLOSEST;
    SphereParams<T> prm;
    memset(&prm, 0, sizeof(prm));
    prm.I = reinterpret_cast<const int32_t *>(sp->dI);
    prm.F = reinterpret_cast<const T *>(sizeof(T) == 4 ? sp->dF32 : sp->dF64);
    prm.q = reinterpret_cast<const T *>(q);
    prm.N = N;
    prm.prims = reinterpret_cast<const DevPrim<T> *>(dev_prims);
    prm.n_prims = n_prims;
    prm.margin = (T)margin;
    prm.out_vals = reinterpret_cast<T *>(out_vals);
    prm.out_jac = reinterpret_cast<T *>(out_jac);
    prm.out_valid = out_valid;
    prm.i_total = (int)sp->I.size();
    prm.f_total = (int)sp->F.size();
    if (KIND == KIND_COLLFREE && host_prims) prm.prim0 = *reinterpret_cast<const DevPrim<T> *>(host_prims);
    for (int r = 0; r < sp->n_rounds && r < MAX_ROUNDS; ++r) {
        const int32_t *rd = &sp->I[sp->I[S_ROUND_I] + 4 * r];
        prm.round_cm[r] = ((unsigned long long)(uint32_t)rd[3] << 32) | (uint32_t)rd[2];
    }
    const int D = p->D, S = p->F;
    prm.D = D; prm.S = S; prm.n_rounds = sp->n_rounds; prm.n_levels = sp->n_levels;
    const bool stage = WITH_JAC && !CLOSEST;
    prm.off_F = align_up(prm.i_total * 4, 16);
    prm.off_prims = align_up(prm.off_F + prm.f_total * (int)sizeof(T), 16);
    prm.off_lv = align_up(prm.off_prims + ((KIND == KIND_COLLFREE && SDFK == SDF_UNION) ? n_prims : 0) * (int)sizeof(DevPrim<T>), 16);
    prm.n_nodes = sp->n_nodes;
    prm.off_warp = align_up(prm.off_lv + sp->n_nodes * 16, 128);
    const int al = 16 / (int)sizeof(T);
    const int PER = al;
    // every round block [s0*ROWS*D, +cnt*ROWS*D) of every configuration must be 16-byte aligned for vector copy-out
    prm.copy_vec = ((S * ROWS * D) % PER == 0 && (32 * ROWS * D) % PER == 0 && ((S % 32) * ROWS * D) % PER == 0 &&
                    (reinterpret_cast<uintptr_t>(out_jac) % 16 == 0)) ? 1 : 0;
    prm.fstride = sp->frame_stride;
    {   // twist slab per configuration, rounded to an odd number of 16-byte quads (conflict-free phase-1 stores)
        int words = PAIR ? (D / 2) * 12 : D * 8;
        int qd = (words + 3) / 4;
        if (qd % 2 == 0) qd += 1;
        prm.tstride = qd * 4;
    }
    const size_t smem_limit = 227 * 1024;
    int warps = p->tune_warps > 0 ? std::min(p->tune_warps, SKB_LB_THREADS / 32) : SKB_LB_THREADS / 32;
    int ctas_target = p->tune_ctas > 0 ? p->tune_ctas : SKB_LB_BLOCKS;
    int G = p->tune_chunk > 0 ? std::min(p->tune_chunk, 8) : 8;
    auto layout = [&](int g) {
        int e = 0;
        prm.woff_q = e;      e += align_up(g * D, al);
        prm.woff_sc = e;     e += align_up(2 * g * D, al);
        prm.woff_frames = e; e += g * prm.fstride;
        prm.woff_tw = e;     e += WITH_JAC ? g * prm.tstride : 0;
        prm.stage_elems = align_up(32 * ROWS * D, al);
        prm.woff_stage = e;  e += stage ? 2 * prm.stage_elems : 0;
        prm.warp_elems = align_up(e, al);
        return (size_t)prm.off_warp + (size_t)warps * prm.warp_elems * sizeof(T);
    };
    // largest power-of-two tile that still lets `ctas_target` CTAs share one SM's shared memory
    while (G > 1 && (G & (G - 1))) G &= G - 1;
    if (p->tune_chunk <= 0)
        while (G > 4 && (layout(G) + 1024) * ctas_target > smem_limit) G >>= 1;
    while (layout(G) > smem_limit && G > 1) G >>= 1;
    while (layout(G) > smem_limit && warps > 1) warps >>= 1;
    const size_t smem = layout(G);
    if (smem > smem_limit) return set_error("sphere kernel: shared-memory footprint exceeds 227 KB");
    prm.G = G;
    prm.logG = 0;
    while ((1 << prm.logG) < G) ++prm.logG;
    prm.slots3 = 32 / (3 * G);
    prm.n_tiles = (N + G - 1) / G;
    auto kern = sphere_kernel<T, KIND, WITH_JAC, CLOSEST, SDFK, VW, MaskT>;
    CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, warps * 32, smem));
    if (per_sm < 1) return set_error("sphere kernel: zero occupancy");
    if (p->tune_ctas > 0) per_sm = std::min(per_sm, p->tune_ctas);
    const long long want = (prm.n_tiles + warps - 1) / warps;
    const int grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)sm_count_cached() * per_sm));
    kern<<<grid, warps * 32, smem, stream>>>(prm);
    CUDA_OK(cudaGetLastError());
    ++g_launch_count;
    return 0;
}

#define SKB_ARGS p, sp, q, N, dev_prims, host_prims, n_prims, margin, out_vals, out_jac, out_valid, stream
#define SKB_SIG const skb_plan *p, const skb_sphere_program *sp, const void *q, int64_t N, const void *dev_prims, const void *host_prims, \
                int n_prims, double margin, void *out_vals, void *out_jac, uint8_t *out_valid, cudaStream_t stream

// all-spheres mode with Jacobian: pick the widest row store that keeps each lane's row start aligned, and the mask width
template <typename T, int KIND, int SDFK>
static int launch_sphere_jac(SKB_SIG) {
    const int D = p->D;
    const int bytes = D * (int)sizeof(T);
    const bool wide = D > 32;
    if (bytes % 16 == 0 && sizeof(T) == 4)
        return wide ? launch_sphere_inst<T, KIND, true, false, SDFK, 4, unsigned long long>(SKB_ARGS)
                    : launch_sphere_inst<T, KIND, true, false, SDFK, 4, unsigned>(SKB_ARGS);
    if ((bytes % 8 == 0 && sizeof(T) == 4) || (bytes % 16 == 0 && sizeof(T) == 8))
        return wide ? launch_sphere_inst<T, KIND, true, false, SDFK, 2, unsigned long long>(SKB_ARGS)
                    : launch_sphere_inst<T, KIND, true, false, SDFK, 2, unsigned>(SKB_ARGS);
    return wide ? launch_sphere_inst<T, KIND, true, false, SDFK, 1, unsigned long long>(SKB_ARGS)
                : launch_sphere_inst<T, KIND, true, false, SDFK, 1, unsigned>(SKB_ARGS);
}

template <typename T, int SDFK>
static int launch_sphere_collfree(SKB_SIG, int mode) {
    const bool wj = out_jac != nullptr;
    if (mode == SKB_MODE_CLOSEST)
        return wj ? launch_sphere_inst<T, KIND_COLLFREE, true, true, SDFK, 1, unsigned long long>(SKB_ARGS)
                  : launch_sphere_inst<T, KIND_COLLFREE, false, true, SDFK, 1, unsigned long long>(SKB_ARGS);
    if (wj) return launch_sphere_jac<T, KIND_COLLFREE, SDFK>(SKB_ARGS);
    return launch_sphere_inst<T, KIND_COLLFREE, false, false, SDFK, 1, unsigned>(SKB_ARGS);
}

template <typename T>
static int launch_sphere_t(SKB_SIG, int kind, int mode) {
    if (kind == KIND_MAP) {
        if (out_jac != nullptr) return launch_sphere_jac<T, KIND_MAP, SDF_SINGLE>(SKB_ARGS);
        return launch_sphere_inst<T, KIND_MAP, false, false, SDF_SINGLE, 1, unsigned>(SKB_ARGS);
    }
    if (n_prims == 1) {
        const DevPrim<T> *pr = reinterpret_cast<const DevPrim<T> *>(host_prims);
        const long long cells = (long long)(pr->dims[0] - 1) * (pr->dims[1] - 1) * (pr->dims[2] - 1);
        if (sizeof(T) == 4 && pr->type == PRIM_GRID && (pr->flags & PF_GRID_CELLS) && (pr->flags & PF_IDENTITY) &&
            !(pr->flags & PF_GRID_F64) && cells < (1ll << 31))
            return launch_sphere_collfree<T, SDF_GRID_FAST>(SKB_ARGS, mode);
        return launch_sphere_collfree<T, SDF_SINGLE>(SKB_ARGS, mode);
    }
    return launch_sphere_collfree<T, SDF_UNION>(SKB_ARGS, mode);
}

int launch_sphere(const skb_plan *p, const skb_sphere_program *sp, int kind, int dtype, const void *q, int64_t N,
                  const void *dev_prims, const void *host_prims, int n_prims, double margin, int mode, void *out_vals,
                  void *out_jac, uint8_t *out_valid, void *stream_) {
    if (N <= 0) return 0;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (dtype == SKB_F32) return launch_sphere_t<float>(SKB_ARGS, kind, mode);
    if (dtype == SKB_F64) return launch_sphere_t<double>(SKB_ARGS, kind, mode);
    return set_error("bad dtype");
}

}  // namespace skb

// skb_sphere2.cu -- the full-output kernel (fp32): every row of values and Jacobians, for
//
//     KIND_COLLFREE   CollFreeConst._evaluate_all               (skmp/constraint.py:335-374)
//     KIND_PAIRWISE   PairWiseSelfCollFreeConst._evaluate       (skmp/constraint.py:749-778, all-pairs mode)
//
// One warp owns a tile of G configurations.
//   phase 1   forward kinematics, level by level, three lanes per (configuration, node)  -> fk_phase (skb_device.cuh)
//   phase 1b  (pairwise) sphere centres of the tile, one lane per sphere, into the warp's shared slab
//   phase 2   "steps" of up to 64 consecutive output rows of ONE configuration: lane l owns row s0 + l and row
//             s0 + 32 + l.  A row is reduced to a value, a wrench (m, g) and a column mask:
//                 collfree   p = frame * offset,  (d, g) = SDF(p),  value = d - r - margin,  m = p x g
//                 pairwise   delta = x_i - x_j,  value = |delta|^2 - (r_i + r_j)^2,  g = 2 delta,  m = x_i x g
//                            (x_i x delta == x_j x delta, so one wrench serves both spheres; columns that move only
//                             sphere j enter with a minus sign)
//             and the Jacobian row is the twist/wrench pairing  J[d] = omega_d . m + v_d . g  over the masked columns,
//             two columns per packed fma.rn.f32x2, the twist pair loaded once (warp-broadcast LDS.128 x3) for BOTH rows
//             of the lane.  The dense [rows][D] block is assembled in shared memory -- it IS the output layout -- and
//             leaves with one TMA bulk store per step (cp.async.bulk shared -> global, L2 evict-first).
// Sphere centres, frames, twists and SDF gradients never touch HBM.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iterator>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <type_traits>

#include "skb_device.cuh"
#include "skb_internal.h"

namespace skb {

#ifndef SKB_CEN_SOA
#define SKB_CEN_SOA 1           // pairwise sphere centres as three planes x[] y[] z[] per configuration (12 bytes per centre through the
                                // shared-memory pipe instead of a padded 16-byte vector)
#endif
#ifndef SKB_S2_PIPE
#define SKB_S2_PIPE 1           // 0: voxel lookups issued and consumed inside the same configuration (round-1 behaviour)
#endif
#ifndef SKB_S2_GRID_BLOCKS
#define SKB_S2_GRID_BLOCKS 3    // launch bounds of the fp32 voxel-grid all-mode kernels
#endif
#ifndef SKB_S2_GRID_THREADS
#define SKB_S2_GRID_THREADS 256
#endif
#ifndef SKB_S2_THREADS
#define SKB_S2_THREADS 256      // max threads per CTA
#define SKB_S2_BLOCKS 3         // min CTAs per SM (register cap = 65536 / product); the fp32 grid all-mode kernels take 4:
                                // they fit 64 registers without spilling and run 2 % faster for it, the others spill
#endif

// Launch-time constants of the fp32 corner-packed grid with identity rotation (the cfg3 hot path), folded on the host:
// origin = world position + lower bound (added in fp64, rounded once), float upper bounds, integer index caps and cell
// strides -- the lookup then needs no int<->float conversion or address arithmetic on descriptor fields.
struct GridFast {
    float ox, oy, oz;          // sample-box origin in world coordinates
    float ix_, iy_, iz_;       // 1 / spacing
    float tx, ty, tz;          // dims - 1 (last sample index, as float)
    int mx, my, mz;            // dims - 2 (last cell index)
    unsigned sy, sz;           // cells per x step = (ny - 1)(nz - 1) is sy * sz; sy = ny - 1, sz = nz - 1
    float fill;
    const float *cells;
    unsigned long long tex;    // texture object over `cells` (float4 texels), or 0: plain 32-byte global loads
};

template <typename T>
struct S2Params {
    const int32_t *I;
    const T *F;
    const unsigned long long *pair_masks;                             // pairwise reducing modes: {plus, minus} per row
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
    int off_F, off_prims, off_warp, warp_elems;                       // bytes, bytes, bytes, elements of T per warp
    int off_lv, n_nodes;                                              // per-CTA level-record table (bytes), node count
    int woff_q, woff_sc, woff_frames, woff_tw, woff_cen, woff_stage;  // element offsets inside a warp slab
    int fstride, tstride, cstride;                                    // elements of T per configuration: frames, twists, centres
    int cplane;                                                       // centres: elements of T per coordinate plane (SKB_CEN_SOA)
    int G, logG, slots3;
    int D, NP, R, n_steps, n_levels, n_cen;                           // R = output rows per configuration
    int copy_vec;                                                     // 1: every step block is 16-byte aligned in out_jac
    // warp-uniform data of the hot loop lives in the kernel parameters (constant bank -> uniform datapath)
    skb_s2_step steps[S2_MAX_STEPS];
    unsigned lanemap[32];                                             // FK walk: lane -> (row, configuration, slot), see fk_lane_map
    int use_lanemap;
    DevPrim<T> prim0;
    GridFast grid;
};

__device__ __forceinline__ void st_stream(float *p, float v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(double *p, double v) { __stcs(p, v); }

// packed fp32x2 FMA (sm_100: SASS FFMA2 / FMUL2): both halves in one issue slot
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

enum { SDF_UNION = 0, SDF_SINGLE = 1, SDF_GRID_FAST = 2 };

// Branch-free: the cell index is clamped and the gather always issued, the fill value / zero gradient are selected
// afterwards -- so the two gathers of a lane's two rows are in flight together instead of being serialised by a
// divergent branch.  Value and gradient are Horner evaluations of the cell's polynomial pack (trilinear_poly).
// The lookup is split in two so that the full-output kernel can issue the NEXT configuration's lookups before it runs
// the current configuration's column loop (the fetch latency then hides behind ~250 issue slots of FFMA2 work).
struct GridPend {              // a lookup in flight: the cell's pack, the in-cell coordinates, the inside flag
    float a[8];
    float fx, fy, fz;
    bool inside;
};
__device__ __forceinline__ void tex_cell(unsigned long long tex, unsigned texel, float &x, float &y, float &z, float &w) {
    // volatile: stays where it is written, ahead of the column loop, instead of being sunk to its first use
    asm volatile("tex.1d.v4.f32.s32 {%0,%1,%2,%3}, [%4, {%5}];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "l"(tex), "r"(texel));
}
__device__ __forceinline__ void s2_grid_issue(const GridFast &gr, float px, float py, float pz, GridPend &pd) {
    const float ux = (px - gr.ox) * gr.ix_, uy = (py - gr.oy) * gr.iy_, uz = (pz - gr.oz) * gr.iz_;
    const int ix = max(0, min((int)floorf(ux), gr.mx)), iy = max(0, min((int)floorf(uy), gr.my)), iz = max(0, min((int)floorf(uz), gr.mz));
    pd.fx = ux - float(ix); pd.fy = uy - float(iy); pd.fz = uz - float(iz);
    // inside the sample box  <=>  0 <= u <= dims - 1 on every axis  <=>  the in-cell coordinate of the CLAMPED cell lies in
    // [0, 1] (u - cell is exact there): two 3-input min / max and two compares instead of six compares whose short-circuit
    // form ptxas turned into predicated re-evaluations of u.  A NaN coordinate is decided by the other two axes and then
    // yields NaN from the polynomial, as numpy's bounds test does.
    pd.inside = fminf(pd.fx, fminf(pd.fy, pd.fz)) >= 0.f && fmaxf(pd.fx, fmaxf(pd.fy, pd.fz)) <= 1.f;
    const unsigned cell = ((unsigned)ix * gr.sy + (unsigned)iy) * gr.sz + (unsigned)iz;
    if (gr.tex != 0ull) {       // (warp-uniform) two 16-byte texel fetches: TEX pipe instead of the LSU data pipe
        tex_cell(gr.tex, 2u * cell, pd.a[0], pd.a[1], pd.a[2], pd.a[3]);
        tex_cell(gr.tex, 2u * cell + 1u, pd.a[4], pd.a[5], pd.a[6], pd.a[7]);
    } else ldg_cell(gr.cells + (size_t)cell * 8, pd.a);
}
__device__ __forceinline__ float s2_grid_finish(const GridFast &gr, const GridPend &pd, float &gx, float &gy, float &gz) {
    float hx, hy, hz;
    const float val = trilinear_poly<float>(pd.a, pd.fx, pd.fy, pd.fz, hx, hy, hz);
    gx = pd.inside ? hx * gr.ix_ : 0.f;
    gy = pd.inside ? hy * gr.iy_ : 0.f;
    gz = pd.inside ? hz * gr.iz_ : 0.f;
    return pd.inside ? val : gr.fill;
}
__device__ __forceinline__ float s2_grid_fast(const GridFast &gr, float px, float py, float pz, float &gx, float &gy, float &gz) {
    GridPend pd;
    s2_grid_issue(gr, px, py, pz, pd);
    return s2_grid_finish(gr, pd, gx, gy, gz);
}

// two adjacent columns of a Jacobian row: packed f32x2 arithmetic for float, two scalar chains for double -- the same
// mul, then five fma per half in both cases (bit-identical to dot6)
template <typename T> struct Pair2;
template <> struct Pair2<float> {
    using type = float2;
    static __device__ __forceinline__ float2 make(float x, float y) { return make_float2(x, y); }
    static __device__ __forceinline__ float2 mul(float2 a, float s) { return mul2(a, make_float2(s, s)); }
    static __device__ __forceinline__ float2 fma(float2 a, float s, float2 c) { return fma2(a, make_float2(s, s), c); }
};
template <> struct Pair2<double> {
    using type = double2;
    static __device__ __forceinline__ double2 make(double x, double y) { return make_double2(x, y); }
    static __device__ __forceinline__ double2 mul(double2 a, double s) { return make_double2(__dmul_rn(a.x, s), __dmul_rn(a.y, s)); }
    static __device__ __forceinline__ double2 fma(double2 a, double s, double2 c) { return make_double2(__fma_rn(a.x, s, c.x), __fma_rn(a.y, s, c.y)); }
};

// the corner-pack fast path exists for fp32 only; the fp64 instantiation of SDF_GRID_FAST is never launched
__device__ __forceinline__ float grid_fast_or_generic(const S2Params<float> &prm, float px, float py, float pz, float &gx, float &gy, float &gz) {
    return s2_grid_fast(prm.grid, px, py, pz, gx, gy, gz);
}
__device__ __forceinline__ double grid_fast_or_generic(const S2Params<double> &prm, double px, double py, double pz, double &gx, double &gy, double &gz) {
    return sdf_prim_world<double>(prm.prim0, px, py, pz, gx, gy, gz);
}

// sphere centre `idx` of a configuration's centre block: three planes (SKB_CEN_SOA) or one padded 4-vector
template <typename T>
__device__ __forceinline__ Vec4<T> load_centre(const T *cc, const int cplane, const int idx) {
    Vec4<T> x;
    if (SKB_CEN_SOA) { x.x = cc[idx]; x.y = cc[cplane + idx]; x.z = cc[2 * cplane + idx]; x.w = T(0); }
    else x = reinterpret_cast<const Vec4<T> *>(cc)[idx];
    return x;
}

template <typename T> struct PendRowT { GridPend g; T px, py, pz; };   // a row whose voxel lookup is in flight
template <typename T, typename MaskT> struct RowState {   // what phase 2 keeps of one output row
    T gx, gy, gz, mx, my, mz;
    MaskT pm, mm;
};

// RED = false: every row is written (values, and with JAC the dense Jacobian block)
// RED = true : per-configuration reduction -- closest row (np.argmin: first index wins ties) and / or validity byte
// NPMAX: compile-time bound on the number of column pairs (the Jacobian loop is fully unrolled against it, so twist
// addresses and column bits become immediates)
// CSC = true (full-output mode with Jacobians only): the Jacobian block of a configuration is written TRANSPOSED, [D][R] -- the
// order of the block-diagonal CSC `data` array of the SQP solver's stacked constraint Jacobian (one column per (waypoint,
// dof), rows of the waypoint's constraint inside it: skmp/solver/nlp_solver/trajectory_constraint.py:157-194).  In that
// layout the rows of a step are consecutive addresses of ONE column, i.e. lane-per-row stores are coalesced as they are:
// no staging block, no bulk store, one 128-byte store per (column, half).
template <typename T, int KIND, int SDFK, int VW, typename MaskT, bool RED, bool JAC, int NPMAX, bool CSC = false>
__global__ void __launch_bounds__((sizeof(T) == 4 && KIND == KIND_COLLFREE && SDFK == SDF_GRID_FAST && JAC && !RED) ? SKB_S2_GRID_THREADS : SKB_S2_THREADS, (sizeof(T) == 4 && KIND == KIND_COLLFREE && SDFK == SDF_GRID_FAST && JAC && !RED) ? SKB_S2_GRID_BLOCKS : SKB_S2_BLOCKS) s2_kernel(const __grid_constant__ S2Params<T> prm) {
    using V4 = Vec4<T>;
    using Rl = Real<T>;
    using P2 = Pair2<T>;
    using V2 = typename P2::type;
    using PendRow = PendRowT<T>;
    // full-output fp32 voxel-grid launches run the lookups one configuration ahead of the column loop
    constexpr bool PIPE = KIND == KIND_COLLFREE && SDFK == SDF_GRID_FAST && !RED && sizeof(T) == 4 && SKB_S2_PIPE;
    extern __shared__ __align__(16) unsigned char smem[];
    int32_t *sI = reinterpret_cast<int32_t *>(smem);
    T *sF = reinterpret_cast<T *>(smem + prm.off_F);
    DevPrim<T> *sP = reinterpret_cast<DevPrim<T> *>(smem + prm.off_prims);

    // the warp index goes through a shuffle so the compiler treats it (and everything derived from it) as warp-uniform
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), n_warps = blockDim.x >> 5;
    int lane = tid & 31;
    if (sizeof(T) == 4) asm volatile("" : "+r"(lane));   // opaque: kept in a register instead of being re-derived from %tid at
                                                         // every use (the fp64 instantiations are register-bound and keep the remat)
    for (int i = tid; i < prm.i_total; i += blockDim.x) sI[i] = prm.I[i];
    for (int i = tid; i < prm.f_total; i += blockDim.x) sF[i] = prm.F[i];
    if (KIND == KIND_COLLFREE && SDFK == SDF_UNION) {
        const int words = prm.n_prims * (int)(sizeof(DevPrim<T>) / 4);
        const int32_t *src = reinterpret_cast<const int32_t *>(prm.prims);
        int32_t *dst = reinterpret_cast<int32_t *>(sP);
        for (int i = tid; i < words; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();

    const int n_levels = prm.n_levels, n_steps = prm.n_steps, D = prm.D, NP = prm.NP, R = prm.R;
    const int fstride = prm.fstride, tstride = prm.tstride, cstride = prm.cstride;
    const int32_t *levelI = sI + sI[Q_LEVEL_I];
    int4 *lvrec = reinterpret_cast<int4 *>(smem + prm.off_lv);
    build_level_records(lvrec, sI + sI[Q_NODE_I], sI + sI[Q_LEVELNODE_I], prm.n_nodes);
    __syncthreads();
    // The lane tables A / B are read once per (tile, step).  The full-output collfree kernels read them from global memory
    // (L1-resident, a few KB) and keep only the tree description in shared memory: the 5 KB they would take per CTA decide
    // between 3 and 4 resident CTAs for a 14-DoF arm pair.
    constexpr bool TAB_GLOBAL = KIND == KIND_COLLFREE && !RED;
    const int4 *tabB = TAB_GLOBAL ? reinterpret_cast<const int4 *>(prm.I + sI[Q_TABB_I]) : reinterpret_cast<const int4 *>(sI + sI[Q_TABB_I]);
    const int4 *tabC = reinterpret_cast<const int4 *>(sI + sI[Q_TABC_I]);
    const int32_t *cenNode = sI + sI[Q_CENNODE_I];
    const V4 *nodeF = reinterpret_cast<const V4 *>(sF + sI[Q_NODE_F]);
    const V4 *tabA = TAB_GLOBAL ? reinterpret_cast<const V4 *>(prm.F + sI[Q_TABA_F]) : reinterpret_cast<const V4 *>(sF + sI[Q_TABA_F]);
    const V4 *cenA = reinterpret_cast<const V4 *>(sF + sI[Q_CENA_F]);
    const T *tabThr = sF + sI[Q_THR_F];               // pairwise, fp64 only: (r_i + r_j)^2 per (step, half, lane)
    const int G = prm.G, logG = prm.logG;

    T *wbase = reinterpret_cast<T *>(smem + prm.off_warp) + (size_t)warp * prm.warp_elems;
    T *wq = wbase + prm.woff_q;
    T *wsc = wbase + prm.woff_sc;
    T *wframes = wbase + prm.woff_frames;
    T *wtw = wbase + prm.woff_tw;
    T *wcen = wbase + prm.woff_cen;
    T *wstage = wbase + prm.woff_stage;
    const uint64_t l2_policy = l2_evict_first_policy();

    // the next tile's joint values are fetched while the current tile is being processed (first 128 of the G * D values
    // of a tile; longer tiles read the rest directly), so the DRAM latency of q never sits on the critical path
    const long long tile_step = (long long)gridDim.x * n_warps;
    T qn[4];
    auto fetch_q = [&](const long long t) {
        const long long b = t * G * D;
        const int cnt = t < prm.n_tiles ? (int)min((long long)G, prm.N - t * G) * D : 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) qn[k] = (lane + 32 * k < cnt) ? __ldg(prm.q + b + lane + 32 * k) : T(0);
    };
    fetch_q((long long)blockIdx.x * n_warps + warp);

    for (long long tile = (long long)blockIdx.x * n_warps + warp; tile < prm.n_tiles; tile += tile_step) {
        const long long n0 = tile * G;
        const int g_live = (int)min((long long)G, prm.N - n0);
        // the staging block shares its shared memory with the q / sin-cos scratch of phase 1: the last bulk store of the
        // previous tile must have finished reading it
        if (JAC && !RED && !CSC && lane == 0) bulk_wait_read<0>();
        __syncwarp();
        // q tile + one fully populated sin/cos pass over all G*D joint values
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = lane + 32 * k;
            if (i < g_live * D) {
                T sn, cs;
                Rl::sincos(qn[k], &sn, &cs);
                wq[i] = qn[k];
                wsc[2 * i] = sn; wsc[2 * i + 1] = cs;
            }
        }
        for (int i = lane + 128; i < g_live * D; i += 32) {
            const T th = prm.q[n0 * D + i];
            T sn, cs;
            Rl::sincos(th, &sn, &cs);
            wq[i] = th;
            wsc[2 * i] = sn; wsc[2 * i + 1] = cs;
        }
        fetch_q(tile + tile_step);
        __syncwarp();

        // ------------------------------------------------------------ phase 1: frames + twists
        fk_phase<T, JAC, true>(lane, G, logG, prm.slots3, g_live, D, n_levels, lvrec, levelI, nodeF, wq, wsc,
                                    wframes, fstride, wtw, tstride, prm.use_lanemap ? prm.lanemap : nullptr);

        // ------------------------------------------------------------ phase 1b: sphere centres (pairwise)
        if (KIND == KIND_PAIRWISE) {
            for (int s = lane; s < prm.n_cen; s += 32) {
                const int node = cenNode[s];
                const V4 A = cenA[s];
                for (int c = 0; c < g_live; ++c) {
                    const V4 *fr = reinterpret_cast<const V4 *>(wframes + c * fstride);
                    const V4 f0 = fr[3 * node], f1 = fr[3 * node + 1], f2 = fr[3 * node + 2];
                    V4 x;
                    x.x = f0.w + f0.x * A.x + f0.y * A.y + f0.z * A.z;
                    x.y = f1.w + f1.x * A.x + f1.y * A.y + f1.z * A.z;
                    x.z = f2.w + f2.x * A.x + f2.y * A.y + f2.z * A.z;
                    x.w = T(0);
                    if (SKB_CEN_SOA) { T *cc = wcen + c * cstride; cc[s] = x.x; cc[prm.cplane + s] = x.y; cc[2 * prm.cplane + s] = x.z; }
                    else reinterpret_cast<V4 *>(wcen + c * cstride)[s] = x;
                }
            }
            __syncwarp();
        }

        // ------------------------------------------------------------ phase 2: steps of up to 64 rows
        // value + wrench + masks of one row of configuration c
        auto front = [&](const int c, const V4 &A, const int4 &B, const int tslot, RowState<T, MaskT> &rs) -> T {
            T val;
            if (KIND == KIND_COLLFREE) {
                const V4 *fr = reinterpret_cast<const V4 *>(wframes + c * fstride);
                const V4 f0 = fr[3 * B.x], f1 = fr[3 * B.x + 1], f2 = fr[3 * B.x + 2];
                const T px = f0.w + f0.x * A.x + f0.y * A.y + f0.z * A.z;
                const T py = f1.w + f1.x * A.x + f1.y * A.y + f1.z * A.z;
                const T pz = f2.w + f2.x * A.x + f2.y * A.y + f2.z * A.z;
                const T d = SDFK == SDF_GRID_FAST ? grid_fast_or_generic(prm, px, py, pz, rs.gx, rs.gy, rs.gz)
                              : SDFK == SDF_SINGLE    ? sdf_prim_world<T>(prm.prim0, px, py, pz, rs.gx, rs.gy, rs.gz)
                                                      : sdf_union<T>(sP, prm.n_prims, px, py, pz, rs.gx, rs.gy, rs.gz);
                val = d - A.w - prm.margin;
                if (JAC) wrench_moment<T>(px, py, pz, rs.gx, rs.gy, rs.gz, rs.mx, rs.my, rs.mz);
                rs.pm = sizeof(MaskT) == 8 ? (MaskT)(((unsigned long long)(unsigned)B.z << 32) | (unsigned)B.y) : (MaskT)(unsigned)B.y;
                rs.mm = 0;
            } else {
                const T *cen = wcen + c * cstride;
                const V4 xi = load_centre<T>(cen, prm.cplane, B.x & 0xffff), xj = load_centre<T>(cen, prm.cplane, (unsigned)B.x >> 16);
                const T dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z;
                val = Rl::fma(dz, dz, Rl::fma(dy, dy, Rl::mul(dx, dx))) - (sizeof(T) == 4 ? (T)__int_as_float(B.w) : tabThr[tslot * 32 + lane]);
                rs.gx = T(2) * dx; rs.gy = T(2) * dy; rs.gz = T(2) * dz;
                if (JAC) wrench_moment<T>(xi.x, xi.y, xi.z, rs.gx, rs.gy, rs.gz, rs.mx, rs.my, rs.mz);
                rs.pm = (MaskT)(unsigned)B.y;
                rs.mm = (MaskT)(unsigned)B.z;
                if (sizeof(MaskT) == 8) {                     // D > 32: the high mask words live in a second table
                    const int4 C = tabC[tslot * 32 + lane];
                    rs.pm |= (MaskT)((unsigned long long)(unsigned)C.x << 32);
                    rs.mm |= (MaskT)((unsigned long long)(unsigned)C.y << 32);
                }
            }
            return val;
        };

        if (!RED) {
            // 64-bit output bases once per tile; (configuration, step) offsets stay 32-bit (G * R * D < 2^31)
            int row_off = lane * D;                         // opaque: computed once, not rematerialised per column pair
            asm volatile("" : "+r"(row_off));
            T *tile_vals = prm.out_vals + n0 * R;
            T *tile_jac = JAC ? prm.out_jac + n0 * R * (long long)D : nullptr;
            int slot = 0;                                   // running index into the per-(half, lane) tables
            for (int st = 0; st < n_steps; ++st) {
                // pair-activity bits of the step (bit jp: some row needs column 2 jp or 2 jp + 1); made opaque so it stays in one
                // register instead of being re-read from the constant bank inside the unrolled loop
                unsigned pact = (unsigned)prm.steps[st].cm;
                asm volatile("" : "+r"(pact));
                const int s0 = prm.steps[st].s0;
                int cntA = prm.steps[st].cntA, cntB = prm.steps[st].cntB;
                // opaque: read from the constant bank once per (tile, step), not before every use inside the configuration loop
                if (sizeof(T) == 4) asm volatile("" : "+r"(cntA), "+r"(cntB));
                // output addresses of the step, formed once per (tile, step): configuration c adds c * R (* D) to them
                T *step_vals = tile_vals + (unsigned)(s0 + lane);
                T *step_jac = JAC ? (CSC ? tile_jac + (unsigned)s0 : tile_jac + (unsigned)(s0 * D)) : nullptr;
                if (sizeof(T) == 4) asm volatile("" : "+l"(step_vals), "+l"(step_jac));
                const int4 B0 = tabB[slot * 32 + lane];
                const int4 B1 = tabB[(slot + (cntB > 0 ? 1 : 0)) * 32 + lane];
                const int slotA = slot, slotB = slot + (cntB > 0 ? 1 : 0);
                V4 A0, A1;
                if (KIND == KIND_COLLFREE) { A0 = tabA[slot * 32 + lane]; A1 = tabA[(slot + (cntB > 0 ? 1 : 0)) * 32 + lane]; }
                slot += cntB > 0 ? 2 : 1;

                // grid lookups split into issue / finish (PIPE): the same arithmetic as front(), in the same order
                auto front_issue = [&](const int c, const V4 &A, const int4 &B, PendRow &pr) {
                    if constexpr (PIPE) {
                        const V4 *fr = reinterpret_cast<const V4 *>(wframes + c * fstride);
                        const V4 f0 = fr[3 * B.x], f1 = fr[3 * B.x + 1], f2 = fr[3 * B.x + 2];
                        pr.px = f0.w + f0.x * A.x + f0.y * A.y + f0.z * A.z;
                        pr.py = f1.w + f1.x * A.x + f1.y * A.y + f1.z * A.z;
                        pr.pz = f2.w + f2.x * A.x + f2.y * A.y + f2.z * A.z;
                        s2_grid_issue(prm.grid, pr.px, pr.py, pr.pz, pr.g);
                    }
                };
                auto front_finish = [&](const PendRow &pr, const V4 &A, const int4 &B, RowState<T, MaskT> &rs) -> T {
                    T val = T(0);
                    if constexpr (PIPE) {
                        const T d = s2_grid_finish(prm.grid, pr.g, rs.gx, rs.gy, rs.gz);
                        val = d - A.w - prm.margin;
                        if (JAC) wrench_moment<T>(pr.px, pr.py, pr.pz, rs.gx, rs.gy, rs.gz, rs.mx, rs.my, rs.mz);
                        rs.pm = sizeof(MaskT) == 8 ? (MaskT)(((unsigned long long)(unsigned)B.z << 32) | (unsigned)B.y) : (MaskT)(unsigned)B.y;
                        rs.mm = 0;
                    }
                    return val;
                };

                // twist pair x wrench for the columns (col, col + 1), masked / signed
                auto contract = [&](const V4 &t0, const V4 &t1, const V4 &t2, const RowState<T, MaskT> &rs, const int col) -> V2 {
                    V2 a = P2::mul(P2::make(t0.x, t0.y), rs.mx);
                    a = P2::fma(P2::make(t0.z, t0.w), rs.my, a);
                    a = P2::fma(P2::make(t1.x, t1.y), rs.mz, a);
                    a = P2::fma(P2::make(t1.z, t1.w), rs.gx, a);
                    a = P2::fma(P2::make(t2.x, t2.y), rs.gy, a);
                    a = P2::fma(P2::make(t2.z, t2.w), rs.gz, a);
                    // column bits are warp-uniform operands: one LOP3 (predicate) + one FSEL per element
                    const MaskT b0 = (MaskT)1 << col, b1 = (MaskT)2 << col;
                    if (KIND == KIND_PAIRWISE) {
                        a.x = (rs.pm & b0) ? a.x : ((rs.mm & b0) ? -a.x : T(0));
                        a.y = (rs.pm & b1) ? a.y : ((rs.mm & b1) ? -a.y : T(0));
                    } else {
                        a.x = (rs.pm & b0) ? a.x : T(0);
                        a.y = (rs.pm & b1) ? a.y : T(0);
                    }
                    return a;
                };

                auto body = [&](auto has_b, const int c, PendRow &pa, PendRow &pb) {
                    constexpr bool HAS_B = decltype(has_b)::value;
                    int tw_off = c * tstride;                      // opaque: one address register for the whole unrolled loop
                    asm volatile("" : "+r"(tw_off));
                    const T *tw = wtw + tw_off;
                    RowState<T, MaskT> ra, rb;
                    T va, vb = T(0);
                    if constexpr (PIPE) {
                        // this configuration's lookups were issued one configuration ago; the next one's go out now and stay
                        // in flight during the column loop below
                        va = front_finish(pa, A0, B0, ra);
                        if (HAS_B) vb = front_finish(pb, A1, B1, rb);
                        if (c + 1 < g_live) {
                            front_issue(c + 1, A0, B0, pa);
                            if (HAS_B) front_issue(c + 1, A1, B1, pb);
                        }
                    } else {
                        va = front(c, A0, B0, slotA, ra);
                        if (HAS_B) vb = front(c, A1, B1, slotB, rb);
                    }
                    if (prm.out_vals != nullptr) {
                        T *gv = step_vals + (unsigned)(c * R);
                        if (lane < cntA) st_stream(gv, va);
                        if (HAS_B && lane < cntB) st_stream(gv + 32, vb);
                    }
                    if (!JAC) return;
                    if (CSC) {
                        // transposed block: element (row, col) of configuration c at  c R D + col R + row.  One warp-uniform
                        // 64-bit base per (configuration, step) and a 32-bit per-lane offset that advances by 2 R per column pair
                        T *g0 = step_jac + (unsigned)(c * R * D) + lane;   // column 0, row s0 + lane of configuration c
                        if (sizeof(T) == 4) asm volatile("" : "+l"(g0));   // opaque: formed once, then advanced by 2 R per column pair
                        const V4 *t4 = reinterpret_cast<const V4 *>(tw);
                        const bool inA = lane < cntA, inB = HAS_B && lane < cntB;
#pragma unroll
                        for (int jp = 0; jp < NPMAX; ++jp) {
                            const int col = 2 * jp;
                            if (col < D) {
                                V2 a = P2::make(T(0), T(0)), b = a;
                                if (pact & (1u << jp)) {
                                    const V4 t0 = t4[3 * jp], t1 = t4[3 * jp + 1], t2 = t4[3 * jp + 2];
                                    a = contract(t0, t1, t2, ra, col);
                                    if (HAS_B) b = contract(t0, t1, t2, rb, col);
                                }
                                if (inA) st_stream(g0, a.x);
                                if (inB) st_stream(g0 + 32, b.x);
                                if (VW == 2 || col + 1 < D) {              // VW == 2: D is even
                                    T *g1 = g0 + (unsigned)R;
                                    if (inA) st_stream(g1, a.y);
                                    if (inB) st_stream(g1 + 32, b.y);
                                }
                                g0 += 2u * (unsigned)R;
                            }
                        }
                        return;
                    }
                    T *rowA = wstage + row_off;
                    T *rowB = rowA + 32 * D;
                    // the bulk store of the previous step must have finished reading the staging block
                    if (lane == 0) bulk_wait_read<0>();
                    __syncwarp();
                    const V4 *t4 = reinterpret_cast<const V4 *>(tw);
#pragma unroll
                    for (int jp = 0; jp < NPMAX; ++jp) {
                        // one uniform test per pair: pairs no row of the step needs (and pairs beyond D) were zero-filled
                        // once per (tile, step) below, so there is nothing to compute or store for them
                        if (pact & (1u << jp)) {
                            const int col = 2 * jp;
                            const V4 t0 = t4[3 * jp], t1 = t4[3 * jp + 1], t2 = t4[3 * jp + 2];
                            const V2 a = contract(t0, t1, t2, ra, col);
                            V2 b = a;
                            if (HAS_B) b = contract(t0, t1, t2, rb, col);
                            if (VW == 2) {
                                *reinterpret_cast<V2 *>(rowA + col) = a;
                                if (HAS_B) *reinterpret_cast<V2 *>(rowB + col) = b;
                            } else {
                                rowA[col] = a.x;
                                if (HAS_B) rowB[col] = b.x;
                                if (col + 1 < D) {
                                    rowA[col + 1] = a.y;
                                    if (HAS_B) rowB[col + 1] = b.y;
                                }
                            }
                        }
                    }
                    // ship: the dense [rows][D] block is contiguous in HBM
                    const int len = (cntA + (HAS_B ? cntB : 0)) * D;
                    T *g = step_jac + (unsigned)(c * R * D);
                    if (prm.copy_vec) {
                        fence_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            bulk_store_hint(g, wstage, (uint32_t)(len * sizeof(T)), l2_policy);
                            bulk_commit();
                        }
                    } else {
                        __syncwarp();
#pragma unroll 4
                        for (int i = lane; i < len; i += 32) st_stream(g + i, wstage[i]);
                        __syncwarp();
                    }
                };

                // (a step whose rows touch every column pair between them rewrites its whole block per configuration:
                // nothing to fill)
                if (JAC && !CSC && pact != (NP >= 32 ? 0xffffffffu : (1u << NP) - 1u)) {
                    // zero fill of the step's staging block, once per (tile, step): the columns no row of the step needs
                    // keep their structural zeros while the configurations of the tile pass through
                    if (lane == 0) bulk_wait_read<0>();
                    __syncwarp();
                    const int len16 = ((cntA + cntB) * D * (int)sizeof(T) + 15) >> 4;
                    float4 *z = reinterpret_cast<float4 *>(wstage);
                    for (int i = lane; i < len16; i += 32) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    __syncwarp();
                }
                PendRow pa, pb;
                if constexpr (PIPE) {
                    front_issue(0, A0, B0, pa);
                    if (cntB > 0) front_issue(0, A1, B1, pb);
                }
                if (cntB > 0) for (int c = 0; c < g_live; ++c) body(std::true_type{}, c, pa, pb);
                else          for (int c = 0; c < g_live; ++c) body(std::false_type{}, c, pa, pb);
            }
        } else if (KIND == KIND_PAIRWISE) {
            // ---- per-configuration reduction over the pairs, lean form: an 8-byte table entry and two centre loads per
            // pair, four half-rounds in flight; only the winning row's wrench and column masks are formed (afterwards)
            const int2 *tabP = reinterpret_cast<const int2 *>(sI + sI[Q_TABP_I]);
            const int n_hr = sI[Q_N_HR];
            for (int c = 0; c < g_live; ++c) {
                const long long n = n0 + c;
                const T *cen = wcen + c * cstride;
                const int cplane = prm.cplane;
                auto pair_value = [&](const int row, V4 &xi, T &dx, T &dy, T &dz) -> T {
                    const int2 P = tabP[row];
                    xi = load_centre<T>(cen, cplane, P.x & 0xffff);
                    const V4 xj = load_centre<T>(cen, cplane, (unsigned)P.x >> 16);
                    dx = xi.x - xj.x; dy = xi.y - xj.y; dz = xi.z - xj.z;
                    const T thr = sizeof(T) == 4 ? (T)__int_as_float(P.y) : (row < R ? tabThr[row] : -Rl::max());
                    return Rl::fma(dz, dz, Rl::fma(dy, dy, Rl::mul(dx, dx))) - thr;
                };
                T best = Rl::inf(), poison = T(0);          // poison: NaN once any row's value is NaN (np.min semantics)
                int bidx = 0x7fffffff;
                for (int h = 0; h < n_hr; h += 4) {
                    T v[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        V4 xi;
                        T dx, dy, dz;
                        v[k] = pair_value((h + k) * 32 + lane, xi, dx, dy, dz);
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        poison = Rl::fma(v[k], T(0), poison);
                        if (v[k] < best) { best = v[k]; bidx = (h + k) * 32 + lane; }
                    }
                    // validity only: one pair at or below zero decides the configuration
                    if (!JAC && prm.out_vals == nullptr && __any_sync(0xffffffffu, best <= T(0))) break;
                }
                T m = best; int mi = bidx;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const T x = __shfl_xor_sync(0xffffffffu, m, o);
                    const int xi = __shfl_xor_sync(0xffffffffu, mi, o);
                    if (x < m || (x == m && xi < mi)) { m = x; mi = xi; }
                }
                const bool nan_cfg = __any_sync(0xffffffffu, poison != poison);
                if (nan_cfg) m = poison + Rl::inf() * T(0);        // NaN on every lane
                if (lane == 0) {
                    if (prm.out_vals != nullptr) prm.out_vals[n] = m;
                    if (prm.out_valid != nullptr) prm.out_valid[n] = m > T(0) ? 1 : 0;
                }
                if (JAC) {
                    // the closest pair's row: g = 2 (x_i - x_j), m = x_i x g; columns of i's chain +, of j's chain -
                    V4 xi;
                    T dx, dy, dz, mx, my, mz;
                    mi = mi < R ? mi : 0;                          // no row won (NaN input): any row, the values are NaN anyway
                    pair_value(mi, xi, dx, dy, dz);
                    const T gx = T(2) * dx, gy = T(2) * dy, gz = T(2) * dz;
                    wrench_moment<T>(xi.x, xi.y, xi.z, gx, gy, gz, mx, my, mz);
                    const unsigned long long pm = __ldg(prm.pair_masks + 2 * mi), mm = __ldg(prm.pair_masks + 2 * mi + 1);
                    const T *tw = wtw + c * tstride;
                    for (int col = lane; col < D; col += 32) {     // one lane per column of the single row
                        const T *t2 = tw + (col >> 1) * 12 + (col & 1);
                        T jv = dot6<T>(t2[0], t2[2], t2[4], t2[6], t2[8], t2[10], mx, my, mz, gx, gy, gz);
                        jv = ((pm >> col) & 1) ? jv : (((mm >> col) & 1) ? -jv : T(0));
                        prm.out_jac[n * D + col] = nan_cfg ? m : jv;
                    }
                }
            }
        } else {
            // ---- per-configuration reduction: closest row (value, Jacobian row) and / or validity byte
            for (int c = 0; c < g_live; ++c) {
                const long long n = n0 + c;
                T best = Rl::inf(), poison = T(0);          // poison: NaN once any row's value is NaN (np.min semantics)
                int bidx = 0x7fffffff;
                RowState<T, MaskT> brs;
                brs.gx = brs.gy = brs.gz = brs.mx = brs.my = brs.mz = T(0); brs.pm = 0; brs.mm = 0;
                int slot = 0;
                for (int st = 0; st < n_steps; ++st) {
                    const int s0 = prm.steps[st].s0, cntA = prm.steps[st].cntA, cntB = prm.steps[st].cntB;
                    // both rows of the lane are evaluated before either is compared, so their gathers overlap
                    RowState<T, MaskT> rsa, rsb;
                    V4 Aa, Ab;
                    const int slot_b = slot + (cntB > 0 ? 1 : 0);
                    if (KIND == KIND_COLLFREE) { Aa = tabA[slot * 32 + lane]; Ab = tabA[slot_b * 32 + lane]; }
                    const T va = front(c, Aa, tabB[slot * 32 + lane], slot, rsa);
                    T vb = Rl::inf();
                    if (cntB > 0) vb = front(c, Ab, tabB[slot_b * 32 + lane], slot_b, rsb);
                    if (lane < cntA) poison = Rl::fma(va, T(0), poison);
                    if (lane < cntB) poison = Rl::fma(vb, T(0), poison);
                    if (lane < cntA && va < best) { best = va; bidx = s0 + lane; brs = rsa; }
                    if (lane < cntB && vb < best) { best = vb; bidx = s0 + 32 + lane; brs = rsb; }
                    slot = slot_b + 1;
                    // validity only: one row at or below zero decides the configuration (the reference's composite
                    // short-circuits the same way, skmp/constraint.py:167-176)
                    if (!JAC && prm.out_vals == nullptr && __any_sync(0xffffffffu, best <= T(0))) break;
                }
                T m = best; int mi = bidx;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const T x = __shfl_xor_sync(0xffffffffu, m, o);
                    const int xi = __shfl_xor_sync(0xffffffffu, mi, o);
                    if (x < m || (x == m && xi < mi)) { m = x; mi = xi; }
                }
                const bool nan_cfg = __any_sync(0xffffffffu, poison != poison);
                if (nan_cfg) m = poison + Rl::inf() * T(0);        // NaN on every lane
                if (lane == 0) {
                    if (prm.out_vals != nullptr) prm.out_vals[n] = m;
                    if (prm.out_valid != nullptr) prm.out_valid[n] = m > T(0) ? 1 : 0;
                }
                if (JAC) {
                    const int src = max(0, __ffs(__ballot_sync(0xffffffffu, bidx == mi && best == m)) - 1);
                    const T gx = __shfl_sync(0xffffffffu, brs.gx, src), gy = __shfl_sync(0xffffffffu, brs.gy, src), gz = __shfl_sync(0xffffffffu, brs.gz, src);
                    const T mx = __shfl_sync(0xffffffffu, brs.mx, src), my = __shfl_sync(0xffffffffu, brs.my, src), mz = __shfl_sync(0xffffffffu, brs.mz, src);
                    const MaskT pm = __shfl_sync(0xffffffffu, brs.pm, src);
                    const MaskT mm = __shfl_sync(0xffffffffu, brs.mm, src);
                    const T *tw = wtw + c * tstride;
                    for (int col = lane; col < D; col += 32) {     // one lane per column of the single row
                        const T *t2 = tw + (col >> 1) * 12 + (col & 1);
                        T jv = dot6<T>(t2[0], t2[2], t2[4], t2[6], t2[8], t2[10], mx, my, mz, gx, gy, gz);
                        jv = ((pm >> col) & 1) ? jv : (((mm >> col) & 1) ? -jv : T(0));
                        prm.out_jac[n * D + col] = nan_cfg ? m : jv;
                    }
                }
            }
        }
    }       // tiles
    if (JAC && !RED && !CSC) bulk_wait_read<0>();   // shared memory must outlive the last bulk reads
}

// ------------------------------------------------------------------ launcher
#define CUDA_OK(expr)                                                                          \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return set_error(std::string(#expr) + ": " + cudaGetErrorString(e__));             \
    } while (0)

static int align_up(int x, int a) { return (x + a - 1) / a * a; }
static int odd_quads(int words) { int q = (words + 3) / 4; if (q % 2 == 0) q += 1; return q * 4; }

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

// One candidate launch of a step program: kernel parameters, shared-memory layout, (G, warps per CTA, CTAs per SM) and the
// score that picked them.
// shared-memory carve-out (KB) the unified L1 / shared array takes for `bytes_per_sm` of resident CTAs (1 KB reserved each)
static int carve_kb(size_t bytes_per_sm) {
    static const int steps[] = {8, 16, 32, 64, 100, 132, 164, 196, 228};
    for (int kb : steps) if (bytes_per_sm <= (size_t)kb * 1024) return kb;
    return 1 << 20;
}

// Lane order of the FK walk.  Lane p of a slot's 3 G lanes may carry any (configuration, row) of the slot's node; its frame
// accesses go to quad  cfg * fsq + 3 node + row  (16-byte units) of the warp's frame area, and a 128-bit shared-memory access
// is processed eight lanes at a time: the order is chosen so that the (cfg * fsq + row) mod 8 of the lanes of one slot inside
// one pass are all different (depth-first search, a few hundred nodes at most; the plain order when none exists).
// Returns false when the plain order is already conflict-free or no better order was found.
static bool fk_lane_map_search(int G, int fsq, unsigned map[32]);
static bool fk_lane_map(int G, int fsq, unsigned map[32]) {          // memoised: a launch pays a map lookup
    struct Entry { bool ok; unsigned map[32]; };
    static std::mutex mu;
    static std::map<std::pair<int, int>, Entry> cache;
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find({G, fsq});
    if (it == cache.end()) {
        Entry e;
        memset(&e, 0, sizeof(e));
        e.ok = fk_lane_map_search(G, fsq, e.map);
        it = cache.emplace(std::make_pair(G, fsq), e).first;
    }
    if (it->second.ok) memcpy(map, it->second.map, sizeof(it->second.map));
    return it->second.ok;
}
static bool fk_lane_map_search(int G, int fsq, unsigned map[32]) {
    const int per = 3 * G, slots3 = 32 / per;
    int lane_cfg[32], lane_row[32], lane_slot[32];
    for (int l = 0; l < 32; ++l) { lane_row[l] = l % 3; lane_cfg[l] = (l / 3) % G; lane_slot[l] = (l / 3) / G; }
    auto conflicts = [&](const int *cfgs, const int *rows) {
        int bad = 0;
        for (int g0 = 0; g0 < 32; g0 += 8)
            for (int s = 0; s < slots3; ++s) {
                int seen = 0;
                for (int l = g0; l < g0 + 8; ++l) {
                    if (l >= per * slots3 || l / per != s) continue;
                    const int o = (cfgs[l] * fsq + rows[l]) & 7;
                    if (seen & (1 << o)) ++bad;
                    seen |= 1 << o;
                }
            }
        return bad;
    };
    const int base_bad = conflicts(lane_cfg, lane_row);
    if (base_bad == 0) return false;
    int best_cfg[32], best_row[32];
    bool found = true;
    memcpy(best_cfg, lane_cfg, sizeof(best_cfg));
    memcpy(best_row, lane_row, sizeof(best_row));
    for (int s = 0; s < slots3 && found; ++s) {
        const int l0 = s * per;
        int used = 0, order[24], budget = 200000;
        // depth-first: position i (lane l0 + i) takes an unused (cfg, row) whose bank group is free in its 8-lane pass
        std::function<bool(int, int)> dfs = [&](int i, int seen) -> bool {
            if (i == per) return true;
            if (--budget < 0) return false;
            if (((l0 + i) & 7) == 0) seen = 0;                       // a new pass starts
            for (int k = 0; k < per; ++k) {
                if (used & (1 << k)) continue;
                const int o = ((k / 3) * fsq + (k % 3)) & 7;
                if (seen & (1 << o)) continue;
                used |= 1 << k; order[i] = k;
                if (dfs(i + 1, seen | (1 << o))) return true;
                used &= ~(1 << k);
            }
            return false;
        };
        if (!dfs(0, 0)) { found = false; break; }
        for (int i = 0; i < per; ++i) { best_cfg[l0 + i] = order[i] / 3; best_row[l0 + i] = order[i] % 3; }
    }
    if (!found || conflicts(best_cfg, best_row) >= base_bad) return false;
    for (int l = 0; l < 32; ++l) {
        if (l >= per * slots3) { map[l] = 7u << 5; continue; }      // idle lane: slot 7 >= slots3
        int s1 = l, s2 = l;
        for (int m = 0; m < per * slots3; ++m)
            if (lane_slot[m] == lane_slot[l] && best_cfg[m] == best_cfg[l]) {
                if (best_row[m] == (best_row[l] + 1) % 3) s1 = m;
                if (best_row[m] == (best_row[l] + 2) % 3) s2 = m;
            }
        map[l] = (unsigned)best_row[l] | (unsigned)best_cfg[l] << 2 | (unsigned)lane_slot[l] << 5 | (unsigned)s1 << 8 | (unsigned)s2 << 16;
    }
    return true;
}

template <typename T>
struct S2Launch {
    S2Params<T> prm;
    int G = 0, warps = 0, per_sm = 0, fstride = 0;
    size_t smem = 0;
    double score = -1.0;
};

template <typename T, int KIND, int SDFK, int VW, typename MaskT, bool RED, bool JAC, int NPMAX, bool CSC = false>
static int configure_s2(const skb_plan *p, const skb_s2_program *sp, const void *q, int64_t N, const void *dev_prims,
                        const void *host_prims, int n_prims, double margin, void *out_vals, void *out_jac, uint8_t *out_valid,
                        S2Launch<T> &L, int force_g = 0, int force_w = 0) {
    S2Params<T> &prm = L.prm;
    memset(&prm, 0, sizeof(prm));
    const bool lean = KIND == KIND_PAIRWISE && RED;      // reducing modes of the pairwise constraint run the lean program
    prm.I = reinterpret_cast<const int32_t *>(lean ? sp->dI_lean : sp->dI);
    prm.pair_masks = reinterpret_cast<const unsigned long long *>(sp->dMasks);
    prm.F = reinterpret_cast<const T *>(sizeof(T) == 4 ? sp->dF32 : sp->dF64);
    prm.q = reinterpret_cast<const T *>(q);
    prm.N = N;
    prm.prims = reinterpret_cast<const DevPrim<T> *>(dev_prims);
    prm.n_prims = n_prims;
    prm.margin = (T)margin;
    prm.out_vals = reinterpret_cast<T *>(out_vals);
    prm.out_jac = reinterpret_cast<T *>(out_jac);
    prm.out_valid = out_valid;
    prm.i_total = (int)(lean ? sp->I_lean.size() : sp->I.size());
    prm.f_total = sizeof(T) == 4 ? sp->I[Q_THR_F] : (int)sp->F.size();     // the fp64-only threshold table is the tail of F
    if (KIND == KIND_COLLFREE && !RED) {        // lane tables stay in global memory: stage the tree description only
        prm.i_total = sp->I[Q_TABB_I];
        prm.f_total = sp->I[Q_TABA_F];
    }
    if (KIND == KIND_COLLFREE && host_prims) prm.prim0 = *reinterpret_cast<const DevPrim<T> *>(host_prims);
    if (KIND == KIND_COLLFREE && SDFK == SDF_GRID_FAST) {
        const DevPrim<T> &pr = prm.prim0;
        GridFast &g = prm.grid;
        g.ox = (float)((double)pr.pos[0] + (double)pr.par[0]); g.oy = (float)((double)pr.pos[1] + (double)pr.par[1]);
        g.oz = (float)((double)pr.pos[2] + (double)pr.par[2]);
        g.ix_ = pr.par[3]; g.iy_ = pr.par[4]; g.iz_ = pr.par[5];
        g.tx = (float)(pr.dims[0] - 1); g.ty = (float)(pr.dims[1] - 1); g.tz = (float)(pr.dims[2] - 1);
        g.mx = pr.dims[0] - 2; g.my = pr.dims[1] - 2; g.mz = pr.dims[2] - 2;
        g.sy = (unsigned)(pr.dims[1] - 1); g.sz = (unsigned)(pr.dims[2] - 1);
        g.fill = pr.par[6];
        g.cells = reinterpret_cast<const float *>(pr.cells);
        g.tex = pr.cells_tex;
    }
    const int D = p->D, R = sp->rows, NP = (D + 1) / 2;
    prm.D = D; prm.NP = NP; prm.R = R; prm.n_steps = (int)sp->steps.size(); prm.n_levels = sp->n_levels; prm.n_cen = sp->n_cen;
    const int al = 16 / (int)sizeof(T);          // elements of T per 16 bytes
    prm.copy_vec = (((long long)R * D) % al == 0 && reinterpret_cast<uintptr_t>(out_jac) % 16 == 0) ? 1 : 0;
    int stage_rows = 32;                         // a program without fused steps stages one half-round
    for (int s = 0; s < prm.n_steps; ++s) {
        prm.steps[s] = sp->steps[s];
        unsigned long long pact = 0;            // the kernel wants pair-activity bits, not column bits
        for (int jp = 0; jp < NP; ++jp) if ((sp->steps[s].cm >> (2 * jp)) & 3ull) pact |= 1ull << jp;
        prm.steps[s].cm = pact;
        const int len = (sp->steps[s].cntA + sp->steps[s].cntB) * D;
        if ((sp->steps[s].s0 * D) % al != 0 || len % al != 0) prm.copy_vec = 0;
        if (sp->steps[s].cntB > 0) stage_rows = 64;
    }
    prm.off_F = align_up(prm.i_total * 4, 16);
    prm.off_prims = align_up(prm.off_F + prm.f_total * (int)sizeof(T), 16);
    prm.off_lv = align_up(prm.off_prims + ((KIND == KIND_COLLFREE && SDFK == SDF_UNION) ? n_prims : 0) * (int)sizeof(DevPrim<T>), 16);
    prm.n_nodes = sp->n_nodes;
    prm.off_warp = align_up(prm.off_lv + sp->n_nodes * 16, 128);
    prm.fstride = sp->frame_stride;
    prm.tstride = odd_quads(NP * 12);
    prm.cplane = align_up(sp->n_cen, 4);
    prm.cstride = KIND == KIND_PAIRWISE ? (SKB_CEN_SOA ? 3 * prm.cplane : odd_quads(4 * sp->n_cen)) : 0;
    const size_t smem_limit = 227 * 1024;
    auto layout = [&](int g, int warps) {
        int e = 0;
        prm.woff_q = e;      e += align_up(g * D, al);
        prm.woff_sc = e;     e += align_up(2 * g * D, al);
        // q / sin-cos are dead after phase 1, the staging block is written in phase 2 only: they share the head of the slab
        const int stage_elems = (JAC && !RED && !CSC) ? align_up(stage_rows * D, al) : 0;
        prm.woff_stage = 0;  e = std::max(e, stage_elems);
        prm.woff_frames = e; e += g * prm.fstride;
        prm.woff_tw = e;     e += JAC ? g * prm.tstride : 0;
        prm.woff_cen = e;    e += g * prm.cstride;
        prm.warp_elems = align_up(e, al);
        return (size_t)prm.off_warp + (size_t)warps * prm.warp_elems * sizeof(T);
    };
    // Configurations per tile G against warps per CTA and CTAs per SM.  A larger G fills more lanes of the tree walk
    // (three lanes per node, `slots3` nodes of a level at once) but costs shared memory, i.e. resident warps.  Every
    // candidate is scored with a small issue-slot model:  (resident warps)^0.75 / (FK slots + phase-2 slots per
    // configuration);  the occupancy comes from the runtime (registers and shared memory).
    auto kern = s2_kernel<T, KIND, SDFK, VW, MaskT, RED, JAC, NPMAX, CSC>;
    static const int cand[][2] = {{8, 8}, {8, 6}, {8, 5}, {8, 4}, {8, 3}, {8, 2}, {4, 8}, {4, 7}, {4, 6}, {4, 5}, {4, 4}, {4, 3}, {4, 2},
                                  {2, 8}, {2, 6}, {2, 5}, {2, 4}, {2, 3}, {2, 2}, {1, 8}, {1, 6}, {1, 4}, {1, 2}, {1, 1}};
    int halves = 0;
    for (auto &st : sp->steps) halves += st.cntB > 0 ? 2 : 1;
    const double phase2 = (JAC && !RED) ? sp->phase2_slots : (RED ? 40.0 : 150.0) * halves;
    auto fk_slots = [&](int g) {
        const int s3 = 32 / (3 * g);
        double it = 0;
        for (int lv = 0; lv < sp->n_levels; ++lv) it += (sp->I[sp->I[Q_LEVEL_I] + 2 * lv + 1] + s3 - 1) / s3;
        return 90.0 * it / g;
    };
    L.G = 0; L.score = -1.0;
    // The choice below costs two occupancy queries per candidate (tens of microseconds in all): it depends only on the
    // program's structure, the kernel instantiation and the tuning knobs, so it is made once and cached -- a planner's
    // N = 1 call pays a map lookup.
    static std::mutex shape_mu;
    static std::map<std::tuple<unsigned long long, const void *, int, int, int, int>, std::tuple<int, int, int, int, double>> shape_cache;
    static const bool lanemap_on = !(getenv("SKB_S2_LANEMAP") && getenv("SKB_S2_LANEMAP")[0] == '0');
    static const int env_g = getenv("SKB_S2_G") ? atoi(getenv("SKB_S2_G")) : 0, env_w = getenv("SKB_S2_WARPS") ? atoi(getenv("SKB_S2_WARPS")) : 0;
    const int want_g = force_g > 0 ? force_g : (p->tune_chunk > 0 ? p->tune_chunk : env_g);      // tuning knobs
    const int want_w = force_w > 0 ? force_w : (p->tune_warps > 0 ? p->tune_warps : env_w);
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    const auto shape_key = std::make_tuple(sp->hash, (const void *)kern, want_g, want_w, p->tune_ctas, cur_dev);
    {
        std::lock_guard<std::mutex> lk(shape_mu);
        auto it = shape_cache.find(shape_key);
        if (it != shape_cache.end()) {
            std::tie(L.G, L.warps, L.per_sm, L.fstride, L.score) = it->second;
            if (L.G == 0) return force_g > 0 ? -1 : set_error("step kernel: shared-memory footprint exceeds 227 KB");
            prm.fstride = L.fstride;
            L.smem = layout(L.G, L.warps);
            prm.G = L.G;
            prm.logG = 0;
            while ((1 << prm.logG) < L.G) ++prm.logG;
            prm.slots3 = 32 / (3 * L.G);
            prm.n_tiles = (N + L.G - 1) / L.G;
            prm.use_lanemap = (sizeof(T) == 4 && lanemap_on && fk_lane_map(L.G, L.fstride / 4, prm.lanemap)) ? 1 : 0;
            return 0;
        }
    }
    // Frame stride of the FK walk: lane -> (row = lane % 3, configuration = lane / 3 % G) reads / writes the 16-byte row
    // `3 node + row` of its configuration's frame block.  Eight consecutive lanes (one pass of a 128-bit shared-memory
    // access) cover configurations c, c+1, c+2: with a stride of 3 (mod 8) quads their rows fall into eight different
    // 16-byte bank groups (PR2 dual arm with the minimal odd stride of 45 quads: 8 wavefronts per parent-row load against
    // 4 ideal, 5.6 against 2.9 for the frame store).  The padding is taken only where it is free: the unified L1 /
    // shared-memory array is carved in steps (... 132, 164, 196, 228 KB), and 3 CTAs of 66.9 instead of 63.9 KB move the
    // SM from the 196 KB to the 228 KB carve-out -- 28 instead of 60 KB of L1 for the voxel-grid gathers cost the full-output
    // PR2 launch 7 %, far more than the conflicts.
    const int fs_min = sp->frame_stride;
    int fs_pad = fs_min / 4;
    while (fs_pad % 8 != 3) ++fs_pad;
    fs_pad *= 4;
    for (auto &c : cand) {
        if (want_g > 0 && c[0] != want_g) continue;
        if (want_w > 0 && c[1] != want_w) continue;
        if (want_w == 0 && c[1] > 1 && (c[1] & 1)) continue;           // odd CTAs load the four sub-partitions unevenly: autotuner only
        prm.fstride = fs_min;
        const size_t sm = layout(c[0], c[1]);
        if (sm > smem_limit) continue;
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, c[1] * 32, sm) != cudaSuccess || occ < 1) continue;
        int fs = fs_min;
        if (fs_pad != fs_min) {                 // the padded stride, when it is free (see above)
            prm.fstride = fs_pad;
            const size_t sm2 = layout(c[0], c[1]);
            int occ2 = 0;
            if (sm2 <= smem_limit && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, kern, c[1] * 32, sm2) == cudaSuccess &&
                occ2 == occ && carve_kb(occ2 * (sm2 + 1024)) == carve_kb(occ * (sm + 1024)))
                fs = fs_pad;
        }
        if (p->tune_ctas > 0) occ = std::min(occ, p->tune_ctas);
        const double score = std::pow((double)std::min(c[1] * occ, 48), 0.75) / (fk_slots(c[0]) + phase2);
        if (score > L.score * 1.0001) { L.score = score; L.G = c[0]; L.warps = c[1]; L.per_sm = occ; L.fstride = fs; }
    }
    {
        std::lock_guard<std::mutex> lk(shape_mu);
        shape_cache[shape_key] = std::make_tuple(L.G, L.warps, L.per_sm, L.fstride, L.score);
    }
    if (L.G == 0) return force_g > 0 ? -1 : set_error("step kernel: shared-memory footprint exceeds 227 KB");
    prm.fstride = L.fstride;
    L.smem = layout(L.G, L.warps);
    prm.G = L.G;
    prm.logG = 0;
    while ((1 << prm.logG) < L.G) ++prm.logG;
    prm.slots3 = 32 / (3 * L.G);
    prm.n_tiles = (N + L.G - 1) / L.G;
    prm.use_lanemap = (sizeof(T) == 4 && lanemap_on && fk_lane_map(L.G, L.fstride / 4, prm.lanemap)) ? 1 : 0;
    return 0;
}

// Launch-configuration cache of the autotuner: (structure hash of the program, kernel instantiation, output class) -> choice.
// The key is CONTENT based (skb_s2_program::hash covers the tree, the row tables and the step list, not the folded
// constants), so a choice survives every plan rebuild after set_q / reflect_skrobot_model and can be persisted:
// SKB_TUNE_CACHE=<file> is read once and appended to whenever a new choice is made.
struct S2Choice { int unfused, G, warps; };
typedef std::tuple<unsigned long long, unsigned, unsigned> S2Key;
static std::mutex g_tune_mu;
static std::map<S2Key, S2Choice> g_tune;
static bool g_tune_loaded = false;
static thread_local int tl_tune_request = 0;        // 1: skb_plan_tune is running on this thread; 2: ... and a choice was made;
                                                    // 3: skb_plan_get_tuned (look the stored choice up, launch nothing)
static thread_local S2Choice tl_tune_result = {-1, 0, 0};

static void tune_cache_load_locked() {
    if (g_tune_loaded) return;
    g_tune_loaded = true;
    const char *path = getenv("SKB_TUNE_CACHE");
    if (!path) return;
    if (FILE *f = fopen(path, "r")) {
        unsigned long long h; unsigned inst, flags; int u, g, w;
        while (fscanf(f, "%llx %x %x %d %d %d", &h, &inst, &flags, &u, &g, &w) == 6)
            if (u >= 0 && u <= 1 && g >= 1 && g <= 8 && w >= 1 && w <= 8) g_tune[S2Key(h, inst, flags)] = S2Choice{u, g, w};
        fclose(f);
    }
}
static bool tune_lookup(const S2Key &key, S2Choice &ch) {
    std::lock_guard<std::mutex> lk(g_tune_mu);
    tune_cache_load_locked();
    auto it = g_tune.find(key);
    if (it == g_tune.end()) return false;
    ch = it->second;
    return true;
}
static void tune_store(const S2Key &key, const S2Choice &ch) {
    std::lock_guard<std::mutex> lk(g_tune_mu);
    tune_cache_load_locked();
    g_tune[key] = ch;
    if (const char *path = getenv("SKB_TUNE_CACHE"))
        if (FILE *f = fopen(path, "a")) {
            fprintf(f, "%llx %x %x %d %d %d\n", std::get<0>(key), std::get<1>(key), std::get<2>(key), ch.unfused, ch.G, ch.warps);
            fclose(f);
        }
}
// skb_plan_tune (skb_api.cpp) brackets one compute call with these: the step-kernel launch inside it times its candidates
// whatever the batch size and reports the choice
void s2_tune_begin(int query_only) { tl_tune_request = query_only ? 3 : 1; tl_tune_result = S2Choice{-1, 0, 0}; }
int s2_tune_end(int32_t out_choice[3]) {
    const int made = tl_tune_request == 2 ? 1 : 0;
    tl_tune_request = 0;
    if (out_choice) { out_choice[0] = tl_tune_result.unfused; out_choice[1] = tl_tune_result.G; out_choice[2] = tl_tune_result.warps; }
    return made;
}

template <typename T, int KIND, int SDFK, int VW, typename MaskT, bool RED, bool JAC, int NPMAX = 1, bool CSC = false>
static int launch_s2_inst(const skb_plan *p, const skb_s2_program *sp, const void *q, int64_t N, const void *dev_prims,
                          const void *host_prims, int n_prims, double margin, void *out_vals, void *out_jac, uint8_t *out_valid,
                          cudaStream_t stream) {
    auto kern = s2_kernel<T, KIND, SDFK, VW, MaskT, RED, JAC, NPMAX, CSC>;
    {
        static std::atomic<unsigned> attr_done{0};                      // bit per device: set once, not on every launch
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev >= 32 || !(attr_done.load(std::memory_order_relaxed) & (1u << dev))) {
            CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            if (dev < 32) attr_done.fetch_or(1u << dev);
        }
    }
    static const bool verbose = getenv("SKB_S2_VERBOSE") != nullptr;
    auto configure = [&](const skb_s2_program *prog, int64_t n, S2Launch<T> &out, int fg, int fw) {
        return configure_s2<T, KIND, SDFK, VW, MaskT, RED, JAC, NPMAX, CSC>(p, prog, q, n, dev_prims, host_prims, n_prims, margin, out_vals,
                                                                            out_jac, out_valid, out, fg, fw);
    };
    auto launch = [&](const S2Launch<T> &L) {
        const long long want = (L.prm.n_tiles + L.warps - 1) / L.warps;
        const int grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)sm_count_cached() * L.per_sm));
        kern<<<grid, L.warps * 32, L.smem, stream>>>(L.prm);
        ++g_launch_count;
    };
    const bool has_alt = JAC && !RED && sp->unfused != nullptr;      // the same rows without fused steps (half the staging block)
    // the issue-slot model's choice: the fused program unless the unfused one raises the resident warps substantially
    auto heuristic = [&](int64_t n, S2Launch<T> &L, bool &alt) {
        alt = false;
        if (configure(sp, n, L, 0, 0) != 0) return -1;
        if (has_alt) {
            S2Launch<T> U;
            if (configure(sp->unfused, n, U, 0, 0) == 0) {
                if (verbose)
                    fprintf(stderr, "[skb] step kernel: fused G=%d warps=%d ctas=%d smem=%zu score=%.5f | unfused G=%d warps=%d ctas=%d smem=%zu score=%.5f\n",
                            L.G, L.warps, L.per_sm, L.smem, L.score, U.G, U.warps, U.per_sm, U.smem, U.score);
                // measured: going from 12 to 18-20 resident warps pays (JAXON env +20 %, PR2 dual arm fp64 +10 %), from 24 to 28
                // it does not (the unfused steps' extra twist loads outweigh it), which the score's warps^0.75 term overrates
                if (U.score > L.score && 4 * U.warps * U.per_sm >= 5 * L.warps * L.per_sm) { L = U; alt = true; }
            }
        }
        return 0;
    };

    // ---- measured choice.  Which (program, G, warps per CTA) is fastest depends on how CTAs and warps land on the SM
    // sub-partitions, the L1 carve-out and the DRAM access pattern -- PR2 dual arm: 4 CTAs of 6 warps run at 0.70 of the
    // HBM roofline, 3 CTAs of 8 warps (the same 24 resident warps) at 0.67, 4 CTAs of 7 warps at 0.65 -- which no
    // occupancy model predicts.  Candidates are timed (every one writes bit-identical results)
    //   * by skb_plan_tune(): explicitly, on scratch outputs of its own, for any batch size, or
    //   * inside a user call only when the batch has at least SKB_S2_AUTOTUNE_MIN configurations (default 2^20: a call
    //     that long does not notice the 30-130 ms) and SKB_S2_AUTOTUNE is not 0;
    // the choice is kept under a content key (see above), reused by every later launch of a structurally equal program and
    // optionally persisted (SKB_TUNE_CACHE).  Smaller batches without a stored choice, launches inside a stream capture and
    // forced configurations use the issue-slot model.
    static const int tune_min = getenv("SKB_S2_AUTOTUNE_MIN") ? atoi(getenv("SKB_S2_AUTOTUNE_MIN")) : (1 << 20);
    static const bool autotune = getenv("SKB_S2_AUTOTUNE") == nullptr || atoi(getenv("SKB_S2_AUTOTUNE")) != 0;
    static const bool env_forced = getenv("SKB_S2_G") || getenv("SKB_S2_WARPS");
    const bool forced = p->tune_chunk > 0 || p->tune_warps > 0 || p->tune_ctas > 0 || env_forced;
    const bool explicit_tune = tl_tune_request != 0;
    if ((autotune || explicit_tune) && !forced) {
        constexpr unsigned inst = (unsigned)sizeof(T) | (unsigned)KIND << 4 | (unsigned)SDFK << 8 | (unsigned)VW << 12 |
                                  (unsigned)sizeof(MaskT) << 16 | (RED ? 1u << 20 : 0u) | (JAC ? 1u << 21 : 0u) | (CSC ? 1u << 22 : 0u) | (unsigned)NPMAX << 24;
        const S2Key key(sp->hash, inst, (unsigned)((reinterpret_cast<uintptr_t>(out_jac) % 16 == 0) | (out_vals != nullptr ? 2 : 0) | (out_valid != nullptr ? 4 : 0)));
        S2Choice ch{-1, 0, 0};
        if (tl_tune_request == 3) {                 // query: report the stored choice, launch nothing
            if (tune_lookup(key, ch)) { tl_tune_result = ch; tl_tune_request = 2; }
            return 0;
        }
        const bool known = !explicit_tune && tune_lookup(key, ch);
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (!known && !explicit_tune && N < tune_min) ch.unfused = -2;           // too small to time inside a user call: model
        if (!known && ch.unfused == -1 && (cudaStreamIsCapturing(stream, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone)) ch.unfused = -2;   // no timing inside a capture
        if (!known && ch.unfused == -1) {
            const int64_t n_try = std::min<int64_t>(N, 262144);
            cudaEvent_t e0 = nullptr, e1 = nullptr;
            CUDA_OK(cudaEventCreate(&e0));
            if (cudaEventCreate(&e1) != cudaSuccess) { cudaEventDestroy(e0); return set_error("step kernel autotuner: cudaEventCreate failed"); }
            static const int gs[] = {8, 4, 2, 1}, ws[] = {8, 7, 6, 5, 4, 3, 2};
            float best_ms = 1e30f, model_ms = 1e30f;
            S2Choice model{-1, 0, 0};                                    // what the issue-slot model would launch
            {
                S2Launch<T> H;
                bool alt = false;
                if (heuristic(n_try, H, alt) == 0) model = S2Choice{alt ? 1 : 0, H.G, H.warps};
            }
            for (int alt = 0; alt < (has_alt ? 2 : 1); ++alt)
                for (int g : gs)
                    for (int w : ws) {
                        S2Launch<T> C;
                        if (configure(alt ? sp->unfused : sp, n_try, C, g, w) != 0) continue;
                        float ms = 1e30f;
                        for (int rep = 0; rep < 2; ++rep) {             // first run warms the instruction cache and L2
                            cudaEventRecord(e0, stream);
                            launch(C);
                            cudaEventRecord(e1, stream);
                            if (cudaEventSynchronize(e1) != cudaSuccess) break;
                            float t = 0.f;
                            if (cudaEventElapsedTime(&t, e0, e1) == cudaSuccess) ms = std::min(ms, t);
                        }
                        if (verbose) fprintf(stderr, "[skb] autotune %s G=%d warps=%d ctas=%d: %.3f ms\n", alt ? "unfused" : "fused", g, w, C.per_sm, ms);
                        if (ms < best_ms) { best_ms = ms; ch = S2Choice{alt, g, w}; }
                        if (alt == model.unfused && g == model.G && w == model.warps) model_ms = ms;
                    }
            // timings of a 0.2 - 1 ms slice scatter by a few per cent: the model's choice stands unless it is clearly beaten
            if (model.unfused >= 0 && best_ms > 0.97f * model_ms) ch = model;
            cudaEventDestroy(e0);
            cudaEventDestroy(e1);
            CUDA_OK(cudaGetLastError());
            if (ch.unfused >= 0) {
                tune_store(key, ch);
                if (explicit_tune) { tl_tune_request = 2; tl_tune_result = ch; }
                if (verbose) fprintf(stderr, "[skb] autotune picked %s G=%d warps=%d\n", ch.unfused ? "unfused" : "fused", ch.G, ch.warps);
            }
        }
        if (ch.unfused >= 0) {
            S2Launch<T> L;
            if (configure((ch.unfused && has_alt) ? sp->unfused : sp, N, L, ch.G, ch.warps) == 0) {
                launch(L);
                CUDA_OK(cudaGetLastError());
                return 0;
            }
        }
    }

    // ---- heuristic choice (small batches, stream capture, autotuning switched off)
    S2Launch<T> L;
    bool alt_unused = false;
    if (heuristic(N, L, alt_unused) != 0) return -1;
    launch(L);
    CUDA_OK(cudaGetLastError());
    return 0;
}

#define S2_ARGS p, sp, q, N, dev_prims, host_prims, n_prims, margin, out_vals, out_jac, out_valid, stream
#define S2_SIG const skb_plan *p, const skb_s2_program *sp, const void *q, int64_t N, const void *dev_prims, const void *host_prims, \
               int n_prims, double margin, void *out_vals, void *out_jac, uint8_t *out_valid, cudaStream_t stream

template <typename T, int KIND, int SDFK>
static int launch_s2_mode(S2_SIG, int mode) {
    const bool even = p->D % 2 == 0, wide = p->D > 32, jac = out_jac != nullptr;
    if (mode == SKB_MODE_CLOSEST || out_valid != nullptr) {
        if (mode != SKB_MODE_CLOSEST && (out_vals != nullptr || jac))
            return set_error("step kernel: validity bytes come with the closest mode or alone");
        if (wide) return jac ? launch_s2_inst<T, KIND, SDFK, 1, unsigned long long, true, true>(S2_ARGS)
                             : launch_s2_inst<T, KIND, SDFK, 1, unsigned long long, true, false>(S2_ARGS);
        return jac ? launch_s2_inst<T, KIND, SDFK, 1, unsigned, true, true>(S2_ARGS)
                   : launch_s2_inst<T, KIND, SDFK, 1, unsigned, true, false>(S2_ARGS);
    }
    if (!jac) return launch_s2_inst<T, KIND, SDFK, 1, unsigned, false, false>(S2_ARGS);
    const int NP = (p->D + 1) / 2;
    if (mode == S2_MODE_ALL_CSC) {              // transposed Jacobian blocks (CollFreeConst rows of an SQP trajectory stack)
        if (KIND != KIND_COLLFREE || NP > 16) return set_error("csc output: CollFreeConst with at most 32 columns");
#define S2_CSC(NPM) (even ? launch_s2_inst<T, KIND_COLLFREE, (KIND == KIND_COLLFREE ? SDFK : SDF_SINGLE), 2, unsigned, false, true, NPM, true>(S2_ARGS) \
                          : launch_s2_inst<T, KIND_COLLFREE, (KIND == KIND_COLLFREE ? SDFK : SDF_SINGLE), 1, unsigned, false, true, NPM, true>(S2_ARGS))
        if (NP <= 4) return S2_CSC(4);
        if (NP <= 8) return S2_CSC(8);
        return S2_CSC(16);
#undef S2_CSC
    }
#define S2_FULL(VWv, MT, NPM) launch_s2_inst<T, KIND, SDFK, VWv, MT, false, true, NPM>(S2_ARGS)
    if (NP <= 4) return even ? S2_FULL(2, unsigned, 4) : S2_FULL(1, unsigned, 4);
    if (NP <= 8) return even ? S2_FULL(2, unsigned, 8) : S2_FULL(1, unsigned, 8);
    if (NP <= 16) return even ? S2_FULL(2, unsigned, 16) : S2_FULL(1, unsigned, 16);
    (void)wide;
    return even ? S2_FULL(2, unsigned long long, 32) : S2_FULL(1, unsigned long long, 32);
#undef S2_FULL
}

template <typename T>
static int launch_s2_t(S2_SIG, int kind, int mode) {
    if (kind == KIND_PAIRWISE) return launch_s2_mode<T, KIND_PAIRWISE, SDF_SINGLE>(S2_ARGS, mode);
    if (n_prims == 1) {
        const DevPrim<T> *pr = reinterpret_cast<const DevPrim<T> *>(host_prims);
        const long long cells = (long long)(pr->dims[0] - 1) * (pr->dims[1] - 1) * (pr->dims[2] - 1);
        if (sizeof(T) == 4 && pr->type == PRIM_GRID && (pr->flags & PF_GRID_CELLS) && (pr->flags & PF_IDENTITY) &&
            !(pr->flags & PF_GRID_F64) && cells < (1ll << 31))
            return launch_s2_mode<T, KIND_COLLFREE, SDF_GRID_FAST>(S2_ARGS, mode);
        return launch_s2_mode<T, KIND_COLLFREE, SDF_SINGLE>(S2_ARGS, mode);
    }
    return launch_s2_mode<T, KIND_COLLFREE, SDF_UNION>(S2_ARGS, mode);
}

int launch_s2(const skb_plan *p, const skb_s2_program *sp, int kind, int dtype, const void *q, int64_t N, const void *dev_prims,
              const void *host_prims, int n_prims, double margin, int mode, void *out_vals, void *out_jac, uint8_t *out_valid,
              void *stream_) {
    if (N <= 0) return 0;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (dtype == SKB_F32) return launch_s2_t<float>(S2_ARGS, kind, mode);
    if (dtype == SKB_F64) return launch_s2_t<double>(S2_ARGS, kind, mode);
    return set_error("bad dtype");
}

}  // namespace skb

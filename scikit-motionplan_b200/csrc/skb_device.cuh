// skb_device.cuh -- device-side building blocks shared by the kernels:
// real-type traits, SDF primitives (value + analytic gradient), bulk-store (TMA 1-D) helpers.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "skb_internal.h"

namespace skb {

template <typename T> struct Real;
template <> struct Real<float> {
    static __device__ __forceinline__ void sincos(float x, float *s, float *c) { sincosf(x, s, c); }
    static __device__ __forceinline__ float sqrt(float x) { return sqrtf(x); }
    static __device__ __forceinline__ float rsqrt(float x) { return rsqrtf(x); }
    // norm and its reciprocal from the squared norm: one MUFU.RSQ (2 ulp) instead of an IEEE sqrt and an IEEE division
    static __device__ __forceinline__ void norm_inv(float n2, float &n, float &inv) { inv = rsqrtf(fmaxf(n2, 1e-36f)); n = n2 * inv; }
    static __device__ __forceinline__ float abs(float x) { return fabsf(x); }
    static __device__ __forceinline__ float floor(float x) { return floorf(x); }
    static __device__ __forceinline__ float atan2(float y, float x) { return atan2f(y, x); }
    static __device__ __forceinline__ float asin(float x) { return asinf(x); }
    static __device__ __forceinline__ float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }   // never contracted
    static __device__ __forceinline__ float inf() { return __int_as_float(0x7f800000); }
    static __device__ __forceinline__ float max() { return __int_as_float(0x7f7fffff); }
};
template <> struct Real<double> {
    static __device__ __forceinline__ void sincos(double x, double *s, double *c) { ::sincos(x, s, c); }
    static __device__ __forceinline__ double sqrt(double x) { return ::sqrt(x); }
    static __device__ __forceinline__ double rsqrt(double x) { return 1.0 / ::sqrt(x); }
    static __device__ __forceinline__ void norm_inv(double n2, double &n, double &inv) { n = ::sqrt(n2); inv = 1.0 / n; }
    static __device__ __forceinline__ double abs(double x) { return ::fabs(x); }
    static __device__ __forceinline__ double floor(double x) { return ::floor(x); }
    static __device__ __forceinline__ double atan2(double y, double x) { return ::atan2(y, x); }
    static __device__ __forceinline__ double asin(double x) { return ::asin(x); }
    static __device__ __forceinline__ double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double inf() { return __longlong_as_double(0x7ff0000000000000LL); }
    static __device__ __forceinline__ double max() { return __longlong_as_double(0x7fefffffffffffffLL); }
};

// Fixed-association building blocks of the chain rule.  Every kernel instantiation (all-spheres, closest, packed
// f32x2) must produce the same bits for the same sphere, so the contraction order is spelled out instead of being
// left to the compiler's FMA contraction: cross = fma(a, b, -(c*d)), dot6 = mul, then five fma in w.xyz, v.xyz order
// (exactly what the packed mul.rn.f32x2 / fma.rn.f32x2 chain computes per half).
template <typename T>
__device__ __forceinline__ void wrench_moment(T px, T py, T pz, T gx, T gy, T gz, T &mx, T &my, T &mz) {
    using R = Real<T>;
    mx = R::fma(py, gz, -R::mul(pz, gy));
    my = R::fma(pz, gx, -R::mul(px, gz));
    mz = R::fma(px, gy, -R::mul(py, gx));
}
template <typename T>
__device__ __forceinline__ T dot6(T wx, T wy, T wz, T vx, T vy, T vz, T mx, T my, T mz, T gx, T gy, T gz) {
    using R = Real<T>;
    T a = R::mul(wx, mx);
    a = R::fma(wy, my, a);
    a = R::fma(wz, mz, a);
    a = R::fma(vx, gx, a);
    a = R::fma(vy, gy, a);
    return R::fma(vz, gz, a);
}

// ------------------------------------------------------------------ SDF primitives (local frame)
// Each returns the signed distance and writes the analytic gradient (local frame).
// Tie-breaking (>= / first index) mirrors oracle/skmp_oracle.c so kinks resolve identically.
template <typename T>
__device__ __forceinline__ T sdf_box_local(const T *h, T x, T y, T z, T &gx, T &gy, T &gz) {
    using R = Real<T>;
    T sx = R::abs(x) - h[0], sy = R::abs(y) - h[1], sz = R::abs(z) - h[2];
    T sgx = x >= T(0) ? T(1) : T(-1), sgy = y >= T(0) ? T(1) : T(-1), sgz = z >= T(0) ? T(1) : T(-1);
    if (sx > T(0) || sy > T(0) || sz > T(0)) {
        T qx = sx > T(0) ? sx : T(0), qy = sy > T(0) ? sy : T(0), qz = sz > T(0) ? sz : T(0);
        T d, inv;
        R::norm_inv(qx * qx + qy * qy + qz * qz, d, inv);
        gx = sgx * qx * inv; gy = sgy * qy * inv; gz = sgz * qz * inv;
        return d;
    }
    int k = 0; T m = sx;
    if (sy > m) { m = sy; k = 1; }
    if (sz > m) { m = sz; k = 2; }
    gx = k == 0 ? sgx : T(0); gy = k == 1 ? sgy : T(0); gz = k == 2 ? sgz : T(0);
    return m;
}

template <typename T>
__device__ __forceinline__ T sdf_cylinder_local(T radius, T hh, T x, T y, T z, T &gx, T &gy, T &gz) {
    using R = Real<T>;
    T rho, irho;
    R::norm_inv(x * x + y * y, rho, irho);
    T s0 = rho - radius, s1 = R::abs(z) - hh;
    T ux = rho > T(0) ? x * irho : T(0), uy = rho > T(0) ? y * irho : T(0);
    T sgz = z >= T(0) ? T(1) : T(-1);
    if (s0 > T(0) || s1 > T(0)) {
        T q0 = s0 > T(0) ? s0 : T(0), q1 = s1 > T(0) ? s1 : T(0);
        T d, inv;
        R::norm_inv(q0 * q0 + q1 * q1, d, inv);
        gx = ux * q0 * inv; gy = uy * q0 * inv; gz = sgz * q1 * inv;
        return d;
    }
    if (s0 >= s1) { gx = ux; gy = uy; gz = T(0); return s0; }
    gx = T(0); gy = T(0); gz = sgz;
    return s1;
}

template <typename T>
__device__ __forceinline__ T sdf_sphere_local(T radius, T x, T y, T z, T &gx, T &gy, T &gz) {
    using R = Real<T>;
    T n, inv;
    R::norm_inv(x * x + y * y + z * z, n, inv);
    inv = n > T(0) ? inv : T(0);
    gx = x * inv; gy = y * inv; gz = z * inv;
    return n - radius;
}

template <typename T>
__device__ __forceinline__ T grid_fetch(const void *grid, bool f64, int64_t idx) {
    if (f64) return (T)__ldg(reinterpret_cast<const double *>(grid) + idx);
    return (T)__ldg(reinterpret_cast<const float *>(grid) + idx);
}

// 32-byte gather of one cell's coefficient pack (sm_100: LDG.E.256, one L1 wavefront per lane)
__device__ __forceinline__ void ldg_cell(const float *p, float c[8]) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(c[0]), "=f"(c[1]), "=f"(c[2]), "=f"(c[3]), "=f"(c[4]), "=f"(c[5]), "=f"(c[6]), "=f"(c[7])
                 : "l"(p));
}
__device__ __forceinline__ void ldg_cell(const double *p, double c[8]) {
    const double2 *q = reinterpret_cast<const double2 *>(p);
    const double2 a = __ldg(q), b = __ldg(q + 1), d = __ldg(q + 2), e = __ldg(q + 3);
    c[0] = a.x; c[1] = a.y; c[2] = b.x; c[3] = b.y; c[4] = d.x; c[5] = d.y; c[6] = e.x; c[7] = e.y;
}

// Trilinear interpolant of one voxel cell as a POLYNOMIAL in the in-cell coordinates (fx, fy, fz) in [0, 1]:
//     f = a0 + a1 fx + a2 fy + a3 fz + a4 fx fy + a5 fx fz + a6 fy fz + a7 fx fy fz
// with a1 = c100 - c000, a4 = c110 - c100 - c010 + c000, ... formed in fp64 when the cell packs are expanded
// (expand_cells_kernel, skb_sphere.cu).  Value and gradient are Horner evaluations: no difference of two interpolated
// O(1 m) numbers is ever taken, so the fp32 gradient carries the rounding of its own (centimetre-sized) coefficients
// instead of 1 ulp of the distance divided by the voxel size (the corner-lerp form lost 1e-5 relative there).
// Explicit fma / no contraction: every kernel instantiation returns the same bits for the same point.
// Returns the value; (gx, gy, gz) = d f / d (fx, fy, fz) (per cell unit -- the caller scales by 1 / spacing).
template <typename T>
__device__ __forceinline__ T trilinear_poly(const T a[8], T fx, T fy, T fz, T &gx, T &gy, T &gz) {
    using R = Real<T>;
    const T p = R::fma(fz, a[7], a[4]);                 // d2f / dfx dfy
    const T r = R::fma(fx, a[7], a[6]);                 // d2f / dfy dfz
    gx = R::fma(fz, a[5], R::fma(fy, p, a[1]));
    gy = R::fma(fz, r, R::fma(fx, a[4], a[2]));
    gz = R::fma(fy, r, R::fma(fx, a[5], a[3]));
    const T t = R::fma(fz, a[6], a[2]);
    return R::fma(fx, gx, R::fma(fy, t, R::fma(fz, a[3], a[0])));
}
// the same coefficients from the eight corner values (grids without cell packs): differences of neighbouring samples
template <typename T>
__device__ __forceinline__ void trilinear_coeffs(T c000, T c001, T c010, T c011, T c100, T c101, T c110, T c111, T a[8]) {
    a[0] = c000;
    a[1] = c100 - c000; a[2] = c010 - c000; a[3] = c001 - c000;
    a[4] = (c110 - c100) - a[2];
    a[5] = (c101 - c100) - a[3];
    a[6] = (c011 - c010) - a[3];
    a[7] = ((c111 - c110) - (c101 - c100)) - ((c011 - c010) - a[3]);
}

// trilinear voxel grid: value + analytic gradient; `fill`, zero gradient outside the sample box
template <typename T>
__device__ __forceinline__ T sdf_grid_local(const DevPrim<T> &pr, T x, T y, T z, T &gx, T &gy, T &gz) {
    using R = Real<T>;
    const T ux = (x - pr.par[0]) * pr.par[3], uy = (y - pr.par[1]) * pr.par[4], uz = (z - pr.par[2]) * pr.par[5];
    const T tx = T(pr.dims[0] - 1), ty = T(pr.dims[1] - 1), tz = T(pr.dims[2] - 1);
    if (!(ux >= T(0) && ux <= tx && uy >= T(0) && uy <= ty && uz >= T(0) && uz <= tz)) {
        gx = gy = gz = T(0);
        return pr.par[6];
    }
    int ix = min((int)R::floor(ux), pr.dims[0] - 2);
    int iy = min((int)R::floor(uy), pr.dims[1] - 2);
    int iz = min((int)R::floor(uz), pr.dims[2] - 2);
    const T fx = ux - T(ix), fy = uy - T(iy), fz = uz - T(iz);
    const bool f64 = (pr.flags & PF_GRID_F64) != 0;
    T a[8];
    // the packs are rounded to the grid's value type: fp64 arithmetic over an fp32 grid reads the corner samples instead
    // (their differences are exact in fp64), so the fp64 build option keeps its 1e-10
    if ((pr.flags & PF_GRID_CELLS) && (sizeof(T) == 4 || f64)) {
        const int64_t cell = ((int64_t)ix * (pr.dims[1] - 1) + iy) * (pr.dims[2] - 1) + iz;
        if (f64) {
            double c[8];
            ldg_cell(reinterpret_cast<const double *>(pr.cells) + cell * 8, c);
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = (T)c[k];
        } else {
            float c[8];
            ldg_cell(reinterpret_cast<const float *>(pr.cells) + cell * 8, c);
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = (T)c[k];
        }
    } else {
        const int64_t sy = pr.dims[2], sx = (int64_t)pr.dims[1] * pr.dims[2];
        const int64_t b = ix * sx + iy * sy + iz;
        trilinear_coeffs<T>(grid_fetch<T>(pr.grid, f64, b), grid_fetch<T>(pr.grid, f64, b + 1),
                            grid_fetch<T>(pr.grid, f64, b + sy), grid_fetch<T>(pr.grid, f64, b + sy + 1),
                            grid_fetch<T>(pr.grid, f64, b + sx), grid_fetch<T>(pr.grid, f64, b + sx + 1),
                            grid_fetch<T>(pr.grid, f64, b + sx + sy), grid_fetch<T>(pr.grid, f64, b + sx + sy + 1), a);
    }
    T hx, hy, hz;
    const T val = trilinear_poly<T>(a, fx, fy, fz, hx, hy, hz);
    gx = hx * pr.par[3]; gy = hy * pr.par[4]; gz = hz * pr.par[5];
    return val;
}

// one primitive: world point -> distance + WORLD gradient
template <typename T>
__device__ __forceinline__ T sdf_prim_world(const DevPrim<T> &pr, T px, T py, T pz, T &gx, T &gy, T &gz) {
    T dx = px - pr.pos[0], dy = py - pr.pos[1], dz = pz - pr.pos[2];
    T lx, ly, lz;
    const bool ident = (pr.flags & PF_IDENTITY) != 0;
    if (ident) { lx = dx; ly = dy; lz = dz; }
    else {   // local = R^T d
        lx = pr.rot[0] * dx + pr.rot[3] * dy + pr.rot[6] * dz;
        ly = pr.rot[1] * dx + pr.rot[4] * dy + pr.rot[7] * dz;
        lz = pr.rot[2] * dx + pr.rot[5] * dy + pr.rot[8] * dz;
    }
    T ax, ay, az, d;
    if (pr.type == PRIM_BOX) d = sdf_box_local<T>(pr.par, lx, ly, lz, ax, ay, az);
    else if (pr.type == PRIM_GRID) d = sdf_grid_local<T>(pr, lx, ly, lz, ax, ay, az);
    else if (pr.type == PRIM_CYLINDER) d = sdf_cylinder_local<T>(pr.par[0], pr.par[1], lx, ly, lz, ax, ay, az);
    else d = sdf_sphere_local<T>(pr.par[0], lx, ly, lz, ax, ay, az);
    if (ident) { gx = ax; gy = ay; gz = az; }
    else {
        gx = pr.rot[0] * ax + pr.rot[1] * ay + pr.rot[2] * az;
        gy = pr.rot[3] * ax + pr.rot[4] * ay + pr.rot[5] * az;
        gz = pr.rot[6] * ax + pr.rot[7] * ay + pr.rot[8] * az;
    }
    return d;
}

// union (min) of primitives: world point -> distance + world gradient of the closest primitive
// (first index wins ties, like np.argmin over the members)
template <typename T>
__device__ __forceinline__ T sdf_union(const DevPrim<T> *prims, int np, T px, T py, T pz, T &gx, T &gy, T &gz) {
    T best = Real<T>::inf();
    gx = gy = gz = T(0);
    for (int i = 0; i < np; ++i) {
        T ax, ay, az;
        const T d = sdf_prim_world<T>(prims[i], px, py, pz, ax, ay, az);
        if (d < best) { best = d; gx = ax; gy = ay; gz = az; }
    }
    return best;
}

// ------------------------------------------------------------------ forward kinematics of one tile (warp-cooperative)
template <typename T> struct alignas(16) Vec4 { T x, y, z, w; };

// Level-by-level tree walk for the G configurations of a warp's tile, THREE lanes per (configuration, node): row r of a
// node frame  Y = parent * pre * Rz(q)  depends only on row r of the parent frame, so each lane carries one row
// [R_r0 R_r1 R_r2 t_r]; the joint twist (omega = third column, v = o x omega) needs the other two rows' (o, omega)
// entries, fetched with four warp shuffles.  lane -> (row = lane % 3, cfg = (lane / 3) % G, slot = lane / 3 / G);
// `slots3` = 32 / (3 G) nodes of a level are processed at once (G <= 8).  Frames go to `wframes` (rows of 4, 12 per node),
// twists to `wtw` in the scalar layout [wx wy wz . | vx vy vz .] per column or, PAIR, component-interleaved per column
// pair [wx_j wx_j1 wy_j wy_j1 | wz_j wz_j1 vx_j vx_j1 | vy_j vy_j1 vz_j vz_j1].
// Per-CTA derived table for the walk, built once in the kernel prologue: entry e of the level-ordered node list ->
// {type, column, 3 * parent (-1 for the world node), 3 * node}, so the level loop does one 16-byte load per node instead
// of chasing levelNodes -> node record and re-deriving the frame row indices.
__device__ __forceinline__ void build_level_records(int4 *lvrec, const int32_t *nodeI, const int32_t *levelNodes, int n_nodes) {
    for (int e = threadIdx.x; e < n_nodes; e += blockDim.x) {
        const int node = levelNodes[e];
        const int4 nd = reinterpret_cast<const int4 *>(nodeI)[node];
        lvrec[e] = make_int4(nd.x, nd.y, nd.z < 0 ? -1 : 3 * nd.z, 3 * node);
    }
}

// Shared-window loads / stores by 32-bit address: the walk below keeps its lane-constant base addresses in registers and
// every access is [base + offset], with no generic-pointer arithmetic for the compiler to re-materialise.
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ int4 lds_i4(uint32_t a) {
    int4 v;
    asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ int2 lds_i2(uint32_t a) {
    int2 v;
    asm volatile("ld.shared.v2.s32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ Vec4<float> lds_v4(uint32_t a, float) {
    Vec4<float> v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ Vec4<double> lds_v4(uint32_t a, double) {
    Vec4<double> v;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.z), "=d"(v.w) : "r"(a + 16u));
    return v;
}
__device__ __forceinline__ void sts_v4(uint32_t a, const Vec4<float> &v) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts_v4(uint32_t a, const Vec4<double> &v) {
    asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(a), "d"(v.x), "d"(v.y) : "memory");
    asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(a + 16u), "d"(v.z), "d"(v.w) : "memory");
}
__device__ __forceinline__ void lds_sc(uint32_t a, float &s, float &c) { asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(s), "=f"(c) : "r"(a)); }
__device__ __forceinline__ void lds_sc(uint32_t a, double &s, double &c) { asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(s), "=d"(c) : "r"(a)); }
__device__ __forceinline__ float lds_1(uint32_t a, float) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ double lds_1(uint32_t a, double) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts_1(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void sts_1(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }

// PLANER / FLOATING base: the chain's last node (NODE_BASE_*, alone in level 1; skb_host.cpp::level_layout) in closed form from
// the base columns col .. col + 2 / col + 5 -- row `row` of  [Rz(yaw) Ry(pitch) Rx(roll) A | x y z]  with the node constant A
// (the z-normalising rotation of the chain's last axis), and the twists (omega, v = o x omega) of all base columns -- instead
// of three / six almost empty passes of the level loop.  Everything below sees the frame the joint-by-joint walk would
// produce.  Not inlined on purpose: robots with a fixed base never run it and must not pay registers for it (the fp32 voxel-
// grid kernel sits exactly at its 64-register budget).
template <typename T, bool WITH_JAC, bool PAIR>
__device__ __noinline__ void fk_base_node(const int type, const int col, const int node3, const int row, const bool on,
                                          const uint32_t a_q, const uint32_t a_sc, const uint32_t a_nodeF, const uint32_t a_nodeFr,
                                          const uint32_t a_fr, const uint32_t a_tw) {
    using V4 = Vec4<T>;
    constexpr uint32_t ES = sizeof(T), VS = 4 * sizeof(T);
    T bx = lds_1(a_q + col * ES, T()), by = lds_1(a_q + (col + 1) * ES, T()), bz = T(0);
    T sy_, cy_, sp_ = T(0), cp_ = T(1), sr_ = T(0), cr_ = T(1);
    if (type == NODE_BASE_FLOATING) {
        bz = lds_1(a_q + (col + 2) * ES, T());
        lds_sc(a_sc + 2 * (col + 3) * ES, sr_, cr_);
        lds_sc(a_sc + 2 * (col + 4) * ES, sp_, cp_);
        lds_sc(a_sc + 2 * (col + 5) * ES, sy_, cy_);
    } else lds_sc(a_sc + 2 * (col + 2) * ES, sy_, cy_);
    const T u = row == 0 ? cy_ : sy_, w = row == 0 ? -sy_ : cy_;           // rows 0 / 1 of Rz(yaw) are (u, w, 0)
    V4 P;
    if (row < 2) { P.x = u * cp_; P.y = u * sp_ * sr_ + w * cr_; P.z = u * sp_ * cr_ - w * sr_; P.w = row == 0 ? bx : by; }
    else { P.x = -sp_; P.y = cp_ * sr_; P.z = cp_ * cr_; P.w = bz; }
    const uint32_t ac = a_nodeF + node3 * VS;
    const V4 c0 = lds_v4(ac, T()), c1 = lds_v4(ac + VS, T()), c2 = lds_v4(ac + 2 * VS, T());
    V4 y;
    y.x = P.x * c0.x + P.y * c1.x + P.z * c2.x;
    y.y = P.x * c0.y + P.y * c1.y + P.z * c2.y;
    y.z = P.x * c0.z + P.y * c1.z + P.z * c2.z;
    y.w = P.w + P.x * c0.w + P.y * c1.w + P.z * c2.w;
    if (!on) return;
    sts_v4(a_fr + node3 * VS, y);
    sts_v4(a_fr, lds_v4(a_nodeFr, T()));                                  // the world node (level 0): its constant row
    if (!WITH_JAC) return;
    auto put = [&](const int c, const T w_r, const T v_r) {
        const uint32_t at = a_tw + (PAIR ? ((c >> 1) * 12 + (c & 1)) : c * 8) * ES;
        sts_1(at, w_r);
        sts_1(at + (PAIR ? 6 : 4) * ES, v_r);
    };
    const T e0 = row == 0 ? T(1) : T(0), e1 = row == 1 ? T(1) : T(0), e2 = row == 2 ? T(1) : T(0);
    put(col, T(0), e0);                                               // x: translation along e_x
    put(col + 1, T(0), e1);                                           // y
    const T vyaw = row == 0 ? by : (row == 1 ? -bx : T(0));           // (x, y, z) x e_z = (y, -x, 0)
    if (type == NODE_BASE_FLOATING) {
        put(col + 2, T(0), e2);                                       // z
        put(col + 5, e2, vyaw);                                       // yaw about e_z
        // pitch about Rz(yaw) e_y = (-sy, cy, 0):  o x omega = (-z cy, -z sy, x cy + y sy)
        put(col + 4, row == 0 ? -sy_ : (row == 1 ? cy_ : T(0)),
            row == 0 ? -bz * cy_ : (row == 1 ? -bz * sy_ : bx * cy_ + by * sy_));
        // roll about Rz(yaw) Ry(pitch) e_x = (cy cp, sy cp, -sp)
        const T wx = cy_ * cp_, wy = sy_ * cp_, wz = -sp_;
        put(col + 3, row == 0 ? wx : (row == 1 ? wy : wz),
            row == 0 ? by * wz - bz * wy : (row == 1 ? bz * wx - bx * wz : bx * wy - by * wx));
    } else put(col + 2, e2, vyaw);                                    // theta about e_z
}

template <typename T, bool WITH_JAC, bool PAIR>
__device__ __forceinline__ void fk_phase(const int lane, const int G, const int logG, const int slots3, const int g_live,
                                         const int D, const int n_levels, const int4 *lvrec, const int32_t *levelI,
                                         const Vec4<T> *nodeF, const T *wq, const T *wsc,
                                         T *wframes, const int fstride, T *wtw, const int tstride,
                                         const unsigned *lanemap = nullptr) {
    using V4 = Vec4<T>;
    constexpr uint32_t ES = sizeof(T), VS = 4 * sizeof(T);     // element / 4-vector size in bytes
    int row = lane % 3, t = lane / 3;
    int cfg = t & (G - 1), slot = t >> logG;
    int src1 = lane - row + (row == 2 ? 0 : row + 1), src2 = lane - row + (row == 0 ? 2 : row - 1);
    // Optional lane permutation (fk_lane_map, host): WHICH lane carries which (row, configuration) of a node is free, and the
    // eight lanes of one 128-bit shared-memory pass should hit eight different 16-byte bank groups of the frame blocks --
    // the plain order does not for most frame strides (PR2 dual arm, 45 quads: 8 wavefronts per parent-row load against 4).
    // Entry: row | cfg << 2 | slot << 5 | lane of row + 1 << 8 | lane of row + 2 << 16.
    if (lanemap != nullptr) {
        const unsigned m = lanemap[lane];
        row = (int)(m & 3u); cfg = (int)((m >> 2) & 7u); slot = (int)((m >> 5) & 7u);
        src1 = (int)((m >> 8) & 31u); src2 = (int)((m >> 16) & 31u);
    }
    const bool act = slot < slots3;
    const int ccfg = min(cfg, g_live - 1);          // dead configurations of a ragged last tile recompute a live one
    // lane-constant shared-window addresses, made opaque so that each is formed once per tile and held in a register
    uint32_t a_lv = smem_addr(lvrec), a_lev = smem_addr(levelI), a_nodeF = smem_addr(nodeF);
    uint32_t a_nodeFr = a_nodeF + row * VS;
    uint32_t a_q = smem_addr(wq) + ccfg * D * ES, a_sc = smem_addr(wsc) + 2 * ccfg * D * ES;
    uint32_t a_fr = smem_addr(wframes) + cfg * fstride * ES + row * VS;     // this lane's row of every node frame
    uint32_t a_tw = smem_addr(wtw) + (cfg * tstride + (PAIR ? 2 * row : row)) * ES;
    if (sizeof(T) == 4)      // the fp64 instantiations are register-bound: let the compiler re-derive there
        asm volatile("" : "+r"(src1), "+r"(src2), "+r"(a_lv), "+r"(a_lev), "+r"(a_nodeF), "+r"(a_nodeFr), "+r"(a_q), "+r"(a_sc),
                     "+r"(a_fr), "+r"(a_tw));
    // PLANER / FLOATING base: the chain's last node (NODE_BASE_*, alone in level 1) is evaluated in closed form by fk_base_node
    int lv_first = 0;
    if (n_levels > 1) {
        const int2 l1 = lds_i2(a_lev + 8u);
        const int4 rec = lds_i4(a_lv + 16u * l1.x);
        if (rec.x >= NODE_BASE_PLANER) {
            lv_first = 2;
            fk_base_node<T, WITH_JAC, PAIR>(rec.x, rec.y, rec.w, row, act && slot == 0, a_q, a_sc, a_nodeF, a_nodeFr, a_fr, a_tw);
            __syncwarp();
        }
    }
    // level 0 is the world node alone (its frame is its constant): stored once, so every node of the loop below has a parent
    if (lv_first == 0) {
        if (act && slot == 0) sts_v4(a_fr, lds_v4(a_nodeFr, T()));
        lv_first = 1;
        __syncwarp();
    }
    for (int lv = lv_first; lv < n_levels; ++lv) {
        const int2 lrec = lds_i2(a_lev + 8u * lv);                          // {first entry of the level, count}
        const int lbeg = lrec.x, lcnt = lrec.y;
        for (int k0 = 0; k0 < lcnt; k0 += slots3) {
            const int k = k0 + slot;
            const bool on = act && k < lcnt;
            const int4 rec = lds_i4(a_lv + 16u * (lbeg + (on ? k : 0)));    // {type, column, 3 parent, 3 node}
            const int type = rec.x, col = rec.y, par3 = rec.z, node3 = rec.w;
            const uint32_t ac = a_nodeF + node3 * VS;
            const V4 c0 = lds_v4(ac, T()), c1 = lds_v4(ac + VS, T()), c2 = lds_v4(ac + 2 * VS, T());
            const V4 P = lds_v4(a_fr + par3 * VS, T());
            V4 y;
            y.x = P.x * c0.x + P.y * c1.x + P.z * c2.x;
            y.y = P.x * c0.y + P.y * c1.y + P.z * c2.y;
            y.z = P.x * c0.z + P.y * c1.z + P.z * c2.z;
            y.w = P.w + P.x * c0.w + P.y * c1.w + P.z * c2.w;
            // (o, omega) entries of the two other rows: joint axis = local +z = third column, origin = fourth
            const T o1 = __shfl_sync(0xffffffffu, y.w, src1), o2 = __shfl_sync(0xffffffffu, y.w, src2);
            const T w1 = __shfl_sync(0xffffffffu, y.z, src1), w2 = __shfl_sync(0xffffffffu, y.z, src2);
            // joint motion without branches: a revolute joint turns (x, y) by its (sin, cos), a prismatic one slides along z
            // by q, anything else is the identity -- (s, c, d) = (0, 1, 0) leaves the row bit for bit as it is
            const bool rev = type == NODE_REVOLUTE, pri = type == NODE_PRISMATIC;
            const int colc = max(col, 0);
            T s, c;
            lds_sc(a_sc + 2 * colc * ES, s, c);
            const T d = pri ? T(1) : T(0);                                       // (selects, not branches: the loads below are unconditional)
            const T qv = lds_1(a_q + colc * ES, T());
            s = rev ? s : T(0);
            c = rev ? c : T(1);
            const T wr = rev ? y.z : T(0);
            const T vr = rev ? o1 * w2 - o2 * w1 : (pri ? y.z : T(0));       // v = o x omega, component `row` / the sliding axis
            const T a = y.x, b = y.y;
            y.x = c * a + s * b; y.y = c * b - s * a;
            y.w += (d * qv) * y.z;
            if (on) {
                sts_v4(a_fr + node3 * VS, y);
                if (WITH_JAC && col >= 0) {
                    const uint32_t at = a_tw + (PAIR ? ((col >> 1) * 12 + (col & 1)) : col * 8) * ES;
                    sts_1(at, wr);
                    sts_1(at + (PAIR ? 6 : 4) * ES, vr);
                }
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------ bulk async store helpers (sm_90+/sm_100a)
// cp.async.bulk shared::cta -> global (1-D TMA): SASS UBLKCP.  16-byte aligned src/dst/size.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes)
                 : "memory");
}
// same with an L2 evict-first policy: the output stream should not displace the L2-resident voxel grid
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_store_hint(void *gdst, const void *ssrc, uint32_t bytes, uint64_t policy) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst), "r"(smem_u32(ssrc)),
                 "r"(bytes), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy before a bulk store reads them
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace skb

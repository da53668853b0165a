// skb_internal.h -- shared between the host-side plan compiler and the CUDA kernels.
// Not part of the public ABI (that is include/skmp_b200.h).
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include <map>
#include <mutex>
#include <atomic>

#include "../../include/skmp_b200.h"

namespace skb {

// ------------------------------------------------------------------ device program layout
// A plan is uploaded as two flat arrays: int32 words I[] and real words F[] (float or double).
// The kernels stage both into shared memory once per CTA.  All indices below are in words.
//
// I[0 .. HDR_WORDS)            header (enum Hdr)
// I[node_i + n*NODE_I ..]      node records   (enum NodeI)
// I[chunk_i + c*CHUNK_I ..]    chunk records  (enum ChunkI)
// I[anc_i ..]                  ancestor column lists (one per node, di+1 entries)
// I[zero_i ..]                 zero-fill column lists (one per chunk)
// F[node_f + n*NODE_F ..]      node constants: pre-transform R(9) t(3) quat(4)
// F[feat_f + f*feat_stride ..] feature constants in processing order: off(3) radius [R(9) quat(4) pad(3)]
enum Hdr {
    H_N_NODES = 0, H_N_CHUNKS, H_N_FEAT, H_D, H_ROWS, H_ROT_TYPE, H_MAX_ACTIVE, H_N_SLOTS,
    H_CHUNK_CAP, H_N_BUF, H_RUNP, H_FEAT_STRIDE,
    H_NODE_I, H_CHUNK_I, H_ANC_I, H_ZERO_I, H_I_TOTAL,
    H_NODE_F, H_FEAT_F, H_F_TOTAL,
    H_VEC_OK,      // 1 when every chunk run is 16-byte aligned in the output (bulk-store path)
    H_VALS_VEC_OK, // 1 when per-chunk value rows are 16-byte aligned
    HDR_WORDS = 24
};
enum NodeI { N_TYPE = 0, N_COL, N_DI, N_SRC, N_SAVE, N_CHUNK_BEGIN, N_CHUNK_END, N_ANC, NODE_I = 8 };
enum NodeF { NF_R = 0, NF_T = 9, NF_Q = 12, NODE_F = 16 };
enum ChunkI { C_FEAT_BEGIN = 0, C_COUNT, C_OUT0, C_ACTIVE, C_ZERO_BEGIN, C_ZERO_COUNT, C_BUF, C_PAD, CHUNK_I = 8 };
enum { FEAT_F_POINT = 4, FEAT_F_POSE = 20 };
enum { SRC_IDENTITY = -1, SRC_CURRENT = -2 };   // >= 0: restore from slot
enum { MAX_SLOTS = 8 };
enum { NODE_NONE = 0, NODE_REVOLUTE = 1, NODE_PRISMATIC = 2,
       // level programs only (skb_host.cpp::level_layout): the chain of a PLANER / FLOATING base -- three / six single-node
       // levels -- is ONE node evaluated in closed form; its other nodes appear in no level
       NODE_BASE_PLANER = 3, NODE_BASE_FLOATING = 4, NODE_BASE_SKIP = 5 };

// pairwise program (separate upload): all-node evaluation order + per-feature node binding
// PI header (enum PHdr), node records reuse NodeI/NodeF (chunk fields unused),
// feature records: I: node index ; F: off(3) radius
// pair records: I: fi, fj, cls ; F: threshold
// class records: I: begin, count into term list ; terms: I: node, col, side(+1: use xi, -1: use xj)
enum PHdr {
    P_N_NODES = 0, P_N_FEAT, P_D, P_N_PAIRS, P_N_CLASSES, P_N_SLOTS,
    P_NODE_I, P_FEATNODE_I, P_PAIR_I, P_CLASS_I, P_TERM_I, P_I_TOTAL,
    P_NODE_F, P_FEAT_F, P_PAIR_F, P_F_TOTAL,
    PHDR_WORDS = 16
};

// device SDF primitive, staged to shared memory by the kernels
template <typename T>
struct DevPrim {
    int32_t type;        // SKB sdf type
    int32_t flags;       // bit0: rotation is identity; bit1: grid data is f64
    int32_t dims[3];
    int32_t pad;
    T pos[3];
    T rot[9];
    T par[8];            // box: half[3]; cyl: radius, half height; sphere: radius; grid: lb[3], inv_spacing[3], fill
    const void *grid;    // device pointer (float or double values), C order [nx][ny][nz]
    const void *cells;   // device pointer or NULL: per-cell polynomial packs [(nx-1)][(ny-1)][(nz-1)][8]
                         // {a0 .. a7} of trilinear_poly (skb_device.cuh), same value type as `grid`
    unsigned long long cells_tex;   // cudaTextureObject_t over `cells` as float4 texels (fp32 packs only), or 0
};
enum { PRIM_BOX = 0, PRIM_CYLINDER = 1, PRIM_SPHERE = 2, PRIM_GRID = 3 };
enum { PF_IDENTITY = 1, PF_GRID_F64 = 2, PF_GRID_CELLS = 4 };

struct Xf {              // rigid transform, quaternion (x,y,z,w) + translation, host fp64
    double q[4];
    double t[3];
};

}  // namespace skb

// ------------------------------------------------------------------ opaque handle definitions
struct skb_model {
    std::vector<int32_t> parent, jtype;
    std::vector<double> origin;   // 6 per link
    std::vector<double> axis;     // 3 per link
    std::vector<double> angle;    // sticky joint value per link
    double base_pose[6];
};

struct skb_host_prim {
    int32_t type;
    int32_t dims[3];
    double pos[3], rot[9], par[8];
    void *dev_grid = nullptr;
    void *dev_cells = nullptr;       // per-cell polynomial packs (one 32/64-byte gather per lookup), may be NULL
    unsigned long long cells_tex = 0;   // texture object over the fp32 packs (two float4 fetches per lookup on the TEX pipe), or 0
    int32_t grid_is_f64 = 0;
    size_t grid_bytes = 0;
};

struct skb_sdf {
    std::vector<skb_host_prim> prims;
    int device = -1;                 // device that holds the grids / primitive tables (-1: nothing uploaded yet)
    std::mutex mu;                   // guards the lazy uploads
    // device copies of the primitive table, built lazily per dtype
    void *dev_f32 = nullptr;
    void *dev_f64 = nullptr;
    std::vector<unsigned char> host_f32, host_f64;   // host copies of the same DevPrim tables
    int dev_count = -1;
};

struct skb_schedule {            // one staged-output schedule for a given rows-per-feature
    int rows = 0;
    std::vector<int32_t> I;
    std::vector<double> F;
    void *dI = nullptr;
    void *dF32 = nullptr;
    void *dF64 = nullptr;
    int max_active = 0;
    int chunk_cap = 0, n_buf = 0, runp = 0;
};

struct skb_pair_program {
    std::vector<int32_t> I;
    std::vector<double> F;
    void *dI = nullptr;
    void *dF32 = nullptr;
    void *dF64 = nullptr;
    int n_pairs = 0;
    int n_nodes = 0;
};

// sphere program (skb_sphere.cu): the fused FK + SDF + Jacobian kernel for point features.
// Phase 1 walks the tree level by level (lane = (configuration, node of the level)), phase 2 assigns
// one lane per sphere ("round" = up to 32 consecutive output spheres of one configuration).
//   I: header (enum SHdr) | node records (4 ints) | level records (2) | level node lists |
//      round records (4) | per-round lane table B: int4 {node, amask_lo, amask_hi, valid}
//   F: node pre-transforms as 3 rows [R_r0 R_r1 R_r2 t_r] (12) | per-round lane table A: {off xyz, radius}
enum SHdr {
    S_N_NODES = 0, S_N_LEVELS, S_N_ROUNDS, S_D, S_N_FEAT, S_FRAME_STRIDE, S_TW_STRIDE,
    S_NODE_I, S_LEVEL_I, S_LEVELNODE_I, S_ROUND_I, S_TABB_I, S_I_TOTAL,
    S_NODE_F, S_TABA_F, S_F_TOTAL, S_MAX_LEVEL_NODES,
    SHDR_WORDS = 20
};

struct skb_sphere_program {
    std::vector<int32_t> I;
    std::vector<double> F;
    void *dI = nullptr;
    void *dF32 = nullptr;
    void *dF64 = nullptr;
    int n_nodes = 0, n_levels = 0, n_rounds = 0;
    int frame_stride = 0, tw_stride = 0;     // elements of T per configuration
    bool built = false, supported = false;
};

// step program (skb_sphere2.cu): the two-rows-per-lane kernel for the full-output modes of CollFreeConst and
// PairWiseSelfCollFreeConst.  A "step" covers up to 64 consecutive output rows (spheres or pairs) of one
// configuration: lane l owns row s0 + l (half A) and row s0 + 32 + l (half B, only when half A is full).
//   I: header (enum QHdr) | node records (4 ints) | level records (2) | level node lists |
//      lane table B: int4 per (step, half, lane)
//         collfree  {node, amask_lo, amask_hi, 0}
//         pairwise  {centre_i | centre_j << 16, plus_mask, minus_mask, float bits of (r_i + r_j)^2}
//      | pairwise only: node per centre
//   F: node pre-transforms (12) | collfree: lane table A {off xyz, radius} per (step, half, lane)
//      | pairwise: centre table {off xyz, radius} per sphere
enum QHdr {
    Q_N_NODES = 0, Q_N_LEVELS, Q_N_STEPS, Q_D, Q_ROWS, Q_N_CEN,
    Q_NODE_I, Q_LEVEL_I, Q_LEVELNODE_I, Q_TABB_I, Q_CENNODE_I, Q_I_TOTAL,
    Q_NODE_F, Q_TABA_F, Q_CENA_F, Q_F_TOTAL,
    Q_THR_F,       // pairwise: (r_i + r_j)^2 per (step, half, lane) as a real (read by the fp64 kernel; fp32 packs it in table B)
    Q_TABC_I,      // pairwise with D > 32: int4 per (step, half, lane) {plus_mask_hi, minus_mask_hi, 0, 0}
    Q_TABP_I,      // pairwise, lean program only: int2 per row {centre_i | centre_j << 16, float bits of (r_i + r_j)^2}
    Q_N_HR,        // pairwise, lean program only: half-rounds of 32 rows in table P, padded to a multiple of 4
    QHDR_WORDS = 20
};
enum { S2_MAX_STEPS = 96 };
struct skb_s2_step { unsigned long long cm; int32_t s0; int16_t cntA, cntB; };   // cm: columns any row of the step touches

struct skb_s2_program {
    std::vector<int32_t> I;
    std::vector<double> F;
    std::vector<skb_s2_step> steps;
    void *dI = nullptr;
    void *dF32 = nullptr;
    void *dF64 = nullptr;
    // pairwise, reducing modes (closest row / validity byte): the "lean" program replaces lane tables B and C by the 8-byte
    // table P (a third of the shared memory); the column masks of the one winning row come from `masks` in global memory
    std::vector<int32_t> I_lean;
    std::vector<uint64_t> masks;     // {plus, minus} per row
    void *dI_lean = nullptr;
    void *dMasks = nullptr;
    int n_nodes = 0, n_levels = 0, rows = 0, n_cen = 0;
    int frame_stride = 0;            // floats per configuration
    unsigned long long hash = 0;     // FNV-1a of the program's STRUCTURE (I, step list, rows, D): the autotuner's cache key
    double phase2_slots = 0;         // issue-slot model of the steps of one configuration (the planner's cost)
    // the same rows without fused steps: half the staging block per warp, i.e. more resident warps when shared memory
    // limits the occupancy (wide rows, deep trees); the launcher scores both and takes the better one
    skb_s2_program *unfused = nullptr;
    bool built = false, supported = false;
};

// centre-of-mass program (skb_com.cu): features = weighted points sorted by the DFS preorder of their node
//   I: header (enum CHdr) | node records (4) | level records (2) | level node lists | node per point |
//      per column {first point, one past the last point} of the joint's subtree
//   F: node pre-transforms (12) | point table {off xyz, weight}
enum CHdr {
    C_N_NODES = 0, C_N_LEVELS, C_N_POINTS, C_D,
    C_NODE_I, C_LEVEL_I, C_LEVELNODE_I, C_PTNODE_I, C_COLRANGE_I, C_I_TOTAL,
    C_NODE_F, C_PT_F, C_F_TOTAL,
    CHDR_WORDS = 16
};
struct skb_com_program {
    std::vector<int32_t> I;
    std::vector<double> F;
    void *dI = nullptr;
    void *dF32 = nullptr;
    void *dF64 = nullptr;
    int n_nodes = 0, n_levels = 0, n_points = 0;
    bool built = false;
};

// pose program (skb_pose.cu): end-effector frames with RPY / XYZW rows on the warp-cooperative FK
//   I: header (enum EHdr) | node records (4) | level records (2) | level node lists | feature records int4 {node, amask_lo, amask_hi, 0}
//   F: node pre-transforms (12) | node pre-rotation quaternions (4) | per feature 20: off(3) pad | R_off(9) pad(3) | quat_off(4)
enum EHdr {
    E_N_NODES = 0, E_N_LEVELS, E_N_FEAT, E_D,
    E_NODE_I, E_LEVEL_I, E_LEVELNODE_I, E_FEAT_I, E_I_TOTAL,
    E_NODE_F, E_NODEQ_F, E_FEAT_F, E_F_TOTAL,
    E_CHAIN_I,     // XYZW: per feature, the quaternion factors of its chain root -> node: CHAIN_LEN columns (-1: constant factor)
    E_CHAINQ_F,    //       and CHAIN_LEN constant quaternions; factor k = chainQ[k] * Qz(q[col_k] / 2), padded with identities
    E_CHAIN_LEN,   //       factors per feature, a multiple of 4 (four lanes multiply a quarter of the chain each)
    EHDR_WORDS = 16
};
struct skb_pose_program {
    std::vector<int32_t> I;
    std::vector<double> F;
    void *dI = nullptr;
    void *dF32 = nullptr;
    void *dF64 = nullptr;
    int n_nodes = 0, n_levels = 0;
    bool built = false, supported = false;
};

struct skb_stream_ctx;           // host-pipeline scratch (skb_hostpipe.cpp)

struct skb_plan {
    int device = 0;
    int D = 0, F = 0, T = 3, rot_type = 0, base_type = 0;
    // compiled tree (host, fp64), shared by all schedules
    struct Node {
        int type, col, parent, di;
        skb::Xf pre;                 // parent-node frame -> joint frame (z-axis normalised)
        std::vector<int> children;
    };
    std::vector<Node> nodes;         // node 0 = pseudo world node
    std::vector<int> order;          // DFS preorder of node indices
    std::vector<int> feat_node;      // per feature (output order)
    std::vector<skb::Xf> feat_off;   // per feature: node frame -> feature frame
    std::vector<double> radii;       // per feature
    int max_active = 0;
    // tuning
    int tune_chunk = 0, tune_nbuf = 0, tune_warps = 0, tune_ctas = 0;
    std::map<int, skb_schedule> schedules;   // keyed by rows-per-feature
    skb_pair_program pairs;
    skb_sphere_program sphere;
    skb_s2_program s2_collfree, s2_pairs;
    skb_com_program com;
    skb_pose_program pose;
    std::vector<double> weights;     // per feature (COM plans), empty otherwise
    bool pairs_set = false;
    std::vector<int32_t> pair_idx;
    std::vector<double> pair_thr;
    skb_stream_ctx *hostpipe = nullptr;
    std::mutex mu;                   // guards the lazily built / uploaded programs
    std::mutex host_mu;              // serialises the *_host entry points of one plan (they share its streams and scratch)
};

namespace skb {

int set_error(const std::string &msg);           // returns -1
const skb_schedule *get_schedule(skb_plan *p, int rows);   // builds + uploads lazily; nullptr on error
const skb_pair_program *get_pair_program(skb_plan *p);
const skb_sphere_program *get_sphere_program(skb_plan *p);   // nullptr when the plan is not eligible (not an error)
void s2_tune_begin(int query_only);                                       // skb_sphere2.cu: brackets of skb_plan_tune (thread local)
int s2_tune_end(int32_t out_choice[3]);                       // 1 when a choice was measured and stored
const skb_s2_program *get_s2_program(skb_plan *p, int kind);   // KIND_COLLFREE / KIND_PAIRWISE; nullptr when not eligible
const void *get_dev_prims(skb_sdf *s, int dtype);
const void *get_host_prims(const skb_sdf *s, int dtype);   // valid after get_dev_prims succeeded for that dtype

// kernel launchers (skb_kernels.cu)
int launch_tree(const skb_plan *p, const skb_schedule *sch, int kind, int dtype, const void *q, int64_t N,
                const void *dev_prims, int n_prims, double margin, int mode, void *out_vals, void *out_jac,
                uint8_t *out_valid, void *stream);
int launch_pairwise(const skb_plan *p, const skb_pair_program *pp, int dtype, const void *q, int64_t N, int mode,
                    void *out_vals, void *out_jac, uint8_t *out_valid, void *stream);
int launch_sdf_eval(const void *dev_prims, int n_prims, int dtype, const void *pts, int64_t M, void *out_vals,
                    void *out_grads, void *stream);
// skb_sphere.cu
int launch_sphere(const skb_plan *p, const skb_sphere_program *sp, int kind, int dtype, const void *q, int64_t N,
                  const void *dev_prims, const void *host_prims, int n_prims, double margin, int mode, void *out_vals,
                  void *out_jac, uint8_t *out_valid, void *stream);
int launch_expand_cells(const void *grid, int is_f64, const int32_t dims[3], void *cells);
// skb_pose.cu
const skb_pose_program *get_pose_program(skb_plan *p);     // nullptr when the plan is not eligible (not an error)
int launch_pose(const skb_plan *p, const skb_pose_program *pp, int dtype, const void *q, int64_t N, void *out_pts, void *out_jac,
                void *stream);
// skb_com.cu
const skb_com_program *get_com_program(skb_plan *p);
int launch_com(const skb_plan *p, const skb_com_program *cp, int dtype, const void *q, int64_t N, const void *dev_prims,
               int n_prims, int what, void *out_vals, void *out_jac, void *stream);
// skb_sphere2.cu (fp32, full-output modes)
int launch_s2(const skb_plan *p, const skb_s2_program *sp, int kind, int dtype, const void *q, int64_t N, const void *dev_prims,
              const void *host_prims, int n_prims, double margin, int mode, void *out_vals, void *out_jac, uint8_t *out_valid,
              void *stream);
enum { KIND_MAP = 0, KIND_COLLFREE = 1, KIND_PAIRWISE = 2 };
enum { S2_MODE_ALL_CSC = 100 };      // internal launch_s2 mode: SKB_MODE_ALL with transposed Jacobian blocks (skb_collfree_csc)

extern std::atomic<int64_t> g_launch_count;
// every compute entry point runs on the device its handles were created on (one handle per device; skb_api.cpp)
int check_plan_device(const skb_plan *p, const char *who);
int check_sdf_device(skb_sdf *s, const char *who);     // binds an SDF without device resources to the current device

}  // namespace skb

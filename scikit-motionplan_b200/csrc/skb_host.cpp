// skb_host.cpp -- host side of the engine: kinematic model, SDF descriptors, and the plan
// compiler that turns (tree, sticky state, features, control joints, base type) into the flat
// program the CUDA kernels interpret.  Native replacement for the bookkeeping half of tinyfk
// (URDF tree, add_new_link, set_q, relevance tables) -- see include/skmp_b200.h for the
// reference call sites.  No kernel code in this file.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <limits>

#include "skb_internal.h"

namespace skb {

static thread_local std::string g_err;
std::atomic<int64_t> g_launch_count{0};

int set_error(const std::string &msg) {
    g_err = msg;
    return -1;
}

#define CUDA_OK(expr)                                                                          \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return set_error(std::string(#expr) + ": " + cudaGetErrorString(e__));             \
    } while (0)

// ------------------------------------------------------------------ fp64 rigid-transform algebra
static void qmul(const double *a, const double *b, double *o) {
    double x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
    double y = a[3] * b[1] - a[0] * b[2] + a[1] * b[3] + a[2] * b[0];
    double z = a[3] * b[2] + a[0] * b[1] - a[1] * b[0] + a[2] * b[3];
    double w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
    o[0] = x; o[1] = y; o[2] = z; o[3] = w;
}
static void qrot(const double *q, const double *v, double *o) {
    double tx = 2 * (q[1] * v[2] - q[2] * v[1]), ty = 2 * (q[2] * v[0] - q[0] * v[2]),
           tz = 2 * (q[0] * v[1] - q[1] * v[0]);
    o[0] = v[0] + q[3] * tx + (q[1] * tz - q[2] * ty);
    o[1] = v[1] + q[3] * ty + (q[2] * tx - q[0] * tz);
    o[2] = v[2] + q[3] * tz + (q[0] * ty - q[1] * tx);
}
static Xf xf_identity() { return Xf{{0, 0, 0, 1}, {0, 0, 0}}; }
static Xf xf_mul(const Xf &a, const Xf &b) {
    Xf r;
    double t[3];
    qrot(a.q, b.t, t);
    for (int i = 0; i < 3; ++i) r.t[i] = a.t[i] + t[i];
    qmul(a.q, b.q, r.q);
    return r;
}
static Xf xf_inv(const Xf &a) {
    Xf r;
    r.q[0] = -a.q[0]; r.q[1] = -a.q[1]; r.q[2] = -a.q[2]; r.q[3] = a.q[3];
    double t[3];
    qrot(r.q, a.t, t);
    for (int i = 0; i < 3; ++i) r.t[i] = -t[i];
    return r;
}
static void quat_from_rpy(double r, double p, double y, double *q) {
    double phi = r / 2, the = p / 2, psi = y / 2;
    q[0] = sin(phi) * cos(the) * cos(psi) - cos(phi) * sin(the) * sin(psi);
    q[1] = cos(phi) * sin(the) * cos(psi) + sin(phi) * cos(the) * sin(psi);
    q[2] = cos(phi) * cos(the) * sin(psi) - sin(phi) * sin(the) * cos(psi);
    q[3] = cos(phi) * cos(the) * cos(psi) + sin(phi) * sin(the) * sin(psi);
}
static Xf xf_from_xyzrpy(const double *o) {
    Xf r;
    r.t[0] = o[0]; r.t[1] = o[1]; r.t[2] = o[2];
    quat_from_rpy(o[3], o[4], o[5], r.q);
    return r;
}
static void quat_to_matrix(const double *q, double *R) {
    double x = q[0], y = q[1], z = q[2], w = q[3];
    double n = x * x + y * y + z * z + w * w;
    double s = n > 0 ? 2.0 / n : 0.0;
    R[0] = 1 - s * (y * y + z * z); R[1] = s * (x * y - z * w);     R[2] = s * (x * z + y * w);
    R[3] = s * (x * y + z * w);     R[4] = 1 - s * (x * x + z * z); R[5] = s * (y * z - x * w);
    R[6] = s * (x * z - y * w);     R[7] = s * (y * z + x * w);     R[8] = 1 - s * (x * x + y * y);
}
static Xf xf_motion(int jtype, const double *axis, double v) {
    Xf r = xf_identity();
    if (jtype == SKB_JOINT_REVOLUTE) {
        double h = 0.5 * v, sh = sin(h);
        r.q[0] = axis[0] * sh; r.q[1] = axis[1] * sh; r.q[2] = axis[2] * sh; r.q[3] = cos(h);
    } else if (jtype == SKB_JOINT_PRISMATIC) {
        r.t[0] = axis[0] * v; r.t[1] = axis[1] * v; r.t[2] = axis[2] * v;
    }
    return r;
}
// rotation A with A * e_z = a  (a unit)
static Xf xf_z_to_axis(const double *a) {
    Xf r = xf_identity();
    double n = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    if (n == 0) return r;
    double az = a[2] / n;
    if (az > 1.0 - 1e-14) return r;
    if (az < -1.0 + 1e-14) { r.q[0] = 1; r.q[1] = 0; r.q[2] = 0; r.q[3] = 0; return r; }
    // rotation axis = e_z x a, angle = acos(az)
    double kx = -a[1] / n, ky = a[0] / n, kn = sqrt(kx * kx + ky * ky);
    kx /= kn; ky /= kn;
    double ang = acos(az), sh = sin(0.5 * ang);
    r.q[0] = kx * sh; r.q[1] = ky * sh; r.q[2] = 0; r.q[3] = cos(0.5 * ang);
    return r;
}

// ------------------------------------------------------------------ plan compiler
static int compile_tree(const skb_model *m, const int32_t *feat_links, int nf, const int32_t *joint_links, int nj,
                        int base_type, int rot_type, skb_plan *p) {
    const int n = (int)m->parent.size();
    if (n == 0) return set_error("plan: empty model");
    if (base_type < 0 || base_type > 2) return set_error("plan: bad base_type");
    if (rot_type < 0 || rot_type > 2) return set_error("plan: bad rot_type");
    for (int f = 0; f < nf; ++f)
        if (feat_links[f] < 0 || feat_links[f] >= n) return set_error("plan: feature link id out of range");
    std::vector<int> ctrl(n, -1);
    for (int j = 0; j < nj; ++j) {
        int l = joint_links[j];
        if (l <= 0 || l >= n) return set_error("plan: joint id out of range (the root link has no joint)");
        if (ctrl[l] >= 0) return set_error("plan: duplicated control joint");
        ctrl[l] = j;
    }
    p->D = nj + (base_type == SKB_BASE_PLANER ? 3 : (base_type == SKB_BASE_FLOATING ? 6 : 0));
    p->F = nf;
    p->rot_type = rot_type;
    p->base_type = base_type;
    p->T = rot_type == SKB_ROT_IGNORE ? 3 : (rot_type == SKB_ROT_RPY ? 6 : 7);

    struct Raw { int type, col, parent; Xf pre; double axis[3]; };
    std::vector<Raw> raw;
    raw.push_back(Raw{NODE_NONE, -1, -1, xf_identity(), {0, 0, 1}});
    int base_node = 0;
    Xf base_const = xf_identity();
    auto add_raw = [&](int type, int col, int parent, const Xf &pre, double ax, double ay, double az) {
        raw.push_back(Raw{type, col, parent, pre, {ax, ay, az}});
        return (int)raw.size() - 1;
    };
    if (base_type == SKB_BASE_FIXED) {
        base_const = xf_from_xyzrpy(m->base_pose);
    } else if (base_type == SKB_BASE_PLANER) {
        // T_base = Trans(x, y, 0) Rz(theta)                       skmp/robot/utils.py:45-50
        int a = add_raw(NODE_PRISMATIC, nj + 0, 0, xf_identity(), 1, 0, 0);
        int b = add_raw(NODE_PRISMATIC, nj + 1, a, xf_identity(), 0, 1, 0);
        base_node = add_raw(NODE_REVOLUTE, nj + 2, b, xf_identity(), 0, 0, 1);
    } else {
        // T_base = Trans(x, y, z) Rz(yaw) Ry(pitch) Rx(roll); q tail = [x y z roll pitch yaw]
        //                                                          skmp/robot/utils.py:51-56
        int a = add_raw(NODE_PRISMATIC, nj + 0, 0, xf_identity(), 1, 0, 0);
        int b = add_raw(NODE_PRISMATIC, nj + 1, a, xf_identity(), 0, 1, 0);
        int c = add_raw(NODE_PRISMATIC, nj + 2, b, xf_identity(), 0, 0, 1);
        int y = add_raw(NODE_REVOLUTE, nj + 5, c, xf_identity(), 0, 0, 1);
        int pt = add_raw(NODE_REVOLUTE, nj + 4, y, xf_identity(), 0, 1, 0);
        base_node = add_raw(NODE_REVOLUTE, nj + 3, pt, xf_identity(), 1, 0, 0);
    }

    std::vector<char> needed(n, 0);
    for (int f = 0; f < nf; ++f)
        for (int l = feat_links[f]; l >= 0 && !needed[l]; l = m->parent[l]) needed[l] = 1;
    std::vector<int> node_of(n, -1);
    std::vector<Xf> C(n, xf_identity());
    for (int l = 0; l < n; ++l) {
        if (!needed[l]) continue;
        int par = m->parent[l];
        if (par < 0) { node_of[l] = base_node; C[l] = base_const; continue; }
        if (par >= l) return set_error("plan: links are not topologically sorted");
        Xf org = xf_from_xyzrpy(&m->origin[6 * l]);
        if (ctrl[l] >= 0 && m->jtype[l] != SKB_JOINT_FIXED) {
            const double *a = &m->axis[3 * l];
            node_of[l] = add_raw(m->jtype[l] == SKB_JOINT_REVOLUTE ? NODE_REVOLUTE : NODE_PRISMATIC, ctrl[l],
                                 node_of[par], xf_mul(C[par], org), a[0], a[1], a[2]);
            C[l] = xf_identity();
        } else {
            node_of[l] = node_of[par];
            C[l] = xf_mul(xf_mul(C[par], org), xf_motion(m->jtype[l], &m->axis[3 * l], m->angle[l]));
        }
    }
    // drop movable nodes that carry no feature below them (e.g. unused base chain never happens, but
    // control joints outside every feature's ancestry do): they contribute only zero columns.
    const int nr = (int)raw.size();
    std::vector<char> used(nr, 0);
    used[0] = 1;
    for (int f = 0; f < nf; ++f)
        for (int k = node_of[feat_links[f]]; k >= 0 && !used[k]; k = raw[k].parent) used[k] = 1;
    std::vector<int> remap(nr, -1);
    std::vector<Xf> A;   // z-normalising rotation per kept node
    p->nodes.clear();
    for (int k = 0; k < nr; ++k) {
        if (!used[k]) continue;
        remap[k] = (int)p->nodes.size();
        skb_plan::Node nd;
        nd.type = raw[k].type; nd.col = raw[k].col;
        nd.parent = raw[k].parent < 0 ? -1 : remap[raw[k].parent];
        nd.di = nd.parent < 0 ? -1 : p->nodes[nd.parent].di + 1;
        Xf a = (raw[k].type == NODE_NONE) ? xf_identity() : xf_z_to_axis(raw[k].axis);
        Xf apar = nd.parent < 0 ? xf_identity() : A[nd.parent];
        nd.pre = xf_mul(xf_mul(xf_inv(apar), raw[k].pre), a);
        A.push_back(a);
        if (nd.parent >= 0) p->nodes[nd.parent].children.push_back(remap[k]);
        p->nodes.push_back(nd);
    }
    p->feat_node.resize(nf);
    p->feat_off.resize(nf);
    for (int f = 0; f < nf; ++f) {
        int k = remap[node_of[feat_links[f]]];
        p->feat_node[f] = k;
        p->feat_off[f] = xf_mul(xf_inv(A[k]), C[feat_links[f]]);
    }
    // DFS preorder
    p->order.clear();
    std::function<void(int)> dfs = [&](int k) {
        p->order.push_back(k);
        for (int c : p->nodes[k].children) dfs(c);
    };
    dfs(0);
    p->max_active = 0;
    for (auto &nd : p->nodes) p->max_active = std::max(p->max_active, nd.di + 1);
    p->radii.assign(nf, 0.0);
    return 0;
}

static int pad_run(int floats) {   // multiple of 4 with an odd quotient: conflict-free 16-byte lanes
    int q = (floats + 3) / 4;
    if (q % 2 == 0) q += 1;
    return q * 4;
}

static int build_schedule(skb_plan *p, int rows, skb_schedule *s) {
    const int D = p->D, nf = p->F;
    const int per_feat = rows * D;
    int C = p->tune_chunk, nbuf = p->tune_nbuf;
    if (nbuf <= 0) nbuf = 2;
    if (C <= 0) {
        C = std::max(1, 64 / std::max(1, per_feat));
        if (C >= 4) C = (C / 4) * 4;
        C = std::min(C, 16);
        if (rows == 1 && C < 4) C = 4;
    }
    // keep one warp's staging under ~48 KB
    while (C > 1 && (size_t)nbuf * 32 * pad_run(C * per_feat) * 8 > 96 * 1024) C = std::max(1, C / 2);
    s->rows = rows; s->chunk_cap = C; s->n_buf = nbuf; s->runp = pad_run(C * per_feat);
    s->max_active = p->max_active;

    const int n_nodes = (int)p->nodes.size();
    // slots: a node is saved when it has a child that does not directly follow it in preorder
    std::vector<int> pos(n_nodes), save(n_nodes, -1), src(n_nodes, SRC_IDENTITY), saved_above(n_nodes, 0);
    for (int i = 0; i < n_nodes; ++i) pos[p->order[i]] = i;
    int n_slots = 0;
    for (int i = 0; i < n_nodes; ++i) {
        int k = p->order[i];
        const auto &nd = p->nodes[k];
        saved_above[k] = nd.parent < 0 ? 0 : saved_above[nd.parent] + (save[nd.parent] >= 0 ? 1 : 0);
        bool need_save = false;
        for (int c : nd.children)
            if (pos[c] != i + 1) need_save = true;
        if (need_save && k != 0) { save[k] = saved_above[k]; n_slots = std::max(n_slots, save[k] + 1); }
        if (nd.parent < 0) src[k] = SRC_IDENTITY;
        else if (nd.parent == 0) src[k] = SRC_IDENTITY;     // world node is the identity frame
        else if (pos[nd.parent] == i - 1) src[k] = SRC_CURRENT;
        else src[k] = save[nd.parent];
    }
    if (n_slots > MAX_SLOTS) return set_error("plan: kinematic tree branches too deeply (more than 8 nested branch points)");

    // chunks per node, in preorder
    struct Chunk { int feat_begin, count, out0, active, node; std::vector<int> zero; int buf; };
    std::vector<Chunk> chunks;
    std::vector<int> feat_order;   // processing order -> output feature index
    std::vector<std::vector<int>> node_feats(n_nodes);
    for (int f = 0; f < nf; ++f) node_feats[p->feat_node[f]].push_back(f);
    std::vector<int> chunk_begin(n_nodes, 0), chunk_end(n_nodes, 0);
    for (int i = 0; i < n_nodes; ++i) {
        int k = p->order[i];
        auto &fs = node_feats[k];   // ascending output index by construction
        chunk_begin[k] = (int)chunks.size();
        size_t a = 0;
        while (a < fs.size()) {
            size_t b = a + 1;
            while (b < fs.size() && fs[b] == fs[b - 1] + 1) ++b;
            // run [fs[a], fs[b-1]] -> chunks cut at multiples of the alignment quantum
            int s0 = fs[a], e = fs[b - 1] + 1;
            const int quantum = (C >= 4) ? 4 : 1;
            while (s0 < e) {
                int cut;
                if (s0 % quantum != 0) cut = std::min(e, (s0 / quantum + 1) * quantum);   // misaligned head
                else {
                    cut = std::min(e, s0 + C);
                    if (cut < e || (cut - s0) % quantum == 0) { /* full or aligned */ }
                    else if (cut - s0 > quantum) cut = s0 + ((cut - s0) / quantum) * quantum;  // leave tail
                }
                Chunk ch;
                ch.feat_begin = (int)feat_order.size();
                ch.count = cut - s0; ch.out0 = s0; ch.active = p->nodes[k].di + 1; ch.node = k; ch.buf = 0;
                for (int f = s0; f < cut; ++f) feat_order.push_back(f);
                chunks.push_back(ch);
                s0 = cut;
            }
            a = b;
        }
        chunk_end[k] = (int)chunks.size();
    }
    // ancestor column lists
    std::vector<int> anc_begin(n_nodes, 0), anc;
    for (int k = 0; k < n_nodes; ++k) {
        anc_begin[k] = (int)anc.size();
        std::vector<int> cols;
        for (int a = k; a > 0; a = p->nodes[a].parent) cols.push_back(p->nodes[a].col);
        std::reverse(cols.begin(), cols.end());
        for (int c : cols) anc.push_back(c);
    }
    // zero-fill schedule: simulate two tiles; the second tile's lists are the steady state
    std::vector<std::vector<std::vector<char>>> dirty(nbuf, std::vector<std::vector<char>>(C, std::vector<char>(D, 0)));
    for (int pass = 0; pass < 2; ++pass) {
        int b = 0;
        for (auto &ch : chunks) {
            ch.buf = b;
            std::vector<char> act(D, 0);
            for (int i = 0; i < ch.active; ++i) act[anc[anc_begin[ch.node] + i]] = 1;
            std::vector<char> z(D, 0);
            for (int r = 0; r < ch.count; ++r)
                for (int c = 0; c < D; ++c)
                    if (dirty[b][r][c] && !act[c]) z[c] = 1;
            std::vector<int> zl;
            for (int c = 0; c < D; ++c) if (z[c]) zl.push_back(c);
            if (pass == 0) ch.zero = zl;
            else { for (int c : zl) if (std::find(ch.zero.begin(), ch.zero.end(), c) == ch.zero.end()) ch.zero.push_back(c); }
            for (int r = 0; r < ch.count; ++r)
                for (int c = 0; c < D; ++c) dirty[b][r][c] = act[c];
            b = (b + 1) % nbuf;
        }
    }
    int max_zero = 0;
    std::vector<int> zero_begin(chunks.size()), zeros;
    for (size_t i = 0; i < chunks.size(); ++i) {
        zero_begin[i] = (int)zeros.size();
        for (int c : chunks[i].zero) zeros.push_back(c);
        max_zero = std::max(max_zero, (int)chunks[i].zero.size());
    }
    (void)max_zero;

    const bool pose = p->rot_type != SKB_ROT_IGNORE;
    const int feat_stride = pose ? FEAT_F_POSE : FEAT_F_POINT;
    // ---- pack
    std::vector<int32_t> &I = s->I;
    std::vector<double> &Fv = s->F;
    I.assign(HDR_WORDS, 0);
    I[H_N_NODES] = n_nodes; I[H_N_CHUNKS] = (int)chunks.size(); I[H_N_FEAT] = nf; I[H_D] = D; I[H_ROWS] = rows;
    I[H_ROT_TYPE] = p->rot_type; I[H_MAX_ACTIVE] = p->max_active; I[H_N_SLOTS] = n_slots;
    I[H_CHUNK_CAP] = C; I[H_N_BUF] = nbuf; I[H_RUNP] = s->runp; I[H_FEAT_STRIDE] = feat_stride;
    I[H_NODE_I] = (int)I.size();
    I.resize(I.size() + (size_t)n_nodes * NODE_I);
    I[H_CHUNK_I] = (int)I.size();
    I.resize(I.size() + chunks.size() * CHUNK_I);
    I[H_ANC_I] = (int)I.size();
    I.insert(I.end(), anc.begin(), anc.end());
    I[H_ZERO_I] = (int)I.size();
    I.insert(I.end(), zeros.begin(), zeros.end());
    I[H_I_TOTAL] = (int)I.size();
    for (int i = 0; i < n_nodes; ++i) {            // node records in preorder
        int k = p->order[i];
        int32_t *r = &I[I[H_NODE_I] + i * NODE_I];
        r[N_TYPE] = p->nodes[k].type; r[N_COL] = p->nodes[k].col; r[N_DI] = p->nodes[k].di;
        r[N_SRC] = src[k]; r[N_SAVE] = save[k];
        r[N_CHUNK_BEGIN] = chunk_begin[k]; r[N_CHUNK_END] = chunk_end[k]; r[N_ANC] = I[H_ANC_I] + anc_begin[k];
    }
    const int ROW = nf * rows * D, VROW = nf * rows;
    int vec_all = 1, vvec_all = 1;
    for (size_t i = 0; i < chunks.size(); ++i) {
        int32_t *r = &I[I[H_CHUNK_I] + i * CHUNK_I];
        const Chunk &ch = chunks[i];
        r[C_FEAT_BEGIN] = ch.feat_begin; r[C_COUNT] = ch.count; r[C_OUT0] = ch.out0; r[C_ACTIVE] = ch.active;
        r[C_ZERO_BEGIN] = I[H_ZERO_I] + zero_begin[i]; r[C_ZERO_COUNT] = (int)ch.zero.size(); r[C_BUF] = ch.buf;
        int jvec = (ROW % 4 == 0) && ((ch.out0 * per_feat) % 4 == 0) && ((ch.count * per_feat) % 4 == 0);
        int vvec = (VROW % 4 == 0) && ((ch.out0 * rows) % 4 == 0) && ((ch.count * rows) % 4 == 0);
        r[C_PAD] = jvec | (vvec << 1);
        vec_all &= jvec; vvec_all &= vvec;
    }
    I[H_VEC_OK] = vec_all; I[H_VALS_VEC_OK] = vvec_all;
    Fv.clear();
    I[H_NODE_F] = 0;
    Fv.resize((size_t)n_nodes * NODE_F, 0.0);
    for (int i = 0; i < n_nodes; ++i) {
        int k = p->order[i];
        double *r = &Fv[(size_t)i * NODE_F];
        quat_to_matrix(p->nodes[k].pre.q, r + NF_R);
        for (int j = 0; j < 3; ++j) r[NF_T + j] = p->nodes[k].pre.t[j];
        for (int j = 0; j < 4; ++j) r[NF_Q + j] = p->nodes[k].pre.q[j];
    }
    I[H_FEAT_F] = (int)Fv.size();
    Fv.resize(Fv.size() + (size_t)nf * feat_stride, 0.0);
    for (int i = 0; i < nf; ++i) {
        int f = feat_order[i];
        double *r = &Fv[I[H_FEAT_F] + (size_t)i * feat_stride];
        for (int j = 0; j < 3; ++j) r[j] = p->feat_off[f].t[j];
        r[3] = p->radii[f];
        if (pose) {
            quat_to_matrix(p->feat_off[f].q, r + 4);
            for (int j = 0; j < 4; ++j) r[13 + j] = p->feat_off[f].q[j];
        }
    }
    I[H_F_TOTAL] = (int)Fv.size();
    return 0;
}

template <typename T>
static int upload_reals(const std::vector<double> &src, void **dst) {
    std::vector<T> tmp(src.size() ? src.size() : 1);
    for (size_t i = 0; i < src.size(); ++i) tmp[i] = (T)src[i];
    CUDA_OK(cudaMalloc(dst, tmp.size() * sizeof(T)));
    CUDA_OK(cudaMemcpy(*dst, tmp.data(), tmp.size() * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

static void free_schedule(skb_schedule &s) {
    if (s.dI) cudaFree(s.dI);
    if (s.dF32) cudaFree(s.dF32);
    if (s.dF64) cudaFree(s.dF64);
    s.dI = s.dF32 = s.dF64 = nullptr;
}

const skb_schedule *get_schedule(skb_plan *p, int rows) {
    std::lock_guard<std::mutex> lk(p->mu);
    auto it = p->schedules.find(rows);
    if (it != p->schedules.end()) return &it->second;
    skb_schedule s;
    if (build_schedule(p, rows, &s) != 0) return nullptr;
    if (cudaMalloc(&s.dI, s.I.size() * sizeof(int32_t)) != cudaSuccess ||
        cudaMemcpy(s.dI, s.I.data(), s.I.size() * sizeof(int32_t), cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error("plan upload failed (no CUDA device?)");
        return nullptr;
    }
    if (upload_reals<float>(s.F, &s.dF32) != 0 || upload_reals<double>(s.F, &s.dF64) != 0) return nullptr;
    auto res = p->schedules.emplace(rows, std::move(s));
    return &res.first->second;
}

// pairwise program: every node evaluated, positions + twists kept per lane in shared memory
static int build_pair_program(skb_plan *p, skb_pair_program *pp) {
    const int n_nodes = (int)p->nodes.size(), nf = p->F, P = (int)p->pair_idx.size() / 2;
    std::vector<int> pos(n_nodes), save(n_nodes, -1), src(n_nodes, SRC_IDENTITY), saved_above(n_nodes, 0);
    for (int i = 0; i < n_nodes; ++i) pos[p->order[i]] = i;
    int n_slots = 0;
    for (int i = 0; i < n_nodes; ++i) {
        int k = p->order[i];
        const auto &nd = p->nodes[k];
        saved_above[k] = nd.parent < 0 ? 0 : saved_above[nd.parent] + (save[nd.parent] >= 0 ? 1 : 0);
        bool need_save = false;
        for (int c : nd.children) if (pos[c] != i + 1) need_save = true;
        if (need_save && k != 0) { save[k] = saved_above[k]; n_slots = std::max(n_slots, save[k] + 1); }
        if (nd.parent <= 0) src[k] = SRC_IDENTITY;
        else if (pos[nd.parent] == i - 1) src[k] = SRC_CURRENT;
        else src[k] = save[nd.parent];
    }
    if (n_slots > MAX_SLOTS) return set_error("plan: kinematic tree branches too deeply");
    // classes keyed by (node_i, node_j)
    std::map<std::pair<int, int>, int> cls_of;
    std::vector<std::vector<int>> cls_terms;   // flattened (node_pos, col, side)
    std::vector<int> pair_cls(P);
    for (int q = 0; q < P; ++q) {
        int fi = p->pair_idx[2 * q], fj = p->pair_idx[2 * q + 1];
        if (fi < 0 || fi >= nf || fj < 0 || fj >= nf) return set_error("pairs: feature index out of range");
        auto key = std::make_pair(p->feat_node[fi], p->feat_node[fj]);
        auto it = cls_of.find(key);
        if (it == cls_of.end()) {
            std::vector<int> ai, aj, terms;
            for (int a = key.first; a > 0; a = p->nodes[a].parent) ai.push_back(a);
            for (int a = key.second; a > 0; a = p->nodes[a].parent) aj.push_back(a);
            for (int a : ai) if (std::find(aj.begin(), aj.end(), a) == aj.end()) { terms.push_back(pos[a]); terms.push_back(p->nodes[a].col); terms.push_back(1); }
            for (int a : aj) if (std::find(ai.begin(), ai.end(), a) == ai.end()) { terms.push_back(pos[a]); terms.push_back(p->nodes[a].col); terms.push_back(-1); }
            it = cls_of.emplace(key, (int)cls_terms.size()).first;
            cls_terms.push_back(terms);
        }
        pair_cls[q] = it->second;
    }
    std::vector<int32_t> &I = pp->I;
    std::vector<double> &Fv = pp->F;
    I.assign(PHDR_WORDS, 0);
    I[P_N_NODES] = n_nodes; I[P_N_FEAT] = nf; I[P_D] = p->D; I[P_N_PAIRS] = P; I[P_N_CLASSES] = (int)cls_terms.size();
    I[P_N_SLOTS] = n_slots;
    // features grouped by node (preorder position): node record holds [begin, end) into this list
    std::vector<int> featlist, fbegin(n_nodes, 0), fend(n_nodes, 0);
    for (int i = 0; i < n_nodes; ++i) {
        fbegin[i] = (int)featlist.size();
        for (int f = 0; f < nf; ++f) if (p->feat_node[f] == p->order[i]) featlist.push_back(f);
        fend[i] = (int)featlist.size();
    }
    I[P_NODE_I] = (int)I.size();
    for (int i = 0; i < n_nodes; ++i) {
        int k = p->order[i];
        int32_t r[NODE_I] = {p->nodes[k].type, p->nodes[k].col, p->nodes[k].di, src[k], save[k], fbegin[i], fend[i], 0};
        I.insert(I.end(), r, r + NODE_I);
    }
    I[P_FEATNODE_I] = (int)I.size();
    I.insert(I.end(), featlist.begin(), featlist.end());
    I[P_PAIR_I] = (int)I.size();
    for (int q = 0; q < P; ++q) { I.push_back(p->pair_idx[2 * q]); I.push_back(p->pair_idx[2 * q + 1]); I.push_back(pair_cls[q]); I.push_back(0); }
    I[P_CLASS_I] = (int)I.size();
    size_t class_base = I.size();
    I.resize(I.size() + cls_terms.size() * 2);
    I[P_TERM_I] = (int)I.size();
    for (size_t c = 0; c < cls_terms.size(); ++c) {
        I[class_base + 2 * c] = (int)I.size();
        I[class_base + 2 * c + 1] = (int)cls_terms[c].size() / 3;
        I.insert(I.end(), cls_terms[c].begin(), cls_terms[c].end());
    }
    I[P_I_TOTAL] = (int)I.size();
    Fv.clear();
    I[P_NODE_F] = 0;
    Fv.resize((size_t)n_nodes * NODE_F, 0.0);
    for (int i = 0; i < n_nodes; ++i) {
        int k = p->order[i];
        double *r = &Fv[(size_t)i * NODE_F];
        quat_to_matrix(p->nodes[k].pre.q, r + NF_R);
        for (int j = 0; j < 3; ++j) r[NF_T + j] = p->nodes[k].pre.t[j];
        for (int j = 0; j < 4; ++j) r[NF_Q + j] = p->nodes[k].pre.q[j];
    }
    I[P_FEAT_F] = (int)Fv.size();
    for (int f = 0; f < nf; ++f) { for (int j = 0; j < 3; ++j) Fv.push_back(p->feat_off[f].t[j]); Fv.push_back(p->radii[f]); }
    I[P_PAIR_F] = (int)Fv.size();
    for (int q = 0; q < P; ++q) Fv.push_back(p->pair_thr[q]);
    I[P_F_TOTAL] = (int)Fv.size();
    pp->n_pairs = P; pp->n_nodes = n_nodes;
    return 0;
}

const skb_pair_program *get_pair_program(skb_plan *p) {
    std::lock_guard<std::mutex> lk(p->mu);
    if (!p->pairs_set) { set_error("pairwise: skb_plan_set_pairs was not called"); return nullptr; }
    if (p->pairs.dI) return &p->pairs;
    if (build_pair_program(p, &p->pairs) != 0) return nullptr;
    skb_pair_program &pp = p->pairs;
    if (cudaMalloc(&pp.dI, pp.I.size() * sizeof(int32_t)) != cudaSuccess ||
        cudaMemcpy(pp.dI, pp.I.data(), pp.I.size() * sizeof(int32_t), cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error("pair program upload failed (no CUDA device?)");
        return nullptr;
    }
    if (upload_reals<float>(pp.F, &pp.dF32) != 0 || upload_reals<double>(pp.F, &pp.dF64) != 0) return nullptr;
    return &pp;
}



// ------------------------------------------------------------------ level layout of the warp-cooperative kernels
// Node records {type, column, parent}, the per-level node lists and the row-form pre-transforms [R_r0 R_r1 R_r2 t_r] shared by
// the sphere / step / centre-of-mass / pose programs.  A node's level is its depth below the world node -- except for the
// chain of a PLANER / FLOATING base (skmp/robot/utils.py:45-56: three / six joints in a row, one node per level, i.e. three
// / six almost empty passes of the level walk): its LAST node becomes one node of type NODE_BASE_PLANER / NODE_BASE_FLOATING
// at level 1 that fk_phase evaluates in closed form from the base columns (frame = [Rz(yaw) Ry(pitch) Rx(roll) A | xyz], and
// the twists of all base columns), the others (NODE_BASE_SKIP) appear in no level.  Everything below the base sees the
// same frame as before.
struct LevelLayout {
    int n_levels = 0;
    std::vector<int> type, col, parent;          // per node, as the kernels see them
    std::vector<std::vector<int>> lv;            // node lists per level
    std::vector<int> skipped;                    // NODE_BASE_SKIP nodes (appended after the last level so the list has n_nodes entries)
    std::vector<double> rows;                    // 12 per node
};

static LevelLayout level_layout(const skb_plan *p) {
    LevelLayout L;
    const int n_nodes = (int)p->nodes.size(), D = p->D;
    L.type.resize(n_nodes); L.col.resize(n_nodes); L.parent.resize(n_nodes);
    for (int k = 0; k < n_nodes; ++k) { L.type[k] = p->nodes[k].type; L.col[k] = p->nodes[k].col; L.parent[k] = p->nodes[k].parent; }
    L.rows.resize((size_t)n_nodes * 12);
    for (int k = 0; k < n_nodes; ++k) {
        double R[9];
        quat_to_matrix(p->nodes[k].pre.q, R);
        for (int r = 0; r < 3; ++r) { double *o = &L.rows[(size_t)k * 12 + 4 * r]; o[0] = R[3 * r]; o[1] = R[3 * r + 1]; o[2] = R[3 * r + 2]; o[3] = p->nodes[k].pre.t[r]; }
    }
    const int nb = p->base_type == SKB_BASE_PLANER ? 3 : (p->base_type == SKB_BASE_FLOATING ? 6 : 0);
    const char *keep = getenv("SKB_BASE_CHAIN");                          // SKB_BASE_CHAIN=1: walk the base joint by joint
    if (nb > 0 && !(keep && atoi(keep) != 0)) {
        const int c0 = D - nb, last_col = nb == 3 ? c0 + 2 : c0 + 3;      // theta / roll close the chain (compile_tree)
        std::vector<int> chain;
        int last = -1;
        for (int k = 1; k < n_nodes; ++k) if (p->nodes[k].col >= c0) chain.push_back(k);
        for (int k : chain) if (p->nodes[k].col == last_col) last = k;
        // the chain must be exactly world -> ... -> last, one child each
        bool ok = (int)chain.size() == nb && last >= 0;
        for (int k : chain) if (k != last && p->nodes[k].children.size() != 1) ok = false;
        if (ok) {
            for (int k : chain) if (k != last) { L.type[k] = NODE_BASE_SKIP; L.skipped.push_back(k); }
            L.type[last] = nb == 3 ? NODE_BASE_PLANER : NODE_BASE_FLOATING;
            L.col[last] = c0;
            L.parent[last] = 0;
            // frame = [R_base A | t_base]: the node's constant is the z-normalising rotation A of its own axis alone
            const double ax[3] = {nb == 3 ? 0.0 : 1.0, 0.0, nb == 3 ? 1.0 : 0.0};
            double R[9];
            quat_to_matrix(xf_z_to_axis(ax).q, R);
            for (int r = 0; r < 3; ++r) { double *o = &L.rows[(size_t)last * 12 + 4 * r]; o[0] = R[3 * r]; o[1] = R[3 * r + 1]; o[2] = R[3 * r + 2]; o[3] = 0.0; }
        }
    }
    std::vector<int> level(n_nodes, 0);
    for (int k = 0; k < n_nodes; ++k) {
        if (L.type[k] == NODE_BASE_SKIP) { level[k] = -1; continue; }
        level[k] = L.parent[k] < 0 ? 0 : level[L.parent[k]] + 1;
        L.n_levels = std::max(L.n_levels, level[k] + 1);
    }
    L.lv.assign(L.n_levels, {});
    for (int k = 0; k < n_nodes; ++k) if (level[k] >= 0) L.lv[level[k]].push_back(k);
    return L;
}

// appends the node / level / level-node tables to I in the layout every level program uses
static void pack_levels(const LevelLayout &L, std::vector<int32_t> &I, int &node_i, int &level_i, int &levelnode_i) {
    node_i = (int)I.size();
    for (size_t k = 0; k < L.type.size(); ++k) { I.push_back(L.type[k]); I.push_back(L.col[k]); I.push_back(L.parent[k]); I.push_back(0); }
    level_i = (int)I.size();
    int cursor = 0;
    for (auto &l : L.lv) { I.push_back(cursor); I.push_back((int)l.size()); cursor += (int)l.size(); }
    levelnode_i = (int)I.size();
    for (auto &l : L.lv) for (int k : l) I.push_back(k);
    for (int k : L.skipped) I.push_back(k);
}

// ------------------------------------------------------------------ sphere program (skb_sphere.cu)
static int odd_quads(int words) {   // round up to a multiple of 4 words whose quotient is odd (bank-conflict-free rows)
    int q = (words + 3) / 4;
    if (q % 2 == 0) q += 1;
    return q * 4;
}

static int build_sphere_program(skb_plan *p, skb_sphere_program *sp) {
    sp->built = true;
    sp->supported = false;
    const int n_nodes = (int)p->nodes.size(), nf = p->F, D = p->D;
    if (p->rot_type != SKB_ROT_IGNORE || D > 64 || D < 1 || n_nodes > 255 || nf > 32 * 32) return 0;
    // levels: node 0 (world) is level 0, a movable node with di controlled ancestors is level di + 1
    const LevelLayout LL = level_layout(p);
    const int n_levels = LL.n_levels;
    const std::vector<std::vector<int>> &lv = LL.lv;
    (void)lv;
    int max_level_nodes = 0;
    for (auto &l : lv) max_level_nodes = std::max(max_level_nodes, (int)l.size());
    // ancestor column masks per node
    std::vector<uint64_t> nmask(n_nodes, 0);
    for (int k = 1; k < n_nodes; ++k) {
        nmask[k] = nmask[p->nodes[k].parent];
        if (p->nodes[k].col >= 0) nmask[k] |= (uint64_t)1 << p->nodes[k].col;
    }
    const int n_rounds = (nf + 31) / 32;
    std::vector<int32_t> &I = sp->I;
    std::vector<double> &Fv = sp->F;
    I.assign(SHDR_WORDS, 0);
    I[S_N_NODES] = n_nodes; I[S_N_LEVELS] = n_levels; I[S_N_ROUNDS] = n_rounds; I[S_D] = D; I[S_N_FEAT] = nf;
    I[S_MAX_LEVEL_NODES] = max_level_nodes;
    sp->frame_stride = odd_quads(n_nodes * 12);
    sp->tw_stride = odd_quads(D * 8);
    I[S_FRAME_STRIDE] = sp->frame_stride; I[S_TW_STRIDE] = sp->tw_stride;
    { int a, b, c; pack_levels(LL, I, a, b, c); I[S_NODE_I] = a; I[S_LEVEL_I] = b; I[S_LEVELNODE_I] = c; }
    while (I.size() % 4) I.push_back(0);
    I[S_ROUND_I] = (int)I.size();
    for (int r = 0; r < n_rounds; ++r) {
        const int s0 = r * 32, cnt = std::min(32, nf - s0);
        uint64_t cm = 0;
        for (int f = s0; f < s0 + cnt; ++f) cm |= nmask[p->feat_node[f]];
        I.push_back(s0); I.push_back(cnt); I.push_back((int32_t)(uint32_t)(cm & 0xffffffffu)); I.push_back((int32_t)(uint32_t)(cm >> 32));
    }
    I[S_TABB_I] = (int)I.size();
    for (int r = 0; r < n_rounds; ++r)
        for (int l = 0; l < 32; ++l) {
            const int f = r * 32 + l;
            if (f < nf) {
                const uint64_t am = nmask[p->feat_node[f]];
                I.push_back(p->feat_node[f]); I.push_back((int32_t)(uint32_t)(am & 0xffffffffu));
                I.push_back((int32_t)(uint32_t)(am >> 32)); I.push_back(1);
            } else { I.push_back(0); I.push_back(0); I.push_back(0); I.push_back(0); }
        }
    I[S_I_TOTAL] = (int)I.size();
    Fv.clear();
    I[S_NODE_F] = 0;
    Fv.insert(Fv.end(), LL.rows.begin(), LL.rows.end());
    I[S_TABA_F] = (int)Fv.size();
    for (int r = 0; r < n_rounds; ++r)
        for (int l = 0; l < 32; ++l) {
            const int f = r * 32 + l;
            if (f < nf) { for (int j = 0; j < 3; ++j) Fv.push_back(p->feat_off[f].t[j]); Fv.push_back(p->radii[f]); }
            else { for (int j = 0; j < 4; ++j) Fv.push_back(0.0); }
        }
    I[S_F_TOTAL] = (int)Fv.size();
    sp->n_nodes = n_nodes; sp->n_levels = n_levels; sp->n_rounds = n_rounds;
    sp->supported = true;
    return 0;
}

static void free_sphere(skb_sphere_program &sp) {
    if (sp.dI) cudaFree(sp.dI);
    if (sp.dF32) cudaFree(sp.dF32);
    if (sp.dF64) cudaFree(sp.dF64);
    sp = skb_sphere_program();
}

const skb_sphere_program *get_sphere_program(skb_plan *p) {
    std::lock_guard<std::mutex> lk(p->mu);
    skb_sphere_program &sp = p->sphere;
    if (!sp.built) {
        const char *force = getenv("SKB_KERNEL");
        if (force && std::string(force) == "tree") { sp.built = true; sp.supported = false; }
        else if (build_sphere_program(p, &sp) != 0) return nullptr;
    }
    if (!sp.supported) return nullptr;
    if (!sp.dI) {
        if (cudaMalloc(&sp.dI, sp.I.size() * sizeof(int32_t)) != cudaSuccess ||
            cudaMemcpy(sp.dI, sp.I.data(), sp.I.size() * sizeof(int32_t), cudaMemcpyHostToDevice) != cudaSuccess) {
            set_error("sphere program upload failed (no CUDA device?)");
            sp.dI = nullptr;
            return nullptr;
        }
        if (upload_reals<float>(sp.F, &sp.dF32) != 0 || upload_reals<double>(sp.F, &sp.dF64) != 0) return nullptr;
    }
    return &sp;
}


// ------------------------------------------------------------------ step program (skb_sphere2.cu)
// Rows (spheres or pairs) are cut into half-rounds of 32; adjacent half-rounds are fused into one two-rows-per-lane
// step when that is cheaper under a simple issue-slot model (a fused step loads each twist pair once for both rows
// but computes the union of the two halves' column sets).
static int build_s2_program(skb_plan *p, int kind, skb_s2_program *sp, bool allow_fuse = true) {
    sp->built = true;
    sp->supported = false;
    const int n_nodes = (int)p->nodes.size(), nf = p->F, D = p->D;
    const bool pairs = kind == KIND_PAIRWISE;
    const int rows = pairs ? (int)p->pair_thr.size() : nf;
    if (p->rot_type != SKB_ROT_IGNORE || D < 1 || D > 64 || n_nodes > 255 || rows < 1) return 0;
    if (pairs && nf > 4096) return 0;
    const char *force = getenv("SKB_KERNEL");
    if (force && (std::string(force) == "tree" || std::string(force) == "sphere1")) return 0;
    const LevelLayout LL = level_layout(p);
    const int n_levels = LL.n_levels;
    const std::vector<std::vector<int>> &lv = LL.lv;
    (void)lv;
    std::vector<uint64_t> nmask(n_nodes, 0);
    for (int k = 1; k < n_nodes; ++k) {
        nmask[k] = nmask[p->nodes[k].parent];
        if (p->nodes[k].col >= 0) nmask[k] |= (uint64_t)1 << p->nodes[k].col;
    }
    // per-row column masks (pairs: plus = columns moving sphere i only, minus = sphere j only)
    std::vector<uint64_t> plus(rows), minus(rows, 0);
    for (int r = 0; r < rows; ++r) {
        if (pairs) {
            const int fi = p->pair_idx[2 * r], fj = p->pair_idx[2 * r + 1];
            if (fi < 0 || fi >= nf || fj < 0 || fj >= nf) return set_error("pairs: feature index out of range");
            const uint64_t mi = nmask[p->feat_node[fi]], mj = nmask[p->feat_node[fj]];
            plus[r] = mi & ~mj; minus[r] = mj & ~mi;
        } else plus[r] = nmask[p->feat_node[r]];
    }
    const int n_hr = (rows + 31) / 32;
    std::vector<uint64_t> hcm(n_hr, 0);
    for (int r = 0; r < rows; ++r) hcm[r / 32] |= plus[r] | minus[r];
    auto npairs = [&](uint64_t m) { int c = 0; for (int j = 0; j < D; j += 2) if ((m >> j) & 3) ++c; return c; };
    // dynamic programme over "single" / "fused with the next half-round"
    std::vector<double> best(n_hr + 2, 0.0);
    std::vector<char> fuse(n_hr + 1, 0);
    double c_s0 = 30.0, c_s1 = 17.0, c_f0 = 40.0, c_f1 = 29.0;      // issue-slot model of one half-round / one fused step
    if (const char *e = getenv("SKB_S2_DP")) sscanf(e, "%lf,%lf,%lf,%lf", &c_s0, &c_s1, &c_f0, &c_f1);   // tuning knob
    for (int i = n_hr - 1; i >= 0; --i) {
        const double single = c_s0 + c_s1 * npairs(hcm[i]) + best[i + 1];
        best[i] = single; fuse[i] = 0;
        if (allow_fuse && i + 1 < n_hr) {
            const double both = c_f0 + c_f1 * npairs(hcm[i] | hcm[i + 1]) + best[i + 2];
            if (both < single) { best[i] = both; fuse[i] = 1; }
        }
    }
    sp->phase2_slots = best[0];
    sp->steps.clear();
    for (int i = 0; i < n_hr;) {
        skb_s2_step st;
        st.s0 = i * 32;
        st.cntA = (int16_t)std::min(32, rows - i * 32);
        if (fuse[i]) { st.cntB = (int16_t)std::min(32, rows - (i + 1) * 32); st.cm = hcm[i] | hcm[i + 1]; i += 2; }
        else { st.cntB = 0; st.cm = hcm[i]; i += 1; }
        sp->steps.push_back(st);
    }
    const int n_steps = (int)sp->steps.size();
    if (n_steps > S2_MAX_STEPS) return 0;

    std::vector<int32_t> &I = sp->I;
    std::vector<double> &Fv = sp->F;
    I.assign(QHDR_WORDS, 0);
    I[Q_N_NODES] = n_nodes; I[Q_N_LEVELS] = n_levels; I[Q_N_STEPS] = n_steps; I[Q_D] = D; I[Q_ROWS] = rows;
    I[Q_N_CEN] = pairs ? nf : 0;
    { int a, b, c; pack_levels(LL, I, a, b, c); I[Q_NODE_I] = a; I[Q_LEVEL_I] = b; I[Q_LEVELNODE_I] = c; }
    while (I.size() % 4) I.push_back(0);
    I[Q_TABB_I] = (int)I.size();
    for (int s = 0; s < n_steps; ++s)
        for (int h = 0; h < (sp->steps[s].cntB > 0 ? 2 : 1); ++h)   // compact: an empty half B has no table slot
            for (int l = 0; l < 32; ++l) {
                const int r = sp->steps[s].s0 + 32 * h + l;
                const int cnt = h == 0 ? sp->steps[s].cntA : sp->steps[s].cntB;
                if (l >= cnt) { for (int j = 0; j < 4; ++j) I.push_back(0); continue; }
                if (pairs) {
                    float thr = (float)p->pair_thr[r];
                    int32_t bits;
                    memcpy(&bits, &thr, 4);
                    I.push_back(p->pair_idx[2 * r] | (p->pair_idx[2 * r + 1] << 16));
                    I.push_back((int32_t)(uint32_t)plus[r]); I.push_back((int32_t)(uint32_t)minus[r]); I.push_back(bits);
                } else {
                    I.push_back(p->feat_node[r]); I.push_back((int32_t)(uint32_t)(plus[r] & 0xffffffffu));
                    I.push_back((int32_t)(uint32_t)(plus[r] >> 32)); I.push_back(0);
                }
            }
    I[Q_CENNODE_I] = (int)I.size();
    if (pairs) for (int f = 0; f < nf; ++f) I.push_back(p->feat_node[f]);
    while (I.size() % 4) I.push_back(0);
    I[Q_TABC_I] = (int)I.size();
    if (pairs && D > 32)
        for (int s = 0; s < n_steps; ++s)
            for (int h = 0; h < (sp->steps[s].cntB > 0 ? 2 : 1); ++h)
                for (int l = 0; l < 32; ++l) {
                    const int r = sp->steps[s].s0 + 32 * h + l;
                    const int cnt = h == 0 ? sp->steps[s].cntA : sp->steps[s].cntB;
                    if (l < cnt) { I.push_back((int32_t)(uint32_t)(plus[r] >> 32)); I.push_back((int32_t)(uint32_t)(minus[r] >> 32)); }
                    else { I.push_back(0); I.push_back(0); }
                    I.push_back(0); I.push_back(0);
                }
    I[Q_I_TOTAL] = (int)I.size();
    Fv.clear();
    I[Q_NODE_F] = 0;
    Fv.insert(Fv.end(), LL.rows.begin(), LL.rows.end());
    I[Q_TABA_F] = (int)Fv.size();
    if (!pairs)
        for (int s = 0; s < n_steps; ++s)
            for (int h = 0; h < (sp->steps[s].cntB > 0 ? 2 : 1); ++h)
                for (int l = 0; l < 32; ++l) {
                    const int r = sp->steps[s].s0 + 32 * h + l;
                    const int cnt = h == 0 ? sp->steps[s].cntA : sp->steps[s].cntB;
                    if (l < cnt) { for (int j = 0; j < 3; ++j) Fv.push_back(p->feat_off[r].t[j]); Fv.push_back(p->radii[r]); }
                    else for (int j = 0; j < 4; ++j) Fv.push_back(0.0);
                }
    I[Q_CENA_F] = (int)Fv.size();
    if (pairs) for (int f = 0; f < nf; ++f) { for (int j = 0; j < 3; ++j) Fv.push_back(p->feat_off[f].t[j]); Fv.push_back(p->radii[f]); }
    I[Q_THR_F] = (int)Fv.size();
    if (pairs)
        for (int s = 0; s < n_steps; ++s)
            for (int h = 0; h < (sp->steps[s].cntB > 0 ? 2 : 1); ++h)
                for (int l = 0; l < 32; ++l) {
                    const int r = sp->steps[s].s0 + 32 * h + l;
                    const int cnt = h == 0 ? sp->steps[s].cntA : sp->steps[s].cntB;
                    Fv.push_back(l < cnt ? p->pair_thr[r] : 0.0);
                }
    I[Q_F_TOTAL] = (int)Fv.size();
    sp->I_lean.clear();
    sp->masks.clear();
    if (pairs) {
        // lean program of the reducing modes: header, nodes, levels | node per centre | table P
        std::vector<int32_t> &L = sp->I_lean;
        L.assign(I.begin(), I.begin() + I[Q_TABB_I]);
        L[Q_TABB_I] = L[Q_TABC_I] = 0;
        L[Q_CENNODE_I] = (int)L.size();
        for (int f = 0; f < nf; ++f) L.push_back(p->feat_node[f]);
        while (L.size() % 4) L.push_back(0);
        const int n_hr4 = (n_hr + 3) / 4 * 4;
        L[Q_N_HR] = n_hr4;
        L[Q_TABP_I] = (int)L.size();
        const float ninf = -std::numeric_limits<float>::max();           // padding rows evaluate to FLT_MAX: finite, never the minimum
        for (int r = 0; r < n_hr4 * 32; ++r) {
            const float thr = r < rows ? (float)p->pair_thr[r] : ninf;
            int32_t bits;
            memcpy(&bits, &thr, 4);
            L.push_back(r < rows ? (p->pair_idx[2 * r] | (p->pair_idx[2 * r + 1] << 16)) : 0);
            L.push_back(bits);
        }
        L[Q_I_TOTAL] = (int)L.size();
        for (int r = 0; r < rows; ++r) { sp->masks.push_back(plus[r]); sp->masks.push_back(minus[r]); }
    }
    sp->n_nodes = n_nodes; sp->n_levels = n_levels; sp->rows = rows; sp->n_cen = pairs ? nf : 0;
    sp->frame_stride = odd_quads(n_nodes * 12);
    {   // structure hash (FNV-1a): tree, row tables, steps -- not the folded constants in F, which change with every
        // reflect_skrobot_model while the best launch configuration does not
        unsigned long long h = 1469598103934665603ull;
        auto mix = [&](const void *data, size_t bytes) {
            const unsigned char *b = reinterpret_cast<const unsigned char *>(data);
            for (size_t i = 0; i < bytes; ++i) { h ^= b[i]; h *= 1099511628211ull; }
        };
        mix(I.data(), I.size() * sizeof(int32_t));
        for (auto &st : sp->steps) { mix(&st.cm, sizeof(st.cm)); mix(&st.s0, sizeof(st.s0)); mix(&st.cntA, sizeof(st.cntA)); mix(&st.cntB, sizeof(st.cntB)); }
        const int meta[4] = {rows, D, kind, allow_fuse ? 1 : 0};
        mix(meta, sizeof(meta));
        sp->hash = h;
    }
    sp->supported = true;
    return 0;
}

static void free_s2(skb_s2_program &sp) {
    if (sp.dI) cudaFree(sp.dI);
    if (sp.dF32) cudaFree(sp.dF32);
    if (sp.dF64) cudaFree(sp.dF64);
    if (sp.dI_lean) cudaFree(sp.dI_lean);
    if (sp.dMasks) cudaFree(sp.dMasks);
    if (sp.unfused) { free_s2(*sp.unfused); delete sp.unfused; }
    sp = skb_s2_program();
}

static int upload_s2(skb_s2_program &sp) {
    if (cudaMalloc(&sp.dI, sp.I.size() * sizeof(int32_t)) != cudaSuccess ||
        cudaMemcpy(sp.dI, sp.I.data(), sp.I.size() * sizeof(int32_t), cudaMemcpyHostToDevice) != cudaSuccess) {
        sp.dI = nullptr;
        return set_error("step program upload failed (no CUDA device?)");
    }
    if (upload_reals<float>(sp.F, &sp.dF32) != 0 || upload_reals<double>(sp.F, &sp.dF64) != 0) return -1;
    if (!sp.I_lean.empty()) {
        if (cudaMalloc(&sp.dI_lean, sp.I_lean.size() * sizeof(int32_t)) != cudaSuccess ||
            cudaMemcpy(sp.dI_lean, sp.I_lean.data(), sp.I_lean.size() * sizeof(int32_t), cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMalloc(&sp.dMasks, sp.masks.size() * sizeof(uint64_t)) != cudaSuccess ||
            cudaMemcpy(sp.dMasks, sp.masks.data(), sp.masks.size() * sizeof(uint64_t), cudaMemcpyHostToDevice) != cudaSuccess)
            return set_error("step program upload failed (lean pair tables)");
    }
    return 0;
}

const skb_s2_program *get_s2_program(skb_plan *p, int kind) {
    std::lock_guard<std::mutex> lk(p->mu);
    if (kind == KIND_PAIRWISE && !p->pairs_set) return nullptr;
    skb_s2_program &sp = kind == KIND_PAIRWISE ? p->s2_pairs : p->s2_collfree;
    if (!sp.built && build_s2_program(p, kind, &sp) != 0) return nullptr;
    if (!sp.supported) return nullptr;
    if (!sp.dI) {
        if (upload_s2(sp) != 0) return nullptr;
        bool fused = false;
        for (auto &st : sp.steps) fused = fused || st.cntB > 0;
        if (fused && getenv("SKB_S2_NO_UNFUSED") == nullptr) {
            sp.unfused = new skb_s2_program();
            if (build_s2_program(p, kind, sp.unfused, false) != 0 || !sp.unfused->supported || upload_s2(*sp.unfused) != 0) {
                free_s2(*sp.unfused);
                delete sp.unfused;
                sp.unfused = nullptr;
            }
        }
    }
    return &sp;
}


// ------------------------------------------------------------------ centre-of-mass program (skb_com.cu)
static int build_com_program(skb_plan *p, skb_com_program *cp) {
    const int n_nodes = (int)p->nodes.size(), nf = p->F, D = p->D;
    if (p->rot_type != SKB_ROT_IGNORE) return set_error("com: the plan must be built with rot_type IGNORE");
    if ((int)p->weights.size() != nf) return set_error("com: skb_plan_set_weights was not called");
    double W = 0;
    for (double w : p->weights) { if (!(w >= 0)) return set_error("com: weights must be non-negative"); W += w; }
    if (!(W > 0)) return set_error("com: the total weight must be positive");
    if (n_nodes > 255) return set_error("com: too many nodes");
    // DFS preorder position of every node and the end of its subtree
    std::vector<int> pos(n_nodes), end(n_nodes);
    for (int i = 0; i < n_nodes; ++i) pos[p->order[i]] = i;
    for (int i = n_nodes - 1; i >= 0; --i) {
        const int k = p->order[i];
        end[k] = i + 1;
        for (int c : p->nodes[k].children) end[k] = std::max(end[k], end[c]);
    }
    std::vector<int> pts(nf);
    for (int f = 0; f < nf; ++f) pts[f] = f;
    std::stable_sort(pts.begin(), pts.end(), [&](int a, int b) { return pos[p->feat_node[a]] < pos[p->feat_node[b]]; });
    std::vector<int> ppos(nf);
    for (int i = 0; i < nf; ++i) ppos[i] = pos[p->feat_node[pts[i]]];
    const LevelLayout LL = level_layout(p);
    const int n_levels = LL.n_levels;
    const std::vector<std::vector<int>> &lv = LL.lv;
    (void)lv;
    std::vector<int32_t> &I = cp->I;
    std::vector<double> &Fv = cp->F;
    I.assign(CHDR_WORDS, 0);
    I[C_N_NODES] = n_nodes; I[C_N_LEVELS] = n_levels; I[C_N_POINTS] = nf; I[C_D] = D;
    { int a, b, c; pack_levels(LL, I, a, b, c); I[C_NODE_I] = a; I[C_LEVEL_I] = b; I[C_LEVELNODE_I] = c; }
    I[C_PTNODE_I] = (int)I.size();
    for (int i = 0; i < nf; ++i) I.push_back(p->feat_node[pts[i]]);
    while (I.size() % 2) I.push_back(0);
    I[C_COLRANGE_I] = (int)I.size();
    std::vector<int> col_node(D, -1);
    for (int k = 0; k < n_nodes; ++k) if (p->nodes[k].col >= 0) col_node[p->nodes[k].col] = k;
    for (int d = 0; d < D; ++d) {
        int a = 0, b = 0;
        if (col_node[d] >= 0) {
            const int k = col_node[d];
            a = (int)(std::lower_bound(ppos.begin(), ppos.end(), pos[k]) - ppos.begin());
            b = (int)(std::lower_bound(ppos.begin(), ppos.end(), end[k]) - ppos.begin());
        }
        I.push_back(a); I.push_back(b);
    }
    I[C_I_TOTAL] = (int)I.size();
    Fv.clear();
    I[C_NODE_F] = 0;
    Fv.insert(Fv.end(), LL.rows.begin(), LL.rows.end());
    I[C_PT_F] = (int)Fv.size();
    for (int i = 0; i < nf; ++i) { for (int j = 0; j < 3; ++j) Fv.push_back(p->feat_off[pts[i]].t[j]); Fv.push_back(p->weights[pts[i]]); }
    I[C_F_TOTAL] = (int)Fv.size();
    cp->n_nodes = n_nodes; cp->n_levels = n_levels; cp->n_points = nf;
    cp->built = true;
    return 0;
}

static void free_com(skb_com_program &cp) {
    if (cp.dI) cudaFree(cp.dI);
    if (cp.dF32) cudaFree(cp.dF32);
    if (cp.dF64) cudaFree(cp.dF64);
    cp = skb_com_program();
}

const skb_com_program *get_com_program(skb_plan *p) {
    std::lock_guard<std::mutex> lk(p->mu);
    skb_com_program &cp = p->com;
    if (!cp.built && build_com_program(p, &cp) != 0) return nullptr;
    if (!cp.dI) {
        if (cudaMalloc(&cp.dI, cp.I.size() * sizeof(int32_t)) != cudaSuccess ||
            cudaMemcpy(cp.dI, cp.I.data(), cp.I.size() * sizeof(int32_t), cudaMemcpyHostToDevice) != cudaSuccess) {
            set_error("com program upload failed (no CUDA device?)");
            cp.dI = nullptr;
            return nullptr;
        }
        if (upload_reals<float>(cp.F, &cp.dF32) != 0 || upload_reals<double>(cp.F, &cp.dF64) != 0) return nullptr;
    }
    return &cp;
}


// ------------------------------------------------------------------ pose program (skb_pose.cu)
static int build_pose_program(skb_plan *p, skb_pose_program *pp) {
    pp->built = true;
    pp->supported = false;
    const int n_nodes = (int)p->nodes.size(), nf = p->F, D = p->D;
    if (p->rot_type == SKB_ROT_IGNORE || D < 1 || D > 64 || n_nodes > 255 || nf > 256) return 0;
    const char *force = getenv("SKB_KERNEL");
    if (force && std::string(force) == "tree") return 0;
    const LevelLayout LL = level_layout(p);
    const int n_levels = LL.n_levels;
    const std::vector<std::vector<int>> &lv = LL.lv;
    (void)lv;
    std::vector<uint64_t> nmask(n_nodes, 0);
    for (int k = 1; k < n_nodes; ++k) {
        nmask[k] = nmask[p->nodes[k].parent];
        if (p->nodes[k].col >= 0) nmask[k] |= (uint64_t)1 << p->nodes[k].col;
    }
    std::vector<int32_t> &I = pp->I;
    std::vector<double> &Fv = pp->F;
    I.assign(EHDR_WORDS, 0);
    I[E_N_NODES] = n_nodes; I[E_N_LEVELS] = n_levels; I[E_N_FEAT] = nf; I[E_D] = D;
    { int a, b, c; pack_levels(LL, I, a, b, c); I[E_NODE_I] = a; I[E_LEVEL_I] = b; I[E_LEVELNODE_I] = c; }
    while (I.size() % 4) I.push_back(0);
    I[E_FEAT_I] = (int)I.size();
    for (int f = 0; f < nf; ++f) {
        const uint64_t am = nmask[p->feat_node[f]];
        I.push_back(p->feat_node[f]); I.push_back((int32_t)(uint32_t)(am & 0xffffffffu)); I.push_back((int32_t)(uint32_t)(am >> 32)); I.push_back(0);
    }
    // XYZW: the quaternion of a feature is the PRODUCT of the factors  Q_pre_k * Qz(q_k / 2)  along its chain, root first, like
    // the reference's chain of link transforms (its sign is the chain's, not that of a matrix conversion).  Constant factors
    // (fixed offsets folded by the plan compiler, prismatic joints) are merged into the next factor here; the chain is padded
    // with identities to a multiple of 4 so that four lanes multiply a quarter each.
    std::vector<std::vector<std::pair<int, Xf>>> chains(nf);
    size_t chain_len = 4;
    for (int f = 0; f < nf; ++f) {
        std::vector<int> path;
        for (int k = p->feat_node[f]; k > 0; k = p->nodes[k].parent) path.push_back(k);
        std::reverse(path.begin(), path.end());
        Xf acc = xf_identity();
        for (int k : path) {
            Xf pre = xf_identity();
            for (int j = 0; j < 4; ++j) pre.q[j] = p->nodes[k].pre.q[j];
            acc = xf_mul(acc, pre);                              // rotations only: the translations stay zero
            if (p->nodes[k].type == NODE_REVOLUTE) { chains[f].push_back({p->nodes[k].col, acc}); acc = xf_identity(); }
        }
        chains[f].push_back({-1, acc});                           // trailing constant (identity when the chain ends in a revolute joint)
        chain_len = std::max(chain_len, (chains[f].size() + 3) / 4 * 4);
    }
    I[E_CHAIN_LEN] = (int)chain_len;
    I[E_CHAIN_I] = (int)I.size();
    for (int f = 0; f < nf; ++f)
        for (size_t k = 0; k < chain_len; ++k) I.push_back(k < chains[f].size() ? chains[f][k].first : -1);
    I[E_I_TOTAL] = (int)I.size();
    Fv.clear();
    I[E_NODE_F] = 0;
    Fv.insert(Fv.end(), LL.rows.begin(), LL.rows.end());
    I[E_NODEQ_F] = (int)Fv.size();
    for (int k = 0; k < n_nodes; ++k) for (int j = 0; j < 4; ++j) Fv.push_back(p->nodes[k].pre.q[j]);
    I[E_CHAINQ_F] = (int)Fv.size();
    for (int f = 0; f < nf; ++f)
        for (size_t k = 0; k < chain_len; ++k)
            for (int j = 0; j < 4; ++j) Fv.push_back(k < chains[f].size() ? chains[f][k].second.q[j] : (j == 3 ? 1.0 : 0.0));
    I[E_FEAT_F] = (int)Fv.size();

    for (int f = 0; f < nf; ++f) {
        double R[9];
        quat_to_matrix(p->feat_off[f].q, R);
        for (int j = 0; j < 3; ++j) Fv.push_back(p->feat_off[f].t[j]);
        Fv.push_back(0.0);
        for (int j = 0; j < 9; ++j) Fv.push_back(R[j]);
        for (int j = 0; j < 3; ++j) Fv.push_back(0.0);
        for (int j = 0; j < 4; ++j) Fv.push_back(p->feat_off[f].q[j]);
    }
    I[E_F_TOTAL] = (int)Fv.size();
    pp->n_nodes = n_nodes; pp->n_levels = n_levels;
    pp->supported = true;
    return 0;
}

static void free_pose(skb_pose_program &pp) {
    if (pp.dI) cudaFree(pp.dI);
    if (pp.dF32) cudaFree(pp.dF32);
    if (pp.dF64) cudaFree(pp.dF64);
    pp = skb_pose_program();
}

const skb_pose_program *get_pose_program(skb_plan *p) {
    std::lock_guard<std::mutex> lk(p->mu);
    skb_pose_program &pp = p->pose;
    if (!pp.built && build_pose_program(p, &pp) != 0) return nullptr;
    if (!pp.supported) return nullptr;
    if (!pp.dI) {
        if (cudaMalloc(&pp.dI, pp.I.size() * sizeof(int32_t)) != cudaSuccess ||
            cudaMemcpy(pp.dI, pp.I.data(), pp.I.size() * sizeof(int32_t), cudaMemcpyHostToDevice) != cudaSuccess) {
            set_error("pose program upload failed (no CUDA device?)");
            pp.dI = nullptr;
            return nullptr;
        }
        if (upload_reals<float>(pp.F, &pp.dF32) != 0 || upload_reals<double>(pp.F, &pp.dF64) != 0) return nullptr;
    }
    return &pp;
}

template <typename T>
static int upload_prims(skb_sdf *s, void **dst) {
    std::vector<DevPrim<T>> v(s->prims.size() ? s->prims.size() : 1);
    for (size_t i = 0; i < s->prims.size(); ++i) {
        const skb_host_prim &h = s->prims[i];
        DevPrim<T> &d = v[i];
        memset(&d, 0, sizeof(d));
        d.type = h.type;
        bool ident = true;
        for (int k = 0; k < 9; ++k) {
            d.rot[k] = (T)h.rot[k];
            if (fabs(h.rot[k] - ((k % 4 == 0) ? 1.0 : 0.0)) > 0) ident = false;
        }
        d.flags = (ident ? PF_IDENTITY : 0) | (h.grid_is_f64 ? PF_GRID_F64 : 0) | (h.dev_cells ? PF_GRID_CELLS : 0);
        d.cells = h.dev_cells;
        d.cells_tex = h.cells_tex;
        for (int k = 0; k < 3; ++k) { d.pos[k] = (T)h.pos[k]; d.dims[k] = h.dims[k]; }
        for (int k = 0; k < 8; ++k) d.par[k] = (T)h.par[k];
        if (h.type == PRIM_GRID) for (int k = 0; k < 3; ++k) d.par[3 + k] = (T)(1.0 / h.par[3 + k]);
        d.grid = h.dev_grid;
    }
    CUDA_OK(cudaMalloc(dst, v.size() * sizeof(DevPrim<T>)));
    CUDA_OK(cudaMemcpy(*dst, v.data(), v.size() * sizeof(DevPrim<T>), cudaMemcpyHostToDevice));
    std::vector<unsigned char> &hc = sizeof(T) == 4 ? s->host_f32 : s->host_f64;
    hc.assign(reinterpret_cast<unsigned char *>(v.data()), reinterpret_cast<unsigned char *>(v.data()) + v.size() * sizeof(DevPrim<T>));
    return 0;
}

const void *get_host_prims(const skb_sdf *s, int dtype) { return dtype == SKB_F32 ? (const void *)s->host_f32.data() : (const void *)s->host_f64.data(); }

const void *get_dev_prims(skb_sdf *s, int dtype) {
    std::lock_guard<std::mutex> lk(s->mu);
    if (s->dev_count != (int)s->prims.size()) {
        if (s->dev_f32) cudaFree(s->dev_f32);
        if (s->dev_f64) cudaFree(s->dev_f64);
        s->dev_f32 = s->dev_f64 = nullptr;
        s->dev_count = (int)s->prims.size();
    }
    if (dtype == SKB_F32) {
        if (!s->dev_f32 && upload_prims<float>(s, &s->dev_f32) != 0) return nullptr;
        return s->dev_f32;
    }
    if (!s->dev_f64 && upload_prims<double>(s, &s->dev_f64) != 0) return nullptr;
    return s->dev_f64;
}

}  // namespace skb

using namespace skb;

// ------------------------------------------------------------------ C ABI: misc
extern "C" {

int skb_version(void) { return 100; }
const char *skb_last_error(void) { return g_err.c_str(); }
int64_t skb_launch_count(void) { return g_launch_count.load(); }

int skb_device_count(int32_t *out_count) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        if (out_count) *out_count = 0;
        return set_error(std::string("no CUDA device: ") + cudaGetErrorString(e));
    }
    if (out_count) *out_count = n;
    return 0;
}

// ------------------------------------------------------------------ model
int skb_model_create(int32_t n_links, const int32_t *parent, const int32_t *joint_type, const double *origin,
                     const double *axis, skb_model **out) {
    if (!out || n_links <= 0 || !parent || !joint_type || !origin || !axis) return set_error("model_create: bad arguments");
    if (parent[0] != -1) return set_error("model_create: link 0 must be the root (parent -1)");
    for (int i = 1; i < n_links; ++i)
        if (parent[i] < 0 || parent[i] >= i) return set_error("model_create: links must be topologically sorted");
    skb_model *m = new skb_model();
    m->parent.assign(parent, parent + n_links);
    m->jtype.assign(joint_type, joint_type + n_links);
    m->origin.assign(origin, origin + 6 * (size_t)n_links);
    m->axis.assign(axis, axis + 3 * (size_t)n_links);
    m->angle.assign(n_links, 0.0);
    for (int k = 0; k < 6; ++k) m->base_pose[k] = 0.0;
    *out = m;
    return 0;
}

int skb_model_add_link(skb_model *m, int32_t parent, const double xyz[3], const double rpy[3], int32_t *out_link_id) {
    if (!m || parent < 0 || parent >= (int)m->parent.size()) return set_error("add_link: bad parent link id");
    m->parent.push_back(parent);
    m->jtype.push_back(SKB_JOINT_FIXED);
    for (int k = 0; k < 3; ++k) m->origin.push_back(xyz ? xyz[k] : 0.0);
    for (int k = 0; k < 3; ++k) m->origin.push_back(rpy ? rpy[k] : 0.0);
    m->axis.push_back(1.0); m->axis.push_back(0.0); m->axis.push_back(0.0);
    m->angle.push_back(0.0);
    if (out_link_id) *out_link_id = (int)m->parent.size() - 1;
    return 0;
}

int skb_model_set_q(skb_model *m, const int32_t *joint_links, int32_t nj, const double *q, int32_t base_type) {
    if (!m || (nj > 0 && (!joint_links || !q))) return set_error("set_q: bad arguments");
    for (int j = 0; j < nj; ++j) {
        if (joint_links[j] <= 0 || joint_links[j] >= (int)m->parent.size()) return set_error("set_q: joint id out of range");
        m->angle[joint_links[j]] = q[j];
    }
    if (base_type == SKB_BASE_PLANER) {
        m->base_pose[0] = q[nj]; m->base_pose[1] = q[nj + 1]; m->base_pose[2] = 0;
        m->base_pose[3] = 0; m->base_pose[4] = 0; m->base_pose[5] = q[nj + 2];
    } else if (base_type == SKB_BASE_FLOATING) {
        for (int k = 0; k < 6; ++k) m->base_pose[k] = q[nj + k];
    } else if (base_type != SKB_BASE_FIXED) {
        return set_error("set_q: bad base_type");
    }
    return 0;
}

int skb_model_get_state(const skb_model *m, double *out_angles, double out_base_pose[6]) {
    if (!m) return set_error("get_state: null model");
    if (out_angles) memcpy(out_angles, m->angle.data(), m->angle.size() * sizeof(double));
    if (out_base_pose) memcpy(out_base_pose, m->base_pose, 6 * sizeof(double));
    return 0;
}

int skb_model_num_links(const skb_model *m, int32_t *out) {
    if (!m || !out) return set_error("num_links: bad arguments");
    *out = (int)m->parent.size();
    return 0;
}

int skb_model_clone(const skb_model *m, skb_model **out) {
    if (!m || !out) return set_error("model_clone: bad arguments");
    *out = new skb_model(*m);
    return 0;
}

void skb_model_destroy(skb_model *m) { delete m; }

// ------------------------------------------------------------------ sdf
int skb_sdf_create(skb_sdf **out) {
    if (!out) return set_error("sdf_create: bad arguments");
    *out = new skb_sdf();
    return 0;
}

static void fill_pose(skb_host_prim &h, const double pos[3], const double rot[9]) {
    static const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int k = 0; k < 3; ++k) h.pos[k] = pos ? pos[k] : 0.0;
    for (int k = 0; k < 9; ++k) h.rot[k] = rot ? rot[k] : I3[k];
    for (int k = 0; k < 8; ++k) h.par[k] = 0.0;
    h.dims[0] = h.dims[1] = h.dims[2] = 0;
}

int skb_sdf_add_box(skb_sdf *s, const double pos[3], const double rot[9], const double half[3]) {
    if (!s || !half) return set_error("sdf_add_box: bad arguments");
    skb_host_prim h; h.type = PRIM_BOX; fill_pose(h, pos, rot);
    for (int k = 0; k < 3; ++k) h.par[k] = half[k];
    s->prims.push_back(h);
    return 0;
}

int skb_sdf_add_cylinder(skb_sdf *s, const double pos[3], const double rot[9], double radius, double half_height) {
    if (!s) return set_error("sdf_add_cylinder: bad arguments");
    skb_host_prim h; h.type = PRIM_CYLINDER; fill_pose(h, pos, rot);
    h.par[0] = radius; h.par[1] = half_height;
    s->prims.push_back(h);
    return 0;
}

int skb_sdf_add_sphere(skb_sdf *s, const double pos[3], double radius) {
    if (!s) return set_error("sdf_add_sphere: bad arguments");
    skb_host_prim h; h.type = PRIM_SPHERE; fill_pose(h, pos, nullptr);
    h.par[0] = radius;
    s->prims.push_back(h);
    return 0;
}

int skb_sdf_add_grid(skb_sdf *s, const double pos[3], const double rot[9], const double lb[3], const double spacing[3],
                     const int32_t dims[3], double fill, const void *host_values, int32_t values_dtype) {
    if (!s || !lb || !spacing || !dims || !host_values) return set_error("sdf_add_grid: bad arguments");
    if (dims[0] < 2 || dims[1] < 2 || dims[2] < 2) return set_error("sdf_add_grid: every grid dimension must be >= 2");
    if (!std::isfinite(fill)) return set_error("sdf_add_grid: fill value must be finite");
    skb_host_prim h; h.type = PRIM_GRID; fill_pose(h, pos, rot);
    for (int k = 0; k < 3; ++k) { h.par[k] = lb[k]; h.par[3 + k] = spacing[k]; h.dims[k] = dims[k]; }
    h.par[6] = fill;
    h.grid_is_f64 = values_dtype == SKB_F64;
    size_t count = (size_t)dims[0] * dims[1] * dims[2];
    h.grid_bytes = count * (h.grid_is_f64 ? 8 : 4);
    if (check_sdf_device(s, "sdf_add_grid") != 0) return -1;     // the grid lives on the current device
    CUDA_OK(cudaMalloc(&h.dev_grid, h.grid_bytes));
    CUDA_OK(cudaMemcpy(h.dev_grid, host_values, h.grid_bytes, cudaMemcpyHostToDevice));
    // corner-packed copy: one 32-byte (fp32) gather per trilinear lookup instead of eight 4-byte ones.
    // 8x the grid's memory, so only while it stays within SKB_GRID_CELLS_MAX_MB (default 1024 MB); SKB_GRID_CELLS=0 disables.
    {
        const char *en = getenv("SKB_GRID_CELLS");
        const char *mx = getenv("SKB_GRID_CELLS_MAX_MB");
        const double max_mb = mx ? atof(mx) : 1024.0;
        const size_t n_cells = (size_t)(dims[0] - 1) * (dims[1] - 1) * (dims[2] - 1);
        const size_t cell_bytes = n_cells * 8 * (h.grid_is_f64 ? 8 : 4);
        if (!(en && en[0] == '0') && cell_bytes <= (size_t)(max_mb * 1048576.0)) {
            if (cudaMalloc(&h.dev_cells, cell_bytes) == cudaSuccess) {
                if (launch_expand_cells(h.dev_grid, h.grid_is_f64, dims, h.dev_cells) != 0) { cudaFree(h.dev_cells); h.dev_cells = nullptr; }
                // fp32 packs are also reachable as a linear texture of float4 texels (two per cell): the step kernel's voxel
                // lookups then travel the TEX pipe of the L1 instead of the LSU data pipe its shared-memory traffic saturates
                const char *tx = getenv("SKB_GRID_TEX");
                if (h.dev_cells && !h.grid_is_f64 && !(tx && tx[0] == '0') && 2 * n_cells <= (size_t)1 << 27) {
                    cudaResourceDesc rd;
                    memset(&rd, 0, sizeof(rd));
                    rd.resType = cudaResourceTypeLinear;
                    rd.res.linear.devPtr = h.dev_cells;
                    rd.res.linear.desc = cudaCreateChannelDesc<float4>();
                    rd.res.linear.sizeInBytes = cell_bytes;
                    cudaTextureDesc td;
                    memset(&td, 0, sizeof(td));
                    td.readMode = cudaReadModeElementType;
                    td.filterMode = cudaFilterModePoint;
                    td.addressMode[0] = cudaAddressModeClamp;
                    cudaTextureObject_t tex = 0;
                    if (cudaCreateTextureObject(&tex, &rd, &td, nullptr) == cudaSuccess) h.cells_tex = (unsigned long long)tex;
                    else cudaGetLastError();
                }
            } else {
                h.dev_cells = nullptr;
                cudaGetLastError();
            }
        }
    }
    s->prims.push_back(h);
    return 0;
}

int skb_sdf_num_prims(const skb_sdf *s, int32_t *out) {
    if (!s || !out) return set_error("sdf_num_prims: bad arguments");
    *out = (int)s->prims.size();
    return 0;
}

void skb_sdf_destroy(skb_sdf *s) {
    if (!s) return;
    for (auto &h : s->prims) {
        if (h.cells_tex) cudaDestroyTextureObject((cudaTextureObject_t)h.cells_tex);
        if (h.dev_grid) cudaFree(h.dev_grid);
        if (h.dev_cells) cudaFree(h.dev_cells);
    }
    if (s->dev_f32) cudaFree(s->dev_f32);
    if (s->dev_f64) cudaFree(s->dev_f64);
    delete s;
}

// ------------------------------------------------------------------ plan
int skb_plan_create(const skb_model *m, const int32_t *feature_links, int32_t nf, const int32_t *joint_links, int32_t nj,
                    int32_t base_type, int32_t rot_type, skb_plan **out) {
    if (!m || !out || nf <= 0 || !feature_links || (nj > 0 && !joint_links)) return set_error("plan_create: bad arguments");
    skb_plan *p = new skb_plan();
    if (compile_tree(m, feature_links, nf, joint_links, nj, base_type, rot_type, p) != 0) { delete p; return -1; }
    if (p->max_active > 48) { delete p; return set_error("plan_create: more than 48 controlled ancestors on one chain is not supported"); }
    cudaGetDevice(&p->device);
    *out = p;
    return 0;
}

int skb_plan_dims(const skb_plan *p, int32_t *D, int32_t *F, int32_t *T) {
    if (!p) return set_error("plan_dims: null plan");
    if (D) *D = p->D;
    if (F) *F = p->F;
    if (T) *T = p->T;
    return 0;
}

static void drop_schedules(skb_plan *p) {
    for (auto &kv : p->schedules) free_schedule(kv.second);
    p->schedules.clear();
    free_sphere(p->sphere);
    free_s2(p->s2_collfree);
    free_s2(p->s2_pairs);
    free_com(p->com);
    free_pose(p->pose);
}
static void drop_pairs(skb_plan *p) {
    free_s2(p->s2_pairs);
    if (p->pairs.dI) cudaFree(p->pairs.dI);
    if (p->pairs.dF32) cudaFree(p->pairs.dF32);
    if (p->pairs.dF64) cudaFree(p->pairs.dF64);
    p->pairs = skb_pair_program();
}

int skb_plan_set_radii(skb_plan *p, const double *radii) {
    if (!p || !radii) return set_error("set_radii: bad arguments");
    std::lock_guard<std::mutex> lk(p->mu);
    p->radii.assign(radii, radii + p->F);
    drop_schedules(p);
    drop_pairs(p);
    return 0;
}

int skb_plan_set_weights(skb_plan *p, const double *weights) {
    if (!p || !weights) return set_error("set_weights: bad arguments");
    std::lock_guard<std::mutex> lk(p->mu);
    p->weights.assign(weights, weights + p->F);
    free_com(p->com);
    return 0;
}

int skb_plan_set_pairs(skb_plan *p, const int32_t *pairs, int32_t n_pairs, const double *rsum_sq) {
    if (!p || n_pairs <= 0 || !pairs || !rsum_sq) return set_error("set_pairs: bad arguments");
    std::lock_guard<std::mutex> lk(p->mu);
    p->pair_idx.assign(pairs, pairs + 2 * (size_t)n_pairs);
    p->pair_thr.assign(rsum_sq, rsum_sq + n_pairs);
    p->pairs_set = true;
    drop_pairs(p);
    return 0;
}

int skb_plan_set_tuning(skb_plan *p, int32_t chunk, int32_t nbuf, int32_t warps, int32_t ctas) {
    if (!p) return set_error("set_tuning: null plan");
    std::lock_guard<std::mutex> lk(p->mu);
    if (nbuf < 0 || nbuf > 4 || chunk < 0 || chunk > 32) return set_error("set_tuning: out of range");
    p->tune_chunk = chunk; p->tune_nbuf = nbuf; p->tune_warps = warps; p->tune_ctas = ctas;
    drop_schedules(p);
    return 0;
}

void skb_plan_destroy_hostpipe(skb_plan *p);   // skb_api.cpp

void skb_plan_destroy(skb_plan *p) {
    if (!p) return;
    skb_plan_destroy_hostpipe(p);
    drop_schedules(p);
    drop_pairs(p);
    delete p;
}

}  // extern "C"

/*
 * skmp_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C, fp64 restatement of the arithmetic that scikit-motionplan's hot path
 * delegates to two un-vendored dependencies (tinyfk >= 0.6.7, scikit-robot >= 0.0.29,
 * reference setup.py:11,13).  Neither package nor any URDF is present in the build
 * container, so this restates their *published algorithms* and anchors on the
 * reference's own call sites:
 *
 *   orc_solve_fk              <- tinyfk.KinematicModel.solve_fk      (skmp/kinematics.py:48-55)
 *   orc_inter_link_sqdists    <- tinyfk compute_inter_link_sqdists   (skmp/constraint.py:715,752-758)
 *   orc_sdf_eval              <- skrobot.sdf Box/Cylinder/Sphere/Union/Grid callables
 *                                (skmp/constraint.py:268,314,353; tests/test_constraint.py:71-75)
 *   orc_collfree              <- CollFreeConst._evaluate_all/_evaluate_closest
 *                                (skmp/constraint.py:287-374), incl. the forward-FD
 *                                gradient (eps=1e-7, :320-331,:362-369) as grad_mode=FD.
 *
 * PARITY UNPINNED: the reference holds no golden vectors for this path and its
 * arithmetic lives in packages that cannot be imported here.  The
 * oracle is pinned only by (i) the reference's own property tests, ported in
 * tests/test_oracle.py, (ii) an independent numpy re-derivation, (iii) committed
 * fixtures under tests/golden/ generated from this file.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (scikit-motionplan_b200/) never does.
 *
 * Deliberately written the way tinyfk does it (per configuration: set joint angles,
 * walk each feature link's parent chain with a per-config transform cache, poses as
 * position + quaternion like urdfdom's urdf::Pose, geometric Jacobian per *relevant*
 * joint) -- not the way the CUDA engine does it (folded constants, twist/wrench
 * pairing, rotation matrices), so that agreement between the two means something.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

enum { JT_FIXED = 0, JT_REVOLUTE = 1, JT_PRISMATIC = 2 };
enum { BASE_FIXED = 0, BASE_PLANER = 1, BASE_FLOATING = 2 };     /* tinyfk.BaseType */
enum { ROT_IGNORE = 0, ROT_RPY = 1, ROT_XYZW = 2 };               /* tinyfk.RotationType */
enum { SDF_BOX = 0, SDF_CYLINDER = 1, SDF_SPHERE = 2, SDF_GRID = 3 };
enum { GRAD_ANALYTIC = 0, GRAD_FD = 1 };

typedef struct {
    int32_t n_links;
    const int32_t *parent;  /* [n_links] parent link index, -1 for the root                 */
    const int32_t *jtype;   /* [n_links] type of the joint that connects parent -> link     */
    const double *origin;   /* [n_links*6] joint origin xyz + rpy in the parent link frame  */
    const double *axis;     /* [n_links*3] joint axis in the joint frame                    */
    const double *angle;    /* [n_links] sticky joint value (tinyfk set_q state)            */
    double base_pose[6];    /* sticky base pose xyz + rpy (skmp/kinematics.py:86-96)        */
} orc_tree;

typedef struct {
    int32_t type;
    int32_t dims[3];        /* grid only                                                     */
    double pos[3];          /* world position of the primitive frame                         */
    double rot[9];          /* world-from-local rotation, row major                          */
    double param[8];        /* box: half extents[3]; cylinder: radius, half height;
                               sphere: radius; grid: lb[3] (local), spacing[3], fill         */
    const float *grid_f32;  /* grid values, C order [nx][ny][nz] (one of the two non-NULL)   */
    const double *grid_f64;
} orc_prim;

/* ------------------------------------------------------------------ small math */
typedef struct { double p[3]; double q[4]; } pose_t;   /* q = (x, y, z, w), as urdf::Rotation */

static void quat_mul(const double *a, const double *b, double *o) {
    double x = a[3]*b[0] + a[0]*b[3] + a[1]*b[2] - a[2]*b[1];
    double y = a[3]*b[1] - a[0]*b[2] + a[1]*b[3] + a[2]*b[0];
    double z = a[3]*b[2] + a[0]*b[1] - a[1]*b[0] + a[2]*b[3];
    double w = a[3]*b[3] - a[0]*b[0] - a[1]*b[1] - a[2]*b[2];
    o[0] = x; o[1] = y; o[2] = z; o[3] = w;
}
static void cross3(const double *a, const double *b, double *o) {
    double x = a[1]*b[2] - a[2]*b[1], y = a[2]*b[0] - a[0]*b[2], z = a[0]*b[1] - a[1]*b[0];
    o[0] = x; o[1] = y; o[2] = z;
}
/* v' = q v q^-1 */
static void quat_rot(const double *q, const double *v, double *o) {
    double t[3], u[3];
    cross3(q, v, t); t[0] *= 2; t[1] *= 2; t[2] *= 2;
    cross3(q, t, u);
    o[0] = v[0] + q[3]*t[0] + u[0];
    o[1] = v[1] + q[3]*t[1] + u[1];
    o[2] = v[2] + q[3]*t[2] + u[2];
}
static void quat_from_rpy(double r, double p, double y, double *q) {  /* urdfdom setFromRPY */
    double phi = r/2, the = p/2, psi = y/2;
    q[0] = sin(phi)*cos(the)*cos(psi) - cos(phi)*sin(the)*sin(psi);
    q[1] = cos(phi)*sin(the)*cos(psi) + sin(phi)*cos(the)*sin(psi);
    q[2] = cos(phi)*cos(the)*sin(psi) - sin(phi)*sin(the)*cos(psi);
    q[3] = cos(phi)*cos(the)*cos(psi) + sin(phi)*sin(the)*sin(psi);
}
static void quat_get_rpy(const double *q, double *rpy) {               /* urdfdom getRPY */
    double sqx = q[0]*q[0], sqy = q[1]*q[1], sqz = q[2]*q[2], sqw = q[3]*q[3];
    rpy[0] = atan2(2*(q[1]*q[2] + q[3]*q[0]), sqw - sqx - sqy + sqz);
    double sarg = -2*(q[0]*q[2] - q[3]*q[1]);
    rpy[1] = sarg <= -1.0 ? -0.5*M_PI : (sarg >= 1.0 ? 0.5*M_PI : asin(sarg));
    rpy[2] = atan2(2*(q[0]*q[1] + q[3]*q[2]), sqw + sqx - sqy - sqz);
}
static void pose_compose(const pose_t *a, const pose_t *b, pose_t *o) {
    double t[3]; pose_t r;
    quat_rot(a->q, b->p, t);
    r.p[0] = a->p[0] + t[0]; r.p[1] = a->p[1] + t[1]; r.p[2] = a->p[2] + t[2];
    quat_mul(a->q, b->q, r.q);
    *o = r;
}

/* ------------------------------------------------------------------ per-config FK state */
typedef struct {
    const orc_tree *t;
    double *angle;      /* working joint values (sticky + this configuration's columns) */
    pose_t base;
    pose_t *world;      /* per-link world pose cache */
    uint8_t *done;
} fkstate;

static int fk_init(fkstate *s, const orc_tree *t) {
    s->t = t;
    s->angle = (double *)malloc(sizeof(double) * t->n_links);
    s->world = (pose_t *)malloc(sizeof(pose_t) * t->n_links);
    s->done = (uint8_t *)malloc(t->n_links);
    return s->angle && s->world && s->done;
}
static void fk_free(fkstate *s) { free(s->angle); free(s->world); free(s->done); }

static void fk_set_config(fkstate *s, const double *q, const int32_t *joint_ids, int nj, int base_type) {
    const orc_tree *t = s->t;
    memcpy(s->angle, t->angle, sizeof(double) * t->n_links);
    for (int j = 0; j < nj; ++j) s->angle[joint_ids[j]] = q[j];
    double bp[6];
    memcpy(bp, t->base_pose, sizeof bp);
    if (base_type == BASE_PLANER) {            /* [x, y, theta]: skmp/robot/utils.py:45-50 */
        bp[0] = q[nj]; bp[1] = q[nj+1]; bp[2] = 0.0; bp[3] = 0.0; bp[4] = 0.0; bp[5] = q[nj+2];
    } else if (base_type == BASE_FLOATING) {   /* [x, y, z, roll, pitch, yaw]: utils.py:51-56 */
        for (int k = 0; k < 6; ++k) bp[k] = q[nj+k];
    }
    s->base.p[0] = bp[0]; s->base.p[1] = bp[1]; s->base.p[2] = bp[2];
    quat_from_rpy(bp[3], bp[4], bp[5], s->base.q);
    memset(s->done, 0, t->n_links);
}

static const pose_t *fk_link_pose(fkstate *s, int link) {
    if (s->done[link]) return &s->world[link];
    const orc_tree *t = s->t;
    /* iterative walk up to the first cached ancestor, then back down */
    int chain[512], n = 0, l = link;
    while (l >= 0 && !s->done[l]) { chain[n++] = l; l = t->parent[l]; }
    for (int i = n - 1; i >= 0; --i) {
        int k = chain[i];
        pose_t par = (t->parent[k] < 0) ? s->base : s->world[t->parent[k]];
        pose_t org, mot, tmp;
        const double *o = &t->origin[6*k];
        org.p[0] = o[0]; org.p[1] = o[1]; org.p[2] = o[2];
        quat_from_rpy(o[3], o[4], o[5], org.q);
        mot.p[0] = mot.p[1] = mot.p[2] = 0; mot.q[0] = mot.q[1] = mot.q[2] = 0; mot.q[3] = 1;
        const double *a = &t->axis[3*k];
        if (t->jtype[k] == JT_REVOLUTE) {
            double h = 0.5 * s->angle[k], sh = sin(h);
            mot.q[0] = a[0]*sh; mot.q[1] = a[1]*sh; mot.q[2] = a[2]*sh; mot.q[3] = cos(h);
        } else if (t->jtype[k] == JT_PRISMATIC) {
            mot.p[0] = a[0]*s->angle[k]; mot.p[1] = a[1]*s->angle[k]; mot.p[2] = a[2]*s->angle[k];
        }
        if (t->parent[k] < 0) {
            /* the root link carries no joint: world(root) = base pose */
            s->world[k] = s->base;
        } else {
            pose_compose(&par, &org, &tmp);
            pose_compose(&tmp, &mot, &s->world[k]);
        }
        s->done[k] = 1;
    }
    return &s->world[link];
}

static int is_ancestor_or_self(const orc_tree *t, int anc, int link) {
    for (int l = link; l >= 0; l = t->parent[l]) if (l == anc) return 1;
    return 0;
}

/* rotational-rate maps: d(rpy)/dt and d(xyzw)/dt for a world angular velocity w */
static void omega_to_rpy_rate(const double *rpy, const double *w, double *o) {
    double cp = cos(rpy[1]), sp = sin(rpy[1]), cy = cos(rpy[2]), sy = sin(rpy[2]);
    double rr = (cy*w[0] + sy*w[1]) / cp;
    o[0] = rr;
    o[1] = -sy*w[0] + cy*w[1];
    o[2] = w[2] + sp*rr;
}
static void omega_to_quat_rate(const double *q, const double *w, double *o) {
    double c[3]; cross3(w, q, c);
    o[0] = 0.5*(q[3]*w[0] + c[0]);
    o[1] = 0.5*(q[3]*w[1] + c[1]);
    o[2] = 0.5*(q[3]*w[2] + c[2]);
    o[3] = -0.5*(w[0]*q[0] + w[1]*q[1] + w[2]*q[2]);
}

static int tdim(int rot_type) { return rot_type == ROT_IGNORE ? 3 : (rot_type == ROT_RPY ? 6 : 7); }

/* one feature: pose row (T) and Jacobian rows (T x D) */
static void feature_eval(fkstate *s, int feat, const int32_t *joint_ids, int nj, int base_type,
                         int rot_type, int with_jac, double *out_p, double *out_j) {
    const orc_tree *t = s->t;
    int T = tdim(rot_type);
    int D = nj + (base_type == BASE_PLANER ? 3 : (base_type == BASE_FLOATING ? 6 : 0));
    const pose_t *pe = fk_link_pose(s, feat);
    double rpy[3] = {0, 0, 0};
    out_p[0] = pe->p[0]; out_p[1] = pe->p[1]; out_p[2] = pe->p[2];
    if (rot_type == ROT_RPY) { quat_get_rpy(pe->q, rpy); out_p[3] = rpy[0]; out_p[4] = rpy[1]; out_p[5] = rpy[2]; }
    if (rot_type == ROT_XYZW) { out_p[3] = pe->q[0]; out_p[4] = pe->q[1]; out_p[5] = pe->q[2]; out_p[6] = pe->q[3]; }
    if (!with_jac) return;
    for (int i = 0; i < T*D; ++i) out_j[i] = 0.0;

    for (int c = 0; c < D; ++c) {
        double w[3] = {0, 0, 0}, dpos[3] = {0, 0, 0};
        int active = 0;
        if (c < nj) {
            int jl = joint_ids[c];
            if (t->jtype[jl] == JT_FIXED || !is_ancestor_or_self(t, jl, feat)) continue;
            const pose_t *pj = fk_link_pose(s, jl);
            double aw[3];
            quat_rot(pj->q, &t->axis[3*jl], aw);
            if (t->jtype[jl] == JT_REVOLUTE) {
                double r[3] = {pe->p[0]-pj->p[0], pe->p[1]-pj->p[1], pe->p[2]-pj->p[2]};
                cross3(aw, r, dpos);
                w[0] = aw[0]; w[1] = aw[1]; w[2] = aw[2];
            } else {
                dpos[0] = aw[0]; dpos[1] = aw[1]; dpos[2] = aw[2];
            }
            active = 1;
        } else {
            int b = c - nj;
            double r[3] = {pe->p[0]-s->base.p[0], pe->p[1]-s->base.p[1], pe->p[2]-s->base.p[2]};
            if (base_type == BASE_PLANER) {
                if (b < 2) dpos[b] = 1.0;
                else { w[2] = 1.0; cross3(w, r, dpos); }
            } else {
                if (b < 3) dpos[b] = 1.0;
                else {
                    /* R = Rz(yaw) Ry(pitch) Rx(roll); columns roll, pitch, yaw */
                    double brpy[3]; quat_get_rpy(s->base.q, brpy);
                    double cp = cos(brpy[1]), sp = sin(brpy[1]), cy = cos(brpy[2]), sy = sin(brpy[2]);
                    if (b == 3) { w[0] = cy*cp; w[1] = sy*cp; w[2] = -sp; }
                    if (b == 4) { w[0] = -sy;   w[1] = cy;    w[2] = 0;  }
                    if (b == 5) { w[0] = 0;     w[1] = 0;     w[2] = 1;  }
                    cross3(w, r, dpos);
                }
            }
            active = 1;
        }
        if (!active) continue;
        out_j[0*D + c] = dpos[0]; out_j[1*D + c] = dpos[1]; out_j[2*D + c] = dpos[2];
        if (rot_type == ROT_RPY) {
            double rr[3]; omega_to_rpy_rate(rpy, w, rr);
            out_j[3*D + c] = rr[0]; out_j[4*D + c] = rr[1]; out_j[5*D + c] = rr[2];
        } else if (rot_type == ROT_XYZW) {
            double qr[4]; omega_to_quat_rate(pe->q, w, qr);
            out_j[3*D + c] = qr[0]; out_j[4*D + c] = qr[1]; out_j[5*D + c] = qr[2]; out_j[6*D + c] = qr[3];
        }
    }
}

/* solve_fk: qs (N,D) -> pts (N,F,T), jac (N,F,T,D)   [skmp/kinematics.py:57,61] */
ORC_API int orc_solve_fk(const orc_tree *t, const double *qs, int64_t N,
                         const int32_t *feat_ids, int nf, const int32_t *joint_ids, int nj,
                         int base_type, int rot_type, int with_jac, double *out_pts, double *out_jac) {
    fkstate s;
    if (!fk_init(&s, t)) return -1;
    int T = tdim(rot_type);
    int D = nj + (base_type == BASE_PLANER ? 3 : (base_type == BASE_FLOATING ? 6 : 0));
    for (int64_t n = 0; n < N; ++n) {
        fk_set_config(&s, qs + n*D, joint_ids, nj, base_type);
        for (int f = 0; f < nf; ++f)
            feature_eval(&s, feat_ids[f], joint_ids, nj, base_type, rot_type, with_jac,
                         out_pts + (n*nf + f)*T, with_jac ? out_jac + (n*nf + f)*(int64_t)T*D : NULL);
    }
    fk_free(&s);
    return 0;
}

/* compute_inter_link_sqdists: sq (N*P), grads (N*P, D)   [skmp/constraint.py:752-778] */
ORC_API int orc_inter_link_sqdists(const orc_tree *t, const double *qs, int64_t N,
                                   const int32_t *pairs, int P, const int32_t *joint_ids, int nj,
                                   int base_type, int with_jac, double *out_sq, double *out_grad) {
    fkstate s;
    if (!fk_init(&s, t)) return -1;
    int D = nj + (base_type == BASE_PLANER ? 3 : (base_type == BASE_FLOATING ? 6 : 0));
    double *ja = (double *)malloc(sizeof(double) * 3 * D * 2), *jb = ja + 3*D;
    for (int64_t n = 0; n < N; ++n) {
        fk_set_config(&s, qs + n*D, joint_ids, nj, base_type);
        for (int p = 0; p < P; ++p) {
            double pa[3], pb[3];
            feature_eval(&s, pairs[2*p], joint_ids, nj, base_type, ROT_IGNORE, with_jac, pa, ja);
            feature_eval(&s, pairs[2*p+1], joint_ids, nj, base_type, ROT_IGNORE, with_jac, pb, jb);
            double d[3] = {pa[0]-pb[0], pa[1]-pb[1], pa[2]-pb[2]};
            out_sq[n*P + p] = d[0]*d[0] + d[1]*d[1] + d[2]*d[2];
            if (with_jac)
                for (int c = 0; c < D; ++c)
                    out_grad[(n*P + p)*(int64_t)D + c] =
                        2.0*(d[0]*(ja[c]-jb[c]) + d[1]*(ja[D+c]-jb[D+c]) + d[2]*(ja[2*D+c]-jb[2*D+c]));
        }
    }
    free(ja);
    fk_free(&s);
    return 0;
}

/* ------------------------------------------------------------------ signed distance fields */
static double sgn(double x) { return x >= 0.0 ? 1.0 : -1.0; }

/* Conditioning of the gradient at the last evaluated point: a Lipschitz bound (1/m) of the analytic gradient field
 * around it.  A point that is off by dp (fp32 forward kinematics: up to 1e-6 m) sees a gradient that is off by up to
 * curv * dp whatever the implementation -- 1/r at distance r from a box edge, a cylinder rim or axis, a sphere centre; the
 * mixed second derivatives of the trilinear cell polynomial for a voxel grid.  The parity tests add that term to the
 * Jacobian tolerance (tests/helpers.py::jac_tol) instead of a blanket slack factor. */
static __thread double tl_curv = 0.0;

/* value, analytic gradient (local frame) and distance to the nearest gradient DISCONTINUITY ("kink": medial axis inside a
 * box / cylinder, edges and rims outside, voxel cell faces, the centre of a sphere) */
static double prim_local(const orc_prim *pr, const double *pl, double *gl, double *kink) {
    gl[0] = gl[1] = gl[2] = 0.0;
    *kink = INFINITY;
    tl_curv = 0.0;
    if (pr->type == SDF_BOX) {
        double s[3], qv[3], nq = 0.0; int outside = 0;
        for (int i = 0; i < 3; ++i) {
            s[i] = fabs(pl[i]) - pr->param[i];
            qv[i] = s[i] > 0.0 ? s[i] : 0.0;
            if (s[i] > 0.0) outside = 1;
            nq += qv[i]*qv[i];
        }
        if (outside) {
            double d = sqrt(nq);
            for (int i = 0; i < 3; ++i) gl[i] = sgn(pl[i]) * qv[i] / d;
            /* nearest edge: at distance d in the edge / corner regions, sqrt(d^2 + (lateral distance to the face's border)^2)
             * above a face (the gradient is constant there until the edge region begins) */
            double a = fmax(s[0], fmax(s[1], s[2])), c = fmin(s[0], fmin(s[1], s[2]));
            double second = s[0] + s[1] + s[2] - a - c;
            double lat = second < 0.0 ? -second : 0.0;
            *kink = sqrt(d*d + lat*lat);
            tl_curv = 1.0 / *kink;
            return d;
        }
        int k = 0;
        if (s[1] > s[k]) k = 1;
        if (s[2] > s[k]) k = 2;
        double second = -INFINITY;
        for (int i = 0; i < 3; ++i) if (i != k && s[i] > second) second = s[i];
        *kink = fmin(s[k] - second, fabs(pl[k]));
        gl[k] = sgn(pl[k]);
        return s[k];
    }
    if (pr->type == SDF_CYLINDER) {
        double rho = sqrt(pl[0]*pl[0] + pl[1]*pl[1]);
        double s0 = rho - pr->param[0], s1 = fabs(pl[2]) - pr->param[1];
        double ux = rho > 0.0 ? pl[0]/rho : 0.0, uy = rho > 0.0 ? pl[1]/rho : 0.0;
        if (s0 > 0.0 || s1 > 0.0) {
            double q0 = s0 > 0.0 ? s0 : 0.0, q1 = s1 > 0.0 ? s1 : 0.0;
            double d = sqrt(q0*q0 + q1*q1);
            gl[0] = ux*q0/d; gl[1] = uy*q0/d; gl[2] = sgn(pl[2])*q1/d;
            *kink = sqrt(s0*s0 + s1*s1);                         /* the rim circle */
            tl_curv = 1.0 / *kink + (s0 > 0.0 && rho > 0.0 ? 1.0 / rho : 0.0);
            return d;
        }
        if (s0 >= s1) { gl[0] = ux; gl[1] = uy; *kink = fmin(s0 - s1, rho); tl_curv = rho > 0.0 ? 1.0 / rho : INFINITY; return s0; }
        gl[2] = sgn(pl[2]); *kink = fmin(s1 - s0, fabs(pl[2]));
        return s1;
    }
    if (pr->type == SDF_SPHERE) {
        double n = sqrt(pl[0]*pl[0] + pl[1]*pl[1] + pl[2]*pl[2]);
        if (n > 0.0) { gl[0] = pl[0]/n; gl[1] = pl[1]/n; gl[2] = pl[2]/n; }
        *kink = n;
        tl_curv = n > 0.0 ? 1.0 / n : INFINITY;
        return n - pr->param[0];
    }
    /* SDF_GRID: scipy RegularGridInterpolator(method="linear", bounds_error=False, fill_value) */
    {
        const double *lb = &pr->param[0], *h = &pr->param[3];
        double fill = pr->param[6];
        int idx[3]; double fr[3], kk = INFINITY;
        for (int i = 0; i < 3; ++i) {
            double u = (pl[i] - lb[i]) / h[i];
            double top = (double)(pr->dims[i] - 1);
            if (!(u >= 0.0 && u <= top)) { *kink = INFINITY; return fill; }
            int c = (int)floor(u);
            if (c > pr->dims[i] - 2) c = pr->dims[i] - 2;
            idx[i] = c; fr[i] = u - (double)c;
            double dk = fmin(fr[i], 1.0 - fr[i]) * h[i];
            if (dk < kk) kk = dk;
        }
        *kink = kk;
        int64_t sy = pr->dims[2], sx = (int64_t)pr->dims[1] * pr->dims[2];
        int64_t b = idx[0]*sx + idx[1]*sy + idx[2];
        double c[8];
        for (int i = 0; i < 8; ++i) {
            int64_t o = b + ((i >> 2) & 1)*sx + ((i >> 1) & 1)*sy + (i & 1);
            c[i] = pr->grid_f64 ? pr->grid_f64[o] : (double)pr->grid_f32[o];
        }
        double fx = fr[0], fy = fr[1], fz = fr[2];
        /* c[4*ix + 2*iy + iz] */
        double c00 = c[0]*(1-fx) + c[4]*fx, c01 = c[1]*(1-fx) + c[5]*fx;
        double c10 = c[2]*(1-fx) + c[6]*fx, c11 = c[3]*(1-fx) + c[7]*fx;
        double c0 = c00*(1-fy) + c10*fy, c1 = c01*(1-fy) + c11*fy;
        double val = c0*(1-fz) + c1*fz;
        double dx = ((c[4]-c[0])*(1-fy) + (c[6]-c[2])*fy)*(1-fz) + ((c[5]-c[1])*(1-fy) + (c[7]-c[3])*fy)*fz;
        double dy = (c10 - c00)*(1-fz) + (c11 - c01)*fz;
        double dz = c1 - c0;
        gl[0] = dx / h[0]; gl[1] = dy / h[1]; gl[2] = dz / h[2];
        {   /* mixed second derivatives of the cell polynomial (the pure ones vanish) */
            double a4 = c[6] - c[4] - c[2] + c[0], a5 = c[5] - c[4] - c[1] + c[0], a6 = c[3] - c[2] - c[1] + c[0];
            double a7 = c[7] - c[6] - c[5] - c[3] + c[4] + c[2] + c[1] - c[0];
            double fxy = (fabs(a4) + fabs(a7)) / (h[0]*h[1]), fxz = (fabs(a5) + fabs(a7)) / (h[0]*h[2]), fyz = (fabs(a6) + fabs(a7)) / (h[1]*h[2]);
            tl_curv = fxy + fxz + fyz;
        }
        return val;
    }
}

static double prim_world(const orc_prim *pr, const double *p, double *g, double *kink) {
    double d[3] = {p[0]-pr->pos[0], p[1]-pr->pos[1], p[2]-pr->pos[2]}, pl[3], gl[3];
    const double *R = pr->rot;
    for (int i = 0; i < 3; ++i) pl[i] = R[0*3+i]*d[0] + R[1*3+i]*d[1] + R[2*3+i]*d[2];   /* R^T d */
    double v = prim_local(pr, pl, gl, kink);
    if (g) for (int i = 0; i < 3; ++i) g[i] = R[i*3+0]*gl[0] + R[i*3+1]*gl[1] + R[i*3+2]*gl[2];
    return v;
}

/* UnionSDF: elementwise min over members (first index wins ties, like np.argmin) */
static double union_eval(const orc_prim *prims, int np_, const double *p, double *g, double *kink) {
    double best = INFINITY, second = INFINITY, bk = INFINITY, bc = 0.0, bg[3] = {0, 0, 0};
    for (int i = 0; i < np_; ++i) {
        double gi[3], ki;
        double v = prim_world(&prims[i], p, gi, &ki);
        if (v < best) { second = best; best = v; bk = ki; bc = tl_curv; bg[0] = gi[0]; bg[1] = gi[1]; bg[2] = gi[2]; }
        else if (v < second) second = v;
    }
    if (g) { g[0] = bg[0]; g[1] = bg[1]; g[2] = bg[2]; }
    if (kink) *kink = fmin(bk, second - best);
    tl_curv = bc;
    return best;
}

ORC_API int orc_sdf_eval(const orc_prim *prims, int np_, const double *pts, int64_t M,
                         double *out_val, double *out_grad, double *out_kink, double *out_curv) {
    for (int64_t i = 0; i < M; ++i) {
        double g[3], k;
        out_val[i] = union_eval(prims, np_, pts + 3*i, g, &k);
        if (out_grad) { out_grad[3*i] = g[0]; out_grad[3*i+1] = g[1]; out_grad[3*i+2] = g[2]; }
        if (out_kink) out_kink[i] = k;
        if (out_curv) out_curv[i] = tl_curv;
    }
    return 0;
}

/* CollFreeConst._evaluate (all / closest).  vals (N,S) or (N,1); jac (N,S,D) or (N,1,D).
 * grad_mode FD reproduces skmp/constraint.py:362-369 (forward difference, eps = 1e-7). */
ORC_API int orc_collfree(const orc_tree *t, const double *qs, int64_t N,
                         const int32_t *feat_ids, int nf, const int32_t *joint_ids, int nj, int base_type,
                         const orc_prim *prims, int np_, const double *radii, double margin,
                         int only_closest, int with_jac, int grad_mode,
                         double *out_vals, double *out_jac, double *out_kink, double *out_curv) {
    fkstate s;
    if (!fk_init(&s, t)) return -1;
    int D = nj + (base_type == BASE_PLANER ? 3 : (base_type == BASE_FLOATING ? 6 : 0));
    double *pts = (double *)malloc(sizeof(double) * nf * 3);
    double *jx = (double *)malloc(sizeof(double) * nf * 3 * D);
    double *sd = (double *)malloc(sizeof(double) * nf);
    const double eps = 1e-7;
    for (int64_t n = 0; n < N; ++n) {
        fk_set_config(&s, qs + n*D, joint_ids, nj, base_type);
        for (int f = 0; f < nf; ++f)   /* colkin.map(qs) is always called with jacobian: constraint.py:305,350 */
            feature_eval(&s, feat_ids[f], joint_ids, nj, base_type, ROT_IGNORE, 1, pts + 3*f, jx + (int64_t)f*3*D);
        int f0 = 0, f1 = nf;
        for (int f = 0; f < nf; ++f) sd[f] = union_eval(prims, np_, pts + 3*f, NULL, NULL);
        if (only_closest) {
            int best = 0;
            for (int f = 1; f < nf; ++f) if (sd[f] - radii[f] < sd[best] - radii[best]) best = f;
            f0 = best; f1 = best + 1;
        }
        for (int f = f0; f < f1; ++f) {
            int64_t row = only_closest ? n : n*nf + f;
            double g[3], kink;
            union_eval(prims, np_, pts + 3*f, g, &kink);
            out_vals[row] = sd[f] - radii[f] - margin;
            if (out_kink) out_kink[row] = kink;
            if (out_curv) out_curv[row] = tl_curv;
            if (!with_jac) continue;
            if (grad_mode == GRAD_FD) {
                for (int i = 0; i < 3; ++i) {
                    double pp[3] = {pts[3*f], pts[3*f+1], pts[3*f+2]};
                    pp[i] += eps;
                    g[i] = (union_eval(prims, np_, pp, NULL, NULL) - sd[f]) / eps;
                }
            }
            const double *J = jx + (int64_t)f*3*D;
            for (int c = 0; c < D; ++c)
                out_jac[row*D + c] = g[0]*J[c] + g[1]*J[D+c] + g[2]*J[2*D+c];
        }
    }
    free(pts); free(jx); free(sd);
    fk_free(&s);
    return 0;
}

ORC_API int orc_version(void) { return 2; }

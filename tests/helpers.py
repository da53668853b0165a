"""Shared scene builders for the parity tests (reference scenes: tests/test_constraint.py:67-78,
tests/solver_tests/utils.py:12-41) and an independent numpy forward-kinematics re-derivation used to
pin the C oracle."""
import numpy as np

from oracle import oracle
from skmp_b200.rotmath import rotation_matrix, urdf_rpy_matrix
from skmp_b200.tinyfk import BaseType, RotationType

BASE = {BaseType.FIXED: oracle.BASE_FIXED, BaseType.PLANER: oracle.BASE_PLANER, BaseType.FLOATING: oracle.BASE_FLOATING}
ROT = {RotationType.IGNORE: oracle.ROT_IGNORE, RotationType.RPY: oracle.ROT_RPY, RotationType.XYZW: oracle.ROT_XYZW}


def oracle_args(kin):
    """(tree, feature link ids, joint child-link ids, base code) of a kinematics map, as plain arrays."""
    s = kin.fksolver
    return oracle.Tree.from_solver(s), np.array(kin.tinyfk_feature_ids), s.joint_child_links(kin.tinyfk_joint_ids), BASE[kin.base_type]


def sample_q(kin, n, rng, box=None, base_lo=None, base_hi=None):
    """Uniform joint samples within the URDF limits (continuous joints +-2 pi), base dofs in +-1."""
    jm = kin.fksolver.urdf.joint_map
    lo, hi = [], []
    for name in kin.control_joint_names:
        lim = jm[name].limit
        lo.append(-2 * np.pi if lim.lower is None else lim.lower)
        hi.append(2 * np.pi if lim.upper is None else lim.upper)
    nb = kin.dim_cspace - len(lo)
    lo += list(-np.ones(nb) if base_lo is None else base_lo)
    hi += list(np.ones(nb) if base_hi is None else base_hi)
    lo, hi = np.array(lo), np.array(hi)
    return rng.uniform(size=(n, len(lo))) * (hi - lo) + lo


def numpy_fk_positions(solver, q_full_angles, base_pose, link_ids):
    """Independent re-derivation: 4x4 homogeneous matrices, Rodrigues joint rotations, recursion to the root."""
    def hom(R, t):
        T = np.eye(4)
        T[:3, :3], T[:3, 3] = R, t
        return T

    cache = {}

    def world(l):
        if l in cache:
            return cache[l]
        par = solver.parent[l]
        if par < 0:
            T = hom(urdf_rpy_matrix(*base_pose[3:]), base_pose[:3])
        else:
            o = solver.origin[l]
            T = world(par) @ hom(urdf_rpy_matrix(o[3], o[4], o[5]), o[:3])
            if solver.jtype[l] == 1:
                T = T @ hom(rotation_matrix(q_full_angles[l], solver.axis[l]), np.zeros(3))
            elif solver.jtype[l] == 2:
                T = T @ hom(np.eye(3), solver.axis[l] * q_full_angles[l])
        cache[l] = T
        return T

    return np.array([world(l)[:3, 3] for l in link_ids]), [world(l)[:3, :3] for l in link_ids]


def random_urdf(rng, n_links: int, path) -> str:
    """A random kinematic tree written as URDF: random parents (branching), revolute / prismatic / fixed joints,
    arbitrary (non axis-aligned) joint axes, origins with full xyz + rpy, random masses.  Returns the file path."""
    lines = ['<robot name="fuzz">']
    for i in range(n_links):
        m = float(rng.uniform(0.2, 3.0)) if rng.uniform() < 0.8 else 0.0
        c = rng.uniform(-0.1, 0.1, 3)
        lines.append(f'  <link name="L{i}"><inertial><origin xyz="{c[0]} {c[1]} {c[2]}" rpy="0 0 0"/><mass value="{m}"/></inertial></link>')
    for i in range(1, n_links):
        par = int(rng.integers(max(0, i - 4), i))
        kind = rng.choice(["revolute", "continuous", "prismatic", "fixed"], p=[0.5, 0.15, 0.15, 0.2])
        xyz, rpy = rng.uniform(-0.3, 0.3, 3), rng.uniform(-np.pi, np.pi, 3)
        ax = rng.normal(size=3)
        ax /= np.linalg.norm(ax)
        lim = '' if kind in ("continuous", "fixed") else f'<limit lower="{-rng.uniform(0.5, 2.5)}" upper="{rng.uniform(0.5, 2.5)}" effort="1" velocity="1"/>'
        lines.append(f'  <joint name="J{i}" type="{kind}"><parent link="L{par}"/><child link="L{i}"/>'
                     f'<origin xyz="{xyz[0]} {xyz[1]} {xyz[2]}" rpy="{rpy[0]} {rpy[1]} {rpy[2]}"/>'
                     f'<axis xyz="{ax[0]} {ax[1]} {ax[2]}"/>{lim}</joint>')
    lines.append('</robot>')
    path.write_text("\n".join(lines))
    return str(path)


# ------------------------------------------------------------------ parity tolerances (BASELINE.json north_star)
# fp32 within 1e-5, fp64 within 1e-10, error measure |got - ref| / max(1, |ref|).  Two conditioning terms are added where
# the REFERENCE FUNCTION itself amplifies input rounding -- they are properties of the function, not of an implementation:
#   * SDF gradients: forward kinematics in fp32 places a sphere centre with an error of up to FK_POS_ERR (measured: 7e-7 m at
#     2.5 m from the origin, profiles/parity_r2.json); the analytic gradient field moves by curv * error, where `curv` is the
#     oracle's Lipschitz bound of the gradient at that point (1/r at distance r from a box edge, cylinder rim / axis, sphere
#     centre; the mixed second derivatives of a voxel cell).  In smooth regions (curv << 1/m) the term vanishes.
#   * RPY rows: roll-pitch-yaw angles and their rate map are singular at pitch = +-pi/2; an fp32 rotation matrix (entries good
#     to RPY_ROT_ERR) yields angles good to err / cos(pitch) and Jacobian rows good to err / cos(pitch)^2.
# Rows closer than KINK_BAND to a gradient DISCONTINUITY (where fp32 and fp64 may legitimately take different sides) are
# excluded from the Jacobian comparison and counted.
TOL = {np.float32: 1e-5, np.float64: 1e-10}
KINK_BAND = {np.float32: 2e-5, np.float64: 1e-9}
FK_POS_ERR = {np.float32: 1e-6, np.float64: 2e-15}      # metres (x lever arm of the Jacobian column, <= 2.5 m)
RPY_ROT_ERR = {np.float32: 4e-7, np.float64: 1e-15}


def _dt(dtype):
    return np.dtype(dtype).type


def scaled_err(got, ref):
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    return np.abs(got - ref) / np.maximum(1.0, np.abs(ref))


def max_err(got, ref) -> float:
    e = scaled_err(got, ref)
    return float(e.max()) if e.size else 0.0


def collfree_jac_excess(jac, ref_jac, kink, curv, dtype, fd_eps=0.0):
    """max over the compared entries of error / tolerance for CollFreeConst Jacobians (N, M, D) against the oracle's, with
    the per-row tolerance TOL + (FK_POS_ERR + fd_eps) * curv; rows within KINK_BAND of a discontinuity are excluded.
    ``fd_eps``: when the oracle differentiates by forward differences like the reference (eps = 1e-7), ITS truncation
    error eps / 2 * |second derivative| <= eps * curv belongs to the tolerance as well.
    -> (excess (<= 1 passes), fraction of rows excluded)."""
    dtype = _dt(dtype)
    e = scaled_err(jac, ref_jac)
    tol = TOL[dtype] + (FK_POS_ERR[dtype] + fd_eps) * np.minimum(curv, 1e12)
    ok = kink > KINK_BAND[dtype]
    ratio = (e / tol[..., None])[ok]
    return (float(ratio.max()) if ratio.size else 0.0), float(1.0 - ok.mean())


def pose_excess(vals, jac, ref_pts, ref_jac, rot_type, dtype):
    """error / tolerance of pose rows: ``vals`` (N, F, T) or (N, F*T) features and ``jac`` (N, F, T, D) against the
    oracle's, TOL for positions, quaternions and IGNORE features; for RPY rows TOL + RPY_ROT_ERR / cos(pitch) (angles) and
    TOL + RPY_ROT_ERR / cos(pitch)^2 (their Jacobian rows).  -> (value excess, Jacobian excess)."""
    dtype = _dt(dtype)
    ref_pts = np.asarray(ref_pts, dtype=np.float64)
    n, f, t = ref_pts.shape
    ev = scaled_err(np.asarray(vals).reshape(n, f, t), ref_pts)
    ej = scaled_err(np.asarray(jac).reshape(n, f, t, -1), ref_jac)
    tv = np.full((n, f, t), TOL[dtype])
    tj = np.full((n, f, t), TOL[dtype])
    if t == 6:      # [x y z roll pitch yaw]: skmp/kinematics.py:153-162
        cp = np.maximum(np.abs(np.cos(ref_pts[:, :, 4])), 1e-12)[:, :, None]
        tv[:, :, 3:] += RPY_ROT_ERR[dtype] / cp
        tj[:, :, 3:] += RPY_ROT_ERR[dtype] / cp ** 2
    return float((ev / tv).max()), float((ej / tj[..., None]).max())

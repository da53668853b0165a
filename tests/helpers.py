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

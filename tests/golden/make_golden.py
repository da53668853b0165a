"""Generates the committed golden fixtures tests/golden/*.npz from the CPU oracle.

    python tests/golden/make_golden.py            # rewrite every fixture

The reference stores no expected values for this path and its arithmetic (tinyfk, skrobot) cannot be
imported in the build container, so these vectors are outputs of oracle/skmp_oracle.c on
the reference's own scenes (tests/test_constraint.py:67-78, tests/solver_tests/utils.py:24-37) and on the
BASELINE.json configurations at reduced N.  They (a) freeze the oracle + robot tables against drift
(tests/test_oracle.py::test_golden_fixtures_reproduce, CPU) and (b) are what the CUDA path is compared
with on the GPU box without re-running the oracle (tests/test_gpu_golden.py).
Every case stores its inputs (qs) and fp64 outputs.
"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
for p in (ROOT, ROOT / "scikit-motionplan_b200", HERE.parent):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

from helpers import ROT, oracle_args, sample_q  # noqa: E402
from oracle import oracle  # noqa: E402


def _pr2():
    from skmp_b200.robot_model import PR2

    r = PR2()
    r.reset_manip_pose()
    return r


def box_scene():
    from skmp_b200.sdf import Box

    box = Box([0.7, 0.5, 1.2])
    box.translate([0.85, -0.2, 0.9])
    return box


def grid_scene(dims=(40, 48, 40)):
    """Union of the reference's box and a rotated box, sampled to an fp32 voxel grid (deterministic)."""
    from skmp_b200.rotmath import rotation_matrix
    from skmp_b200.sdf import Box, GridSDF, UnionSDF

    b2 = Box([0.4, 0.4, 0.3])
    b2.translate([0.5, 0.5, 0.6])
    b2.rotate_with_matrix(rotation_matrix(0.5, [0.2, 1.0, 0.1]))
    union = UnionSDF([box_scene(), b2])
    lb, ub = np.array([-0.5, -1.5, 0.0]), np.array([1.5, 1.5, 2.0])
    spacing = (ub - lb) / (np.array(dims) - 1)
    axes = [lb[i] + spacing[i] * np.arange(dims[i]) for i in range(3)]
    pts = np.stack(np.meshgrid(*axes, indexing="ij"), axis=-1).reshape(-1, 3)
    vals = oracle.Sdf(union.primitives())(pts).reshape(dims).astype(np.float32)
    return GridSDF(vals, lb, spacing, fill_value=1.0e3)


def cfg1():
    """BASELINE config 1: PR2 right arm vs the reference's box; the start configuration of
    tests/solver_tests/utils.py:24 is row 0."""
    from skmp_b200.robot.pr2 import PR2Config

    kin = PR2Config().get_collision_kin()
    kin.reflect_skrobot_model(_pr2())
    tree, feat, jl, base = oracle_args(kin)
    rng = np.random.default_rng(100)
    qs = sample_q(kin, 32, rng).astype(np.float32).astype(np.float64)
    qs[0] = np.float32([0.564, 0.35, -0.74, -0.7, -0.7, -0.17, -0.63])
    sdf = oracle.Sdf(box_scene().primitives())
    v, j, kink, curv = oracle.collfree(tree, qs, feat, jl, sdf, kin.get_radius_list(), margin=0.05, with_kink=True, with_curv=True)
    vc, jc = oracle.collfree(tree, qs, feat, jl, sdf, kin.get_radius_list(), margin=0.05, only_closest=True)
    pts, pj = oracle.solve_fk(tree, qs[:8], feat, jl, base)
    return {"qs": qs, "values": v, "jac": j, "kink": kink, "curv": curv, "closest_values": vc, "closest_jac": jc, "points": pts, "points_jac": pj}


def cfg2():
    """BASELINE config 2 at reduced N: Fetch pairwise self-collision, pair list from the q = 0 filter."""
    from skmp_b200.robot.fetch import FetchConfig
    from skmp_b200.robot_model import Fetch
    import itertools

    fetch = Fetch()
    fetch.reset_pose()
    kin = FetchConfig().get_collision_kin()
    kin.reflect_skrobot_model(fetch)
    tree, feat, jl, base = oracle_args(kin)
    radii = dict(zip(feat.tolist(), kin.get_radius_list()))
    pairs = np.array(list(itertools.combinations(feat.tolist(), 2)))
    rs = np.array([radii[a] + radii[b] for a, b in pairs])
    sq0, _ = oracle.inter_link_sqdists(tree, np.zeros((1, 8)), pairs, jl, base, with_jacobian=False)
    keep = np.sqrt(sq0[0]) - 3.0 * rs >= 0
    pairs, thr = pairs[keep], rs[keep] ** 2
    rng = np.random.default_rng(101)
    qs = sample_q(kin, 16, rng).astype(np.float32).astype(np.float64)
    sq, gr = oracle.inter_link_sqdists(tree, qs, pairs, jl, base)
    local = {f: i for i, f in enumerate(feat.tolist())}
    return {"qs": qs, "pairs_local": np.array([[local[a], local[b]] for a, b in pairs], dtype=np.int32),
            "thresholds": thr, "values": sq - thr, "jac": gr}


def cfg3():
    """BASELINE config 3 at reduced N and grid size: PR2 dual arm vs a voxel-grid SDF."""
    from skmp_b200.robot.pr2 import PR2Config

    kin = PR2Config(control_arm="dual").get_collision_kin()
    kin.reflect_skrobot_model(_pr2())
    tree, feat, jl, base = oracle_args(kin)
    grid = grid_scene()
    rng = np.random.default_rng(102)
    qs = sample_q(kin, 16, rng).astype(np.float32).astype(np.float64)
    v, j, kink, curv = oracle.collfree(tree, qs, feat, jl, oracle.Sdf(grid.primitives()), kin.get_radius_list(), with_kink=True, with_curv=True)
    return {"qs": qs, "values": v, "jac": j, "kink": kink, "curv": curv}


def cfg5():
    """BASELINE config 5's pose part: JAXON floating base, 4 end effectors, XYZW."""
    from skmp_b200.robot.jaxon import JaxonConfig
    from skmp_b200.robot_model import Jaxon
    from skmp_b200.tinyfk import RotationType

    jaxon = Jaxon()
    jaxon.reset_manip_pose()
    kin = JaxonConfig().get_endeffector_kin(rot_type=RotationType.XYZW)
    kin.reflect_skrobot_model(jaxon)
    tree, feat, jl, base = oracle_args(kin)
    rng = np.random.default_rng(103)
    qs = (sample_q(kin, 16, rng) * 0.5).astype(np.float32).astype(np.float64)
    pts, jac = oracle.solve_fk(tree, qs, feat, jl, base, ROT[RotationType.XYZW])
    return {"qs": qs, "points": pts, "jac": jac}


CASES = {"cfg1_pr2_rarm_box": cfg1, "cfg2_fetch_pairwise": cfg2, "cfg3_pr2_dual_grid": cfg3, "cfg5_jaxon_pose_xyzw": cfg5}


if __name__ == "__main__":
    for name, fn in CASES.items():
        out = fn()
        np.savez_compressed(HERE / f"{name}.npz", **out)
        print(name, {k: v.shape for k, v in out.items()})

"""Random kinematic trees (branching, arbitrary joint axes, prismatic / fixed joints, non-controlled joints baked as
sticky state, control joints in shuffled order, all three base types) through every CUDA entry point, against the oracle.
The named robots exercise the realistic shapes; this file exercises the plan compiler and the kernels' generality."""
import numpy as np
import pytest

from helpers import BASE, ROT, collfree_jac_excess, pose_excess, random_urdf
from oracle import oracle

pytestmark = pytest.mark.gpu


def _close(a, b, tol):
    return float((np.abs(a - b) / np.maximum(1.0, np.abs(b))).max()) if a.size else 0.0


@pytest.mark.parametrize("seed", range(8))
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_random_tree_all_entry_points(seed, dtype, tmp_path):
    import torch

    from skmp_b200.collision import SphereCollection
    from skmp_b200.constraint import CollFreeConst, PairWiseSelfCollFreeConst
    from skmp_b200.kinematics import CollSphereKinematicsMap, EndEffectorKinematicsMap
    from skmp_b200.robot_model import RobotModel
    from skmp_b200.sdf import Box, Cylinder, Sphere, UnionSDF
    from skmp_b200.tinyfk import BaseType, KinematicModel, RotationType

    rng = np.random.default_rng(500 + seed)
    n_links = int(rng.integers(5, 26))
    urdf = random_urdf(rng, n_links, tmp_path / "fuzz.urdf")
    probe = KinematicModel(urdf)
    movable = [j for j in probe.joint_names if probe.jtype[probe._joint_child[probe._joint_index[j]]] != 0]
    if len(movable) < 2:
        pytest.skip("degenerate draw")
    ctrl = [str(j) for j in rng.permutation(movable)[: max(1, len(movable) * 2 // 3)]]
    base_type = [BaseType.FIXED, BaseType.PLANER, BaseType.FLOATING][seed % 3]
    robot = RobotModel(urdf)
    robot.angle_vector(rng.uniform(-1, 1, len(robot.joint_list)))               # sticky state of the other joints
    if base_type == BaseType.FIXED:
        from skmp_b200.rotmath import rpy_matrix
        robot.newcoords(rng.uniform(-0.5, 0.5, 3), rpy_matrix(*rng.uniform(-0.6, 0.6, 3)))
    D = len(ctrl) + {BaseType.FIXED: 0, BaseType.PLANER: 3, BaseType.FLOATING: 6}[base_type]
    tol = 1e-5 if dtype == np.float32 else 1e-10
    n = 257
    qs = rng.uniform(-1.2, 1.2, size=(n, D)).astype(dtype).astype(np.float64)
    q_dev = torch.as_tensor(qs.astype(dtype), device="cuda")

    # ---- collision spheres on random links -> CollFreeConst (union of three primitive types) and pairwise
    spheres = {}
    for l in rng.choice(n_links, size=min(n_links, 7), replace=False):
        k = int(rng.integers(1, 9))
        spheres[f"L{int(l)}"] = SphereCollection([rng.uniform(-0.15, 0.15, 3) for _ in range(k)], list(rng.uniform(0.02, 0.1, k)),
                                                 [f"s{int(l)}_{i}" for i in range(k)])
    colkin = CollSphereKinematicsMap(urdf, ctrl, spheres, base_type)
    box = Box(rng.uniform(0.2, 0.8, 3), pos=rng.uniform(-0.5, 0.5, 3))
    box.rotate_with_matrix(np.linalg.qr(rng.normal(size=(3, 3)))[0] * 1.0)
    cyl = Cylinder(0.2, 0.7, pos=rng.uniform(-0.8, 0.8, 3))
    sph = Sphere(0.25, pos=rng.uniform(-0.8, 0.8, 3))
    sdf = UnionSDF([box, cyl, sph])
    const = CollFreeConst(colkin, sdf, robot, distance_margin=0.01)
    s = colkin.fksolver
    tree = oracle.Tree.from_solver(s)
    feat, jl = np.array(colkin.tinyfk_feature_ids), s.joint_child_links(colkin.tinyfk_joint_ids)
    rv, rj, kink, curv = oracle.collfree(tree, qs, feat, jl, oracle.Sdf(sdf.primitives()), colkin.get_radius_list(), margin=0.01,
                                         base_type=BASE[base_type], with_kink=True, with_curv=True)
    v, j = const.evaluate(q_dev, True)
    assert v.shape == rv.shape and j.shape == rj.shape
    assert _close(v.cpu().numpy(), rv, tol) < tol
    excess, excluded = collfree_jac_excess(j.cpu().numpy(), rj, kink, curv, dtype)     # per-row tolerance: tol + position error x curvature
    assert excluded < 0.1 and excess <= 1.0
    vc, jc = CollFreeConst(colkin, sdf, robot, only_closest_feature=True, distance_margin=0.01).evaluate(q_dev, True)
    assert torch.equal(vc[:, 0], v.min(dim=1).values)
    assert torch.equal(const.is_valid_batch(q_dev), (v > 0).all(dim=1))
    if colkin.n_feature >= 3:
        ids = colkin.tinyfk_feature_ids
        pairs = [(ids[a], ids[b]) for a in range(len(ids)) for b in range(a + 1, len(ids))]
        pairs = [pairs[i] for i in rng.permutation(len(pairs))[: min(len(pairs), 150)]]
        pw = PairWiseSelfCollFreeConst(colkin, robot, id_pairs=pairs)
        sq, gr = oracle.inter_link_sqdists(tree, qs, np.array(pairs), jl, BASE[base_type])
        pv, pj = pw.evaluate(q_dev, True)
        assert _close(pv.cpu().numpy(), sq - pw.check_sphere_pair_sqdists, tol) < tol
        assert _close(pj.cpu().numpy(), gr, tol) < tol
        pc, pcj = PairWiseSelfCollFreeConst(colkin, robot, id_pairs=pairs, only_closest_feature=True).evaluate(q_dev, True)
        assert torch.equal(pc[:, 0], pv.min(dim=1).values)

    # ---- pose features on random links, all three rotation types
    ee = [f"L{int(l)}" for l in rng.choice(n_links, size=min(n_links, 3), replace=False)]
    for rot in (RotationType.IGNORE, RotationType.RPY, RotationType.XYZW):
        efkin = EndEffectorKinematicsMap(urdf, ctrl, ee, base_type, rot)
        efkin.reflect_skrobot_model(robot)
        s2 = efkin.fksolver
        rp, rpj = oracle.solve_fk(oracle.Tree.from_solver(s2), qs, np.array(efkin.tinyfk_feature_ids),
                                  s2.joint_child_links(efkin.tinyfk_joint_ids), BASE[base_type], ROT[rot])
        pts, jac = efkin.map(q_dev)
        ev, ej = pose_excess(pts.cpu().numpy(), jac.cpu().numpy(), rp, rpj, rot, dtype)    # RPY rows: + err / cos(pitch)^k
        assert ev <= 1.0 and ej <= 1.0, rot

    # ---- centre of mass of the same tree
    com_solver = KinematicModel(urdf)
    com_solver.set_q(list(range(len(com_solver.joint_names))), [robot.__dict__[nm].joint_angle() if nm in robot.__dict__ else 0.0
                                                                 for nm in com_solver.joint_names], BaseType.FIXED)
    jids = com_solver.get_joint_ids(ctrl)
    feats, w = com_solver.com_points()
    xs, jt = com_solver.solve_com_fk(q_dev, jids, [], [], base_type, True)
    rc, rcj = oracle.com_fk(oracle.Tree.from_solver(com_solver), qs, feats, np.array(w), com_solver.joint_child_links(jids), BASE[base_type])
    assert _close(xs.cpu().numpy(), rc, tol) < tol and _close(jt.cpu().numpy().reshape(rcj.shape), rcj, tol) < tol

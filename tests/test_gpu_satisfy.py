"""Batched multi-seed IK / projection (skmp_b200.satisfy) on the reference's own scenarios (tests/test_satisfy.py):
solutions are verified with the CPU oracle, not with the code that produced them."""
import numpy as np
import pytest

from helpers import ROT, oracle_args
from oracle import oracle

pytestmark = pytest.mark.gpu


def _scene(box_pos):
    from skmp_b200.constraint import CollFreeConst
    from skmp_b200.robot.pr2 import PR2Config
    from skmp_b200.robot_model import PR2
    from skmp_b200.sdf import Box
    from skmp_b200.tinyfk import BaseType

    pr2 = PR2(); pr2.reset_manip_pose()
    conf = PR2Config(base_type=BaseType.FIXED)
    colkin, efkin = conf.get_collision_kin(), conf.get_endeffector_kin()
    box = Box([0.7, 0.5, 1.2]); box.translate(box_pos)
    return pr2, conf, colkin, efkin, box, CollFreeConst(colkin, box.sdf, pr2), conf.get_box_const()


def _oracle_collfree(colkin, box, q):
    tree, feat, jl, base = oracle_args(colkin)
    v, _ = oracle.collfree(tree, np.atleast_2d(q), feat, jl, oracle.Sdf(box.primitives()), colkin.get_radius_list(), with_jacobian=False)
    return v


def test_pose_ik_with_collision_many_seeds():
    """tests/test_satisfy.py:22-53: reach (0.7, -0.6, 1.0) with the right arm next to the 0.7 x 0.5 x 1.2 box."""
    from skmp_b200.constraint import ConfigPointConst, PoseConstraint
    from skmp_b200.satisfy import SatisfactionConfig, satisfy_by_optimization, satisfy_by_optimization_batch

    pr2, conf, colkin, efkin, box, ineq, box_const = _scene([0.85, -0.2, 0.9])
    eq1 = ConfigPointConst(np.array([-0.78, 0.055, -1.37, -0.59, -0.494, -0.20, 1.87]))
    eq1.reflect_skrobot_model(pr2)
    r1 = satisfy_by_optimization(eq1, box_const, ineq, None)
    assert r1.success and np.allclose(r1.q, eq1.desired_angles)
    from skmp_b200.robot_model import Coordinates

    eq2 = PoseConstraint.from_skrobot_coords([Coordinates(pos=[0.7, -0.6, 1.0])], efkin, pr2)      # RPY: 6 rows
    res = satisfy_by_optimization_batch(eq2, box_const, ineq, None, SatisfactionConfig(n_seeds=256), rng=np.random.default_rng(0))
    assert res.qs.shape == (256, 7) and res.success.sum() >= 32          # a good share of the restarts converge at once
    tree, feat, jl, base = oracle_args(efkin)
    good = res.qs[res.success]
    pts, _ = oracle.solve_fk(tree, good, feat, jl, base, ROT[efkin.rot_type], with_jacobian=False)
    err = pts.reshape(len(good), -1) - np.array([0.7, -0.6, 1.0, 0.0, 0.0, 0.0])
    assert ((err * err).sum(axis=1) < 1e-6).all()                         # the reference's acceptance test (:108-112)
    assert (_oracle_collfree(colkin, box, good) > 0).all()
    assert ((good >= box_const.lb) & (good <= box_const.ub)).all()
    single = satisfy_by_optimization(eq2, box_const, ineq, None)
    assert single.success and single.q.shape == (7,)
    v, _ = eq2.evaluate_single(single.q, False)
    assert float(np.dot(v, v)) < 1e-6 and ineq.is_valid(single.q)


def test_projection_without_eq_const():
    """tests/test_satisfy.py:82-105: the manipulation pose collides with the box at (0.68, -0.3, 0.9); project out."""
    from skmp_b200.robot.utils import get_robot_state
    from skmp_b200.satisfy import satisfy_by_optimization

    pr2, conf, colkin, efkin, box, ineq, box_const = _scene([0.68, -0.3, 0.9])
    q_now = get_robot_state(pr2, conf._get_control_joint_names())
    assert not ineq.is_valid(q_now)
    res = satisfy_by_optimization(None, box_const, ineq, q_now)
    assert res.success and ineq.is_valid(res.q)
    assert (_oracle_collfree(colkin, box, res.q) > 0).all()
    assert np.linalg.norm(res.q - q_now) < 2.0                            # stays near the seed (objective |q - q_seed|^2)


def test_relative_pose_dual_arm():
    """tests/test_satisfy.py:56-79: RelativePoseConstraint between the two grippers, no inequality."""
    from skmp_b200.constraint import RelativePoseConstraint
    from skmp_b200.robot.pr2 import PR2Config
    from skmp_b200.robot_model import PR2
    from skmp_b200.satisfy import satisfy_by_optimization_with_budget
    from skmp_b200.tinyfk import BaseType

    pr2 = PR2(); pr2.reset_manip_pose()
    conf = PR2Config(base_type=BaseType.FIXED, control_arm="dual")
    efkin = conf.get_endeffector_kin()
    rel = RelativePoseConstraint(np.array([0.0, 0.0, -0.2]), efkin, pr2)
    res = satisfy_by_optimization_with_budget(rel, conf.get_box_const(), None, None, n_trial_budget=30)
    assert res.success
    v, _ = rel.evaluate_single(res.q, False)
    assert float(np.dot(v, v)) < 1e-6


@pytest.mark.parametrize("D,M", [(7, 6), (7, 82), (37, 29), (64, 10)])
def test_lm_step_kernel_matches_numpy(D, M):
    """skb_lm_step against numpy.linalg.solve on random systems (rank-deficient J included: M < D relies on the damping)."""
    import torch

    from skmp_b200.satisfy import _lm_step

    rng = np.random.default_rng(D * 100 + M)
    B = 33
    J = rng.normal(size=(B, M, D))
    J[:, M // 2] = 0.0                                   # an inactive hinge row
    r = rng.normal(size=(B, M))
    r[:, M // 2] = 0.0
    lam = 10.0 ** rng.uniform(-6, 1, size=B)
    q = rng.normal(size=(B, D))
    lb, ub = -np.ones(D) * 1.5, np.ones(D) * 1.5
    t = lambda a: torch.as_tensor(a, dtype=torch.float64, device="cuda")
    got = _lm_step(t(J), t(r), t(lam), t(q), t(lb), t(ub)).cpu().numpy()
    for b in range(B):
        A = J[b].T @ J[b]
        A = A + lam[b] * (np.eye(D) + np.diag(np.diag(A)))
        dq = np.linalg.solve(A, -J[b].T @ r[b])
        ref = np.clip(q[b] + dq, lb, ub)
        assert np.abs(got[b] - ref).max() < 1e-7 * max(1.0, np.abs(dq).max())

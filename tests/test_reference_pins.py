"""The only numbers for this path that the REFERENCE ITSELF holds (tinyfk / scikit-robot / the URDFs are not
available offline and the reference stores no golden vectors) -- checked here against the oracle and the bundled stand-in
PR2 description, on the CPU.  tests/test_gpu_coverage.py::test_reference_pins_on_gpu repeats them through the CUDA path.

1. ``q_goal_cand = [-0.78, 0.055, -1.37, -0.59, -0.494, -0.20, 1.87]`` (tests/solver_tests/test_nlp_solver.py:40,
   tests/test_satisfy.py:35) is a right-arm IK answer the reference hard-codes as a goal candidate.  On the stand-in PR2 (with
   ``reset_manip_pose``'s torso height) its end-effector pose is (0.799, -0.605, 1.091), rpy = (-0.004, -0.006, -0.011):
   the reference's own pose target ``Coordinates(pos=[0.8, -0.6, 1.1])`` with identity rotation (tests/test_constraint.py:131)
   to 9.4 mm and 0.011 rad.  (It is NOT the (0.7, -0.6, 1.0) goal of tests/solver_tests/utils.py:27 -- 13 cm away -- the
   candidate evidently dates from the (0.8, -0.6, 1.1) target.)  An identity end-effector rotation from seven generic joint
   angles pins the joint axes, their order and signs, the RPY convention and the chain of fixed offsets all at once.
2. ``reset_manip_pose`` collides with Box(0.7, 0.5, 1.2) at (0.68, -0.3, 0.9)   (tests/test_satisfy.py:92-100: ``assert not
   ineq_const.is_valid(q_now)``).
3. The standard problem's start configuration and the manipulation pose are collision-free against the same box at
   (0.85, -0.2, 0.9) (tests/solver_tests/utils.py:24-41: solvers are started from it).
"""
import numpy as np

from helpers import ROT, oracle_args
from oracle import oracle

Q_GOAL_CAND = np.array([-0.78, 0.055, -1.37, -0.59, -0.494, -0.20, 1.87])
START = np.array([0.564, 0.35, -0.74, -0.7, -0.7, -0.17, -0.63])


def _pr2():
    from skmp_b200.robot.pr2 import PR2Config
    from skmp_b200.robot_model import PR2

    pr2 = PR2()
    pr2.reset_manip_pose()
    return pr2, PR2Config()


def test_ik_answer_reaches_the_reference_pose_target():
    from skmp_b200.tinyfk import RotationType

    pr2, conf = _pr2()
    efkin = conf.get_endeffector_kin()
    efkin.reflect_skrobot_model(pr2)
    tree, feat, jl, base = oracle_args(efkin)
    pose, _ = oracle.solve_fk(tree, Q_GOAL_CAND[None], feat, jl, base, ROT[RotationType.RPY])
    pos, rpy = pose[0, 0, :3], pose[0, 0, 3:]
    assert np.linalg.norm(pos - np.array([0.8, -0.6, 1.1])) < 0.012, pos
    assert np.abs(rpy).max() < 0.015, rpy
    # XYZW rows: the same pose as a quaternion close to identity (x, y, z, w)
    quat, _ = oracle.solve_fk(tree, Q_GOAL_CAND[None], feat, jl, base, ROT[RotationType.XYZW])
    assert np.abs(quat[0, 0, 3:6]).max() < 0.008 and quat[0, 0, 6] > 0.9999


def test_reference_collision_facts():
    from skmp_b200.robot_model import get_robot_state
    from skmp_b200.sdf import Box

    pr2, conf = _pr2()
    colkin = conf.get_collision_kin()
    colkin.reflect_skrobot_model(pr2)
    assert colkin.n_feature == 76                                     # SURVEY 8 size table, cfg1
    tree, feat, jl, base = oracle_args(colkin)
    q_now = get_robot_state(pr2, conf._get_control_joint_names())

    def min_value(box_pos, q):
        v, _ = oracle.collfree(tree, q[None], feat, jl, oracle.Sdf(Box([0.7, 0.5, 1.2], pos=box_pos).primitives()),
                               colkin.get_radius_list(), with_jacobian=False)
        return v.min()

    assert min_value([0.68, -0.3, 0.9], q_now) < 0                    # tests/test_satisfy.py:96
    assert min_value([0.85, -0.2, 0.9], START) > 0                    # the standard problem is feasible at its start
    assert min_value([0.85, -0.2, 0.9], q_now) > 0

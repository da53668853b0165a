"""Host-side mirror of the reference interface (no GPU needed): names <-> ids, sticky state, kinematics-map
constructors, the torch/numpy-only constraints, composite bookkeeping and error behaviour
(skmp/kinematics.py:86-114, skmp/constraint.py:41-56,116-180,183-263,408-429)."""
import copy

import numpy as np
import pytest

from skmp_b200.constraint import (AbstractEqConst, BoxConst, CollFreeConst, ConfigPointConst, EqCompositeConst,
                                  IneqCompositeConst, PoseConstraint)
from skmp_b200.kinematics import ArticulatedKinematicsMap, KinematicsMapBase
from skmp_b200.robot.fetch import FetchConfig
from skmp_b200.robot.jaxon import JaxonConfig
from skmp_b200.robot.pr2 import PR2Config
from skmp_b200.robot_model import PR2, Fetch
from skmp_b200.sdf import Box
from skmp_b200.tinyfk import BaseType, KinematicModel, RotationType


def test_alias_named_by_north_star():
    assert ArticulatedKinematicsMap is KinematicsMapBase


def test_kinematic_model_ids_and_mutation():
    s = KinematicModel(str(PR2Config.urdf_path()))
    ids = s.get_link_ids(["r_gripper_tool_frame", "l_gripper_tool_frame"])
    assert len(set(ids)) == 2
    with pytest.raises(ValueError):
        s.get_link_ids(["no_such_link"])
    with pytest.raises(ValueError):
        s.get_joint_ids(["no_such_joint"])
    n0 = len(s.link_names)
    s.add_new_link("probe", ids[0], [0.1, 0.0, 0.0])
    assert s.get_link_ids(["probe"]) == [n0]
    with pytest.raises(ValueError):
        s.add_new_link("probe", ids[0], [0.0, 0.0, 0.0])
    jids = s.get_joint_ids(PR2Config.rarm_joint_names())
    s.set_q(jids, np.arange(7) * 0.1 + np.array([0.0] * 7))
    assert np.allclose(s.angle[s.joint_child_links(jids)], np.arange(7) * 0.1)
    s.set_q(jids, list(np.zeros(7)) + [1.0, 2.0, 0.5], base_type=BaseType.PLANER)
    assert np.allclose(s.base_pose, [1.0, 2.0, 0.0, 0.0, 0.0, 0.5])
    c = copy.deepcopy(s)                       # constraint.py:518,567 deep-copy their kinematics map
    c.add_new_link("probe2", ids[1], [0.0, 0.1, 0.0])
    assert len(c.link_names) == len(s.link_names) + 1


def test_reflect_bakes_every_joint_and_the_base():
    pr2 = PR2(); pr2.reset_manip_pose()
    pr2.newcoords([0.3, -0.2, 0.0])
    kin = PR2Config().get_collision_kin()
    kin.reflect_skrobot_model(pr2)
    s = kin.fksolver
    assert np.allclose(s.base_pose[:3], [0.3, -0.2, 0.0])
    torso = s.joint_child_links(s.get_joint_ids(["torso_lift_joint"]))[0]
    assert s.angle[torso] == pytest.approx(pr2.torso_lift_joint.joint_angle())
    l_sh = s.joint_child_links(s.get_joint_ids(["l_shoulder_pan_joint"]))[0]
    assert s.angle[l_sh] == pytest.approx(pr2.l_shoulder_pan_joint.joint_angle())


def test_collision_map_selection_rules():
    """skmp/robot/pr2.py:267-285: a moving base forces whole-body spheres; FIXED right arm drops l_* and base."""
    assert PR2Config().get_collision_kin().n_feature == 76
    assert PR2Config(control_arm="larm").get_collision_kin().n_feature == 76
    assert PR2Config(control_arm="dual").get_collision_kin().n_feature == 152
    assert PR2Config(base_type=BaseType.PLANER).get_collision_kin().n_feature == 161
    k = PR2Config().get_collision_kin()
    assert all(("r_" in n) for n in k.sphere_name_list)
    assert len(k.get_radius_list()) == 76 and min(k.get_radius_list()) > 0
    assert FetchConfig().get_collision_kin().n_feature == 51
    assert JaxonConfig().get_collision_kin().n_feature == 85
    ef = JaxonConfig().get_endeffector_kin()
    assert (ef.n_feature, ef.dim_cspace, ef.dim_tspace, ef.rot_type) == (4, 37, 7, RotationType.XYZW)
    ef.update_rotation_type(RotationType.RPY)
    assert ef.dim_tspace == 6


def test_box_and_configpoint_constraints_numpy_and_torch():
    import torch

    box = PR2Config().get_box_const()
    assert box.lb.shape == (7,) and (box.ub > box.lb).all()
    q = box.sample()
    v, j = box.evaluate_single(q, True)
    assert v.shape == (14,) and j.shape == (14, 7) and (v >= 0).all() and box.is_valid(q)
    assert not box.is_valid(box.ub + 0.1)
    vt, jt = box.evaluate(torch.as_tensor(q[None]), True)
    assert np.allclose(vt.numpy(), v[None]) and np.allclose(jt.numpy(), j[None])
    _, dummy = box.evaluate(q[None], False)
    assert dummy.shape == (1, 1) and np.isnan(dummy).all()        # AbstractConst.dummy_jacobian, constraint.py:55-56
    cp = ConfigPointConst(np.arange(7.0))
    v, j = cp.evaluate(np.ones((3, 7)), True)
    assert np.allclose(v, 1.0 - np.arange(7.0)) and np.allclose(j, np.eye(7))
    assert cp.is_equality() and not box.is_equality()


def test_composite_dedupes_by_id_and_checks_kind():
    b1 = PR2Config().get_box_const()
    b2 = PR2Config().get_box_const()
    comp = IneqCompositeConst([b1, b2, b1])
    assert [c.id_value for c in comp.const_list] == [b1.id_value, b2.id_value]     # constraint.py:153-158
    v, j = comp.evaluate(np.zeros((2, 7)), True)
    assert v.shape == (2, 28) and j.shape == (2, 28, 7)
    with pytest.raises(AssertionError):
        EqCompositeConst([b1])
    e = EqCompositeConst([ConfigPointConst(np.zeros(7)), ConfigPointConst(np.ones(7))])
    assert e.evaluate(np.zeros((1, 7)), False)[0].shape == (1, 14)
    assert isinstance(e, AbstractEqConst)


def test_evaluate_before_reflect_raises_runtime_error():
    """skmp/constraint.py:42-46."""
    pr2 = PR2(); pr2.reset_manip_pose()
    const = CollFreeConst(PR2Config().get_collision_kin(), Box([0.7, 0.5, 1.2]).sdf, pr2)
    const.reflect_robot_flag = False
    with pytest.raises(RuntimeError):
        const.evaluate(np.zeros((1, 7)), True)
    with pytest.raises(RuntimeError):
        const.is_valid_batch(np.zeros((1, 7)))


def test_pose_constraint_target_encoding():
    """PoseConstraint.from_skrobot_coords: [pos, flip(ypr)] for RPY, [pos, xyzw] for XYZW (constraint.py:474-498)."""
    from skmp_b200.robot_model import Coordinates
    from skmp_b200.rotmath import rpy_matrix

    pr2 = PR2(); pr2.reset_manip_pose()
    co = Coordinates([0.7, -0.6, 1.0], rpy_matrix(0.3, -0.2, 0.1))      # yaw, pitch, roll
    c = PoseConstraint.from_skrobot_coords([co], PR2Config().get_endeffector_kin(RotationType.RPY), pr2)
    assert np.allclose(c.get_description(), [0.7, -0.6, 1.0, 0.1, -0.2, 0.3])
    c = PoseConstraint.from_skrobot_coords([co], PR2Config().get_endeffector_kin(RotationType.XYZW), pr2)
    d = c.get_description()
    assert d.shape == (7,) and np.linalg.norm(d[3:]) == pytest.approx(1.0)
    c = PoseConstraint.from_skrobot_coords([co], PR2Config().get_endeffector_kin(RotationType.IGNORE), pr2)
    assert np.allclose(c.get_description(), [0.7, -0.6, 1.0])


def test_fetch_reachability_box_and_body_obstacles():
    conf = FetchConfig()
    lb, ub = conf.get_reachability_box()
    assert (ub > lb).all()
    assert len(conf.get_self_body_obstacles()) >= 2
    f = Fetch(); f.reset_pose()


def test_satisfy_shortcuts_and_no_cpu_fallback():
    """skmp/satisfy.py:55-56: a ConfigPointConst equality is answered without optimisation; anything else needs the GPU."""
    import torch

    from skmp_b200.constraint import BoxConst, ConfigPointConst
    from skmp_b200.satisfy import SatisfactionConfig, satisfy_by_optimization, satisfy_by_optimization_batch

    box = BoxConst(-np.ones(3), np.ones(3))
    eq = ConfigPointConst(np.array([0.1, 0.2, 0.3]))
    res = satisfy_by_optimization(eq, box, None, None)
    assert res.success and np.allclose(res.q, [0.1, 0.2, 0.3])
    cfg = SatisfactionConfig()
    assert (cfg.ftol, cfg.n_max_eval, cfg.acceptable_error) == (1e-6, 200, 1e-6)      # the reference's defaults (:29-34)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            satisfy_by_optimization_batch(None, box, None, np.zeros((2, 3)))


def test_skrobot_coordinate_api_of_primitives_and_targets():
    """The slice of skrobot's Coordinates API the reference's scenes use on primitives and pose targets
    (example/humanoid_dualarm.py:31-42,63; constrained_dualarm_planning.py:47-60): translate (local / world), rotate by an
    axis name, rotate_with_matrix in the world frame leaves the position alone, newcoords takes a coordinates object or
    (rotation, position)."""
    from skmp_b200.robot_model import Coordinates, PR2
    from skmp_b200.rotmath import rotation_matrix
    from skmp_b200.sdf import Box

    Rz = rotation_matrix(np.pi * 0.5, [0.0, 0.0, 1.0])
    table = Box([0.6, 1.0, 0.8])
    table.rotate(np.pi * 0.5, "z")
    table.translate([0.7, 0.0, 0.4])                       # local frame: x is now the world's y
    assert np.allclose(table.worldrot(), Rz) and np.allclose(table.worldpos(), Rz @ [0.7, 0.0, 0.4])
    table.translate([0.0, 0.0, 0.1], wrt="world")
    assert np.allclose(table.worldpos(), Rz @ [0.7, 0.0, 0.4] + [0.0, 0.0, 0.1])
    pos = table.worldpos()
    table.rotate_with_matrix(Rz, wrt="world")
    assert np.allclose(table.worldpos(), pos) and np.allclose(table.worldrot(), Rz @ Rz)
    co = Coordinates([0.6, -0.25, 0.25]).rotate(np.pi * 0.5, "z")
    assert np.allclose(co.worldrot(), Rz) and np.allclose(co.worldpos(), [0.6, -0.25, 0.25])
    other = Box([0.1, 0.1, 0.1])
    other.newcoords(co.copy_worldcoords())
    assert np.allclose(other.worldpos(), co.worldpos()) and np.allclose(other.worldrot(), Rz)
    other.newcoords(np.eye(3), [1.0, 2.0, 3.0])            # skrobot's (rotation, position)
    assert np.allclose(other.worldrot(), np.eye(3)) and np.allclose(other.worldpos(), [1.0, 2.0, 3.0])
    prims = other.primitives()
    assert prims[0][0] == "box" and np.allclose(prims[0][1], [1.0, 2.0, 3.0])
    robot = PR2()
    robot.newcoords(co)
    assert np.allclose(robot.translation, co.worldpos()) and np.allclose(robot.rotation, Rz)
    robot.newcoords(Rz.T, [0.1, 0.2, 0.0])
    assert np.allclose(robot.rotation, Rz.T) and np.allclose(robot.translation, [0.1, 0.2, 0.0])
    robot.newcoords([0.3, -0.2, 0.0])                      # position only (this package's earlier call shape)
    assert np.allclose(robot.translation, [0.3, -0.2, 0.0]) and np.allclose(robot.rotation, np.eye(3))


def test_jaxon_close_box_const():
    """skmp/robot/jaxon.py:346-372."""
    from skmp_b200.robot.jaxon import JaxonConfig

    conf = JaxonConfig()
    full = conf.get_box_const()
    assert conf.get_close_box_const().lb.shape == (37,)
    q = 0.5 * (full.lb + full.ub)
    q[-6:-3] = [5.0, -5.0, 9.0]                                  # base position far outside the global box: not clipped
    close = conf.get_close_box_const(q, joint_margin=0.3, base_pos_margin=0.8, base_rot_margin=10.0)
    assert np.allclose(close.ub[:-6], np.minimum(q[:-6] + 0.3, full.ub[:-6])) and np.allclose(close.lb[:-6], np.maximum(q[:-6] - 0.3, full.lb[:-6]))
    assert np.allclose(close.lb[-6:-3], q[-6:-3] - 0.8) and np.allclose(close.ub[-6:-3], q[-6:-3] + 0.8)
    assert np.allclose(close.lb[-3:], full.lb[-3:]) and np.allclose(close.ub[-3:], full.ub[-3:])     # clipped to the global box
    assert close.is_valid(q)

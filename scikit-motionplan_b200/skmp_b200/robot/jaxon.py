"""JAXON configuration factory -- mirror of skmp/robot/jaxon.py (kinematics / sphere parts).
The URDF is the bundled structure-faithful stand-in (robot_descriptions is not available offline)."""
import uuid
from dataclasses import dataclass
from pathlib import Path
from typing import Dict, List, Optional

import numpy as np

from ..collision import SphereCollection
from ..constraint import BoxConst, COMStabilityConst
from ..kinematics import AttachedObstacleCollPointsKinematicsMap, CollSphereKinematicsMap, EndEffectorKinematicsMap
from ..robot_model import bundled_urdf
from ..rotmath import rotation_matrix, rpy_angle
from ..sdf import Box
from ..tinyfk import BaseType, KinematicModel, RotationType

# sphere table of skmp/robot/jaxon.py:143-327 (reference input data, values verbatim): link -> [x, y, z, r]
_SOLE = lambda s: [[-0.08, 0.035 * s, -0.075, 0.04], [-0.08, -0.05 * s, -0.075, 0.04], [0.105, 0.035 * s, -0.075, 0.04],
                   [0.105, -0.05 * s, -0.075, 0.04], [-0.0, 0.0, -0.075, 0.04], [0.0, 0.0, 0.0, 0.07]]
_LEG3 = lambda s: [[0.0, -0.015 * s, 0.0, 0.09], [0.02, -0.015 * s, -0.07, 0.09], [0.02, -0.015 * s, -0.14, 0.09],
                   [0.02, -0.015 * s, -0.21, 0.09], [0.0, -0.015 * s, -0.3, 0.08], [0.0, -0.015 * s, -0.35, 0.08]]
_LEG2 = lambda s: [[-0.02, -0.01 * s, 0.0, 0.12], [-0.02, -0.01 * s, -0.08, 0.12], [0.02, -0.01 * s, -0.16, 0.12],
                   [0.02, -0.02 * s, -0.24, 0.1], [0.02, -0.02 * s, -0.32, 0.1]]
_FINGER0 = [[0.04, 0.01, 0.0, 0.02], [0.06, 0.02, 0.0, 0.02], [0.1, 0.02, 0.0, 0.02], [0.14, 0.01, 0.0, 0.02]]
_FINGER1 = [[0.04, -0.01, 0.02, 0.02], [0.06, -0.02, 0.02, 0.02], [0.1, -0.02, 0.02, 0.02], [0.14, -0.01, 0.02, 0.02],
            [0.04, -0.01, -0.02, 0.02], [0.06, -0.02, -0.02, 0.02], [0.1, -0.02, -0.02, 0.02], [0.14, -0.01, -0.02, 0.02]]
_LINK7 = [[0.0, 0.04, -0.16, 0.06], [0.0, -0.04, -0.16, 0.06], [0.0, 0.0, -0.10, 0.06], [0.03, 0.0, -0.15, 0.06]]
_LINK5 = [[0.0, 0.0, 0.14, 0.08], [0.0, 0.0, 0.04, 0.08], [0.0, 0.0, -0.06, 0.09], [0.0, 0.0, -0.16, 0.09]]
_LINK3 = [[0.0, 0.0, 0.16, 0.08], [0.0, 0.0, 0.06, 0.08], [0.0, 0.0, -0.04, 0.08]]
_CHEST = [[-0.06, 0.0, -0.05, 0.25], [-0.17, 0.0, -0.03, 0.25], [-0.08, 0.0, -0.27, 0.25]]
_HEAD = [[0.0, 0.0, 0.1, 0.15], [0.0, 0.0, 0.2, 0.15]]


def _collection(rows) -> SphereCollection:
    arr = np.asarray(rows, dtype=np.float64)
    return SphereCollection([r[:3].copy() for r in arr], [float(r) for r in arr[:, 3]], [str(uuid.uuid4()) for _ in rows])


@dataclass
class JaxonConfig:
    @classmethod
    def urdf_path(cls) -> Path:
        return bundled_urdf("jaxon")

    @staticmethod
    def add_end_coords(kin: KinematicModel) -> None:
        rarm_id, larm_id = kin.get_link_ids(["RARM_LINK7", "LARM_LINK7"])
        rpy = np.flip(rpy_angle(rotation_matrix(np.pi * 0.5, [0, 0, 1.0]))[0])
        kin.add_new_link("rarm_end_coords", rarm_id, np.array([0, 0, -0.220]), rpy=rpy)
        kin.add_new_link("larm_end_coords", larm_id, np.array([0, 0, -0.220]), rpy=rpy)
        kin.add_new_link("rarm_tip_coords", rarm_id, np.array([0, 0, -0.30]), rpy=rpy)
        rleg_id, lleg_id = kin.get_link_ids(["RLEG_LINK5", "LLEG_LINK5"])
        kin.add_new_link("rleg_end_coords", rleg_id, np.array([0, 0, -0.1]))
        kin.add_new_link("lleg_end_coords", lleg_id, np.array([0, 0, -0.1]))

    def _get_control_joint_names(self) -> List[str]:
        names = []
        for limb, n in (("ARM", 8), ("LEG", 6)):
            for i in range(n):
                names += [f"R{limb}_JOINT{i}", f"L{limb}_JOINT{i}"]
        return names + [f"CHEST_JOINT{i}" for i in range(3)]

    def get_endeffector_kin(self, rleg: bool = True, lleg: bool = True, rarm: bool = True, larm: bool = True,
                            rot_type: RotationType = RotationType.XYZW):
        names = [n for flag, n in ((rleg, "rleg_end_coords"), (lleg, "lleg_end_coords"), (rarm, "rarm_end_coords"),
                                   (larm, "larm_end_coords")) if flag]
        return EndEffectorKinematicsMap(self.urdf_path(), self._get_control_joint_names(), names, base_type=BaseType.FLOATING,
                                        rot_type=rot_type, fksolver_init_hook=self.add_end_coords)

    def get_attached_obstacle_kin(self, relative_position: np.ndarray, shape: Box) -> AttachedObstacleCollPointsKinematicsMap:
        return AttachedObstacleCollPointsKinematicsMap(self.urdf_path(), self._get_control_joint_names(), "rarm_end_coords",
                                                       relative_position, shape, base_type=BaseType.FLOATING,
                                                       fksolver_init_hook=self.add_end_coords)

    def get_collision_kin(self, rsole: bool = True, lsole: bool = True, rgripper: bool = True,
                          lgripper: bool = True) -> CollSphereKinematicsMap:
        table: Dict[str, SphereCollection] = {}
        if rsole:
            table["RLEG_LINK5"] = _collection(_SOLE(1.0))
        if lsole:
            table["LLEG_LINK5"] = _collection(_SOLE(-1.0))
        table["RLEG_LINK3"] = _collection(_LEG3(1.0))
        table["LLEG_LINK3"] = _collection(_LEG3(-1.0))
        table["RLEG_LINK2"] = _collection(_LEG2(1.0))
        table["LLEG_LINK2"] = _collection(_LEG2(-1.0))
        for flag, side in ((rgripper, "R"), (lgripper, "L")):
            if flag:
                table[f"{side}ARM_FINGER0"] = _collection(_FINGER0)
                table[f"{side}ARM_FINGER1"] = _collection(_FINGER1)
                table[f"{side}ARM_LINK7"] = _collection(_LINK7)
        for link, rows in (("RARM_LINK5", _LINK5), ("LARM_LINK5", _LINK5), ("RARM_LINK3", _LINK3), ("LARM_LINK3", _LINK3),
                           ("CHEST_LINK2", _CHEST), ("HEAD_LINK0", _HEAD)):
            table[link] = _collection(rows)
        return CollSphereKinematicsMap(self.urdf_path(), self._get_control_joint_names(), table, base_type=BaseType.FLOATING)

    def get_box_const(self) -> BoxConst:
        base_bounds = np.array([-1.0, -1.0, 0.0, -1.0, -1.0, -1.0]), np.array([2.0, 1.0, 3.0, 1.0, 1.0, 1.0])
        return BoxConst.from_urdf(self.urdf_path(), self._get_control_joint_names(), base_bounds=base_bounds)

    def get_close_box_const(self, q: Optional[np.ndarray] = None, joint_margin: float = 1.0, base_pos_margin: float = 0.8,
                            base_rot_margin: float = 1.0) -> BoxConst:
        """skmp/robot/jaxon.py:346-372: the joint box shrunk to a neighbourhood of ``q`` (joints and base rotation clipped to
        the global box, base position not)."""
        box_const = self.get_box_const()
        if q is None:
            return box_const
        q = np.asarray(q, dtype=np.float64)
        q_max, q_min = np.zeros(q.shape), np.zeros(q.shape)
        for sl, margin, do_clip in ((slice(None, -6), joint_margin, True), (slice(-6, -3), base_pos_margin, False),
                                    (slice(-3, None), base_rot_margin, True)):
            if do_clip:
                q_max[sl] = np.minimum(q[sl] + margin, box_const.ub[sl])
                q_min[sl] = np.maximum(q[sl] - margin, box_const.lb[sl])
            else:
                q_max[sl] = q[sl] + margin
                q_min[sl] = q[sl] - margin
        return BoxConst(q_min, q_max)

    def get_motion_step_box(self) -> np.ndarray:
        width = {jn: 0.2 for jn in self._get_control_joint_names()}
        for limb in ("RARM", "LARM", "RLEG", "LLEG", "CHEST"):
            for i in range(3):
                width[f"{limb}_JOINT{i}"] = 0.1
        return np.hstack([np.array(list(width.values())), np.ones(6) * 0.1])

    def get_com_stability_const(self, robot_model, com_box: Box, action_link_names: Optional[List[str]] = None,
                                action_forces: Optional[List[float]] = None) -> COMStabilityConst:
        """skmp/robot/jaxon.py:407-426"""
        return COMStabilityConst(self.urdf_path(), self._get_control_joint_names(), BaseType.FLOATING, robot_model, com_box,
                                 action_link_names, action_forces)

"""Fetch configuration factory -- mirror of skmp/robot/fetch.py (sphere/pose parts; the FCL mesh
self-collision constraint of :111-125 is out of scope)."""
from dataclasses import dataclass
from pathlib import Path
from typing import List, Tuple

import numpy as np

from ..constraint import BoxConst, PairWiseSelfCollFreeConst
from ..kinematics import CollSphereKinematicsMap, EndEffectorKinematicsMap
from ..robot_model import resolve_urdf
from ..sdf import Box, Cylinder
from ..tinyfk import BaseType, RotationType
from .utils import load_collision_spheres, sphere_table_path


@dataclass
class FetchConfig:
    base_type: BaseType = BaseType.FIXED
    use_torso: bool = True

    @classmethod
    def urdf_path(cls) -> Path:
        return resolve_urdf("~/.skrobot/fetch_description/fetch.urdf", "fetch")

    def get_control_joint_names(self) -> List[str]:
        names = ["shoulder_pan_joint", "shoulder_lift_joint", "upperarm_roll_joint", "elbow_flex_joint",
                 "forearm_roll_joint", "wrist_flex_joint", "wrist_roll_joint"]
        return (["torso_lift_joint"] + names) if self.use_torso else names

    def get_box_const(self, eps: float = 1e-4) -> BoxConst:
        bounds = BoxConst.from_urdf(self.urdf_path(), self.get_control_joint_names())
        bounds.lb -= eps
        bounds.ub += eps
        return bounds

    def get_endeffector_kin(self, rot_type: RotationType = RotationType.RPY) -> EndEffectorKinematicsMap:
        return EndEffectorKinematicsMap(self.urdf_path(), self.get_control_joint_names(), ["gripper_link"],
                                        base_type=self.base_type, rot_type=rot_type)

    def get_collision_kin(self) -> CollSphereKinematicsMap:
        spheres = load_collision_spheres(sphere_table_path("fetch"))
        return CollSphereKinematicsMap(self.urdf_path(), self.get_control_joint_names(), spheres, base_type=self.base_type)

    def get_self_body_obstacles(self) -> List:
        """Primitive stand-ins for the robot's own body (skmp/robot/fetch.py:74-93)."""
        spec = [
            (Cylinder(0.29, 0.32), [0.005, 0.0, 0.2]), (Box([0.16, 0.16, 1.0]), [-0.12, 0.0, 0.5]),
            (Box([0.1, 0.18, 0.08]), [0.0, 0.0, 0.97]), (Box([0.05, 0.17, 0.15]), [-0.035, 0.0, 0.92]),
            (Cylinder(0.1, 1.0), [-0.143, 0.09, 0.5]), (Cylinder(0.1, 1.0), [-0.143, -0.09, 0.5]),
            (Cylinder(0.235, 0.12), [0.0, 0.0, 1.04]),
        ]
        return [prim.translate(pos) for prim, pos in spec]

    def get_reachability_box(self, use_precomputed: bool = True) -> Tuple[np.ndarray, np.ndarray]:
        if use_precomputed:
            return np.array([-0.60046263, -1.08329689, -0.18025853]), np.array([1.10785484, 1.08329689, 2.12170273])
        # 8^8 = 16.7M position-only FK evaluations in one map() call (skmp/robot/fetch.py:100-109)
        box = self.get_box_const()
        axes = [np.linspace(a, b, 8) for a, b in zip(box.lb, box.ub)]
        grid = np.stack([m.ravel() for m in np.meshgrid(*axes)], axis=1).astype(np.float32)
        X, _ = self.get_endeffector_kin(RotationType.IGNORE).map(grid, with_jacobian=False)
        X = np.asarray(X).reshape(-1, 3)
        return X.min(axis=0), X.max(axis=0)

    def get_pairwise_selcol_consts(self, robot_model, only_closest_feature: bool = False) -> PairWiseSelfCollFreeConst:
        """Sphere-pair self-collision for BASELINE config 2 (pair list from the constructor's q=0 filter)."""
        return PairWiseSelfCollFreeConst(self.get_collision_kin(), robot_model, only_closest_feature=only_closest_feature)

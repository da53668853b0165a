"""PR2 configuration factory -- mirror of skmp/robot/pr2.py (names, defaults, selection rules)."""
import copy
from dataclasses import dataclass
from enum import Enum
from itertools import product
from pathlib import Path
from typing import Dict, List, Literal, Optional, Tuple

import numpy as np

from ..constraint import BoxConst, PairWiseSelfCollFreeConst
from ..kinematics import CollSphereKinematicsMap, EndEffectorKinematicsMap
from ..robot_model import resolve_urdf
from ..tinyfk import BaseType, RotationType
from ..urdf import parse_urdf
from .utils import load_collision_spheres, sphere_table_path


class CollisionMode(Enum):
    DEFAULT = 0
    RARM = 1
    LARM = 2
    RGRIPPER = 3
    LGRIPPER = 4


_ARM_JOINTS = ["shoulder_pan", "shoulder_lift", "upper_arm_roll", "elbow_flex", "forearm_roll", "wrist_flex", "wrist_roll"]
_ARM_COLL_LINKS = ["shoulder_pan_link", "shoulder_lift_link", "upper_arm_link", "forearm_link", "gripper_palm_link",
                   "gripper_r_finger_link", "gripper_l_finger_link"]
_STEP = {"shoulder_pan": 0.05, "shoulder_lift": 0.05, "upper_arm_roll": 0.1, "elbow_flex": 0.1,
         "forearm_roll": 0.2, "wrist_flex": 0.2, "wrist_roll": 0.2}


@dataclass
class PR2Config:
    control_arm: Literal["rarm", "larm", "dual"] = "rarm"
    collision_mode: CollisionMode = CollisionMode.DEFAULT
    selcol_mode: Literal["easy", "normal"] = "easy"
    base_type: BaseType = BaseType.FIXED
    use_torso: bool = False

    @classmethod
    def urdf_path(cls) -> Path:
        return resolve_urdf("~/.skrobot/pr2_description/pr2.urdf", "pr2")

    @classmethod
    def get_default_config_table(cls) -> Dict[str, float]:
        d = np.deg2rad
        return {
            "torso_lift_joint": 0.3,
            "l_shoulder_pan_joint": d(75), "l_shoulder_lift_joint": d(50), "l_upper_arm_roll_joint": d(110),
            "l_elbow_flex_joint": d(-110), "l_forearm_roll_joint": d(-20), "l_wrist_flex_joint": d(-10),
            "l_wrist_roll_joint": d(-10),
            "r_shoulder_pan_joint": d(-75), "r_shoulder_lift_joint": d(50), "r_upper_arm_roll_joint": d(-110),
            "r_elbow_flex_joint": d(-110), "r_forearm_roll_joint": d(20), "r_wrist_flex_joint": d(-10),
            "r_wrist_roll_joint": d(-10),
            "head_pan_joint": 0.0, "head_tilt_joint": d(50),
        }

    @classmethod
    def rarm_joint_names(cls) -> List[str]:
        return [f"r_{j}_joint" for j in _ARM_JOINTS]

    @classmethod
    def larm_joint_names(cls) -> List[str]:
        return [f"l_{j}_joint" for j in _ARM_JOINTS]

    @classmethod
    def rarm_collision_link_names(cls) -> List[str]:
        return [f"r_{l}" for l in _ARM_COLL_LINKS]

    @classmethod
    def larm_collision_link_names(cls) -> List[str]:
        return [f"l_{l}" for l in _ARM_COLL_LINKS]

    @classmethod
    def rgripper_collision_link_names(cls) -> List[str]:
        return [f"r_{l}" for l in _ARM_COLL_LINKS[4:]]

    @classmethod
    def lgripper_collision_link_names(cls) -> List[str]:
        return [f"l_{l}" for l in _ARM_COLL_LINKS[4:]]

    @classmethod
    def base_collision_link_names(cls) -> List[str]:
        return ["base_link"]

    def _get_control_joint_names(self) -> List[str]:
        return self.get_control_joint_names()

    def get_control_joint_names(self) -> List[str]:
        names = {"rarm": self.rarm_joint_names(), "larm": self.larm_joint_names(),
                 "dual": self.rarm_joint_names() + self.larm_joint_names()}[self.control_arm]
        if self.use_torso:
            names.append("torso_lift_joint")
        return names

    def _get_endeffector_names(self) -> List[str]:
        return {"rarm": ["r_gripper_tool_frame"], "larm": ["l_gripper_tool_frame"],
                "dual": ["r_gripper_tool_frame", "l_gripper_tool_frame"]}[self.control_arm]

    def get_endeffector_kin(self, rot_type: RotationType = RotationType.RPY) -> EndEffectorKinematicsMap:
        return EndEffectorKinematicsMap(self.urdf_path(), self.get_control_joint_names(), self._get_endeffector_names(),
                                        base_type=self.base_type, rot_type=rot_type)

    def get_default_motion_step_box(self) -> np.ndarray:
        box = []
        for name in self.get_control_joint_names():
            box.append(0.05 if name == "torso_lift_joint" else _STEP[name[2:-6]])
        box += [0.05] * {BaseType.FIXED: 0, BaseType.PLANER: 3, BaseType.FLOATING: 6}[self.base_type]
        return np.array(box)

    def get_box_const(self, base_bound: Optional[Tuple[np.ndarray, np.ndarray]] = None) -> BoxConst:
        if self.base_type == BaseType.PLANER:
            if base_bound is None:
                base_bound = np.array([-1.0, -2.0, -1.0]), np.array([2.0, 2.0, 1.0])
        elif self.base_type == BaseType.FLOATING:
            if base_bound is None:
                base_bound = -np.ones(6), np.ones(6)
        else:
            assert base_bound is None
        joint_map = parse_urdf(self.urdf_path()).joint_map
        lb, ub = [], []
        for name in self.get_control_joint_names():
            lim = joint_map[name].limit
            lb.append(-2 * np.pi if lim.lower is None else lim.lower)
            ub.append(2 * np.pi if lim.upper is None else lim.upper)
        if base_bound is not None:
            lb += list(base_bound[0])
            ub += list(base_bound[1])
        return BoxConst(np.array(lb), np.array(ub))

    def _get_collision_link_names(self):
        m = self.collision_mode
        if m == CollisionMode.DEFAULT:
            return self.rarm_collision_link_names() + self.larm_collision_link_names() + self.base_collision_link_names()
        return {CollisionMode.RARM: self.rarm_collision_link_names, CollisionMode.LARM: self.larm_collision_link_names,
                CollisionMode.RGRIPPER: self.rgripper_collision_link_names,
                CollisionMode.LGRIPPER: self.lgripper_collision_link_names}[m]()

    def get_collision_kin(self, whole_body: bool = False) -> CollSphereKinematicsMap:
        spheres = load_collision_spheres(sphere_table_path("pr2"))
        if self.base_type != BaseType.FIXED:
            whole_body = True
        if not whole_body:
            # spheres that cannot move relative to the obstacle are irrelevant
            drop = {"rarm": "l_", "larm": "r_"}.get(self.control_arm)
            for key in list(spheres):
                if (drop and key.startswith(drop)) or key.startswith("base_link"):
                    del spheres[key]
        return CollSphereKinematicsMap(self.urdf_path(), self.get_control_joint_names(), spheres, base_type=self.base_type)

    def get_pairwise_selcol_consts(self, robot_model) -> PairWiseSelfCollFreeConst:
        colkin = self.get_collision_kin(whole_body=True)
        names = colkin.sphere_name_list
        rarm = [n for n in names if "r_gripper" in n or "r_forearm" in n]
        larm = [n for n in names if "l_gripper" in n or "l_forearm" in n]
        if self.selcol_mode == "easy":
            shoulders = [n for n in names if n.startswith("l_shoulder_pan_link") or n.startswith("r_shoulder_pan_link")]
            base = [n for n in names if n.startswith("base_link")]
            anti_r = anti_l = shoulders + base
        elif self.selcol_mode == "normal":
            anti_r = [n for n in names if not n.startswith("r_")]
            anti_l = [n for n in names if not n.startswith("l_")]
        else:
            assert False
        table = dict(zip(names, colkin.tinyfk_feature_ids))
        pairs = []
        seen = set()
        groups = []
        if self.control_arm in ("rarm", "dual"):
            groups.append((rarm, anti_r))
        if self.control_arm in ("larm", "dual"):
            groups.append((larm, anti_l))
        for g, anti in groups:
            for n1, n2 in product(g, anti):
                a, b = table[n1], table[n2]
                key = (min(a, b), max(a, b))
                if key not in seen:
                    seen.add(key)
                    pairs.append(key)
        return PairWiseSelfCollFreeConst(colkin, robot_model, id_pairs=pairs, only_closest_feature=True)

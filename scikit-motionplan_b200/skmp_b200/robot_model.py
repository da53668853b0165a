"""Light stand-in for ``skrobot.model.RobotModel`` (scikit-robot is not installable offline).

The hot path only reads *state* from the robot model (skmp/kinematics.py:86-96): ``translation``,
``rotation``, ``joint_names`` and ``robot.__dict__[joint_name].joint_angle()``.  A real
``skrobot.model.RobotModel`` satisfies the same duck type and can be passed instead.
"""
from pathlib import Path
from typing import Dict, List, Optional

import numpy as np

from .rotmath import rpy_matrix
from .urdf import UrdfModel, parse_urdf

_URDF_DIR = Path(__file__).resolve().parent / "robot" / "urdf"


def bundled_urdf(name: str) -> Path:
    return _URDF_DIR / f"{name}.urdf"


def resolve_urdf(preferred: Path, bundled_name: str) -> Path:
    """The reference's on-disk location if it exists (real description), else the bundled stand-in."""
    p = Path(preferred).expanduser()
    return p if p.exists() else bundled_urdf(bundled_name)


class Joint:
    def __init__(self, name: str, type_name: str, lower: Optional[float], upper: Optional[float]):
        self.name = name
        self.type = type_name
        self.min_angle = -np.inf if lower is None else lower
        self.max_angle = np.inf if upper is None else upper
        self._angle = 0.0

    def joint_angle(self, v: Optional[float] = None) -> float:
        if v is not None:
            self._angle = float(v)
        return self._angle


class RobotModel:
    def __init__(self, urdf_file):
        self.urdf_path = Path(urdf_file)
        self.urdf: UrdfModel = parse_urdf(self.urdf_path)
        self.joint_list: List[Joint] = []
        for j in self.urdf.joints:
            if j.type_name == "fixed":
                continue
            joint = Joint(j.name, j.type_name, j.limit.lower, j.limit.upper)
            self.joint_list.append(joint)
            self.__dict__[j.name] = joint
        self.translation = np.zeros(3)
        self.rotation = np.eye(3)

    @property
    def joint_names(self) -> List[str]:
        return [j.name for j in self.joint_list]

    def newcoords(self, c, pos=None) -> None:
        """``robot.newcoords(c, pos=None)`` as in skrobot: ``c`` is a coordinates object, or a 3x3 rotation with ``pos`` the
        translation; a 3-vector first (and optionally a 3x3 rotation second) is understood as well."""
        if hasattr(c, "worldpos") and hasattr(c, "worldrot"):
            p, r = c.worldpos(), c.worldrot()
        else:
            a = np.asarray(c, dtype=np.float64)
            if a.shape == (3, 3):
                r, p = a, (np.zeros(3) if pos is None else pos)
            else:
                p, r = a, (np.eye(3) if pos is None else pos)
        self.translation = np.asarray(p, dtype=np.float64).reshape(3).copy()
        self.rotation = np.asarray(r, dtype=np.float64).reshape(3, 3).copy()

    def angle_vector(self, av=None) -> np.ndarray:
        if av is not None:
            for j, v in zip(self.joint_list, av):
                j.joint_angle(v)
        return np.array([j.joint_angle() for j in self.joint_list])

    def set_angles(self, table: Dict[str, float]) -> None:
        for name, v in table.items():
            self.__dict__[name].joint_angle(v)


class Coordinates:
    """``skrobot.coordinates.Coordinates(pos=, rot=)`` stand-in for pose targets (constraint.py:474-498) with the methods the
    reference's examples chain on it (translate / rotate / copy_worldcoords)."""

    def __init__(self, pos=None, rot=None):
        self._pos = np.zeros(3) if pos is None else np.asarray(pos, dtype=np.float64).copy()
        self._rot = np.eye(3) if rot is None else np.asarray(rot, dtype=np.float64).copy()

    def worldpos(self):
        return self._pos.copy()

    def worldrot(self):
        return self._rot.copy()

    def translate(self, vec, wrt: str = "local"):
        vec = np.asarray(vec, dtype=np.float64)
        self._pos = self._pos + (self._rot @ vec if wrt == "local" else vec)
        return self

    def rotate(self, theta: float, axis, wrt: str = "local"):
        from .rotmath import rotation_matrix

        ax = {"x": [1.0, 0.0, 0.0], "y": [0.0, 1.0, 0.0], "z": [0.0, 0.0, 1.0]}[axis] if isinstance(axis, str) else axis
        mat = rotation_matrix(float(theta), np.asarray(ax, dtype=np.float64))
        self._rot = self._rot @ mat if wrt == "local" else mat @ self._rot
        return self

    def copy_worldcoords(self):
        return Coordinates(self._pos, self._rot)


class PR2(RobotModel):
    def __init__(self):
        super().__init__(resolve_urdf("~/.skrobot/pr2_description/pr2.urdf", "pr2"))

    def reset_manip_pose(self):
        d = np.deg2rad
        self.set_angles({
            "torso_lift_joint": 0.3,
            "l_shoulder_pan_joint": d(75), "l_shoulder_lift_joint": d(50), "l_upper_arm_roll_joint": d(110),
            "l_elbow_flex_joint": d(-110), "l_forearm_roll_joint": d(-20), "l_wrist_flex_joint": d(-10),
            "l_wrist_roll_joint": d(-10),
            "r_shoulder_pan_joint": d(-75), "r_shoulder_lift_joint": d(50), "r_upper_arm_roll_joint": d(-110),
            "r_elbow_flex_joint": d(-110), "r_forearm_roll_joint": d(20), "r_wrist_flex_joint": d(-10),
            "r_wrist_roll_joint": d(-10),
            "head_pan_joint": 0.0, "head_tilt_joint": d(50),
        })
        return self.angle_vector()


class Fetch(RobotModel):
    def __init__(self):
        super().__init__(resolve_urdf("~/.skrobot/fetch_description/fetch.urdf", "fetch"))

    def reset_pose(self):
        d = np.deg2rad
        self.set_angles({
            "torso_lift_joint": 0.02, "shoulder_pan_joint": d(75.6304), "shoulder_lift_joint": d(80.2141),
            "upperarm_roll_joint": d(-11.4592), "elbow_flex_joint": d(98.5487), "forearm_roll_joint": 0.0,
            "wrist_flex_joint": d(95.111), "wrist_roll_joint": 0.0, "head_pan_joint": 0.0, "head_tilt_joint": 0.0,
        })
        return self.angle_vector()


class Jaxon(RobotModel):
    def __init__(self):
        super().__init__(bundled_urdf("jaxon"))

    def reset_manip_pose(self):
        table = {
            "RLEG": [0.0, 0.0, -0.349066, 0.698132, -0.349066, 0.0],
            "LLEG": [0.0, 0.0, -0.349066, 0.698132, -0.349066, 0.0],
            "CHEST": [0.0, 0.0, 0.0],
            "RARM": [0.0, 0.959931, -0.349066, -0.261799, -1.74533, -0.436332, 0.0, -0.785398],
            "LARM": [0.0, 0.959931, 0.349066, 0.261799, -1.74533, 0.436332, 0.0, -0.785398],
        }
        for key, values in table.items():
            for i, v in enumerate(values):
                self.__dict__[f"{key}_JOINT{i}"].joint_angle(v)
        return self.angle_vector()


def set_robot_state(robot_model, joint_names, angles, base_type=None) -> None:
    """skmp/robot/utils.py:39-59 -- q layout is joints then [x,y,theta] or [x,y,z,roll,pitch,yaw]."""
    from .tinyfk import BaseType

    base_type = BaseType.FIXED if base_type is None else base_type
    angles = np.asarray(angles, dtype=np.float64)
    if base_type == BaseType.PLANER:
        av, (x, y, th) = angles[:-3], angles[-3:]
        robot_model.newcoords([x, y, 0.0], rpy_matrix(th, 0.0, 0.0))
    elif base_type == BaseType.FLOATING:
        av, xyz, rpy = angles[:-6], angles[-6:-3], angles[-3:]
        robot_model.newcoords(xyz, rpy_matrix(rpy[2], rpy[1], rpy[0]))
    else:
        av = angles
    assert len(av) == len(joint_names)
    for name, v in zip(joint_names, av):
        robot_model.__dict__[name].joint_angle(v)


def get_robot_state(robot_model, joint_names, base_type=None) -> np.ndarray:
    """skmp/robot/utils.py:62-83"""
    from .rotmath import rpy_angle
    from .tinyfk import BaseType

    base_type = BaseType.FIXED if base_type is None else base_type
    av = np.array([robot_model.__dict__[n].joint_angle() for n in joint_names])
    if base_type == BaseType.PLANER:
        x, y, _ = robot_model.translation
        return np.hstack([av, [x, y, rpy_angle(robot_model.rotation)[0][0]]])
    if base_type == BaseType.FLOATING:
        return np.hstack([av, robot_model.translation, np.flip(rpy_angle(robot_model.rotation)[0])])
    return av

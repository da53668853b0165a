"""Kinematics maps: joint configurations -> feature points (+ Jacobians), on the GPU.

Mirror of skmp/kinematics.py (class and method names, argument meaning, shapes):
  KinematicsMapBase.map                     :43-63   qs (N, D) -> points (N, F, T), jac (N, F, T, D)
  KinematicsMapBase.reflect_skrobot_model   :86-96   bake all joint angles + base pose as sticky state
  KinematicsMapBase.add_new_feature_point   :98-114
  EndEffectorKinematicsMap                  :117-162
  AttachedObstacleCollPointsKinematicsMap   :171-231
  CollSphereKinematicsMap                   :234-300
``ArticulatedKinematicsMap`` (the name used by newer upstream versions) is an alias of
``KinematicsMapBase``.  Arrays may be numpy (host path, results numpy) or CUDA torch tensors
(resident path, results stay on the device).
"""
import uuid
from abc import ABC, abstractmethod
from pathlib import Path
from typing import Callable, Dict, List, Optional, Tuple, Union

import numpy as np

from .collision import SphereCollection
from .rotmath import rpy_angle
from .sdf import Box
from .tinyfk import BaseType, KinematicModel, RotationType, base_dim, tspace_dim


class KinematicsMapBase:
    dim_cspace: int
    dim_tspace: int
    n_feature: int
    fksolver: KinematicModel
    tinyfk_joint_ids: List[int]
    tinyfk_feature_ids: List[int]
    control_joint_names: List[str]
    _rot_type: RotationType
    _base_type: BaseType

    @property
    def rot_type(self) -> RotationType:
        return self._rot_type

    @property
    def base_type(self) -> BaseType:
        return self._base_type

    def map(self, points_cspace, with_jacobian: bool = True):
        n_point, n_dim_cspace = points_cspace.shape
        points, jacobians = self.fksolver.solve_fk(
            points_cspace, self.tinyfk_feature_ids, self.tinyfk_joint_ids,
            base_type=self._base_type, with_jacobian=with_jacobian, rot_type=self._rot_type,
        )
        return points, jacobians

    def _state_vector(self, robot_model) -> np.ndarray:
        av = np.array([robot_model.__dict__[n].joint_angle() for n in self.control_joint_names])
        if self._base_type == BaseType.PLANER:
            x, y, _ = robot_model.translation
            return np.hstack([av, [x, y, rpy_angle(robot_model.rotation)[0][0]]])
        if self._base_type == BaseType.FLOATING:
            return np.hstack([av, robot_model.translation, np.flip(rpy_angle(robot_model.rotation)[0])])
        return av

    def map_skrobot_model(self, robot_model, with_jacobian: bool = True):
        return self.map(np.expand_dims(self._state_vector(robot_model), axis=0), with_jacobian=with_jacobian)

    def reflect_skrobot_model(self, robot_model) -> None:
        """Copy the robot model's configuration (every joint + base pose) into the solver's sticky
        state; non-controlled joints and the base are folded into kernel constants from here on."""
        base_pose = np.hstack([robot_model.translation, np.flip(rpy_angle(robot_model.rotation)[0])])
        names = [n for n in robot_model.joint_names if n in self.fksolver._joint_index]
        angles = [robot_model.__dict__[n].joint_angle() for n in names]
        joint_ids = self.fksolver.get_joint_ids(names)
        self.fksolver.set_q(joint_ids, angles + base_pose.tolist(), base_type=BaseType.FLOATING)

    def add_new_feature_point(self, link_like: Union[str, int], position, rotation=None) -> None:
        parent_id = self.fksolver.get_link_ids([link_like])[0] if isinstance(link_like, str) else link_like
        name = "new_feature_{}".format(uuid.uuid4())
        self.fksolver.add_new_link(name, parent_id, position, rotation)
        self.tinyfk_feature_ids.append(self.fksolver.get_link_ids([name])[0])
        self.n_feature += 1


ArticulatedKinematicsMap = KinematicsMapBase


def _make_solver(urdfpath, hook) -> KinematicModel:
    solver = KinematicModel(str(Path(urdfpath).expanduser()))
    if hook is not None:
        hook(solver)
    return solver


class EndEffectorKinematicsMap(KinematicsMapBase):
    def __init__(self, urdfpath: Path, joint_names: List[str], end_effector_names: List[str],
                 base_type: BaseType = BaseType.FIXED, rot_type: RotationType = RotationType.RPY,
                 fksolver_init_hook: Optional[Callable[[KinematicModel], None]] = None):
        self.fksolver = _make_solver(urdfpath, fksolver_init_hook)
        self.tinyfk_feature_ids = self.fksolver.get_link_ids(end_effector_names)
        self.tinyfk_joint_ids = self.fksolver.get_joint_ids(joint_names)
        self.n_feature = len(self.tinyfk_feature_ids)
        self.dim_cspace = len(joint_names) + base_dim(base_type)
        self.control_joint_names = joint_names
        self._base_type = base_type
        self.update_rotation_type(rot_type)

    def update_rotation_type(self, rot_type: RotationType) -> None:
        self.dim_tspace = tspace_dim(rot_type)
        self._rot_type = rot_type


class CollSpheresKinematicsMapBase(KinematicsMapBase, ABC):
    @abstractmethod
    def get_radius_list(self) -> List[float]:
        ...


class AttachedObstacleCollPointsKinematicsMap(CollSpheresKinematicsMapBase):
    """Surface points of a box attached to a link (e.g. a grasped object); radius = ``margin``."""

    def __init__(self, urdfpath: Path, joint_names: List[str], attach_link_name: str, relative_position,
                 shape: Box, n_grid: int = 6, base_type: BaseType = BaseType.FIXED,
                 fksolver_init_hook: Optional[Callable[[KinematicModel], None]] = None, margin=0.0):
        self.fksolver = _make_solver(urdfpath, fksolver_init_hook)
        half = 0.5 * np.asarray(shape._extents)
        lin = [np.linspace(-half[i], half[i], n_grid) for i in range(3)]
        gx, gy, gz = np.meshgrid(lin[0], lin[1], lin[2])
        local = np.stack([gx.flatten(), gy.flatten(), gz.flatten()], axis=1)
        # keep lattice points within 1 cm of the surface: box SDF in the local frame is exact here
        s = np.abs(local) - half
        sd = np.linalg.norm(np.maximum(s, 0.0), axis=1) + np.minimum(np.max(s, axis=1), 0.0)
        world = shape.transform_vector(local)[sd > -1e-2]
        points_from_link = (world - shape.worldpos()) + np.asarray(relative_position)
        attach_id = self.fksolver.get_link_ids([attach_link_name])[0]
        names = []
        for pt in points_from_link:
            name = "pt_{}".format(uuid.uuid4())
            names.append(name)
            self.fksolver.add_new_link(name, attach_id, pt)
        self.dim_cspace = len(joint_names) + base_dim(base_type)
        self.dim_tspace = 3
        self.n_feature = len(names)
        self.radius_list = [margin] * len(names)
        self.tinyfk_joint_ids = self.fksolver.get_joint_ids(joint_names)
        self.control_joint_names = joint_names
        self._base_type = base_type
        self._rot_type = RotationType.IGNORE
        self.tinyfk_feature_ids = self.fksolver.get_link_ids(names)

    def get_radius_list(self) -> List[float]:
        return self.radius_list


class CollSphereKinematicsMap(CollSpheresKinematicsMapBase):
    radius_list: List[float]
    sphere_name_list: List[str]
    sphere_center_list: List[np.ndarray]

    def __init__(self, urdfpath: Path, joint_names: List[str],
                 link_wise_sphere_collection: Dict[str, SphereCollection],
                 base_type: BaseType = BaseType.FIXED,
                 fksolver_init_hook: Optional[Callable[[KinematicModel], None]] = None):
        self.fksolver = _make_solver(urdfpath, fksolver_init_hook)
        self.radius_list, self.sphere_name_list, self.sphere_center_list = [], [], []
        for link_name, spheres in link_wise_sphere_collection.items():
            link_id = self.fksolver.get_link_ids([link_name])[0]
            for center, radius, name in zip(spheres.center_list, spheres.radius_list, spheres.name_list):
                self.radius_list.append(float(radius))
                self.sphere_name_list.append(name)
                self.sphere_center_list.append(np.asarray(center, dtype=np.float64))
                self.fksolver.add_new_link(name, link_id, center)
        self.n_feature = len(self.radius_list)
        self.dim_cspace = len(joint_names) + base_dim(base_type)
        self.dim_tspace = 3
        self.tinyfk_joint_ids = self.fksolver.get_joint_ids(joint_names)
        self.tinyfk_feature_ids = self.fksolver.get_link_ids(self.sphere_name_list)
        self.control_joint_names = joint_names
        self._base_type = base_type
        self._rot_type = RotationType.IGNORE

    def get_radius_list(self) -> List[float]:
        return self.radius_list

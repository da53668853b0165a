"""skmp_b200 -- B200-native collision-constraint engine behind scikit-motionplan's Python API.

Modules mirror the reference package layout for the hot path only:
  skmp_b200.kinematics  <- skmp/kinematics.py      skmp_b200.constraint <- skmp/constraint.py
  skmp_b200.collision   <- skmp/collision.py       skmp_b200.robot.*    <- skmp/robot/*.py
  skmp_b200.tinyfk      <- the tinyfk surface the reference binds (C-ABI: include/skmp_b200.h)
  skmp_b200.sdf         <- skrobot Box/Cylinder/UnionSDF/GridSDF as GPU-consumable descriptors
"""
from ._array import config  # noqa: F401
from ._lib import LIB_PATH, SkbError  # noqa: F401
from .tinyfk import BaseType, KinematicModel, RotationType  # noqa: F401

__all__ = ["config", "BaseType", "RotationType", "KinematicModel", "SkbError", "LIB_PATH"]

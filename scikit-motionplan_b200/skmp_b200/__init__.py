"""skmp_b200 -- B200-native collision-constraint engine behind scikit-motionplan's Python API.

Modules mirror the reference package layout for the hot path only:
  skmp_b200.kinematics  <- skmp/kinematics.py      skmp_b200.constraint <- skmp/constraint.py
  skmp_b200.collision   <- skmp/collision.py       skmp_b200.robot.*    <- skmp/robot/*.py
  skmp_b200.tinyfk      <- the tinyfk surface the reference binds (C-ABI: include/skmp_b200.h)
  skmp_b200.sdf         <- skrobot Box/Cylinder/UnionSDF/GridSDF as GPU-consumable descriptors
  skmp_b200.satisfy     <- skmp/satisfy.py (all restarts as one batch on the GPU)
  skmp_b200.batched     <- batched callers: edge validity (solver/motion_step_box.py), SQP constraint stacking
  skmp_b200.sharded     <- row sharding of a batch over the GPUs of a box (torch.distributed, no data-path collective)
"""
from ._array import config  # noqa: F401
from ._lib import LIB_PATH, SkbError  # noqa: F401
from .tinyfk import BaseType, KinematicModel, RotationType  # noqa: F401

__all__ = ["config", "BaseType", "RotationType", "KinematicModel", "SkbError", "LIB_PATH"]

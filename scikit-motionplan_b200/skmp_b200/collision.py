"""Collision-sphere container (the one type of skmp/collision.py the hot path uses, :10-17).
Sphere *fitting* (create_sphere_collection, PCA over meshes) is offline preprocessing and out of scope."""
from dataclasses import dataclass
from typing import List

import numpy as np


@dataclass
class SphereCollection:
    center_list: List[np.ndarray]
    radius_list: List[float]
    name_list: List[str]

    def __len__(self) -> int:
        return len(self.center_list)

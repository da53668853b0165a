"""Sphere-table loading (skmp/robot/utils.py:19-36) and robot-state helpers (:39-84)."""
import json
import uuid
from pathlib import Path
from typing import Dict

import numpy as np

from ..collision import SphereCollection
from ..robot_model import get_robot_state, set_robot_state  # noqa: F401

_DATA = Path(__file__).resolve().parent


def load_collision_spheres(table_path) -> Dict[str, SphereCollection]:
    """link name -> spheres; sphere names are ``link_name + uuid4()[:13]`` so that name-prefix tests such
    as ``"r_gripper" in name`` (skmp/robot/pr2.py:301-317) keep working.  Accepts the bundled JSON
    tables or the reference's YAML layout."""
    path = Path(table_path)
    if path.suffix in (".yaml", ".yml"):
        import yaml

        with open(path, "r") as f:
            table = {k: v["spheres"] for k, v in yaml.safe_load(f)["collision_spheres"].items()}
    else:
        table = json.loads(path.read_text())["links"]
    out: Dict[str, SphereCollection] = {}
    for link_name, specs in table.items():
        arr = np.asarray(specs, dtype=np.float64)
        names = [link_name + str(uuid.uuid4())[:13] for _ in specs]
        out[link_name] = SphereCollection([row[:3].copy() for row in arr], [float(r) for r in arr[:, 3]], names)
    return out


def sphere_table_path(robot: str) -> Path:
    return _DATA / f"coll_spheres_{robot}.json"

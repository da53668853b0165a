"""Minimal URDF reader (stdlib xml.etree only; lxml / urdfpy / skrobot are absent).

Replaces the URDF handling the reference gets from tinyfk's C++ parser and skrobot's
``URDF.load`` (skmp/utils.py:13-32, skmp/constraint.py:218-233).  Produces a flat,
topologically sorted link table: every non-root link carries the joint that connects it to its
parent (type, origin xyz+rpy, axis, limits).
"""
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field
from pathlib import Path
from typing import Dict, List, Optional, Union

import numpy as np

JT_FIXED, JT_REVOLUTE, JT_PRISMATIC = 0, 1, 2
_TYPE = {"fixed": JT_FIXED, "revolute": JT_REVOLUTE, "continuous": JT_REVOLUTE, "prismatic": JT_PRISMATIC}


@dataclass
class JointLimit:
    lower: Optional[float] = None
    upper: Optional[float] = None


@dataclass
class UrdfJoint:
    name: str
    type_name: str
    parent: str
    child: str
    xyz: np.ndarray
    rpy: np.ndarray
    axis: np.ndarray
    limit: JointLimit

    @property
    def jtype(self) -> int:
        return _TYPE[self.type_name]


@dataclass
class UrdfLink:
    name: str
    mass: float = 0.0
    com: np.ndarray = field(default_factory=lambda: np.zeros(3))


@dataclass
class UrdfModel:
    name: str
    links: List[UrdfLink]          # topologically sorted, root first
    joints: List[UrdfJoint]        # in file order
    root: str

    @property
    def joint_map(self) -> Dict[str, UrdfJoint]:
        return {j.name: j for j in self.joints}

    @property
    def link_map(self) -> Dict[str, UrdfLink]:
        return {l.name: l for l in self.links}


def _vec(text: Optional[str], default) -> np.ndarray:
    if text is None:
        return np.array(default, dtype=np.float64)
    return np.array([float(v) for v in text.split()], dtype=np.float64)


def parse_urdf(source: Union[str, Path]) -> UrdfModel:
    """``source`` is a path or an XML string."""
    text = str(source)
    if not text.lstrip().startswith("<"):
        text = Path(text).expanduser().read_text()
    robot = ET.fromstring(text)
    links: Dict[str, UrdfLink] = {}
    for le in robot.findall("link"):
        link = UrdfLink(le.get("name"))
        ine = le.find("inertial")
        if ine is not None:
            me = ine.find("mass")
            if me is not None:
                link.mass = float(me.get("value", "0"))
            oe = ine.find("origin")
            if oe is not None:
                link.com = _vec(oe.get("xyz"), [0, 0, 0])
        links[link.name] = link
    joints: List[UrdfJoint] = []
    for je in robot.findall("joint"):
        tname = je.get("type")
        if tname not in _TYPE:
            if tname in ("floating", "planar"):
                raise ValueError(f"joint {je.get('name')}: URDF joint type '{tname}' is not supported")
            raise ValueError(f"joint {je.get('name')}: unknown joint type '{tname}'")
        oe, ae, lime = je.find("origin"), je.find("axis"), je.find("limit")
        axis = _vec(ae.get("xyz") if ae is not None else None, [1, 0, 0])
        nrm = np.linalg.norm(axis)
        if nrm > 0:
            axis = axis / nrm
        lim = JointLimit()
        if lime is not None and tname != "continuous":
            lo, up = lime.get("lower"), lime.get("upper")
            lim = JointLimit(float(lo) if lo is not None else None, float(up) if up is not None else None)
        joints.append(
            UrdfJoint(
                je.get("name"), tname, je.find("parent").get("link"), je.find("child").get("link"),
                _vec(oe.get("xyz") if oe is not None else None, [0, 0, 0]),
                _vec(oe.get("rpy") if oe is not None else None, [0, 0, 0]),
                axis, lim,
            )
        )
    children = {j.child for j in joints}
    roots = [n for n in links if n not in children]
    if len(roots) != 1:
        raise ValueError(f"URDF must have exactly one root link, found {roots}")
    # topological order (BFS from the root, stable in file order)
    by_parent: Dict[str, List[UrdfJoint]] = {}
    for j in joints:
        by_parent.setdefault(j.parent, []).append(j)
    order, queue = [], [roots[0]]
    while queue:
        n = queue.pop(0)
        order.append(links[n])
        queue.extend(j.child for j in by_parent.get(n, []))
    if len(order) != len(links):
        raise ValueError("URDF link graph is not a tree")
    return UrdfModel(robot.get("name", "robot"), order, joints, roots[0])

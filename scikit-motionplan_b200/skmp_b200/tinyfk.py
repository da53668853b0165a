"""Facade with the surface of ``tinyfk`` that the reference uses (tinyfk is an un-vendored C++
dependency of the reference; setup.py:11).  Call sites mirrored: KinematicModel(urdf)
kinematics.py:135; get_link_ids :106,139; get_joint_ids :95,142; add_new_link :110,211,280 and
robot/jaxon.py:78-84 (rpy= keyword); set_q :96; solve_fk :48-55; compute_inter_link_sqdists
constraint.py:715,752.

The arithmetic runs in the sm_100a kernels behind include/skmp_b200.h; this class only keeps
names <-> ids and the authoritative copy of the sticky state, and caches compiled plans.
"""
import copy
import ctypes as C
from enum import Enum
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._array import Batch, current_device_index, device_guard
from .urdf import JT_FIXED, UrdfModel, parse_urdf


class BaseType(Enum):
    FIXED = 0
    PLANER = 1
    FLOATING = 2


class RotationType(Enum):
    IGNORE = 0
    RPY = 1
    XYZW = 2


def base_dim(base_type: BaseType) -> int:
    return {BaseType.FIXED: 0, BaseType.PLANER: 3, BaseType.FLOATING: 6}[base_type]


def tspace_dim(rot_type: RotationType) -> int:
    return {RotationType.IGNORE: 3, RotationType.RPY: 6, RotationType.XYZW: 7}[rot_type]


class Plan:
    """Owns one ``skb_plan`` handle."""

    def __init__(self, handle: int, dim_cspace: int, n_feature: int, dim_tspace: int):
        self.handle = C.c_void_p(handle)
        self.dim_cspace, self.n_feature, self.dim_tspace = dim_cspace, n_feature, dim_tspace

    def __del__(self):
        try:
            if self.handle:
                _lib.lib().skb_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class KinematicModel:
    def __init__(self, urdf):
        """``urdf`` is a path to a URDF file or an XML string."""
        model: UrdfModel = urdf if isinstance(urdf, UrdfModel) else parse_urdf(urdf)
        self.urdf = model
        self.link_names: List[str] = [l.name for l in model.links]
        index = {n: i for i, n in enumerate(self.link_names)}
        n = len(self.link_names)
        self.parent = np.full(n, -1, dtype=np.int32)
        self.jtype = np.zeros(n, dtype=np.int32)
        self.origin = np.zeros((n, 6), dtype=np.float64)
        self.axis = np.tile(np.array([1.0, 0.0, 0.0]), (n, 1))
        self.angle = np.zeros(n, dtype=np.float64)
        self.base_pose = np.zeros(6, dtype=np.float64)
        self.mass = np.array([l.mass for l in model.links], dtype=np.float64)
        self.com = np.array([l.com for l in model.links], dtype=np.float64).reshape(n, 3)
        self.joint_names: List[str] = []
        self._joint_child: List[int] = []      # joint id -> child link index
        for j in model.joints:
            c = index[j.child]
            self.parent[c] = index[j.parent]
            self.jtype[c] = j.jtype
            self.origin[c, :3] = j.xyz
            self.origin[c, 3:] = j.rpy
            self.axis[c] = j.axis
            self.joint_names.append(j.name)
            self._joint_child.append(c)
        self._link_index: Dict[str, int] = index
        self._joint_index: Dict[str, int] = {n_: i for i, n_ in enumerate(self.joint_names)}
        self._version = 0
        self._handle = None
        self._handle_version = -1
        self._plans: Dict[tuple, Plan] = {}

    # ------------------------------------------------------------------ names <-> ids
    def get_link_ids(self, link_names: Sequence[str]) -> List[int]:
        try:
            return [self._link_index[n] for n in link_names]
        except KeyError as e:
            raise ValueError(f"no link named {e.args[0]!r}") from None

    def get_joint_ids(self, joint_names: Sequence[str]) -> List[int]:
        try:
            return [self._joint_index[n] for n in joint_names]
        except KeyError as e:
            raise ValueError(f"no joint named {e.args[0]!r}") from None

    def joint_child_links(self, joint_ids: Sequence[int]) -> np.ndarray:
        return np.array([self._joint_child[j] for j in joint_ids], dtype=np.int32)

    # ------------------------------------------------------------------ mutation
    def add_new_link(self, link_name: str, parent_id: int, position, rotation=None, rpy=None) -> None:
        if link_name in self._link_index:
            raise ValueError(f"link {link_name!r} already exists")
        if not (0 <= parent_id < len(self.link_names)):
            raise ValueError("bad parent link id")
        if rpy is None:
            rpy = rotation
        rpy = np.zeros(3) if rpy is None else np.asarray(rpy, dtype=np.float64)
        pos = np.asarray(position, dtype=np.float64)
        self.link_names.append(link_name)
        self._link_index[link_name] = len(self.link_names) - 1
        self.parent = np.append(self.parent, np.int32(parent_id))
        self.jtype = np.append(self.jtype, np.int32(JT_FIXED))
        self.origin = np.vstack([self.origin, np.hstack([pos, rpy])])
        self.axis = np.vstack([self.axis, [1.0, 0.0, 0.0]])
        self.angle = np.append(self.angle, 0.0)
        self.mass = np.append(self.mass, 0.0)
        self.com = np.vstack([self.com, np.zeros(3)])
        self._touch()

    def set_q(self, joint_ids: Sequence[int], q, base_type: BaseType = BaseType.FIXED) -> None:
        q = np.asarray(q, dtype=np.float64)
        nj = len(joint_ids)
        assert len(q) == nj + base_dim(base_type)
        for j, v in zip(joint_ids, q[:nj]):
            self.angle[self._joint_child[j]] = v
        if base_type == BaseType.PLANER:
            self.base_pose[:] = [q[nj], q[nj + 1], 0.0, 0.0, 0.0, q[nj + 2]]
        elif base_type == BaseType.FLOATING:
            self.base_pose[:] = q[nj:]
        self._touch()

    def _touch(self):
        self._version += 1
        self._plans.clear()

    # ------------------------------------------------------------------ native handles
    def native(self) -> C.c_void_p:
        """skb_model mirroring the current Python-side state (rebuilt after a mutation)."""
        if self._handle is None or self._handle_version != self._version:
            self._drop_handle()
            L = _lib.lib()
            h = C.c_void_p()
            parent = np.ascontiguousarray(self.parent, dtype=np.int32)
            jtype = np.ascontiguousarray(self.jtype, dtype=np.int32)
            origin = np.ascontiguousarray(self.origin, dtype=np.float64)
            axis = np.ascontiguousarray(self.axis, dtype=np.float64)
            _lib.check(L.skb_model_create(len(parent), _lib.dptr(parent, C.c_int32), _lib.dptr(jtype, C.c_int32),
                                          _lib.dptr(origin), _lib.dptr(axis), C.byref(h)))
            links = np.arange(1, len(parent), dtype=np.int32)
            q = np.ascontiguousarray(np.hstack([self.angle[1:], self.base_pose]))
            _lib.check(L.skb_model_set_q(h, _lib.dptr(links, C.c_int32), len(links), _lib.dptr(q), BaseType.FLOATING.value))
            self._handle, self._handle_version = h, self._version
        return self._handle

    def _drop_handle(self):
        if self._handle is not None:
            try:
                _lib.lib().skb_model_destroy(self._handle)
            except Exception:
                pass
            self._handle = None

    def __del__(self):
        self._plans.clear()
        self._drop_handle()

    def __deepcopy__(self, memo):
        new = object.__new__(KinematicModel)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k in ("_handle", "_plans"):
                continue
            setattr(new, k, copy.deepcopy(v, memo))
        new._handle, new._handle_version, new._plans = None, -1, {}
        return new

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_handle"], d["_handle_version"], d["_plans"] = None, -1, {}
        return d

    def get_plan(self, feature_ids: Sequence[int], joint_ids: Sequence[int], base_type: BaseType,
                 rot_type: RotationType, radii: Optional[Sequence[float]] = None,
                 pairs: Optional[np.ndarray] = None, pair_sqdists: Optional[np.ndarray] = None,
                 weights: Optional[Sequence[float]] = None, device_index: Optional[int] = None) -> Plan:
        """Compiled plan for (features, joints, base, rotation type[, radii / pairs / weights]) on ONE device: the program
        tables are uploaded to the device that is current when the plan is first used, so the cache is keyed by device."""
        dev = current_device_index() if device_index is None else int(device_index)
        key = (dev, tuple(feature_ids), tuple(joint_ids), base_type, rot_type,
               None if radii is None else tuple(float(r) for r in radii),
               None if pairs is None else (np.asarray(pairs).tobytes(), np.asarray(pair_sqdists).tobytes()),
               None if weights is None else tuple(float(w) for w in weights))
        plan = self._plans.get(key)
        if plan is None:
            L = _lib.lib()
            feats = np.asarray(feature_ids, dtype=np.int32)
            jl = self.joint_child_links(joint_ids)
            h = C.c_void_p()
            with device_guard(dev):
                _lib.check(L.skb_plan_create(self.native(), _lib.dptr(feats, C.c_int32), len(feats),
                                             _lib.dptr(jl, C.c_int32), len(jl), base_type.value, rot_type.value, C.byref(h)))
            plan = Plan(h.value, len(jl) + base_dim(base_type), len(feats), tspace_dim(rot_type))
            if radii is not None:
                r = np.ascontiguousarray(radii, dtype=np.float64)
                assert len(r) == len(feats)
                _lib.check(L.skb_plan_set_radii(plan.handle, _lib.dptr(r)))
            if pairs is not None:
                pr = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
                sq = np.ascontiguousarray(pair_sqdists, dtype=np.float64)
                _lib.check(L.skb_plan_set_pairs(plan.handle, _lib.dptr(pr, C.c_int32), len(pr), _lib.dptr(sq)))
            if weights is not None:
                w = np.ascontiguousarray(weights, dtype=np.float64)
                assert len(w) == len(feats)
                _lib.check(L.skb_plan_set_weights(plan.handle, _lib.dptr(w)))
            self._plans[key] = plan
        return plan

    # ------------------------------------------------------------------ tinyfk compute surface
    def solve_fk(self, qs, feature_ids, joint_ids, base_type: BaseType = BaseType.FIXED,
                 with_jacobian: bool = False, rot_type: RotationType = RotationType.IGNORE):
        """-> points (N, F, T), jacobians (N, F, T, D) (``np.empty(0)``-like when not requested)."""
        b = Batch(qs, len(joint_ids) + base_dim(base_type))
        plan = self.get_plan(feature_ids, joint_ids, base_type, rot_type, device_index=b.device_index)
        pts = b.empty((b.n, plan.n_feature, plan.dim_tspace))
        jac = b.empty((b.n, plan.n_feature, plan.dim_tspace, plan.dim_cspace)) if with_jacobian else None
        L = _lib.lib()
        with b.device_ctx():
            if b.on_device:
                _lib.check(L.skb_fk(plan.handle, b.dtype_code, b.q_ptr, b.n, int(with_jacobian), b.ptr(pts), b.ptr(jac), b.stream))
            else:
                _lib.check(L.skb_fk_host(plan.handle, b.dtype_code, b.q_ptr, b.n, int(with_jacobian), b.ptr(pts), b.ptr(jac)))
        return b.out(pts), (b.out(jac) if with_jacobian else b.empty0())

    def compute_inter_link_sqdists(self, qs, pairs, joint_ids, base_type: BaseType = BaseType.FIXED,
                                   with_jacobian: bool = False):
        """-> sqdists (N*P,), grads (N*P, D); pairs are (link_id, link_id) tuples."""
        pairs = [tuple(p) for p in pairs]
        feats = sorted({i for p in pairs for i in p})
        local = {f: k for k, f in enumerate(feats)}
        pr = np.array([[local[a], local[b]] for a, b in pairs], dtype=np.int32)
        b = Batch(qs, len(joint_ids) + base_dim(base_type))
        plan = self.get_plan(feats, joint_ids, base_type, RotationType.IGNORE, None, pr, np.zeros(len(pr)), device_index=b.device_index)
        P = len(pr)
        vals = b.empty((b.n, P))
        jac = b.empty((b.n, P, plan.dim_cspace)) if with_jacobian else None
        L = _lib.lib()
        with b.device_ctx():
            if b.on_device:
                _lib.check(L.skb_pairwise(plan.handle, b.dtype_code, b.q_ptr, b.n, _lib.MODE_ALL, b.ptr(vals), b.ptr(jac), None, b.stream))
            else:
                _lib.check(L.skb_pairwise_host(plan.handle, b.dtype_code, b.q_ptr, b.n, _lib.MODE_ALL, b.ptr(vals), b.ptr(jac), None))
        vals_o = b.out(vals).reshape(-1)
        return vals_o, (b.out(jac).reshape(b.n * P, plan.dim_cspace) if with_jacobian else b.empty0())

    # ------------------------------------------------------------------ centre of mass
    GRAVITY = 9.8      # action forces (newton) enter the weighted mean as force / GRAVITY (kg)

    def com_points(self, action_link_ids: Sequence[int] = (), action_forces: Sequence[float] = ()):
        """(feature link ids, weights) of the centre-of-mass plan: one fixed child link at the inertial origin of
        every link with mass > 0 (created on first use, named ``<link>_com``), plus the action links themselves
        weighted by force / g."""
        feats, weights = [], []
        for i, m in enumerate(list(self.mass)):
            if m <= 0:
                continue
            name = self.link_names[i] + "_com"
            if name not in self._link_index:
                self.add_new_link(name, i, self.com[i])
            feats.append(self._link_index[name])
            weights.append(float(m))
        for lid, f in zip(action_link_ids, action_forces):
            feats.append(int(lid))
            weights.append(float(f) / self.GRAVITY)
        if not feats:
            raise ValueError("the model has no link with a positive mass")
        return feats, weights

    def _com_call(self, qs, joint_ids, action_link_ids, action_forces, base_type, with_jacobian, what, sdf):
        feats, weights = self.com_points(action_link_ids, action_forces)
        b = Batch(qs, len(joint_ids) + base_dim(base_type))
        plan = self.get_plan(feats, joint_ids, base_type, RotationType.IGNORE, weights=weights, device_index=b.device_index)
        m = 3 if what == _lib.COM_POINT else 1
        vals = b.empty((b.n, m))
        jac = b.empty((b.n, m, plan.dim_cspace)) if with_jacobian else None
        L = _lib.lib()
        sdf_h = sdf.native(b.device_index) if sdf is not None else None
        with b.device_ctx():
            if b.on_device:
                _lib.check(L.skb_com(plan.handle, sdf_h, b.dtype_code, b.q_ptr, b.n, what, b.ptr(vals), b.ptr(jac), b.stream))
            else:
                _lib.check(L.skb_com_host(plan.handle, sdf_h, b.dtype_code, b.q_ptr, b.n, what, b.ptr(vals), b.ptr(jac)))
        return b, vals, jac

    def solve_com_fk(self, qs, joint_ids, action_link_ids=(), action_forces=(), base_type: BaseType = BaseType.FIXED,
                     with_jacobian: bool = False):
        """-> com (N, 3), jacobian (N*3, D) (call site skmp/constraint.py:901-909, which reshapes it to (N,1,3,D)).
        com = (sum_l m_l x_l + sum_a (f_a / g) x_a) / (sum_l m_l + sum_a f_a / g)."""
        b, vals, jac = self._com_call(qs, joint_ids, list(action_link_ids or ()), list(action_forces or ()), base_type,
                                      with_jacobian, _lib.COM_POINT, None)
        return b.out(vals), (b.out(jac).reshape(b.n * 3, -1) if with_jacobian else b.empty0())

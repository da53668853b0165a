"""Constraints evaluated on the GPU:  evaluate(qs) -> (values (N, M), jacobians (N, M, D)).

Mirror of the hot-path part of skmp/constraint.py (same class / method names, argument order,
shapes and error behaviour):
  AbstractConst.evaluate / evaluate_single / dummy_jacobian      :41-56
  AbstractIneqConst.is_valid                                     :89-91   all(value > 0)
  _CompositeConst / IneqCompositeConst / EqCompositeConst        :116-180 (is_valid: not any(value < 0))
  BoxConst                                                       :183-263
  CollFreeConst (all spheres / only closest)                     :266-378
  ConfigPointConst                                               :408-429
  PoseConstraint                                                 :432-505
  RelativePoseConstraint / FixedZAxisConstraint                  :508-607
  PairWiseSelfCollFreeConst                                      :681-782
Differences that are deliberate and documented in DESIGN.md: SDF gradients are analytic (the
reference uses forward differences with eps=1e-7, :320-331,:362-369); arrays may be numpy (host
path) or CUDA torch tensors (device-resident path); ``sdf`` must be an ``skmp_b200.sdf`` descriptor;
batched ``is_valid_batch`` is added for planners that validate many configurations at once.
Mesh / FCL / neural self-collision constraints (:610-678,:785-842) are out of scope (no Jacobian,
per-configuration Python loops, third-party engines).
"""
import copy
import itertools
import uuid
from abc import ABC, abstractmethod
from pathlib import Path
from typing import Generic, List, Optional, Protocol, Sequence, Tuple, TypeVar, Union, runtime_checkable

import numpy as np

from . import _lib
from ._array import Batch, _is_torch, current_device_index, torch
from .kinematics import CollSpheresKinematicsMapBase, EndEffectorKinematicsMap
from .rotmath import matrix2quaternion, rpy_angle, wxyz2xyzw
from .sdf import SignedDistanceFunction, as_descriptor
from .tinyfk import BaseType, RotationType
from .urdf import parse_urdf


_CONST_CACHE: dict = {}


def _to_kind(ref, arr: np.ndarray):
    """numpy constant -> same array kind / dtype / device as ``ref``.  Device copies of small constants (targets, bounds,
    identity blocks) are cached by content: no host-to-device copy per call, and nothing a CUDA-graph capture of the
    caller would have to record."""
    if _is_torch(ref):
        arr = np.ascontiguousarray(arr)
        if arr.nbytes <= 1 << 16:
            key = (arr.tobytes(), arr.shape, str(arr.dtype), ref.dtype, str(ref.device))
            t = _CONST_CACHE.get(key)
            if t is None:
                if len(_CONST_CACHE) > 256:
                    _CONST_CACHE.clear()
                t = _CONST_CACHE[key] = torch.as_tensor(arr, dtype=ref.dtype, device=ref.device)
            return t
        return torch.as_tensor(arr, dtype=ref.dtype, device=ref.device)
    return arr


def _cat(arrays, axis: int):
    if _is_torch(arrays[0]):
        return torch.cat(list(arrays), dim=axis)
    return np.concatenate(list(arrays), axis=axis)


class _SingleCall:
    """Pre-bound state of the planners' call shape -- ONE configuration as a 1-D numpy array, a blocking answer
    (skmp/solver/ompl_solver.py:74-80, motion_step_box.py:50-56, satisfy.py:67-82) -- so that a call is: copy q into a
    scratch row, one ``skb_*_host`` call with pre-built ctypes arguments (the library serves a batch this small through one
    mapped, pinned block: one launch, one stream synchronisation), read the scratch outputs.  One instance per
    (constraint, dtype, device, solver version)."""

    __slots__ = ("fn", "head", "q", "qp", "vals", "vp", "jac", "jp", "valid", "bp", "tail_all", "mode_closest_min", "m")

    def __init__(self, fn, head, dtype, d, m, tail_mode, closest_mode):
        import ctypes as C

        self.fn, self.head, self.m = fn, head, m
        self.q = np.zeros((1, d), dtype=dtype)
        self.vals = np.zeros((1, m), dtype=dtype)
        self.jac = np.zeros((1, m, d), dtype=dtype)
        self.valid = np.zeros(1, dtype=np.uint8)
        self.qp, self.vp, self.jp, self.bp = (C.c_void_p(a.ctypes.data) for a in (self.q, self.vals, self.jac, self.valid))
        self.tail_all = tail_mode               # (N, [margin,] mode) arguments between q and the outputs
        self.mode_closest_min = closest_mode

    def run(self, q, tail, vp, jp, bp):
        self.q[0] = q
        rc = self.fn(*self.head, self.qp, *tail, vp, jp, bp)
        if rc != 0:
            _lib.check(rc)


class AbstractConst(ABC):
    reflect_robot_flag: bool = False
    id_value: str

    def __getstate__(self):
        # native plan handles are per-object caches: never copied or pickled (copy.deepcopy / ParallelSolver forks)
        d = dict(self.__dict__)
        d.pop("_plan_cache", None)
        d.pop("_single_cache", None)
        return d

    def assign_id_value(self):
        self.id_value = str(uuid.uuid4())

    def evaluate(self, qs, with_jacobian: bool):
        if not self.reflect_robot_flag:
            raise RuntimeError("{}: you need to call reflect_skrobot_model beforehand".format(type(self).__name__))
        return self._evaluate(qs, with_jacobian)

    def evaluate_single(self, q, with_jacobian: bool):
        assert q.ndim == 1
        f, jac = self.evaluate(q[None, :], with_jacobian)
        return f[0], jac[0]

    def dummy_jacobian(self) -> np.ndarray:
        return np.array([[np.nan]])

    @abstractmethod
    def _evaluate(self, qs, with_jacobian: bool):
        ...

    def reflect_skrobot_model(self, robot_model) -> None:
        self._reflect_skrobot_model(robot_model)
        self.reflect_robot_flag = True

    @abstractmethod
    def _reflect_skrobot_model(self, robot_model) -> None:
        ...

    @classmethod
    @abstractmethod
    def is_equality(cls) -> bool:
        ...


class AbstractIneqConst(AbstractConst):
    @classmethod
    def is_equality(cls) -> bool:
        return False

    def is_valid(self, q) -> bool:
        value, _ = self.evaluate_single(q, False)
        return bool((value > 0).all())

    def is_valid_batch(self, qs):
        """all(values > 0) per configuration, for many configurations at once -> (N,) bool."""
        values, _ = self.evaluate(qs, False)
        return (values > 0).all(axis=1) if not _is_torch(values) else (values > 0).all(dim=1)


class AbstractEqConst(AbstractConst):
    @classmethod
    def is_equality(cls) -> bool:
        return True

    def is_approx_satisfied(self, q, eps: float = 0.01) -> bool:
        values, _ = self.evaluate_single(q, False)
        return bool((abs(values) < eps).all())


@runtime_checkable
class VectorDescriptable(Protocol):
    def get_description(self) -> np.ndarray:
        ...


InnerConstT = TypeVar("InnerConstT", bound=Union[AbstractIneqConst, AbstractEqConst])


class _CompositeConst(AbstractConst, Generic[InnerConstT]):
    const_list: List[InnerConstT]

    def __init__(self, const_list: List[InnerConstT]) -> None:
        for const in const_list:
            assert const.is_equality() == self.is_equality()
        # members already reflected the robot state; dedupe by id keeping first-seen order
        self.const_list = list({const.id_value: const for const in const_list}.values())
        self.reflect_robot_flag = True
        self.assign_id_value()

    def _evaluate(self, qs, with_jacobian: bool):
        parts = [const.evaluate(qs, with_jacobian=with_jacobian) for const in self.const_list]
        values = _cat([p[0] for p in parts], axis=1)
        if not with_jacobian:
            return values, self.dummy_jacobian()
        return values, _cat([p[1] for p in parts], axis=1)

    def _reflect_skrobot_model(self, robot_model) -> None:
        for const in self.const_list:
            const.reflect_skrobot_model(robot_model)


class IneqCompositeConst(AbstractIneqConst, _CompositeConst[AbstractIneqConst]):
    """Order matters: ``is_valid`` returns at the first violated member."""

    def is_valid(self, q) -> bool:
        lean = type(q) is np.ndarray and q.ndim == 1
        for const in self.const_list:
            if lean and hasattr(const, "_min_value_single") and const.reflect_robot_flag:
                if const._min_value_single(q) < 0.0:         # the collision kernels reduce to the row minimum themselves
                    return False
                continue
            values, _ = const.evaluate_single(q, False)
            if bool((values < 0.0).any()):
                return False
        return True

    def is_valid_batch(self, qs):
        """``is_valid`` for every row: invalid iff some member has a value < 0 (:167-176).  Like the reference's loop it
        short-circuits in ``const_list`` order -- rows already rejected are not shown to the later members -- and the
        collision members reduce to their row minimum on the GPU (one float per configuration leaves the kernel)."""
        def bad_rows(const, q):
            if hasattr(const, "_min_values"):
                m = const._min_values(q)
                return (m < 0.0).reshape(-1)
            values, _ = const.evaluate(q, False)
            return (values < 0.0).any(dim=1) if _is_torch(values) else (values < 0.0).any(axis=1)

        if _is_torch(qs):
            ok = torch.ones(qs.shape[0], dtype=torch.bool, device=qs.device)
            alive = None                                    # indices of the rows still valid (None = all)
            for const in self.const_list:
                sub = qs if alive is None else qs[alive]
                if sub.shape[0] == 0:
                    break
                bad = bad_rows(const, sub)
                bad = torch.as_tensor(bad, device=qs.device)
                if alive is None:
                    ok &= ~bad
                    alive = torch.nonzero(ok, as_tuple=False).reshape(-1)
                else:
                    ok[alive[bad]] = False
                    alive = alive[~bad]
            return ok
        qs = np.asarray(qs)
        ok = np.ones(len(qs), dtype=bool)
        for const in self.const_list:
            alive = np.nonzero(ok)[0]
            if len(alive) == 0:
                break
            bad = np.asarray(bad_rows(const, qs[alive]))
            ok[alive[bad]] = False
        return ok


class EqCompositeConst(AbstractEqConst, _CompositeConst[AbstractEqConst]):
    ...


class BoxConst(AbstractIneqConst):
    lb: np.ndarray
    ub: np.ndarray
    names: Optional[List[str]]

    def __init__(self, lb: np.ndarray, ub: np.ndarray, names: Optional[List[str]] = None) -> None:
        self.lb, self.ub = np.asarray(lb, dtype=np.float64), np.asarray(ub, dtype=np.float64)
        self.names = ["<unnamed>"] * len(self.lb) if names is None else names
        assert len(self.names) == len(self.lb)
        self.reflect_skrobot_model(None)
        self.assign_id_value()

    @classmethod
    def from_urdf(cls, urdf_path: Path, joint_names: List[str],
                  base_bounds: Optional[Tuple[np.ndarray, np.ndarray]] = None):
        joint_map = parse_urdf(Path(urdf_path).expanduser()).joint_map
        b_min, b_max, names = [], [], list(joint_names)
        for name in joint_names:
            lim = joint_map[name].limit
            b_min.append(-2 * np.pi if lim.lower is None or not np.isfinite(lim.lower) else lim.lower)
            b_max.append(2 * np.pi if lim.upper is None or not np.isfinite(lim.upper) else lim.upper)
        if base_bounds is not None:
            lb, ub = base_bounds
            assert len(lb) in (3, 6)
            names += ["x", "y", "z", "roll", "pitch", "yaw"][: len(lb)] if len(lb) == 6 else ["x", "y", "theta"]
            b_min += list(lb)
            b_max += list(ub)
        return cls(np.array(b_min), np.array(b_max), names)

    def _evaluate(self, qs, with_jacobian: bool):
        n_point, dim = qs.shape
        lb, ub = _to_kind(qs, self.lb), _to_kind(qs, self.ub)
        f = _cat([qs - lb, ub - qs], axis=1)
        if not with_jacobian:
            return f, self.dummy_jacobian()
        single = _to_kind(qs, np.vstack((np.eye(dim), -np.eye(dim))))
        jac = single[None].expand(n_point, -1, -1).clone() if _is_torch(qs) else np.tile(single, (n_point, 1, 1))
        return f, jac

    def _reflect_skrobot_model(self, robot_model) -> None:
        pass

    def sample(self) -> np.ndarray:
        w = self.ub - self.lb
        return np.random.rand(len(w)) * w + self.lb


class CollFreeConst(AbstractIneqConst):
    """values[n, s] = sdf(x_s(q_n)) - radius_s - distance_margin, one fused FK + SDF kernel launch."""

    colkin: CollSpheresKinematicsMapBase
    sdf: SignedDistanceFunction
    only_closest_feature: bool
    distance_margin: float

    def __init__(self, colkin: CollSpheresKinematicsMapBase, sdf, robot_model,
                 only_closest_feature: bool = False, distance_margin: float = 0.0) -> None:
        self.colkin = colkin
        self.sdf = as_descriptor(sdf)
        self.reflect_skrobot_model(robot_model)
        self.only_closest_feature = only_closest_feature
        self.distance_margin = distance_margin
        self.assign_id_value()

    def _plan(self, device_index: Optional[int] = None):
        # the plan is a snapshot of the solver's sticky state: rebuilt only after the solver changed (set_q / add_new_link);
        # one plan per device (its tables live in that device's memory)
        k = self.colkin
        dev = current_device_index() if device_index is None else device_index
        cache = self.__dict__.setdefault("_plan_cache", {})
        cached = cache.get(dev)
        if cached is not None and cached[0] is k.fksolver and cached[1] == k.fksolver._version:
            return cached[2]
        plan = k.fksolver.get_plan(k.tinyfk_feature_ids, k.tinyfk_joint_ids, k.base_type, RotationType.IGNORE,
                                   radii=k.get_radius_list(), device_index=dev)
        cache[dev] = (k.fksolver, k.fksolver._version, plan)
        return plan

    def _run(self, qs, want_vals: bool, with_jacobian: bool, want_valid: bool, only_closest: Optional[bool] = None):
        assert self.sdf is not None
        b = Batch(qs, self.colkin.dim_cspace)
        plan = self._plan(b.device_index)
        closest = self.only_closest_feature if only_closest is None else only_closest
        m = 1 if closest else plan.n_feature
        vals = b.empty((b.n, m)) if want_vals else None
        jac = b.empty((b.n, m, plan.dim_cspace)) if with_jacobian else None
        valid = b.empty_bytes(b.n) if want_valid else None
        mode = _lib.MODE_CLOSEST if closest else _lib.MODE_ALL
        L = _lib.lib()
        sdf_h = self.sdf.native(b.device_index)
        with b.device_ctx():
            if b.on_device:
                _lib.check(L.skb_collfree(plan.handle, sdf_h, b.dtype_code, b.q_ptr, b.n,
                                          float(self.distance_margin), mode, b.ptr(vals), b.ptr(jac), b.ptr(valid), b.stream))
            else:
                _lib.check(L.skb_collfree_host(plan.handle, sdf_h, b.dtype_code, b.q_ptr, b.n,
                                               float(self.distance_margin), mode, b.ptr(vals), b.ptr(jac), b.ptr(valid)))
        return b, vals, jac, valid

    def _evaluate(self, qs, with_jacobian: bool):
        b, vals, jac, _ = self._run(qs, True, with_jacobian, False)
        return b.out(vals), (b.out(jac) if with_jacobian else self.dummy_jacobian())

    # ---- one configuration per call (numpy 1-D in, numpy / bool out): the planners' call shape, kept lean
    def _single(self, q):
        dt = q.dtype.type if q.dtype.type in (np.float32, np.float64) else np.float64
        dev = current_device_index()
        key = (dt, dev)
        cache = self.__dict__.setdefault("_single_cache", {})
        sc = cache.get(key)
        solver = self.colkin.fksolver
        if sc is None or sc[1] != solver._version or sc[2] is not solver or sc[3] != (self.distance_margin, self.only_closest_feature):
            plan = self._plan(dev)
            code = _lib.F32 if dt == np.float32 else _lib.F64
            mode = _lib.MODE_CLOSEST if self.only_closest_feature else _lib.MODE_ALL
            m = 1 if self.only_closest_feature else plan.n_feature
            call = _SingleCall(_lib.lib().skb_collfree_host, (plan.handle, self.sdf.native(dev), code), dt, plan.dim_cspace, m,
                               (1, float(self.distance_margin), mode), (1, float(self.distance_margin), _lib.MODE_CLOSEST))
            sc = cache[key] = (call, solver._version, solver, (self.distance_margin, self.only_closest_feature), plan)
        return sc[0]

    def evaluate_single(self, q, with_jacobian: bool):
        if type(q) is not np.ndarray or q.ndim != 1 or not self.reflect_robot_flag:
            return AbstractConst.evaluate_single(self, q, with_jacobian)      # (no zero-argument super: shared with the pairwise class)
        sc = self._single(q)
        sc.run(q, sc.tail_all, sc.vp, sc.jp if with_jacobian else None, None)
        return sc.vals[0].copy(), (sc.jac[0].copy() if with_jacobian else self.dummy_jacobian()[0])

    def is_valid(self, q) -> bool:
        if type(q) is not np.ndarray or q.ndim != 1 or not self.reflect_robot_flag:
            return AbstractIneqConst.is_valid(self, q)
        sc = self._single(q)
        sc.run(q, sc.tail_all, None, None, sc.bp)          # the validity byte: all(values > 0), computed by the kernel
        return bool(sc.valid[0])

    def _min_value_single(self, q) -> float:
        """min over the rows for one configuration (what a composite's ``is_valid`` compares with 0)."""
        sc = self._single(q)
        sc.run(q, sc.mode_closest_min, sc.vp, None, None)
        return float(sc.vals[0, 0])

    def evaluate_csc(self, qs):
        """values (N, M) and the Jacobian block of every configuration TRANSPOSED, (N, D, M): the order in which the SQP
        solver's stacked, block-diagonal CSC Jacobian stores a waypoint's rows (one column per (waypoint, dof);
        skmp/solver/nlp_solver/trajectory_constraint.py:157-194).  For CUDA tensors in all-spheres mode with D <= 32 the
        kernel writes that layout directly (``skb_collfree_csc``: lane-per-row stores are coalesced in it, no staging);
        otherwise the (N, M, D) blocks are transposed afterwards.  Same numbers as ``evaluate`` bit for bit."""
        if not self.reflect_robot_flag:
            raise RuntimeError("{}: you need to call reflect_skrobot_model beforehand".format(type(self).__name__))
        b = Batch(qs, self.colkin.dim_cspace)
        plan = self._plan(b.device_index)
        if not b.on_device or self.only_closest_feature or plan.dim_cspace > 32:
            v, j = self._evaluate(qs, True)
            jt = j.transpose(1, 2).contiguous() if _is_torch(j) else np.ascontiguousarray(np.swapaxes(j, 1, 2))
            return v, jt
        vals = b.empty((b.n, plan.n_feature))
        data = b.empty((b.n, plan.dim_cspace, plan.n_feature))
        with b.device_ctx():
            _lib.check(_lib.lib().skb_collfree_csc(plan.handle, self.sdf.native(b.device_index), b.dtype_code, b.q_ptr, b.n,
                                                   float(self.distance_margin), b.ptr(vals), b.ptr(data), b.stream))
        return vals, data

    def _min_values(self, qs):
        """Row minimum of the values, (N, 1): the closest-mode kernel without Jacobian (what a composite needs)."""
        if not self.reflect_robot_flag:
            raise RuntimeError("{}: you need to call reflect_skrobot_model beforehand".format(type(self).__name__))
        b, vals, _, _ = self._run(qs, True, False, False, only_closest=True)      # the mode is an argument: re-entrant
        return b.out(vals)

    def is_valid_batch(self, qs):
        """Validity bytes straight from the kernel (1 byte per configuration leaves the GPU)."""
        if not self.reflect_robot_flag:
            raise RuntimeError("{}: you need to call reflect_skrobot_model beforehand".format(type(self).__name__))
        b, _, _, valid = self._run(qs, False, False, True)
        out = b.out(valid)
        return out.bool() if _is_torch(out) else out.astype(bool)

    def tune(self, qs, with_jacobian: bool = True, values: bool = True, validity: bool = False, csc: bool = False):
        """Measure the kernel launch configuration for this constraint's (mode, dtype, outputs) once, explicitly, on up to
        262 144 of the sample configurations ``qs`` (a CUDA tensor) and on scratch outputs of the library's own
        (``skb_plan_tune``).  The choice is keyed by the plan's structure: it survives ``reflect_skrobot_model`` and is
        shared by every structurally equal constraint; ``SKB_TUNE_CACHE=<file>`` persists it.  Calls of fewer than 2^20
        configurations never tune on their own.  Results do not depend on the choice.
        -> (unfused, G, warps) or None when the serving kernel has no launch candidates."""
        import ctypes as C

        if not self.reflect_robot_flag:
            raise RuntimeError("{}: you need to call reflect_skrobot_model beforehand".format(type(self).__name__))
        b = Batch(qs, self.colkin.dim_cspace)
        if not b.on_device:
            raise ValueError("tune() needs sample configurations on the GPU (a CUDA tensor)")
        plan = self._plan(b.device_index)
        outputs = (1 if values else 0) | (2 if with_jacobian else 0) | (4 if validity else 0) | (8 if csc else 0)   # csc: evaluate_csc's kernel
        mode = _lib.MODE_CLOSEST if self.only_closest_feature else _lib.MODE_ALL
        choice = (C.c_int32 * 3)()
        sdf = getattr(self, "sdf", None)
        with b.device_ctx():
            _lib.check(_lib.lib().skb_plan_tune(plan.handle, sdf.native(b.device_index) if sdf is not None else None, b.dtype_code,
                                                b.q_ptr, b.n, float(getattr(self, "distance_margin", 0.0)), mode, outputs, b.stream, choice))
        return None if choice[0] < 0 else (int(choice[0]), int(choice[1]), int(choice[2]))

    def tuned(self, dtype=np.float32, with_jacobian: bool = True, values: bool = True, validity: bool = False, device_index=None):
        """The stored launch choice for (mode, dtype, outputs), or None (``skb_plan_get_tuned``)."""
        import ctypes as C

        dev = current_device_index() if device_index is None else device_index
        plan = self._plan(dev)
        outputs = (1 if values else 0) | (2 if with_jacobian else 0) | (4 if validity else 0)
        mode = _lib.MODE_CLOSEST if self.only_closest_feature else _lib.MODE_ALL
        choice = (C.c_int32 * 3)()
        sdf = getattr(self, "sdf", None)
        from ._array import device_guard
        with device_guard(dev):
            _lib.check(_lib.lib().skb_plan_get_tuned(plan.handle, sdf.native(dev) if sdf is not None else None,
                                                     _lib.F32 if np.dtype(dtype) == np.float32 else _lib.F64, mode, outputs, choice))
        return None if choice[0] < 0 else (int(choice[0]), int(choice[1]), int(choice[2]))

    def _reflect_skrobot_model(self, robot_model) -> None:
        assert robot_model, "robot_model must not be None"
        self.colkin.reflect_skrobot_model(robot_model)


class ConfigPointConst(AbstractEqConst):
    desired_angles: np.ndarray

    def __init__(self, desired_angles: np.ndarray) -> None:
        self.desired_angles = np.asarray(desired_angles, dtype=np.float64)
        self.reflect_skrobot_model(None)
        self.assign_id_value()

    def _evaluate(self, qs, with_jacobian: bool):
        n_point, dim = qs.shape
        val = qs - _to_kind(qs, self.desired_angles)
        if not with_jacobian:
            return val, self.dummy_jacobian()
        eye = _to_kind(qs, np.eye(dim))
        jac = eye[None].expand(n_point, -1, -1).clone() if _is_torch(qs) else np.tile(eye, (n_point, 1, 1))
        return val, jac

    def _reflect_skrobot_model(self, robot_model) -> None:
        pass

    def get_description(self) -> np.ndarray:
        return self.desired_angles


class PoseConstraint(AbstractEqConst):
    """values = flatten(efkin.map(qs)) - hstack(desired_poses): raw component-wise residual, no angle
    wrapping and no quaternion sign handling, like the reference (:460-461)."""

    efkin: EndEffectorKinematicsMap
    desired_poses: List[np.ndarray]
    debug_rank_deficiency: bool = False

    def __init__(self, desired_poses: List[np.ndarray], efkin: EndEffectorKinematicsMap, robot_model,
                 debug_rank_deficiency: bool = False) -> None:
        assert len(desired_poses) == efkin.n_feature
        self.desired_poses = desired_poses
        self.efkin = efkin
        self.reflect_skrobot_model(robot_model)
        self.debug_rank_deficiency = debug_rank_deficiency
        self.assign_id_value()

    def _evaluate(self, qs, with_jacobian: bool = False):
        n_point, n_dim = qs.shape
        xs, jacs = self.efkin.map(qs)       # always with jacobian, like the reference (:455)
        m = xs.shape[1] * xs.shape[2]       # explicit: reshape(.., -1) is ambiguous for an empty batch
        values = xs.reshape(n_point, m) - _to_kind(xs, np.hstack(self.desired_poses))
        jacs = jacs.reshape(n_point, m, n_dim)
        if self.debug_rank_deficiency:
            jn = jacs.cpu().numpy() if _is_torch(jacs) else jacs
            for jac in jn:
                rank_diff = jac.shape[0] - np.linalg.matrix_rank(jac)
                assert rank_diff <= 0, "rank diference: {}, row sums: {}".format(rank_diff, np.sum(jac, axis=1))
        return values, jacs

    @classmethod
    def from_skrobot_coords(cls, co_list, efkin: EndEffectorKinematicsMap, robot_model,
                            debug_rank_deficiency: bool = False) -> "PoseConstraint":
        vectors = []
        for co in co_list:
            pos = co.worldpos()
            if efkin.rot_type == RotationType.RPY:
                vectors.append(np.hstack([pos, np.flip(rpy_angle(co.worldrot())[0])]))
            elif efkin.rot_type == RotationType.XYZW:
                vectors.append(np.hstack([pos, wxyz2xyzw(matrix2quaternion(co.worldrot()))]))
            elif efkin.rot_type == RotationType.IGNORE:
                vectors.append(pos)
            else:
                assert False
        return cls(vectors, efkin, robot_model, debug_rank_deficiency)

    def _reflect_skrobot_model(self, robot_model) -> None:
        assert robot_model is not None
        self.efkin.reflect_skrobot_model(robot_model)

    def get_description(self) -> np.ndarray:
        return np.hstack(self.desired_poses)


class RelativePoseConstraint(AbstractEqConst):
    """Feature 2 must coincide with feature 1 displaced by ``desired_relative_position`` (in feature
    1's frame).  Works on a private deep copy of ``efkin`` with the displaced point added."""

    desired_relative_position: np.ndarray
    efkin: EndEffectorKinematicsMap

    def __init__(self, desired_relative_position: np.ndarray, efkin: EndEffectorKinematicsMap, robot_model):
        efkin = copy.deepcopy(efkin)
        assert efkin.n_feature == 2
        efkin.add_new_feature_point(efkin.tinyfk_feature_ids[0], desired_relative_position, None)
        assert efkin.n_feature == 3
        self.desired_relative_position = desired_relative_position
        self.efkin = efkin
        self.reflect_skrobot_model(robot_model)
        self.assign_id_value()

    def _evaluate(self, qs, with_jacobian: bool):
        xs, jacs = self.efkin.map(qs)
        diffs = xs[:, 1, :] - xs[:, 2, :]
        if not with_jacobian:
            return diffs, self.dummy_jacobian()
        return diffs, jacs[:, 1, :, :] - jacs[:, 2, :, :]

    def _reflect_skrobot_model(self, robot_model) -> None:
        assert robot_model is not None
        self.efkin.reflect_skrobot_model(robot_model)


class FixedZAxisConstraint(AbstractEqConst):
    """Keeps each end-effector's local x and y axes horizontal: z(p + x_hat) - z(p) = 0 and
    z(p + y_hat) - z(p) = 0, via two extra feature points per end effector."""

    efkin: EndEffectorKinematicsMap
    n_feature: int

    def __init__(self, efkin: EndEffectorKinematicsMap, robot_model):
        efkin = copy.deepcopy(efkin)
        base_ids = list(efkin.tinyfk_feature_ids)
        for fid in base_ids:
            efkin.add_new_feature_point(fid, np.array([1, 0, 0]), None)
        for fid in base_ids:
            efkin.add_new_feature_point(fid, np.array([0, 1, 0]), None)
        self.n_feature = len(base_ids)
        self.efkin = efkin
        self.reflect_skrobot_model(robot_model)
        self.assign_id_value()

    def _evaluate(self, qs, with_jacobian: bool):
        n_point, n_dim = qs.shape
        nf = self.n_feature
        xs, jacs = self.efkin.map(qs)
        z = xs[:, :, 2]                                   # (N, 3*nf): origins, +x points, +y points
        diffs = _cat([z[:, nf:2 * nf] - z[:, :nf], z[:, 2 * nf:] - z[:, :nf]], axis=1)
        if not with_jacobian:
            return diffs, self.dummy_jacobian()
        jz = jacs[:, :, 2, :]                             # (N, 3*nf, D)
        return diffs, _cat([jz[:, nf:2 * nf] - jz[:, :nf], jz[:, 2 * nf:] - jz[:, :nf]], axis=1)

    def _reflect_skrobot_model(self, robot_model) -> None:
        assert robot_model is not None
        self.efkin.reflect_skrobot_model(robot_model)


class PairWiseSelfCollFreeConst(AbstractIneqConst):
    """values[n, p] = |x_i - x_j|^2 - (r_i + r_j)^2 over the checked sphere pairs (squared metres).

    ``check_sphere_id_pairs`` (tinyfk link-id tuples) is the single source of truth for the column
    order.  With ``id_pairs=None`` every pair is taken except those closer than 3 (r_i + r_j) at
    q = 0; the reference builds that list through ``list(set(...))`` (:732-733) whose order is
    unspecified -- here the surviving pairs keep itertools.combinations order.
    """

    colkin: CollSpheresKinematicsMapBase
    check_sphere_id_pairs: List[Tuple[int, int]]
    check_sphere_pair_sqdists: np.ndarray
    only_closest_feature: bool

    def __init__(self, colkin: CollSpheresKinematicsMapBase, robot_model,
                 id_pairs: Optional[List[Tuple[int, int]]] = None, only_closest_feature: bool = False) -> None:
        radius_of = dict(zip(colkin.tinyfk_feature_ids, colkin.get_radius_list()))
        if id_pairs is None:
            all_pairs = list(itertools.combinations(colkin.tinyfk_feature_ids, 2))
            rs = np.array([radius_of[a] + radius_of[b] for a, b in all_pairs])
            sqdists, _ = colkin.fksolver.compute_inter_link_sqdists(
                np.zeros((1, colkin.dim_cspace)), all_pairs, colkin.tinyfk_joint_ids, base_type=colkin.base_type)
            keep = np.sqrt(np.asarray(sqdists)) - 3.0 * rs >= 0
            id_pairs = [p for p, k in zip(all_pairs, keep) if k]
            pair_dists = rs[keep]
        else:
            id_pairs = [tuple(p) for p in id_pairs]
            pair_dists = np.array([radius_of[a] + radius_of[b] for a, b in id_pairs])
        self.colkin = colkin
        self.check_sphere_id_pairs = id_pairs
        self.check_sphere_pair_sqdists = pair_dists**2
        self.reflect_skrobot_model(robot_model)
        self.only_closest_feature = only_closest_feature
        self.assign_id_value()

    def _plan(self, device_index: Optional[int] = None):
        k = self.colkin
        dev = current_device_index() if device_index is None else device_index
        cache = self.__dict__.setdefault("_plan_cache", {})
        cached = cache.get(dev)
        if (cached is not None and cached[0] is k.fksolver and cached[1] == k.fksolver._version
                and cached[3] is self.check_sphere_id_pairs and cached[4] is self.check_sphere_pair_sqdists):
            return cached[2]
        local = {fid: i for i, fid in enumerate(k.tinyfk_feature_ids)}
        pairs = np.array([[local[a], local[b]] for a, b in self.check_sphere_id_pairs], dtype=np.int32)
        plan = k.fksolver.get_plan(k.tinyfk_feature_ids, k.tinyfk_joint_ids, k.base_type, RotationType.IGNORE,
                                   radii=None, pairs=pairs, pair_sqdists=self.check_sphere_pair_sqdists, device_index=dev)
        cache[dev] = (k.fksolver, k.fksolver._version, plan, self.check_sphere_id_pairs, self.check_sphere_pair_sqdists)
        return plan

    def _run(self, qs, want_vals: bool, with_jacobian: bool, want_valid: bool, only_closest: Optional[bool] = None):
        b = Batch(qs, self.colkin.dim_cspace)
        plan = self._plan(b.device_index)
        closest = self.only_closest_feature if only_closest is None else only_closest
        m = 1 if closest else len(self.check_sphere_id_pairs)
        vals = b.empty((b.n, m)) if want_vals else None
        jac = b.empty((b.n, m, plan.dim_cspace)) if with_jacobian else None
        valid = b.empty_bytes(b.n) if want_valid else None
        mode = _lib.MODE_CLOSEST if closest else _lib.MODE_ALL
        L = _lib.lib()
        with b.device_ctx():
            if b.on_device:
                _lib.check(L.skb_pairwise(plan.handle, b.dtype_code, b.q_ptr, b.n, mode, b.ptr(vals), b.ptr(jac), b.ptr(valid), b.stream))
            else:
                _lib.check(L.skb_pairwise_host(plan.handle, b.dtype_code, b.q_ptr, b.n, mode, b.ptr(vals), b.ptr(jac), b.ptr(valid)))
        return b, vals, jac, valid

    def _evaluate(self, qs, with_jacobian: bool):
        b, vals, jac, _ = self._run(qs, True, with_jacobian, False)
        return b.out(vals), (b.out(jac) if with_jacobian else self.dummy_jacobian())

    def _single(self, q):
        dt = q.dtype.type if q.dtype.type in (np.float32, np.float64) else np.float64
        dev = current_device_index()
        key = (dt, dev)
        cache = self.__dict__.setdefault("_single_cache", {})
        sc = cache.get(key)
        solver = self.colkin.fksolver
        if (sc is None or sc[1] != solver._version or sc[2] is not solver or sc[3] != self.only_closest_feature
                or sc[5] is not self.check_sphere_id_pairs):
            plan = self._plan(dev)
            code = _lib.F32 if dt == np.float32 else _lib.F64
            mode = _lib.MODE_CLOSEST if self.only_closest_feature else _lib.MODE_ALL
            m = 1 if self.only_closest_feature else len(self.check_sphere_id_pairs)
            call = _SingleCall(_lib.lib().skb_pairwise_host, (plan.handle, code), dt, plan.dim_cspace, m, (1, mode), (1, _lib.MODE_CLOSEST))
            sc = cache[key] = (call, solver._version, solver, self.only_closest_feature, plan, self.check_sphere_id_pairs)
        return sc[0]

    evaluate_single = CollFreeConst.evaluate_single
    is_valid = CollFreeConst.is_valid
    _min_value_single = CollFreeConst._min_value_single

    _min_values = CollFreeConst._min_values
    tune = CollFreeConst.tune
    tuned = CollFreeConst.tuned

    def is_valid_batch(self, qs):
        if not self.reflect_robot_flag:
            raise RuntimeError("{}: you need to call reflect_skrobot_model beforehand".format(type(self).__name__))
        b, _, _, valid = self._run(qs, False, False, True)
        out = b.out(valid)
        return out.bool() if _is_torch(out) else out.astype(bool)

    def _reflect_skrobot_model(self, robot_model) -> None:
        assert robot_model is not None
        self.colkin.reflect_skrobot_model(robot_model)


class COMStabilityConst(AbstractIneqConst):
    """value = -com_box.sdf(com(q)) (positive while the centre of mass projects inside the box), one row.

    Reference: skmp/constraint.py:845-931.  The centre of mass is the weighted mean of the links' inertial origins
    (URDF masses) and, optionally, of "action" links weighted by force / g -- one fused CUDA launch (FK, weighted
    prefix sums, box SDF with analytic gradient, chain rule).  The reference differentiates the box by forward
    differences (eps 1e-7, :912-920); see DESIGN.md "Gradient semantics".
    """

    def __init__(self, urdfpath, joint_names: List[str], base_type: BaseType, robot_model, com_box,
                 action_link_names: Optional[List[str]] = None, action_forces: Optional[List[float]] = None,
                 fksolver_init_hook=None):
        from .kinematics import _make_solver

        self.dim_cspace = len(joint_names) + (base_type == BaseType.PLANER) * 3 + (base_type == BaseType.FLOATING) * 6
        self.fksolver = _make_solver(urdfpath, fksolver_init_hook)
        self.tinyfk_joint_ids = self.fksolver.get_joint_ids(joint_names)
        self.control_joint_names = list(joint_names)
        if action_link_names is None:
            assert action_forces is None
            self.action_link_ids, self.action_forces = [], []
        else:
            assert action_forces is not None and len(action_forces) == len(action_link_names)
            self.action_link_ids = self.fksolver.get_link_ids(action_link_names)
            self.action_forces = list(action_forces)
        self.base_type = base_type
        self.model = robot_model
        self.com_box = as_descriptor(com_box)
        self.reflect_skrobot_model(robot_model)
        self.assign_id_value()

    def _evaluate(self, qs, with_jacobian: bool = False):
        b, vals, jac = self.fksolver._com_call(qs, self.tinyfk_joint_ids, self.action_link_ids, self.action_forces,
                                               self.base_type, with_jacobian, _lib.COM_STABILITY, self.com_box)
        return b.out(vals), (b.out(jac) if with_jacobian else self.dummy_jacobian())

    def com(self, qs, with_jacobian: bool = False):
        """Centre of mass (N,3) [+ Jacobian (N,3,D)] of the same weighted point set."""
        b, vals, jac = self.fksolver._com_call(qs, self.tinyfk_joint_ids, self.action_link_ids, self.action_forces,
                                               self.base_type, with_jacobian, _lib.COM_POINT, None)
        return b.out(vals), (b.out(jac) if with_jacobian else None)

    def _reflect_skrobot_model(self, robot_model) -> None:
        pass   # as the reference (:930-931): the solver keeps its own sticky state

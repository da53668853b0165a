"""World signed-distance fields as *descriptors* the GPU kernels consume.

The reference passes an opaque numpy callable ``sdf(points (M,3)) -> (M,)`` (skmp/constraint.py:268)
produced by scikit-robot: ``Box(...).sdf`` / ``Cylinder(...).sdf`` (tests/test_constraint.py:71-75,
skmp/robot/fetch.py:74-93) or ``UnionSDF`` (example/fetch_selcol_plan.py:42).  A Python callable cannot
be fused into a CUDA kernel, so here the same objects carry their parameters; every class is still
callable on (M, 3) points (evaluated on the GPU, no CPU fallback).

``Box`` / ``Cylinder`` / ``Sphere`` mimic the part of skrobot.model.primitives the reference touches:
``translate``, ``rotate_with_matrix``, ``worldpos``, ``worldrot``, ``_extents``, ``.sdf``.
"""
import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._array import Batch, current_device_index, device_guard, torch


class SignedDistanceFunction:
    """Base descriptor: a list of primitives whose union (min) is the field."""

    def __init__(self):
        self._handle = None
        self._keepalive = []

    # each entry: ("box", pos, rot, half) | ("cylinder", pos, rot, radius, half_height)
    #           | ("sphere", pos, radius) | ("grid", pos, rot, lb, spacing, values, fill)
    def primitives(self) -> List[tuple]:
        raise NotImplementedError

    def native(self, device_index: Optional[int] = None) -> C.c_void_p:
        """The ``skb_sdf`` handle of this field on one device (default: the current one).  Grids and primitive tables live
        in that device's memory, so there is one handle per device."""
        dev = current_device_index() if device_index is None else int(device_index)
        if self._handle is None:
            self._handle = {}
        if dev not in self._handle:
            _lib.require_cuda()
            L = _lib.lib()
            h = C.c_void_p()
            _lib.check(L.skb_sdf_create(C.byref(h)))
            d = lambda a: _lib.dptr(np.ascontiguousarray(a, dtype=np.float64))
            with device_guard(dev):
                for prim in self.primitives():
                    kind = prim[0]
                    if kind == "box":
                        _, pos, rot, half = prim
                        _lib.check(L.skb_sdf_add_box(h, d(pos), d(rot), d(half)))
                    elif kind == "cylinder":
                        _, pos, rot, radius, hh = prim
                        _lib.check(L.skb_sdf_add_cylinder(h, d(pos), d(rot), float(radius), float(hh)))
                    elif kind == "sphere":
                        _, pos, radius = prim
                        _lib.check(L.skb_sdf_add_sphere(h, d(pos), float(radius)))
                    elif kind == "grid":
                        _, pos, rot, lb, spacing, values, fill = prim
                        vals = np.ascontiguousarray(values)
                        if vals.dtype not in (np.float32, np.float64):
                            vals = vals.astype(np.float64)
                        dims = np.array(vals.shape, dtype=np.int32)
                        code = _lib.F32 if vals.dtype == np.float32 else _lib.F64
                        _lib.check(L.skb_sdf_add_grid(h, d(pos), d(rot), d(lb), d(spacing), _lib.dptr(dims, C.c_int32),
                                                      float(fill), C.c_void_p(vals.ctypes.data), code))
                    else:
                        raise ValueError(kind)
            self._handle[dev] = h
        return self._handle[dev]

    def invalidate(self):
        handles, self._handle = self._handle, None
        for h in (handles or {}).values():
            try:
                _lib.lib().skb_sdf_destroy(h)
            except Exception:
                pass

    def __del__(self):
        self.invalidate()

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_handle"] = None
        return d

    def __call__(self, points, with_gradient: bool = False):
        """sdf(points (M, 3)) -> (M,) signed distances [, (M, 3) analytic gradients]."""
        b = Batch(points, 3)
        vals = b.empty((b.n,))
        grads = b.empty((b.n, 3)) if with_gradient else None
        L = _lib.lib()
        if b.on_device:
            with b.device_ctx():
                _lib.check(L.skb_sdf_eval(self.native(b.device_index), b.dtype_code, b.q_ptr, b.n, b.ptr(vals), b.ptr(grads), b.stream))
        else:
            if torch is None:
                raise _lib.SkbError("evaluating an SDF on host arrays needs torch for device staging")
            dev = torch.from_numpy(b.q).cuda() if not b.is_torch else b.q.cuda()
            dv = torch.empty((b.n,), dtype=dev.dtype, device=dev.device)
            dg = torch.empty((b.n, 3), dtype=dev.dtype, device=dev.device) if with_gradient else None
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(L.skb_sdf_eval(self.native(b.device_index), b.dtype_code, C.c_void_p(dev.data_ptr()), b.n, C.c_void_p(dv.data_ptr()),
                                      C.c_void_p(dg.data_ptr()) if dg is not None else None, st))
            vals = dv.cpu() if b.is_torch else dv.cpu().numpy()
            grads = None if dg is None else (dg.cpu() if b.is_torch else dg.cpu().numpy())
            return (vals, grads) if with_gradient else vals
        return (b.out(vals), b.out(grads)) if with_gradient else b.out(vals)


class _Posed(SignedDistanceFunction):
    def __init__(self, pos=None, rot=None):
        super().__init__()
        self._pos = np.zeros(3) if pos is None else np.asarray(pos, dtype=np.float64).copy()
        self._rot = np.eye(3) if rot is None else np.asarray(rot, dtype=np.float64).copy()

    # --- the slice of skrobot's CascadedCoords API that the reference uses on primitives
    def translate(self, vec, wrt: str = "local"):
        vec = np.asarray(vec, dtype=np.float64)
        self._pos = self._pos + (self._rot @ vec if wrt == "local" else vec)
        self.invalidate()
        return self

    def rotate_with_matrix(self, mat, wrt: str = "local"):
        """``skrobot.coordinates.Coordinates.rotate_with_matrix``: local -> R <- R mat, world -> R <- mat R; the position
        stays where it is in both cases."""
        mat = np.asarray(mat, dtype=np.float64)
        self._rot = self._rot @ mat if wrt == "local" else mat @ self._rot
        self.invalidate()
        return self

    def rotate(self, theta: float, axis, wrt: str = "local"):
        """``Coordinates.rotate(theta, axis)``: axis 'x' / 'y' / 'z' or a 3-vector (example/humanoid_dualarm.py:34)."""
        from .rotmath import rotation_matrix

        ax = {"x": [1.0, 0.0, 0.0], "y": [0.0, 1.0, 0.0], "z": [0.0, 0.0, 1.0]}.get(axis, axis) if isinstance(axis, str) else axis
        return self.rotate_with_matrix(rotation_matrix(float(theta), np.asarray(ax, dtype=np.float64)), wrt)

    def newcoords(self, c, pos=None):
        """``Coordinates.newcoords(c, pos=None)``: ``c`` is a coordinates object (anything with worldpos / worldrot), or a
        3x3 rotation with ``pos`` the translation.  (A 3-vector first and a 3x3 matrix second -- this package's earlier
        argument order -- is still understood: the shapes are unambiguous.)"""
        if hasattr(c, "worldpos") and hasattr(c, "worldrot"):
            p, r = c.worldpos(), c.worldrot()
        else:
            a = np.asarray(c, dtype=np.float64)
            if a.shape == (3, 3):
                r, p = a, (self._pos if pos is None else pos)
            else:
                p, r = a, (self._rot if pos is None else pos)
        self._pos = np.asarray(p, dtype=np.float64).reshape(3).copy()
        self._rot = np.asarray(r, dtype=np.float64).reshape(3, 3).copy()
        self.invalidate()
        return self

    def copy_worldcoords(self):
        from .robot_model import Coordinates

        return Coordinates(self._pos, self._rot)

    def worldpos(self) -> np.ndarray:
        return self._pos.copy()

    def worldrot(self) -> np.ndarray:
        return self._rot.copy()

    def transform_vector(self, v) -> np.ndarray:
        return np.asarray(v) @ self._rot.T + self._pos

    @property
    def sdf(self):
        return self


class Box(_Posed):
    """``skrobot.model.primitives.Box(extents, with_sdf=True)`` stand-in; ``box.sdf`` is the descriptor."""

    def __init__(self, extents, pos=None, rot=None, with_sdf: bool = True, **_ignored):
        super().__init__(pos, rot)
        self._extents = np.asarray(extents, dtype=np.float64).copy()

    @property
    def extents(self):
        return self._extents

    def primitives(self):
        return [("box", self._pos, self._rot, 0.5 * self._extents)]


class Cylinder(_Posed):
    """``Cylinder(radius, height, with_sdf=True)``: axis along local z, centred."""

    def __init__(self, radius: float, height: float, pos=None, rot=None, with_sdf: bool = True, **_ignored):
        super().__init__(pos, rot)
        self.radius, self.height = float(radius), float(height)

    def primitives(self):
        return [("cylinder", self._pos, self._rot, self.radius, 0.5 * self.height)]


class Sphere(_Posed):
    def __init__(self, radius: float, pos=None, with_sdf: bool = True, **_ignored):
        super().__init__(pos, None)
        self.radius = float(radius)

    def primitives(self):
        return [("sphere", self._pos, self.radius)]


BoxSDF, CylinderSDF, SphereSDF = Box, Cylinder, Sphere


class GridSDF(_Posed):
    """Voxel-grid SDF: ``values[ix, iy, iz]`` sampled at ``lb + (ix, iy, iz) * spacing`` in the grid's
    local frame; trilinear inside, ``fill_value`` (finite) and zero gradient outside."""

    def __init__(self, values, lb, spacing, fill_value: float = 1.0e3, pos=None, rot=None):
        super().__init__(pos, rot)
        self.values = np.ascontiguousarray(values)
        assert self.values.ndim == 3 and min(self.values.shape) >= 2
        self.lb = np.asarray(lb, dtype=np.float64).copy()
        sp = np.asarray(spacing, dtype=np.float64)
        self.spacing = np.full(3, float(sp)) if sp.ndim == 0 else sp.copy()
        if not np.isfinite(fill_value):
            raise ValueError("fill_value must be finite (the reference's inf fill would make the FD gradient nan)")
        self.fill_value = float(fill_value)

    @classmethod
    def from_sdf(cls, sdf: SignedDistanceFunction, lb, ub, dims=(128, 128, 128), dtype=np.float32, fill_value=1.0e3):
        """Sample another descriptor on a regular lattice (on the GPU)."""
        lb, ub = np.asarray(lb, dtype=np.float64), np.asarray(ub, dtype=np.float64)
        dims = np.asarray(dims, dtype=np.int64)
        spacing = (ub - lb) / (dims - 1)
        axes = [lb[i] + spacing[i] * np.arange(dims[i]) for i in range(3)]
        pts = np.stack(np.meshgrid(*axes, indexing="ij"), axis=-1).reshape(-1, 3)
        vals = np.asarray(sdf(pts)).reshape(tuple(dims)).astype(dtype)
        return cls(vals, lb, spacing, fill_value)

    def primitives(self):
        return [("grid", self._pos, self._rot, self.lb, self.spacing, self.values, self.fill_value)]


class UnionSDF(SignedDistanceFunction):
    """``skrobot.sdf.UnionSDF(sdf_list)``: elementwise minimum over members."""

    def __init__(self, sdf_list: Sequence[SignedDistanceFunction]):
        super().__init__()
        self.sdf_list = [s.sdf if hasattr(s, "sdf") and not isinstance(s, SignedDistanceFunction) else s for s in sdf_list]
        for s in self.sdf_list:
            if not isinstance(s, SignedDistanceFunction):
                raise TypeError("UnionSDF members must be skmp_b200.sdf descriptors (arbitrary Python callables cannot run on the GPU)")

    def primitives(self):
        out = []
        for s in self.sdf_list:
            out.extend(s.primitives())
        return out

    @property
    def sdf(self):
        return self


def as_descriptor(sdf) -> SignedDistanceFunction:
    """Accept a descriptor or an object with a ``.sdf`` descriptor (``Box(...).sdf`` style)."""
    if isinstance(sdf, SignedDistanceFunction):
        return sdf
    inner = getattr(sdf, "sdf", None)
    if isinstance(inner, SignedDistanceFunction):
        return inner
    raise TypeError(
        "CollFreeConst needs an skmp_b200.sdf descriptor (Box / Cylinder / Sphere / UnionSDF / GridSDF); "
        "an opaque Python callable cannot be fused into the CUDA kernel and there is no CPU fallback"
    )

"""ctypes front end of the CPU oracle (oracle/skmp_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
module; the product package (scikit-motionplan_b200/) never does.  PARITY UNPINNED: see the header
of skmp_oracle.c.
"""
import ctypes as C
import subprocess
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from typing import Optional, Sequence

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "_build" / "libskmp_oracle.so"

BASE_FIXED, BASE_PLANER, BASE_FLOATING = 0, 1, 2
ROT_IGNORE, ROT_RPY, ROT_XYZW = 0, 1, 2
GRAD_ANALYTIC, GRAD_FD = 0, 1
_SDF_TYPES = {"box": 0, "cylinder": 1, "sphere": 2, "grid": 3}


def build(force: bool = False) -> Path:
    src = _HERE / "skmp_oracle.c"
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE)], check=True, capture_output=True)
    return LIB_PATH


class _Tree(C.Structure):
    _fields_ = [("n_links", C.c_int32), ("parent", C.POINTER(C.c_int32)), ("jtype", C.POINTER(C.c_int32)),
                ("origin", C.POINTER(C.c_double)), ("axis", C.POINTER(C.c_double)), ("angle", C.POINTER(C.c_double)),
                ("base_pose", C.c_double * 6)]


class _Prim(C.Structure):
    _fields_ = [("type", C.c_int32), ("dims", C.c_int32 * 3), ("pos", C.c_double * 3), ("rot", C.c_double * 9),
                ("param", C.c_double * 8), ("grid_f32", C.POINTER(C.c_float)), ("grid_f64", C.POINTER(C.c_double))]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            build()
        _lib = C.CDLL(str(LIB_PATH))
        _lib.orc_version.restype = C.c_int
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class Tree:
    """Raw link table + sticky state, as plain arrays (see orc_tree in skmp_oracle.c)."""

    def __init__(self, parent, jtype, origin, axis, angle, base_pose):
        self.parent = np.ascontiguousarray(parent, dtype=np.int32)
        self.jtype = np.ascontiguousarray(jtype, dtype=np.int32)
        self.origin = np.ascontiguousarray(origin, dtype=np.float64).reshape(-1, 6)
        self.axis = np.ascontiguousarray(axis, dtype=np.float64).reshape(-1, 3)
        self.angle = np.ascontiguousarray(angle, dtype=np.float64)
        self.base_pose = np.ascontiguousarray(base_pose, dtype=np.float64)
        self.c = _Tree(len(self.parent), _p(self.parent, C.c_int32), _p(self.jtype, C.c_int32), _p(self.origin, C.c_double),
                       _p(self.axis, C.c_double), _p(self.angle, C.c_double), (C.c_double * 6)(*self.base_pose))

    @classmethod
    def from_solver(cls, solver) -> "Tree":
        """Snapshot of an skmp_b200.tinyfk.KinematicModel's arrays (plain data, no product code runs)."""
        return cls(solver.parent, solver.jtype, solver.origin, solver.axis, solver.angle, solver.base_pose)


class Sdf:
    """Union of primitives; ``prims`` uses the tuple format of skmp_b200.sdf.*.primitives()."""

    def __init__(self, prims: Sequence[tuple]):
        self._keep = []
        arr = (_Prim * max(1, len(prims)))()
        for i, pr in enumerate(prims):
            kind = pr[0]
            c = arr[i]
            c.type = _SDF_TYPES[kind]
            pos = np.asarray(pr[1], dtype=np.float64)
            rot = np.eye(3) if kind == "sphere" else np.asarray(pr[2], dtype=np.float64)
            c.pos[:] = list(pos)
            c.rot[:] = list(rot.reshape(-1))
            param = np.zeros(8)
            if kind == "box":
                param[:3] = pr[3]
            elif kind == "cylinder":
                param[0], param[1] = pr[3], pr[4]
            elif kind == "sphere":
                param[0] = pr[2]
            else:
                _, _, _, lb, spacing, values, fill = pr
                vals = np.ascontiguousarray(values)
                if vals.dtype not in (np.float32, np.float64):
                    vals = vals.astype(np.float64)
                self._keep.append(vals)
                param[:3], param[3:6], param[6] = lb, spacing, fill
                c.dims[:] = list(vals.shape)
                if vals.dtype == np.float32:
                    c.grid_f32 = _p(vals, C.c_float)
                else:
                    c.grid_f64 = _p(vals, C.c_double)
            c.param[:] = list(param)
        self.arr, self.n = arr, len(prims)

    def __call__(self, pts, with_grad=False, with_kink=False, with_curv=False):
        """values [, gradients] [, distance to the nearest gradient discontinuity] [, Lipschitz bound of the gradient (1/m)]"""
        pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 3)
        m = len(pts)
        val = np.empty(m)
        grad = np.empty((m, 3))
        kink = np.empty(m)
        curv = np.empty(m)
        lib().orc_sdf_eval(self.arr, self.n, _p(pts, C.c_double), C.c_int64(m), _p(val, C.c_double), _p(grad, C.c_double),
                           _p(kink, C.c_double), _p(curv, C.c_double))
        out = (val,)
        if with_grad:
            out += (grad,)
        if with_kink:
            out += (kink,)
        if with_curv:
            out += (curv,)
        return out[0] if len(out) == 1 else out


def _dim(nj, base_type):
    return nj + (3 if base_type == BASE_PLANER else 6 if base_type == BASE_FLOATING else 0)


def _tdim(rot_type):
    return {ROT_IGNORE: 3, ROT_RPY: 6, ROT_XYZW: 7}[rot_type]


def solve_fk(tree: Tree, qs, feat_ids, joint_links, base_type=BASE_FIXED, rot_type=ROT_IGNORE, with_jacobian=True):
    qs = np.ascontiguousarray(qs, dtype=np.float64)
    feat = np.ascontiguousarray(feat_ids, dtype=np.int32)
    jl = np.ascontiguousarray(joint_links, dtype=np.int32)
    n, d, t = len(qs), _dim(len(jl), base_type), _tdim(rot_type)
    assert qs.shape == (n, d)
    pts = np.empty((n, len(feat), t))
    jac = np.empty((n, len(feat), t, d)) if with_jacobian else np.empty(0)
    rc = lib().orc_solve_fk(C.byref(tree.c), _p(qs, C.c_double), C.c_int64(n), _p(feat, C.c_int32), len(feat),
                            _p(jl, C.c_int32), len(jl), base_type, rot_type, int(with_jacobian), _p(pts, C.c_double),
                            _p(jac, C.c_double) if with_jacobian else None)
    assert rc == 0
    return pts, jac


def inter_link_sqdists(tree: Tree, qs, pairs, joint_links, base_type=BASE_FIXED, with_jacobian=True):
    qs = np.ascontiguousarray(qs, dtype=np.float64)
    pr = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
    jl = np.ascontiguousarray(joint_links, dtype=np.int32)
    n, d = len(qs), _dim(len(jl), base_type)
    sq = np.empty((n, len(pr)))
    grad = np.empty((n, len(pr), d)) if with_jacobian else np.empty(0)
    rc = lib().orc_inter_link_sqdists(C.byref(tree.c), _p(qs, C.c_double), C.c_int64(n), _p(pr, C.c_int32), len(pr),
                                      _p(jl, C.c_int32), len(jl), base_type, int(with_jacobian), _p(sq, C.c_double),
                                      _p(grad, C.c_double) if with_jacobian else None)
    assert rc == 0
    return sq, grad


def collfree(tree: Tree, qs, feat_ids, joint_links, sdf: Sdf, radii, margin=0.0, base_type=BASE_FIXED, only_closest=False,
             with_jacobian=True, grad_mode=GRAD_ANALYTIC, with_kink=False, with_curv=False):
    """CollFreeConst._evaluate restated (skmp/constraint.py:287-374).  with_kink / with_curv append, per output row, the
    distance of the sphere centre to the nearest gradient discontinuity of the SDF and a Lipschitz bound of its gradient."""
    qs = np.ascontiguousarray(qs, dtype=np.float64)
    feat = np.ascontiguousarray(feat_ids, dtype=np.int32)
    jl = np.ascontiguousarray(joint_links, dtype=np.int32)
    rad = np.ascontiguousarray(radii, dtype=np.float64)
    n, d = len(qs), _dim(len(jl), base_type)
    assert qs.shape == (n, d) and len(rad) == len(feat)
    m = 1 if only_closest else len(feat)
    vals = np.empty((n, m))
    jac = np.zeros((n, m, d)) if with_jacobian else np.empty(0)
    kink = np.empty((n, m))
    curv = np.empty((n, m))
    rc = lib().orc_collfree(C.byref(tree.c), _p(qs, C.c_double), C.c_int64(n), _p(feat, C.c_int32), len(feat),
                            _p(jl, C.c_int32), len(jl), base_type, sdf.arr, sdf.n, _p(rad, C.c_double), C.c_double(margin),
                            int(only_closest), int(with_jacobian), grad_mode, _p(vals, C.c_double),
                            _p(jac, C.c_double) if with_jacobian else None, _p(kink, C.c_double), _p(curv, C.c_double))
    assert rc == 0
    out = (vals, jac)
    if with_kink:
        out += (kink,)
    if with_curv:
        out += (curv,)
    return out


def collfree_threaded(n_threads: int, tree: Tree, qs, *args, keep: bool = True, block: int = 2048, **kw):
    """Shard rows over ``n_threads`` host threads (ctypes releases the GIL): the all-cores CPU baseline.
    With keep=False each thread walks its shard in blocks and drops the outputs (timing only; a 1M-config
    Jacobian set would not fit in host memory), returning the number of rows processed."""
    qs = np.ascontiguousarray(qs, dtype=np.float64)
    shards = np.array_split(qs, max(1, n_threads))

    def work(shard):
        if keep:
            return collfree(tree, shard, *args, **kw)
        done = 0
        for i in range(0, len(shard), block):
            collfree(tree, shard[i:i + block], *args, **kw)
            done += len(shard[i:i + block])
        return done

    if n_threads <= 1:
        parts = [work(qs)]
    else:
        with ThreadPoolExecutor(n_threads) as ex:
            parts = list(ex.map(work, shards))
    if not keep:
        return int(sum(parts))
    return tuple(np.concatenate([p[i] for p in parts], axis=0) for i in range(len(parts[0])))


def com_fk(tree: Tree, qs, point_links, weights, joint_links, base_type=BASE_FIXED, with_jacobian=True):
    """tinyfk.solve_com_fk restated (call site skmp/constraint.py:901-909): the weighted mean of the point features'
    positions -- link masses at the inertial origins, action links weighted by force / g -- and of their geometric
    Jacobians.  [UNVERIFIED-UPSTREAM] for the action-force weighting."""
    w = np.asarray(weights, dtype=np.float64)
    pts, jac = solve_fk(tree, qs, point_links, joint_links, base_type, ROT_IGNORE, with_jacobian)
    com = np.einsum("f,nfk->nk", w, pts) / w.sum()
    jc = np.einsum("f,nfkd->nkd", w, jac) / w.sum() if with_jacobian else np.empty(0)
    return com, jc


def com_stability(tree: Tree, qs, point_links, weights, joint_links, box: Sdf, base_type=BASE_FIXED, with_jacobian=True,
                  grad_mode=GRAD_ANALYTIC, with_kink=False):
    """COMStabilityConst._evaluate restated (skmp/constraint.py:899-928): value = -sdf(com); Jacobian = -grad . J_com
    with the analytic gradient or the reference's forward difference (eps = 1e-7)."""
    com, jc = com_fk(tree, qs, point_links, weights, joint_links, base_type, with_jacobian)
    d, g, kink = box(com, with_grad=True, with_kink=True)
    vals = -d.reshape(-1, 1)
    if not with_jacobian:
        return vals, np.empty(0)
    if grad_mode == GRAD_FD:
        eps = 1e-7
        g = np.zeros_like(com)
        for i in range(3):
            xp = com.copy()
            xp[:, i] += eps
            g[:, i] = (box(xp) - d) / eps
    jac = -np.einsum("nk,nkd->nd", g, jc)[:, None, :]
    return (vals, jac, kink) if with_kink else (vals, jac)

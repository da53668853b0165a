"""GPU tests of the classes and code paths that round 1 built but never exercised:

  * AttachedObstacleCollPointsKinematicsMap               (skmp/kinematics.py:171-231; JAXON's 152-point cloud)
  * RelativePoseConstraint / FixedZAxisConstraint         (skmp/constraint.py:508-607; tests/test_constraint.py:146-169)
  * the standalone ``sdf(points)`` callable               (skmp/constraint.py:268,314,353,397,401)
  * the fallback kernels: lane-per-configuration ``tree_kernel`` / ``pairwise_kernel`` (SKB_KERNEL=tree; what robots with
    D > 64 or > 255 nodes get), the one-row-per-lane ``sphere_kernel`` for every collfree mode (SKB_KERNEL=sphere1) and voxel
    grids without cell packs (SKB_GRID_CELLS=0; what grids too large for the packs get)
  * the two voxel-lookup paths of the step kernel (texture fetches / 256-bit global loads), bit for bit, on ragged tiles
  * the reference-held scene facts of tests/test_reference_pins.py, on the CUDA path.
"""
import copy

import numpy as np
import pytest

from helpers import ROT, RPY_ROT_ERR, TOL, collfree_jac_excess, oracle_args, pose_excess, sample_q, scaled_err
from oracle import oracle

pytestmark = pytest.mark.gpu


def _torch():
    import torch

    return torch


def err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float((np.abs(a - b) / np.maximum(1.0, np.abs(b))).max()) if a.size else 0.0


def dev(x, dtype):
    return _torch().as_tensor(np.asarray(x).astype(dtype), device="cuda")


# ------------------------------------------------------------------ AttachedObstacleCollPointsKinematicsMap
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_attached_obstacle_points_jaxon(dtype):
    """A box held in JAXON's right hand: 6^3 lattice minus the 4^3 interior = 152 points (SURVEY 8 size table), radius =
    margin; map() and CollFreeConst against the oracle, the point cloud against an independent numpy construction."""
    from skmp_b200.constraint import CollFreeConst
    from skmp_b200.robot.jaxon import JaxonConfig
    from skmp_b200.robot_model import Jaxon
    from skmp_b200.rotmath import rotation_matrix
    from skmp_b200.sdf import Box, UnionSDF

    jaxon = Jaxon(); jaxon.reset_manip_pose()
    held = Box([0.2, 0.3, 0.25])
    held.translate([0.4, -0.3, 1.0])
    held.rotate_with_matrix(rotation_matrix(0.4, [0.3, 1.0, -0.2]))
    rel = np.array([0.05, 0.0, -0.12])
    kin = JaxonConfig().get_attached_obstacle_kin(rel, held)
    assert kin.n_feature == 152 and len(kin.tinyfk_feature_ids) == 152 and kin.dim_cspace == 37 and kin.dim_tspace == 3
    # the point cloud, restated: lattice in the box frame, rotated by the box, interior (deeper than 1 cm) dropped;
    # numpy.meshgrid's default 'xy' indexing fixes the order (kinematics.py:195-204)
    e = np.array([0.2, 0.3, 0.25])
    gx, gy, gz = np.meshgrid(*[np.linspace(-0.5 * e[i], 0.5 * e[i], 6) for i in range(3)])
    local = np.stack([gx.flatten(), gy.flatten(), gz.flatten()], axis=1)
    depth = np.max(np.abs(local) - 0.5 * e, axis=1)            # box SDF inside the box
    keep = depth > -1e-2
    expect = local[keep] @ held.worldrot().T + rel
    origins = kin.fksolver.origin[kin.tinyfk_feature_ids, :3]
    assert np.abs(origins - expect).max() < 1e-12
    attach = kin.fksolver.get_link_ids(["rarm_end_coords"])[0]
    assert (kin.fksolver.parent[kin.tinyfk_feature_ids] == attach).all()

    kin_m = copy.deepcopy(kin)
    kin_m.radius_list = [0.03] * kin_m.n_feature                 # margin = 3 cm (kinematics.py:222)
    env = UnionSDF([Box([4.0, 4.0, 0.2], pos=[0.0, 0.0, -0.1]), Box([0.6, 1.0, 0.8], pos=[0.9, 0.0, 0.4])])
    rng = np.random.default_rng(31)
    qs = sample_q(kin, 300, rng) * 0.5
    qs[:, 33] += 1.0
    qs = qs.astype(dtype).astype(np.float64)
    kin.reflect_skrobot_model(jaxon)
    pts, jac = kin.map(dev(qs, dtype))
    tree, feat, jl, base = oracle_args(kin)
    rp, rj = oracle.solve_fk(tree, qs, feat, jl, base)
    assert err(pts.cpu().numpy(), rp) < TOL[dtype] and err(jac.cpu().numpy(), rj) < TOL[dtype]
    for k, closest in ((kin, False), (kin_m, False), (kin_m, True)):
        const = CollFreeConst(k, env, jaxon, only_closest_feature=closest)
        v, j = const.evaluate(dev(qs, dtype), True)
        tree, feat, jl, base = oracle_args(k)
        rv, rjac, kink, curv = oracle.collfree(tree, qs, feat, jl, oracle.Sdf(env.primitives()), k.get_radius_list(), base_type=base,
                                               only_closest=closest, with_kink=True, with_curv=True)
        assert err(v.cpu().numpy(), rv) < TOL[dtype]
        # the held box's surface points graze the table's edges: within a millimetre of an edge the gradient turns by
        # 1 / distance per metre, which the per-row tolerance accounts for (helpers.collfree_jac_excess)
        excess, excluded = collfree_jac_excess(j.cpu().numpy(), rjac, kink, curv, dtype)
        assert excess <= 1.0 and excluded < 0.01
    # radius = margin shifts every value by exactly the margin
    v0, _ = CollFreeConst(kin, env, jaxon).evaluate(dev(qs, np.float64), False)
    v1, _ = CollFreeConst(kin_m, env, jaxon).evaluate(dev(qs, np.float64), False)
    assert float((v0 - v1 - 0.03).abs().max()) < 1e-12


# ------------------------------------------------------------------ derived pose constraints
def _fd_jacobian(const, q, eps=1e-7):
    """tests/test_constraint.py:26-38 (forward difference of evaluate_single), in fp64."""
    f0, _ = const.evaluate_single(q, False)
    f0 = np.asarray(f0.cpu() if hasattr(f0, "cpu") else f0)
    jac = np.zeros((len(f0), len(q)))
    for i in range(len(q)):
        q1 = q.copy()
        q1[i] += eps
        f1, _ = const.evaluate_single(q1, False)
        jac[:, i] = (np.asarray(f1.cpu() if hasattr(f1, "cpu") else f1) - f0) / eps
    return jac


@pytest.mark.parametrize("rot", ["RPY", "XYZW", "IGNORE"])
def test_relative_pose_constraint(rot):
    """tests/test_constraint.py:146-161: Jacobian = forward difference to 4 decimals, the caller's efkin is not modified;
    plus values / Jacobians against the oracle's FK of the three features (xs[:,1] - xs[:,2], constraint.py:534-548)."""
    from skmp_b200.constraint import RelativePoseConstraint
    from skmp_b200.robot.pr2 import PR2Config
    from skmp_b200.robot_model import PR2
    from skmp_b200.tinyfk import RotationType

    pr2 = PR2()
    efkin = PR2Config(control_arm="dual").get_endeffector_kin(RotationType[rot])
    const = RelativePoseConstraint(np.ones(3) * 0.1, efkin, pr2)
    assert efkin.n_feature == 2 and len(efkin.tinyfk_feature_ids) == 2
    assert const.efkin.n_feature == 3 and isinstance(const.id_value, str)
    rng = np.random.default_rng(41)
    for _ in range(3):
        q = rng.standard_normal(14)
        _, jac = const.evaluate_single(q, True)
        np.testing.assert_almost_equal(np.asarray(jac), _fd_jacobian(const, q), decimal=4)
    for dtype in (np.float32, np.float64):
        qs = rng.standard_normal((200, 14)).astype(dtype).astype(np.float64)
        v, j = const.evaluate(dev(qs, dtype), True)
        tree, feat, jl, base = oracle_args(const.efkin)
        xs, jacs = oracle.solve_fk(tree, qs, feat, jl, base, ROT[RotationType[rot]])
        # the rotational rows are differences of two RPY angle sets / their rate maps: tolerance of the worse-conditioned one
        # (helpers.pose_excess: TOL + RPY_ROT_ERR / cos(pitch)^k), twice for the difference
        cp = np.minimum(np.abs(np.cos(xs[:, 1, 4])), np.abs(np.cos(xs[:, 2, 4]))) if xs.shape[2] == 6 else np.ones(len(qs))
        tv = np.full(xs.shape[2], TOL[dtype])[None, :].repeat(len(qs), 0)
        tj = tv.copy()
        if xs.shape[2] == 6:
            tv[:, 3:] += 2 * RPY_ROT_ERR[dtype] / cp[:, None]
            tj[:, 3:] += 2 * RPY_ROT_ERR[dtype] / cp[:, None] ** 2
        assert (scaled_err(v.cpu().numpy(), xs[:, 1, :] - xs[:, 2, :]) / tv).max() <= 1.0
        assert (scaled_err(j.cpu().numpy(), jacs[:, 1] - jacs[:, 2]) / tj[:, :, None]).max() <= 1.0
        v2, j2 = const.evaluate(qs.astype(dtype), False)            # host path, no Jacobian
        assert (scaled_err(v2, xs[:, 1, :] - xs[:, 2, :]) / tv).max() <= 1.0 and np.isnan(np.asarray(j2)).all()


@pytest.mark.parametrize("rot", ["RPY", "IGNORE"])
def test_fixed_zaxis_constraint(rot):
    """tests/test_constraint.py:164-169 + the reference's own formula (constraint.py:582-603: reshape to
    (n, 3, n_feature, T), z of (+x point - origin) then of (+y point - origin)) evaluated on the oracle's FK."""
    from skmp_b200.constraint import FixedZAxisConstraint
    from skmp_b200.robot.pr2 import PR2Config
    from skmp_b200.robot_model import PR2
    from skmp_b200.tinyfk import RotationType

    pr2 = PR2()
    efkin = PR2Config(control_arm="dual").get_endeffector_kin(RotationType[rot])
    const = FixedZAxisConstraint(efkin, pr2)
    assert efkin.n_feature == 2 and const.efkin.n_feature == 6 and const.n_feature == 2
    rng = np.random.default_rng(43)
    for _ in range(3):
        q = rng.standard_normal(14)
        val, jac = const.evaluate_single(q, True)
        assert np.asarray(val).shape == (4,) and np.asarray(jac).shape == (4, 14)
        np.testing.assert_almost_equal(np.asarray(jac), _fd_jacobian(const, q), decimal=4)
    for dtype in (np.float32, np.float64):
        qs = rng.standard_normal((200, 14)).astype(dtype).astype(np.float64)
        v, j = const.evaluate(dev(qs, dtype), True)
        tree, feat, jl, base = oracle_args(const.efkin)
        xs, jacs = oracle.solve_fk(tree, qs, feat, jl, base, ROT[RotationType[rot]])
        n, nf = len(qs), 2
        tmp = xs.reshape(n, 3, nf, -1)
        ref_v = np.concatenate([(tmp[:, 1] - tmp[:, 0])[:, :, 2], (tmp[:, 2] - tmp[:, 0])[:, :, 2]], axis=1).reshape(n, -1)
        tj = jacs.reshape(n, 3, nf, -1, 14)
        ref_j = np.concatenate([(tj[:, 1] - tj[:, 0])[:, :, 2, :], (tj[:, 2] - tj[:, 0])[:, :, 2, :]], axis=1).reshape(n, -1, 14)
        assert err(v.cpu().numpy(), ref_v) < TOL[dtype] and err(j.cpu().numpy(), ref_j) < TOL[dtype]      # z rows only: no RPY term
        # the unit offsets make the values the (3,1) / (3,2) entries of the end-effector rotation: |value| <= 1
        assert float(v.abs().max()) <= 1.0 + 1e-5


# ------------------------------------------------------------------ standalone sdf(points)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_sdf_callable_direct(dtype):
    """sdf(points (M,3)) -> (M,) [+ analytic gradients (M,3)] for every descriptor, CUDA tensors and numpy, against the
    oracle's primitives (``skb_sdf_eval``).  The reference calls it at skmp/constraint.py:314,329,353,368,397,401."""
    from skmp_b200.rotmath import rotation_matrix
    from skmp_b200.sdf import Box, Cylinder, GridSDF, Sphere, UnionSDF

    rng = np.random.default_rng(51)
    box = Box([0.7, 0.5, 1.2]); box.translate([0.85, -0.2, 0.9]); box.rotate_with_matrix(rotation_matrix(0.3, [0.1, 0.2, 1.0]))
    cyl = Cylinder(0.15, 0.8); cyl.translate([0.4, -0.6, 1.0]); cyl.rotate_with_matrix(rotation_matrix(1.0, [1.0, 0.0, 0.0]))
    sph = Sphere(0.2, pos=[0.6, 0.0, 1.6])
    union = UnionSDF([box, cyl, sph])
    lb, ub = np.array([-0.5, -1.5, 0.0]), np.array([1.5, 1.5, 2.0])
    dims = np.array([40, 56, 48])
    spacing = (ub - lb) / (dims - 1)
    axes = [lb[i] + spacing[i] * np.arange(dims[i]) for i in range(3)]
    lattice = np.stack(np.meshgrid(*axes, indexing="ij"), axis=-1).reshape(-1, 3)
    grid_vals = oracle.Sdf(union.primitives())(lattice).reshape(tuple(dims)).astype(np.float32)   # sampled by the ORACLE, not by the engine
    grid = GridSDF(grid_vals, lb, spacing, fill_value=7.5)
    grid_rot = GridSDF(grid_vals.astype(np.float64), lb, spacing, fill_value=7.5, pos=[0.1, -0.05, 0.02], rot=rotation_matrix(0.2, [0.0, 0.0, 1.0]))
    pts = rng.uniform([-0.8, -1.8, -0.3], [1.8, 1.8, 2.3], size=(6000, 3)).astype(dtype).astype(np.float64)   # some outside the grid
    for sdf in (box, cyl, sph, union, grid, grid_rot):
        rv, rg, kink = oracle.Sdf(sdf.primitives())(pts, with_grad=True, with_kink=True)
        ok = kink > (1e-4 if dtype == np.float32 else 1e-9)
        v, g = sdf(dev(pts, dtype), with_gradient=True)
        assert err(v.cpu().numpy(), rv) < TOL[dtype]
        assert err(g.cpu().numpy()[ok], rg[ok]) < 3 * TOL[dtype]       # unit-length gradients from rsqrt (2 ulp) in fp32
        vh = sdf(pts.astype(dtype))                                        # numpy in -> numpy out, values only
        assert isinstance(vh, np.ndarray) and err(vh, rv) < TOL[dtype]
        vh2, gh2 = sdf(pts.astype(dtype), with_gradient=True)
        assert np.array_equal(vh2, v.cpu().numpy()) and np.array_equal(gh2, g.cpu().numpy())
    outside = ((pts < lb) | (pts > ub)).any(axis=1)
    vg, gg = grid(dev(pts, dtype), with_gradient=True)
    assert outside.sum() > 100 and (vg.cpu().numpy()[outside] == 7.5).all() and (gg.cpu().numpy()[outside] == 0).all()


# ------------------------------------------------------------------ fallback kernels
@pytest.mark.parametrize("env", [{"SKB_KERNEL": "tree"}, {"SKB_KERNEL": "sphere1"}, {"SKB_GRID_CELLS": "0"}],
                         ids=["tree_kernels", "sphere1_kernel", "grid_without_cell_packs"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_fallback_kernels(monkeypatch, env, dtype):
    """cfg1 (PR2 right arm vs box, all / closest / validity), a union + voxel-grid scene, cfg2 (Fetch pairwise, all / closest)
    and pose rows, computed by the fallback code paths (selected when the plans / SDFs are created under the switch) and
    compared with the oracle to the same tolerances as the default kernels."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    from skmp_b200.constraint import CollFreeConst, PoseConstraint
    from skmp_b200.robot.fetch import FetchConfig
    from skmp_b200.robot.pr2 import PR2Config
    from skmp_b200.robot_model import PR2, Fetch
    from skmp_b200.rotmath import rotation_matrix
    from skmp_b200.sdf import Box, Cylinder, GridSDF, UnionSDF
    from skmp_b200.tinyfk import RotationType

    torch = _torch()
    pr2 = PR2(); pr2.reset_manip_pose()
    rng = np.random.default_rng(61)
    box = Box([0.7, 0.5, 1.2], pos=[0.85, -0.2, 0.9])
    cyl = Cylinder(0.15, 0.8); cyl.translate([0.4, -0.6, 1.0]); cyl.rotate_with_matrix(rotation_matrix(1.0, [1.0, 0.0, 0.0]))
    union = UnionSDF([box, cyl])
    lb, ub, dims = np.array([-0.5, -1.5, 0.0]), np.array([1.5, 1.5, 2.0]), np.array([48, 64, 48])
    spacing = (ub - lb) / (dims - 1)
    axes = [lb[i] + spacing[i] * np.arange(dims[i]) for i in range(3)]
    lattice = np.stack(np.meshgrid(*axes, indexing="ij"), axis=-1).reshape(-1, 3)
    grid = GridSDF(oracle.Sdf(union.primitives())(lattice).reshape(tuple(dims)).astype(np.float32), lb, spacing)
    colkin = PR2Config().get_collision_kin()            # fresh solver: its plans are compiled under the switch
    qs = sample_q(colkin, 500, rng).astype(dtype).astype(np.float64)
    tree, feat, jl, base = oracle_args(colkin)
    for sdf in (box, union, grid):
        osdf = oracle.Sdf(sdf.primitives())
        for closest in (False, True):
            const = CollFreeConst(colkin, sdf, pr2, only_closest_feature=closest)
            v, j = const.evaluate(dev(qs, dtype), True)
            rv, rj, kink, curv = oracle.collfree(tree, qs, feat, jl, osdf, colkin.get_radius_list(), only_closest=closest,
                                                 with_kink=True, with_curv=True)
            assert err(v.cpu().numpy(), rv) < TOL[dtype]
            excess, excluded = collfree_jac_excess(j.cpu().numpy(), rj, kink, curv, dtype)
            assert excess <= 1.0 and excluded < 0.02
            v2, j2 = const.evaluate(dev(qs, dtype), False)        # another kernel instantiation: same values to rounding
            assert err(v2.cpu().numpy(), v.cpu().numpy()) < TOL[dtype] * 0.1 and np.isnan(np.asarray(j2)).all()
        const = CollFreeConst(colkin, sdf, pr2)
        valid = const.is_valid_batch(dev(qs, dtype)).cpu().numpy()
        rv, _ = oracle.collfree(tree, qs, feat, jl, osdf, colkin.get_radius_list(), with_jacobian=False)
        decisive = np.abs(rv).min(axis=1) > 1e-6
        assert (valid == (rv > 0).all(axis=1))[decisive].all() and 0.02 < valid.mean() < 0.98
    # Fetch pairwise
    fetch = Fetch(); fetch.reset_pose()
    for closest in (False, True):
        pw = FetchConfig().get_pairwise_selcol_consts(fetch, only_closest_feature=closest)
        q2 = sample_q(pw.colkin, 300, rng).astype(dtype).astype(np.float64)
        v, j = pw.evaluate(dev(q2, dtype), True)
        t2, _, jl2, b2 = oracle_args(pw.colkin)
        sq, gr = oracle.inter_link_sqdists(t2, q2, np.array(pw.check_sphere_id_pairs), jl2, b2)
        rv = sq - pw.check_sphere_pair_sqdists
        if closest:
            idx, n = rv.argmin(axis=1), np.arange(len(rv))
            assert err(v.cpu().numpy()[:, 0], rv[n, idx]) < TOL[dtype] and err(j.cpu().numpy()[:, 0], gr[n, idx]) < TOL[dtype]
        else:
            assert err(v.cpu().numpy(), rv) < TOL[dtype] and err(j.cpu().numpy(), gr) < TOL[dtype]
    # pose rows (the tree kernel's RPY / XYZW epilogue under SKB_KERNEL=tree)
    for rt in (RotationType.RPY, RotationType.XYZW):
        efkin = PR2Config(control_arm="dual").get_endeffector_kin(rt)
        efkin.reflect_skrobot_model(pr2)
        q3 = sample_q(efkin, 200, rng).astype(dtype).astype(np.float64)
        pts, jac = efkin.map(dev(q3, dtype))
        t3, f3, jl3, b3 = oracle_args(efkin)
        rp, rjac = oracle.solve_fk(t3, q3, f3, jl3, b3, ROT[rt])
        ev, ej = pose_excess(pts.cpu().numpy(), jac.cpu().numpy(), rp, rjac, rt, dtype)
        assert ev <= 1.0 and ej <= 1.0


# ------------------------------------------------------------------ voxel lookups: TEX pipe vs LDG.256
@pytest.mark.parametrize("n", [1, 3, 501, 4099])
def test_grid_lookup_paths_bit_identical(monkeypatch, n):
    """The fp32 cell packs are read either by two texel fetches (default; the step kernel issues them one configuration
    ahead of the column loop) or by one 256-bit global load (SKB_GRID_TEX=0, what grids beyond 2^27 texels get).  Both read
    the same 32 bytes, so every mode returns the same bits -- on batch sizes that leave ragged last tiles (the look-ahead
    must not run past the last live configuration of a tile) and on the dual-arm PR2 of BASELINE config 3."""
    from skmp_b200.constraint import CollFreeConst
    from skmp_b200.robot.pr2 import PR2Config
    from skmp_b200.robot_model import PR2
    from skmp_b200.sdf import Box, Cylinder, GridSDF, UnionSDF

    pr2 = PR2(); pr2.reset_manip_pose()
    rng = np.random.default_rng(7)
    union = UnionSDF([Box([0.7, 0.5, 1.2], pos=[0.85, -0.2, 0.9]), Cylinder(0.15, 0.8)])
    lb, ub, dims = np.array([-0.6, -1.5, 0.0]), np.array([1.5, 1.5, 2.0]), np.array([40, 56, 44])
    spacing = (ub - lb) / (dims - 1)
    axes = [lb[i] + spacing[i] * np.arange(dims[i]) for i in range(3)]
    lattice = np.stack(np.meshgrid(*axes, indexing="ij"), axis=-1).reshape(-1, 3)
    values = oracle.Sdf(union.primitives())(lattice).reshape(tuple(dims)).astype(np.float32)
    colkin = PR2Config(control_arm="dual").get_collision_kin()
    qs = dev(sample_q(colkin, n, rng), np.float32)
    results = []
    for tex in ("1", "0"):
        monkeypatch.setenv("SKB_GRID_TEX", tex)              # read when the grid is uploaded (first use of the SDF object)
        grid = GridSDF(values, lb, spacing, fill_value=2.5)
        out = []
        for closest in (False, True):
            const = CollFreeConst(colkin, grid, pr2, only_closest_feature=closest)
            v, j = const.evaluate(qs, True)
            v2, _ = const.evaluate(qs, False)
            out += [v.cpu().numpy(), j.cpu().numpy(), v2.cpu().numpy()]
        out.append(CollFreeConst(colkin, grid, pr2).is_valid_batch(qs).cpu().numpy())
        results.append(out)
    for a, b in zip(*results):
        assert a.shape == b.shape and np.array_equal(a, b, equal_nan=True)
    # and the look-ahead changes nothing against the oracle: first and last configuration of the batch
    tree, feat, jl, base = oracle_args(colkin)
    pick = np.unique([0, n - 1])
    rv, rj = oracle.collfree(tree, qs.cpu().numpy().astype(np.float64)[pick], feat, jl, oracle.Sdf(grid.primitives()),
                             colkin.get_radius_list())
    assert err(results[0][0][pick], rv) < TOL[np.float32]


# ------------------------------------------------------------------ measurement aids of the C ABI
def test_peak_microbenchmarks_run():
    """skb_fma_peak / skb_store_peak (the roofline denominators bench_kernels.py reports) return plausible figures and
    reject bad arguments."""
    import ctypes

    from skmp_b200 import _lib

    _torch().cuda.init()
    lib = _lib.lib()
    t = ctypes.c_double(0.0)
    _lib.check(lib.skb_fma_peak(0, 512, ctypes.byref(t)))
    assert 5.0 < t.value < 200.0                                    # TFLOP/s fp32
    for kind, blk in ((0, 0), (1, 3584)):
        _lib.check(lib.skb_store_peak(kind, 1 << 30, blk, ctypes.byref(t)))
        assert 500.0 < t.value < 20000.0                            # GB/s written
    assert lib.skb_store_peak(1, 1 << 30, 100, ctypes.byref(t)) != 0     # block size not a multiple of 16
    assert lib.skb_store_peak(2, 1 << 30, 0, ctypes.byref(t)) != 0       # unknown kind


# ------------------------------------------------------------------ reference-held scene facts on the CUDA path
def test_reference_pins_on_gpu():
    """The facts of tests/test_reference_pins.py through the product API (numpy in, like the reference's callers)."""
    from skmp_b200.constraint import CollFreeConst, PoseConstraint
    from skmp_b200.robot.pr2 import PR2Config
    from skmp_b200.robot_model import PR2, Coordinates, get_robot_state
    from skmp_b200.sdf import Box

    pr2 = PR2(); pr2.reset_manip_pose()
    conf = PR2Config()
    colkin, efkin = conf.get_collision_kin(), conf.get_endeffector_kin()
    q_goal_cand = np.array([-0.78, 0.055, -1.37, -0.59, -0.494, -0.20, 1.87])
    start = np.array([0.564, 0.35, -0.74, -0.7, -0.7, -0.17, -0.63])
    pose = PoseConstraint.from_skrobot_coords([Coordinates(pos=[0.8, -0.6, 1.1])], efkin, pr2)
    val, _ = pose.evaluate_single(q_goal_cand, False)
    assert np.abs(np.asarray(val)[:3]).max() < 0.012 and np.abs(np.asarray(val)[3:]).max() < 0.015
    q_now = get_robot_state(pr2, conf._get_control_joint_names())
    near = CollFreeConst(colkin, Box([0.7, 0.5, 1.2], pos=[0.68, -0.3, 0.9]).sdf, pr2)
    assert not near.is_valid(q_now)                                   # tests/test_satisfy.py:92-100
    std = CollFreeConst(colkin, Box([0.7, 0.5, 1.2], pos=[0.85, -0.2, 0.9]).sdf, pr2)
    assert std.is_valid(start) and std.is_valid(q_now)                # tests/solver_tests/utils.py:24-37

"""GPU parity: CUDA path (through the C ABI) vs the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): fp32 within 1e-5 relative (scale = max(1, |ref|)), fp64 within
1e-10; validity identical outside a +-1e-6 m band around 0.  Gradient kinks of the SDF (box medial axis,
voxel cell faces, union switch-over) are measure-zero sets where fp32 and fp64 may legitimately pick
different sides; the oracle reports each sample's distance to the nearest kink and samples closer than
KINK_BAND are excluded from the *Jacobian* comparison only.
"""
import numpy as np
import pytest

from helpers import BASE, KINK_BAND, ROT, TOL, collfree_jac_excess, oracle_args, pose_excess, sample_q
from oracle import oracle

pytestmark = pytest.mark.gpu

VALID_BAND = 1e-6


def close(a, b, tol):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    err = np.abs(a - b) / np.maximum(1.0, np.abs(b))
    return err.max() if err.size else 0.0


def _torch():
    import torch

    return torch


@pytest.fixture(scope="module")
def pr2():
    from skmp_b200.robot_model import PR2

    r = PR2()
    r.reset_manip_pose()
    return r


def box_scene():
    from skmp_b200.sdf import Box

    box = Box([0.7, 0.5, 1.2])
    box.translate([0.85, -0.2, 0.9])
    return box


def run_collfree(const, qs, dtype, device):
    torch = _torch()
    if device:
        q = torch.as_tensor(qs.astype(dtype), device="cuda")
        v, j = const.evaluate(q, True)
        torch.cuda.synchronize()
        return v.cpu().numpy(), j.cpu().numpy()
    v, j = const.evaluate(qs.astype(dtype), True)
    return np.asarray(v), np.asarray(j)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("device", [True, False])
@pytest.mark.parametrize("margin", [0.0, 0.05])
def test_collfree_box_config1(pr2, dtype, device, margin):
    """BASELINE config 1: PR2 right arm 7-DoF vs box, 1k random configs, values + Jacobians."""
    from skmp_b200.constraint import CollFreeConst
    from skmp_b200.robot.pr2 import PR2Config

    colkin = PR2Config().get_collision_kin()
    box = box_scene()
    rng = np.random.default_rng(1)
    qs = sample_q(colkin, 1000, rng).astype(dtype).astype(np.float64)
    for closest in (False, True):
        const = CollFreeConst(colkin, box.sdf, pr2, only_closest_feature=closest, distance_margin=margin)
        v, j = run_collfree(const, qs, dtype, device)
        tree, feat, jl, base = oracle_args(colkin)
        rv, rj, kink, curv = oracle.collfree(tree, qs, feat, jl, oracle.Sdf(box.primitives()), colkin.get_radius_list(),
                                             margin=margin, only_closest=closest, with_kink=True, with_curv=True)
        assert v.shape == rv.shape and j.shape == rj.shape
        assert close(v, rv, TOL[dtype]) < TOL[dtype]
        excess, excluded = collfree_jac_excess(j, rj, kink, curv, dtype)
        assert excluded < 0.001 and excess <= 1.0
        band = np.abs(rv) > VALID_BAND
        assert ((v > 0) == (rv > 0))[band].all()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_validity_bytes(pr2, dtype):
    from skmp_b200.constraint import CollFreeConst
    from skmp_b200.robot.pr2 import PR2Config

    torch = _torch()
    colkin = PR2Config(control_arm="dual").get_collision_kin()
    box = box_scene()
    const = CollFreeConst(colkin, box.sdf, pr2)
    rng = np.random.default_rng(2)
    qs = sample_q(colkin, 4099, rng).astype(dtype).astype(np.float64)
    valid = const.is_valid_batch(torch.as_tensor(qs.astype(dtype), device="cuda")).cpu().numpy()
    tree, feat, jl, base = oracle_args(colkin)
    rv, _ = oracle.collfree(tree, qs, feat, jl, oracle.Sdf(box.primitives()), colkin.get_radius_list(), with_jacobian=False)
    decisive = np.abs(rv).min(axis=1) > VALID_BAND
    assert decisive.mean() > 0.99
    assert (valid == (rv > 0).all(axis=1))[decisive].all()
    assert 0.05 < valid.mean() < 0.95          # the scene produces both outcomes
    for i in range(5):                          # N = 1 path agrees with the batch
        assert const.is_valid(qs[i].astype(dtype)) == bool(valid[i]) or not decisive[i]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("base_type", ["FIXED", "PLANER", "FLOATING"])
def test_map_positions(pr2, dtype, base_type):
    """KinematicsMapBase.map, IGNORE rotation, all base types (whole-body spheres when the base moves)."""
    from skmp_b200.robot.pr2 import PR2Config
    from skmp_b200.tinyfk import BaseType

    torch = _torch()
    kin = PR2Config(control_arm="dual", base_type=BaseType[base_type], use_torso=True).get_collision_kin()
    kin.reflect_skrobot_model(pr2)
    rng = np.random.default_rng(3)
    qs = sample_q(kin, 333, rng).astype(dtype).astype(np.float64)
    pts, jac = kin.map(torch.as_tensor(qs.astype(dtype), device="cuda"))
    tree, feat, jl, base = oracle_args(kin)
    rp, rj = oracle.solve_fk(tree, qs, feat, jl, base)
    assert pts.shape == rp.shape and jac.shape == rj.shape
    assert close(pts.cpu().numpy(), rp, TOL[dtype]) < TOL[dtype]
    assert close(jac.cpu().numpy(), rj, TOL[dtype]) < TOL[dtype]
    pts2, jac2 = kin.map(qs.astype(dtype), with_jacobian=False)     # host path, no jacobian
    assert close(pts2, rp, TOL[dtype]) < TOL[dtype] and np.asarray(jac2).size == 0


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("rot", ["IGNORE", "RPY", "XYZW"])
def test_map_pose_features(pr2, dtype, rot):
    """EndEffectorKinematicsMap (K4 epilogue): PR2 dual arm fixed base and JAXON floating base."""
    from skmp_b200.robot.jaxon import JaxonConfig
    from skmp_b200.robot.pr2 import PR2Config
    from skmp_b200.robot_model import Jaxon
    from skmp_b200.tinyfk import RotationType

    torch = _torch()
    rt = RotationType[rot]
    jaxon = Jaxon()
    jaxon.reset_manip_pose()
    cases = [(PR2Config(control_arm="dual").get_endeffector_kin(rt), pr2, 1.0),
             (JaxonConfig().get_endeffector_kin(rot_type=rt), jaxon, 0.5)]
    for kin, robot, scale in cases:
        kin.reflect_skrobot_model(robot)
        rng = np.random.default_rng(4)
        qs = (sample_q(kin, 257, rng) * scale).astype(dtype).astype(np.float64)
        pts, jac = kin.map(torch.as_tensor(qs.astype(dtype), device="cuda"))
        tree, feat, jl, base = oracle_args(kin)
        rp, rj = oracle.solve_fk(tree, qs, feat, jl, base, ROT[rt])
        ev, ej = pose_excess(pts.cpu().numpy(), jac.cpu().numpy(), rp, rj, rt, dtype)    # RPY rows: + err / cos(pitch)^k
        assert ev <= 1.0 and ej <= 1.0


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_collfree_union_and_grid(pr2, dtype):
    """Union of rotated primitives, and a voxel-grid SDF (BASELINE config 3 scene at reduced N)."""
    from skmp_b200.constraint import CollFreeConst
    from skmp_b200.robot.pr2 import PR2Config
    from skmp_b200.rotmath import rotation_matrix
    from skmp_b200.sdf import Box, Cylinder, GridSDF, Sphere, UnionSDF

    colkin = PR2Config(control_arm="dual").get_collision_kin()
    b1 = box_scene()
    b2 = Box([0.4, 0.4, 0.3]); b2.translate([0.5, 0.5, 0.6]); b2.rotate_with_matrix(rotation_matrix(0.5, [0.2, 1.0, 0.1]))
    cy = Cylinder(0.15, 0.8); cy.translate([0.4, -0.6, 1.0]); cy.rotate_with_matrix(rotation_matrix(1.0, [1.0, 0.0, 0.0]))
    sp = Sphere(0.2, pos=[0.6, 0.0, 1.6])
    union = UnionSDF([b1, b2, cy, sp])
    grid = GridSDF.from_sdf(union, lb=[-0.5, -1.5, 0.0], ub=[1.5, 1.5, 2.0], dims=(64, 80, 64), dtype=np.float32)
    rng = np.random.default_rng(5)
    qs = sample_q(colkin, 600, rng).astype(dtype).astype(np.float64)
    tree, feat, jl, base = oracle_args(colkin)
    for sdf in (union, grid):
        const = CollFreeConst(colkin, sdf, pr2)
        v, j = run_collfree(const, qs, dtype, True)
        rv, rj, kink, curv = oracle.collfree(tree, qs, feat, jl, oracle.Sdf(sdf.primitives()), colkin.get_radius_list(),
                                             with_kink=True, with_curv=True)
        assert close(v, rv, TOL[dtype]) < TOL[dtype]
        # voxel grid: value and gradient are Horner evaluations of the cell polynomial (no cancellation), so the grid
        # meets the same per-row tolerance as the analytic primitives
        excess, excluded = collfree_jac_excess(j, rj, kink, curv, dtype)
        assert excluded < 0.01 and excess <= 1.0


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_pairwise_fetch_config2(dtype):
    """BASELINE config 2 at reduced N: Fetch 8-DoF PairWiseSelfCollFreeConst, pair list from the q=0 filter."""
    from skmp_b200.robot.fetch import FetchConfig
    from skmp_b200.robot_model import Fetch

    torch = _torch()
    fetch = Fetch()
    fetch.reset_pose()
    conf = FetchConfig()
    const = conf.get_pairwise_selcol_consts(fetch)
    colkin = const.colkin
    P = len(const.check_sphere_id_pairs)
    assert 100 < P <= 1275
    v0, _ = const.evaluate_single(np.zeros(8), False)
    assert (np.asarray(v0) > 0).all()                      # tests/test_constraint.py:178-180
    rng = np.random.default_rng(6)
    qs = sample_q(colkin, 500, rng).astype(dtype).astype(np.float64)
    v, j = const.evaluate(torch.as_tensor(qs.astype(dtype), device="cuda"), True)
    tree, _, jl, base = oracle_args(colkin)
    sq, gr = oracle.inter_link_sqdists(tree, qs, np.array(const.check_sphere_id_pairs), jl, base)
    rv = sq - const.check_sphere_pair_sqdists
    assert close(v.cpu().numpy(), rv, TOL[dtype]) < TOL[dtype]
    assert close(j.cpu().numpy(), gr, TOL[dtype]) < TOL[dtype]
    closest = conf.get_pairwise_selcol_consts(fetch, only_closest_feature=True)
    closest.check_sphere_id_pairs, closest.check_sphere_pair_sqdists = const.check_sphere_id_pairs, const.check_sphere_pair_sqdists
    vc, jc = closest.evaluate(qs.astype(dtype), True)
    idx = rv.argmin(axis=1)
    assert close(np.asarray(vc)[:, 0], rv.min(axis=1), TOL[dtype]) < TOL[dtype]
    decided = np.sort(rv, axis=1)[:, 1] - np.sort(rv, axis=1)[:, 0] > 1e-4
    assert close(np.asarray(jc)[decided, 0], gr[np.arange(len(qs)), idx][decided], TOL[dtype]) < TOL[dtype]


def test_min_all_equals_closest(pr2):
    """tests/test_constraint.py:85-89: np.min(values_all) == value_closest, exact."""
    from skmp_b200.constraint import CollFreeConst
    from skmp_b200.robot.pr2 import PR2Config

    torch = _torch()
    colkin = PR2Config().get_collision_kin()
    box = box_scene()
    a = CollFreeConst(colkin, box.sdf, pr2, only_closest_feature=False, distance_margin=0.05)
    c = CollFreeConst(colkin, box.sdf, pr2, only_closest_feature=True, distance_margin=0.05)
    q = torch.randn(777, 7, device="cuda", generator=torch.Generator("cuda").manual_seed(0))
    va, _ = a.evaluate(q, False)
    vc, _ = c.evaluate(q, False)
    assert torch.equal(va.min(dim=1).values, vc[:, 0])


def test_pairwise_jaxon_floating_base_wide_masks():
    """JAXON 37-DoF floating base self collision (D > 32: 64-bit plus / minus column masks, > 2000 pairs):
    all-pairs values + Jacobians and the closest row against the oracle, validity == min > 0."""
    from skmp_b200.constraint import PairWiseSelfCollFreeConst
    from skmp_b200.robot.jaxon import JaxonConfig
    from skmp_b200.robot_model import Jaxon

    torch = _torch()
    jaxon = Jaxon(); jaxon.reset_manip_pose()
    colkin = JaxonConfig().get_collision_kin()
    full = PairWiseSelfCollFreeConst(colkin, jaxon)
    closest = PairWiseSelfCollFreeConst(colkin, jaxon, only_closest_feature=True)
    assert closest.check_sphere_id_pairs == full.check_sphere_id_pairs and len(full.check_sphere_id_pairs) > 1000
    rng = np.random.default_rng(12)
    qs = sample_q(colkin, 300, rng).astype(np.float32).astype(np.float64)
    q_dev = torch.as_tensor(qs.astype(np.float32), device="cuda")
    v, j = full.evaluate(q_dev, True)
    tree, _, jl, base = oracle_args(colkin)
    sq, gr = oracle.inter_link_sqdists(tree, qs, np.array(full.check_sphere_id_pairs), jl, base)
    rv = sq - full.check_sphere_pair_sqdists
    tol = TOL[np.float32]
    assert close(v.cpu().numpy(), rv, tol) < tol
    assert close(j.cpu().numpy(), gr, tol) < tol
    vc, jc = closest.evaluate(q_dev, True)
    assert torch.equal(vc[:, 0], v.min(dim=1).values)                      # same arithmetic in both modes
    idx = v.argmin(dim=1)
    assert torch.equal(jc[:, 0], j[torch.arange(len(qs), device="cuda"), idx])
    assert torch.equal(closest.is_valid_batch(q_dev), vc[:, 0] > 0)
    v2, _ = full.evaluate(q_dev, False)
    assert torch.equal(v2, v)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_nan_configuration_is_never_valid(pr2, dtype):
    """A NaN joint value poisons its configuration like numpy does (np.min -> NaN, all(values > 0) -> False): the
    closest value and its Jacobian row are NaN, the validity byte is 0, the neighbouring configurations are untouched."""
    from skmp_b200.constraint import CollFreeConst
    from skmp_b200.robot.pr2 import PR2Config

    torch = _torch()
    colkin = PR2Config().get_collision_kin()
    box = box_scene()
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    q = torch.randn(97, 7, device="cuda", generator=torch.Generator("cuda").manual_seed(4)).to(tdt) * 0.3
    bad = q.clone()
    bad[11, 6] = float("nan")                                  # the wrist: only the gripper spheres depend on it
    for closest in (False, True):
        const = CollFreeConst(colkin, box.sdf, pr2, only_closest_feature=closest)
        v0, j0 = const.evaluate(q, True)
        v1, j1 = const.evaluate(bad, True)
        keep = torch.ones(len(q), dtype=torch.bool, device="cuda"); keep[11] = False
        assert torch.equal(v0[keep], v1[keep]) and torch.equal(j0[keep], j1[keep])
        assert torch.isnan(v1[11]).any() and (not closest or torch.isnan(j1[11]).all())
        ok = const.is_valid_batch(bad)
        assert not bool(ok[11]) and torch.equal(ok[keep], const.is_valid_batch(q)[keep])


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n_pairs", [1, 31, 32, 33, 128, 129, 700])
def test_pairwise_reducing_modes_any_pair_count(dtype, n_pairs):
    """The closest row / validity byte of an explicit pair list (robot/pr2.py:338-350 passes one) for pair counts around
    the scan's 32-lane and 4-half-round granularity: bit-identical to min / argmin over the all-pairs launch, shuffled
    pair order included (np.argmin: the first minimal row wins), and a NaN configuration stays NaN / invalid."""
    from skmp_b200.constraint import PairWiseSelfCollFreeConst
    from skmp_b200.robot.fetch import FetchConfig
    from skmp_b200.robot_model import Fetch

    torch = _torch()
    fetch = Fetch(); fetch.reset_pose()
    colkin = FetchConfig().get_collision_kin()
    every = PairWiseSelfCollFreeConst(colkin, fetch).check_sphere_id_pairs
    rng = np.random.default_rng(100 + n_pairs)
    pick = rng.permutation(len(every))[:n_pairs]
    id_pairs = [every[i] for i in pick]
    if n_pairs > 2:
        id_pairs[1] = id_pairs[0]                              # a duplicated pair: an exact tie, the lower row must win
    full = PairWiseSelfCollFreeConst(colkin, fetch, id_pairs=id_pairs)
    closest = PairWiseSelfCollFreeConst(colkin, fetch, id_pairs=id_pairs, only_closest_feature=True)
    qs = torch.as_tensor(sample_q(colkin, 257, rng).astype(dtype), device="cuda")
    v, j = full.evaluate(qs, True)
    vc, jc = closest.evaluate(qs, True)
    idx = v.argmin(dim=1)
    if n_pairs > 2:
        assert not (idx == 1).any()
    assert torch.equal(vc[:, 0], v.min(dim=1).values)
    assert torch.equal(jc[:, 0], j[torch.arange(len(qs), device="cuda"), idx])
    vc2, _ = closest.evaluate(qs, False)
    assert torch.equal(vc2, vc)
    assert torch.equal(closest.is_valid_batch(qs), vc[:, 0] > 0)
    bad = qs.clone()
    bad[3] = float("nan")
    vb, jb = closest.evaluate(bad, True)
    assert torch.isnan(vb[3]).all() and torch.equal(vb[:3], vc[:3]) and torch.equal(vb[4:], vc[4:])
    assert not bool(closest.is_valid_batch(bad)[3])


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("with_forces", [False, True])
def test_com_stability_jaxon(dtype, with_forces):
    """COMStabilityConst (skmp/constraint.py:845-931) and solve_com_fk against the oracle: JAXON 37-DoF floating base,
    38 massive links (+ 2 action links with 5 N each), box 0.2 x 0.6 x 5."""
    from skmp_b200.robot.jaxon import JaxonConfig
    from skmp_b200.robot_model import Jaxon
    from skmp_b200.sdf import Box
    from skmp_b200.tinyfk import BaseType

    torch = _torch()
    jaxon = Jaxon(); jaxon.reset_manip_pose()
    conf = JaxonConfig()
    box = Box([0.2, 0.6, 5.0])
    const = conf.get_com_stability_const(jaxon, box, *( (["RARM_LINK7", "LARM_LINK7"], [5.0, 5.0]) if with_forces else ()))
    assert isinstance(const.id_value, str)
    s = const.fksolver
    feats, w = s.com_points(const.action_link_ids, const.action_forces)
    tree = oracle.Tree.from_solver(s)
    jl = s.joint_child_links(const.tinyfk_joint_ids)
    rng = np.random.default_rng(21)
    qs = (rng.normal(size=(1000, 37)) * 0.3).astype(dtype).astype(np.float64)
    q_dev = torch.as_tensor(qs.astype(dtype), device="cuda")
    tol = TOL[dtype]
    com, jc = const.com(q_dev, True)
    rcom, rjc = oracle.com_fk(tree, qs, feats, np.array(w), jl, oracle.BASE_FLOATING)
    assert com.shape == (1000, 3) and jc.shape == (1000, 3, 37)
    assert close(com.cpu().numpy(), rcom, tol) < tol and close(jc.cpu().numpy(), rjc, tol) < tol
    xs, jt = s.solve_com_fk(qs.astype(dtype), const.tinyfk_joint_ids, const.action_link_ids, const.action_forces,
                            BaseType.FLOATING, True)                  # numpy in -> numpy out, the reference's call shape
    assert isinstance(xs, np.ndarray) and xs.shape == (1000, 3) and jt.shape == (3000, 37)
    assert close(xs, rcom, tol) < tol and close(jt.reshape(1000, 3, 37), rjc, tol) < tol
    v, j = const.evaluate(q_dev, True)
    rv, rj, kink = oracle.com_stability(tree, qs, feats, np.array(w), jl, oracle.Sdf(box.primitives()), oracle.BASE_FLOATING,
                                        with_kink=True)
    assert v.shape == (1000, 1) and j.shape == (1000, 1, 37)
    assert close(v.cpu().numpy(), rv, tol) < tol
    ok = kink > (2e-5 if dtype == np.float32 else 1e-9)
    assert ok.mean() > 0.95 and close(j.cpu().numpy()[ok], rj[ok], tol) < tol
    v2, j2 = const.evaluate(q_dev, False)
    assert torch.equal(v2, v) and np.isnan(np.asarray(j2)).all()
    assert const.is_valid(qs[0].astype(dtype)) == bool(rv[0, 0] > 0)


def test_ragged_and_empty_batches_bit_exact(pr2):
    """Tile tails: batches of 0, 1, 3, 5, 33 rows give exactly the rows of a large batch (every mode of both
    collision constraints, pose map, centre of mass), and N = 0 returns empty arrays of the right shape."""
    from skmp_b200.constraint import CollFreeConst, PoseConstraint
    from skmp_b200.robot.fetch import FetchConfig
    from skmp_b200.robot.jaxon import JaxonConfig
    from skmp_b200.robot.pr2 import PR2Config
    from skmp_b200.robot_model import Fetch, Jaxon
    from skmp_b200.sdf import Box

    torch = _torch()
    conf = PR2Config()
    colkin = conf.get_collision_kin()
    box = box_scene()
    fetch = Fetch(); fetch.reset_pose()
    jaxon = Jaxon(); jaxon.reset_manip_pose()
    cases = [
        (CollFreeConst(colkin, box.sdf, pr2), 7),
        (CollFreeConst(colkin, box.sdf, pr2, only_closest_feature=True), 7),
        (FetchConfig().get_pairwise_selcol_consts(fetch), 8),
        (FetchConfig().get_pairwise_selcol_consts(fetch, only_closest_feature=True), 8),
        (PoseConstraint([np.zeros(6)], conf.get_endeffector_kin(), pr2), 7),
        (JaxonConfig().get_com_stability_const(jaxon, Box([0.2, 0.6, 5.0])), 37),
    ]
    g = torch.Generator("cuda").manual_seed(3)
    for const, d in cases:
        qs = (torch.rand((133, d), generator=g, device="cuda") - 0.5) * 1.5
        v, j = const.evaluate(qs, True)
        for n in (1, 3, 5, 33):
            v1, j1 = const.evaluate(qs[40:40 + n].contiguous(), True)
            assert torch.equal(v1, v[40:40 + n]) and torch.equal(j1, j[40:40 + n]), (type(const).__name__, n)
        v0, j0 = const.evaluate(qs[:0], True)
        assert v0.shape == (0,) + v.shape[1:] and j0.shape == (0,) + j.shape[1:]
        if hasattr(const, "is_valid_batch"):
            assert torch.equal(const.is_valid_batch(qs), (v > 0).all(dim=1))
            assert const.is_valid_batch(qs[:0]).shape == (0,)


@pytest.mark.parametrize("GUARD", [4096, 4099])       # 16-byte aligned outputs (TMA bulk stores) / misaligned (fallback stores)
def test_outputs_stay_inside_their_buffers(pr2, GUARD):
    """Canary test through the raw C ABI (compute-sanitizer is not available on the pool): every output sits in the middle
    of a larger buffer pre-filled with a sentinel; after the launch the guard bands must be untouched and the payload fully
    written, for batch sizes that leave ragged tiles and for TMA-stored as well as fallback-stored Jacobians."""
    import ctypes as C

    from skmp_b200 import _lib
    from skmp_b200.constraint import CollFreeConst, PairWiseSelfCollFreeConst, PoseConstraint
    from skmp_b200.robot.fetch import FetchConfig
    from skmp_b200.robot.jaxon import JaxonConfig
    from skmp_b200.robot.pr2 import PR2Config
    from skmp_b200.robot_model import Fetch, Jaxon
    from skmp_b200.sdf import Box, UnionSDF

    torch = _torch()
    L = _lib.lib()
    SENT = -777.25
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ptr = lambda t: C.c_void_p(t.data_ptr())

    def guarded(n_elems, dtype):
        buf = torch.full((GUARD + n_elems + GUARD,), SENT, dtype=dtype, device="cuda")
        return buf, buf[GUARD: GUARD + n_elems]

    def check(buf, n_elems, what):
        assert (buf[:GUARD] == SENT).all() and (buf[GUARD + n_elems:] == SENT).all(), what + ": guard band overwritten"
        assert not (buf[GUARD: GUARD + n_elems] == SENT).any(), what + ": payload not fully written"

    fetch = Fetch(); fetch.reset_pose()
    jaxon = Jaxon(); jaxon.reset_manip_pose()
    jk = JaxonConfig().get_collision_kin()
    box = box_scene()
    env = UnionSDF([Box([4.0, 4.0, 0.2], pos=[0.0, 0.0, -0.1]), Box([0.6, 1.0, 0.8], pos=[0.9, 0.0, 0.4])])
    coll = [(CollFreeConst(PR2Config().get_collision_kin(), box.sdf, pr2), 7),            # TMA-aligned rows
            (CollFreeConst(jk, env, jaxon), 37)]                                            # unaligned rows: fallback stores
    g = torch.Generator("cuda").manual_seed(9)
    for dtype, code in ((torch.float32, _lib.F32), (torch.float64, _lib.F64)):
        for n in (1, 5, 131):
            for const, d in coll:
                plan, S = const._plan(), const.colkin.n_feature
                q = ((torch.rand((n, d), generator=g, device="cuda") - 0.5) * 1.2).to(dtype)
                for mode, m in ((_lib.MODE_ALL, S), (_lib.MODE_CLOSEST, 1)):
                    vb, v = guarded(n * m, dtype)
                    jb, j = guarded(n * m * d, dtype)
                    _lib.check(L.skb_collfree(plan.handle, const.sdf.native(), code, ptr(q), n, 0.0, mode, ptr(v), ptr(j), None, st))
                    torch.cuda.synchronize()
                    check(vb, n * m, f"collfree values mode {mode} n {n}"); check(jb, n * m * d, f"collfree jac mode {mode} n {n}")
                ob = torch.full((GUARD + n + GUARD,), 7, dtype=torch.uint8, device="cuda")
                _lib.check(L.skb_collfree(plan.handle, const.sdf.native(), code, ptr(q), n, 0.0, _lib.MODE_ALL, None, None,
                                          C.c_void_p(ob.data_ptr() + GUARD), st))
                torch.cuda.synchronize()
                assert (ob[:GUARD] == 7).all() and (ob[GUARD + n:] == 7).all() and (ob[GUARD: GUARD + n] <= 1).all()
            pw = FetchConfig().get_pairwise_selcol_consts(fetch)
            plan, P = pw._plan(), len(pw.check_sphere_id_pairs)
            q = ((torch.rand((n, 8), generator=g, device="cuda") - 0.5) * 1.2).to(dtype)
            vb, v = guarded(n * P, dtype); jb, j = guarded(n * P * 8, dtype)
            _lib.check(L.skb_pairwise(plan.handle, code, ptr(q), n, _lib.MODE_ALL, ptr(v), ptr(j), None, st))
            torch.cuda.synchronize()
            check(vb, n * P, "pairwise values"); check(jb, n * P * 8, "pairwise jac")
            ef = PR2Config().get_endeffector_kin()
            ef.reflect_skrobot_model(pr2)
            plan = ef.fksolver.get_plan(ef.tinyfk_feature_ids, ef.tinyfk_joint_ids, ef.base_type, ef.rot_type)
            q = ((torch.rand((n, 7), generator=g, device="cuda") - 0.5) * 1.2).to(dtype)
            vb, v = guarded(n * 6, dtype); jb, j = guarded(n * 6 * 7, dtype)
            _lib.check(L.skb_fk(plan.handle, code, ptr(q), n, 1, ptr(v), ptr(j), st))
            torch.cuda.synchronize()
            check(vb, n * 6, "pose values"); check(jb, n * 6 * 7, "pose jac")

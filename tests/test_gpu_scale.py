"""BASELINE.json configurations at (per-GPU) full size, checked through size-independent properties:
batch-split invariance (tiles are independent), permutation equivariance, closest == min(all),
validity == all(values > 0), structural zeros of the Jacobian, determinism, and spot checks of a
sub-sample against the CPU oracle."""
import numpy as np
import pytest

from helpers import ROT, oracle_args, sample_q
from oracle import oracle

pytestmark = pytest.mark.gpu


def _torch():
    import torch

    return torch


def _device_samples(kin, n, seed, scale=1.0):
    """Uniform samples within the joint limits, generated on the device (numpy would take a while for 10M x 37)."""
    torch = _torch()
    lo = sample_q(kin, 1, np.random.default_rng(0)) * 0       # only to get shapes; limits below
    jm = kin.fksolver.urdf.joint_map
    lo = [(-2 * np.pi if jm[n_].limit.lower is None else jm[n_].limit.lower) for n_ in kin.control_joint_names]
    hi = [(2 * np.pi if jm[n_].limit.upper is None else jm[n_].limit.upper) for n_ in kin.control_joint_names]
    nb = kin.dim_cspace - len(lo)
    lo, hi = np.array(lo + [-1.0] * nb) * scale, np.array(hi + [1.0] * nb) * scale
    g = torch.Generator(device="cuda").manual_seed(seed)
    lo_t, hi_t = torch.tensor(lo, dtype=torch.float32, device="cuda"), torch.tensor(hi, dtype=torch.float32, device="cuda")
    return torch.rand((n, kin.dim_cspace), generator=g, device="cuda", dtype=torch.float32) * (hi_t - lo_t) + lo_t


def test_cfg3_pr2_dual_grid_full_shard():
    """1.25 M configurations (the per-GPU shard of the 10 M of config 3) vs a 128^3 voxel grid, values + Jacobians."""
    import bench
    from skmp_b200.constraint import CollFreeConst

    torch = _torch()
    pr2, colkin, grid = bench.make_problem()
    const = CollFreeConst(colkin, grid, pr2)
    n = 1_250_000
    qs = _device_samples(colkin, n, 7)
    vals, jac = const.evaluate(qs, True)
    assert vals.shape == (n, 152) and jac.shape == (n, 152, 14)
    assert torch.isfinite(vals).all() and torch.isfinite(jac[:: 97]).all()
    # structural zeros: right-arm spheres never depend on left-arm joints and vice versa
    r_s = torch.tensor([i for i, nm in enumerate(colkin.sphere_name_list) if nm.startswith("r_")], device="cuda")
    l_s = torch.tensor([i for i, nm in enumerate(colkin.sphere_name_list) if nm.startswith("l_")], device="cuda")
    r_j = torch.tensor([i for i, nm in enumerate(colkin.control_joint_names) if nm.startswith("r_")], device="cuda")
    l_j = torch.tensor([i for i, nm in enumerate(colkin.control_joint_names) if nm.startswith("l_")], device="cuda")
    assert len(r_s) == 76 and len(l_s) == 76 and len(r_j) == 7 and len(l_j) == 7
    sub = jac[::7]
    assert (sub[:, r_s][:, :, l_j] == 0).all() and (sub[:, l_s][:, :, r_j] == 0).all()
    assert (sub[:, r_s][:, :, r_j] != 0).any() and (sub[:, l_s][:, :, l_j] != 0).any()
    # batch-split invariance, bit exact: odd-sized pieces shift every tile boundary
    cut = 333_337
    v1, j1 = const.evaluate(qs[:cut], True)
    v2, j2 = const.evaluate(qs[cut: cut + 100_003], True)
    assert torch.equal(v1, vals[:cut]) and torch.equal(j1, jac[:cut])
    assert torch.equal(v2, vals[cut: cut + 100_003]) and torch.equal(j2, jac[cut: cut + 100_003])
    del v1, j1, v2, j2
    # closest == min over all (tests/test_constraint.py:85-89) and validity bytes == all(values > 0)
    closest = CollFreeConst(colkin, grid, pr2, only_closest_feature=True)
    vc, jc = closest.evaluate(qs, True)
    assert torch.equal(vc[:, 0], vals.min(dim=1).values)
    idx = vals.argmin(dim=1)
    rows = torch.arange(0, n, 1013, device="cuda")
    assert torch.equal(jc[rows, 0], jac[rows, idx[rows]])
    valid = const.is_valid_batch(qs)
    assert torch.equal(valid, (vals > 0).all(dim=1))
    # spot check against the oracle
    pick = np.arange(0, n, n // 300)[:300]
    q64 = qs[pick].cpu().numpy().astype(np.float64)
    tree, feat, jl, base = oracle_args(colkin)
    rv, rj, kink = oracle.collfree(tree, q64, feat, jl, oracle.Sdf(grid.primitives()), colkin.get_radius_list(), with_kink=True)
    ev = np.abs(vals[pick].cpu().numpy() - rv) / np.maximum(1, np.abs(rv))
    assert ev.max() < 4e-5
    ok = kink > 2e-5
    ej = (np.abs(jac[pick].cpu().numpy() - rj) / np.maximum(1, np.abs(rj)))[ok]
    assert ej.max() < 2e-4


def test_launch_autotuner_is_result_neutral():
    """The first launch of >= 65 536 configurations of a plan times every (program, G, warps per CTA) candidate on a slice
    of the caller's own buffers (`launch_s2_inst`, csrc/skb_sphere2.cu).  Whatever the candidates left behind, and the
    tuned launch itself, must equal -- bit for bit -- what a separate plan computes with the model's configuration on
    pieces too small to be tuned, in every mode."""
    import bench
    from skmp_b200.constraint import CollFreeConst

    torch = _torch()
    pr2, colkin, grid = bench.make_problem()
    n = 70_001
    qs = _device_samples(colkin, n, 11)
    tuned = CollFreeConst(colkin, grid, pr2)
    vals, jac = tuned.evaluate(qs, True)            # tunes on the first 70 001 rows
    vals_b, jac_b = tuned.evaluate(qs, True)        # reuses the choice
    assert torch.equal(vals, vals_b) and torch.equal(jac, jac_b)
    valid = tuned.is_valid_batch(qs)
    closest = CollFreeConst(colkin, grid, pr2, only_closest_feature=True)
    vc, jc = closest.evaluate(qs, True)
    # a second solver: its plans carry no tuned entry, and 9 973 rows per call stay below the tuning threshold
    pr2_f, colkin_f, grid_f = bench.make_problem()
    fresh = CollFreeConst(colkin_f, grid_f, pr2_f)
    fresh_closest = CollFreeConst(colkin_f, grid_f, pr2_f, only_closest_feature=True)
    for a in range(0, n, 9_973):
        b = min(n, a + 9_973)
        v, j = fresh.evaluate(qs[a:b], True)
        assert torch.equal(v, vals[a:b]) and torch.equal(j, jac[a:b])
        assert torch.equal(fresh.is_valid_batch(qs[a:b]), valid[a:b])
        v, j = fresh_closest.evaluate(qs[a:b], True)
        assert torch.equal(v, vc[a:b]) and torch.equal(j, jc[a:b])


def test_cfg2_fetch_pairwise_1m():
    """Fetch 8-DoF PairWiseSelfCollFreeConst, 1 M configurations (closest mode at full size, all-pairs on a slice)."""
    from skmp_b200.robot.fetch import FetchConfig
    from skmp_b200.robot_model import Fetch

    torch = _torch()
    fetch = Fetch(); fetch.reset_pose()
    conf = FetchConfig()
    full = conf.get_pairwise_selcol_consts(fetch)
    closest = conf.get_pairwise_selcol_consts(fetch, only_closest_feature=True)
    kin = full.colkin
    n = 1_000_000
    qs = _device_samples(kin, n, 8)
    vc, jc = closest.evaluate(qs, True)
    assert vc.shape == (n, 1) and jc.shape == (n, 1, 8) and torch.isfinite(vc).all() and torch.isfinite(jc).all()
    perm = torch.randperm(n, device="cuda", generator=torch.Generator("cuda").manual_seed(1))
    vp, jp = closest.evaluate(qs[perm].contiguous(), True)
    assert torch.equal(vp, vc[perm]) and torch.equal(jp, jc[perm])          # permutation equivariance, bit exact
    sl = slice(123_457, 123_457 + 60_000)
    va, ja = full.evaluate(qs[sl], True)
    assert va.shape == (60_000, len(full.check_sphere_id_pairs))
    assert torch.equal(va.min(dim=1).values, vc[sl, 0])
    valid = closest.is_valid_batch(qs)
    assert torch.equal(valid, vc[:, 0] > 0)
    pick = np.arange(0, 60_000, 200)
    tree, _, jl, base = oracle_args(kin)
    sq, gr = oracle.inter_link_sqdists(tree, qs[sl][pick].cpu().numpy().astype(np.float64), np.array(full.check_sphere_id_pairs), jl, base)
    rv = sq - full.check_sphere_pair_sqdists
    assert (np.abs(va[pick].cpu().numpy() - rv) / np.maximum(1, np.abs(rv))).max() < 1e-5
    assert (np.abs(ja[pick].cpu().numpy() - gr) / np.maximum(1, np.abs(gr))).max() < 2e-5


def test_cfg4_sqp_trajectory_batch():
    """4096 PR2 trajectories x 100 waypoints: collision blocks on every waypoint + pose block at the goal."""
    from skmp_b200.batched import TrajectoryConstraintStack
    from skmp_b200.constraint import CollFreeConst, PoseConstraint
    from skmp_b200.robot.pr2 import PR2Config
    from skmp_b200.robot_model import PR2
    from skmp_b200.sdf import Box
    from skmp_b200.tinyfk import RotationType

    torch = _torch()
    pr2 = PR2(); pr2.reset_manip_pose()
    conf = PR2Config()
    colkin = conf.get_collision_kin()
    ineq = CollFreeConst(colkin, Box([0.7, 0.5, 1.2], pos=[0.85, -0.2, 0.9]).sdf, pr2)
    eq = PoseConstraint([np.array([0.7, -0.6, 1.0, 0.0, 0.0, 0.0])], conf.get_endeffector_kin(RotationType.RPY), pr2)
    B, n_wp, D = 4096, 100, 7
    a, b = _device_samples(colkin, B, 9), _device_samples(colkin, B, 10)
    t = torch.linspace(0, 1, n_wp, device="cuda")[None, :, None]
    trajs = (a[:, None, :] * (1 - t) + b[:, None, :] * t).reshape(B, n_wp * D)       # straight lines (trajectory.py:166-169)
    blocks = TrajectoryConstraintStack(D, n_wp, {i: ineq for i in range(n_wp)}).evaluate_batch(trajs)
    goal = TrajectoryConstraintStack(D, n_wp, {n_wp - 1: eq}).evaluate_batch(trajs)
    assert len(blocks) == n_wp and blocks[0][0].shape == (B, 76) and blocks[0][1].shape == (B, 76, 7)
    assert goal[n_wp - 1][0].shape == (B, 6) and goal[n_wp - 1][1].shape == (B, 6, 7)
    for bi, wi in ((0, 0), (17, 50), (4095, 99)):
        q = trajs[bi].reshape(n_wp, D)[wi]
        v, j = ineq.evaluate(q[None], True)
        assert torch.equal(v[0], blocks[wi][0][bi]) and torch.equal(j[0], blocks[wi][1][bi])
    v, j = eq.evaluate(trajs[5].reshape(n_wp, D)[-1:], True)
    assert torch.equal(v[0], goal[n_wp - 1][0][5]) and torch.equal(j[0], goal[n_wp - 1][1][5])


def test_cfg5_jaxon_validity_10m():
    """JAXON floating base (37-DoF): environment collision validity for 10 M configurations (RRT edge validity
    scale; only N bytes leave the kernel), plus env + self collision + 4 XYZW pose rows on a 200 k slice."""
    from skmp_b200.constraint import CollFreeConst, IneqCompositeConst, PairWiseSelfCollFreeConst, PoseConstraint
    from skmp_b200.robot.jaxon import JaxonConfig
    from skmp_b200.robot_model import Jaxon
    from skmp_b200.sdf import Box, UnionSDF
    from skmp_b200.tinyfk import RotationType

    torch = _torch()
    jaxon = Jaxon(); jaxon.reset_manip_pose()
    conf = JaxonConfig()
    colkin = conf.get_collision_kin()
    ground = Box([4.0, 4.0, 0.2], pos=[0.0, 0.0, -0.1])
    table = Box([0.6, 1.0, 0.8], pos=[0.9, 0.0, 0.4])
    env = CollFreeConst(colkin, UnionSDF([ground, table]), jaxon)
    n = 10_000_000
    qs = _device_samples(colkin, n, 11, scale=0.5)
    qs[:, 33] += 1.0                                                    # lift the base: a mix of valid / invalid
    valid = env.is_valid_batch(qs)
    assert valid.shape == (n,) and 0.001 < valid.float().mean().item() < 0.999
    sl = slice(5_000_000, 5_200_000)
    vals, jac = env.evaluate(qs[sl], True)
    assert vals.shape == (200_000, 85) and jac.shape == (200_000, 85, 37)
    assert torch.equal(valid[sl], (vals > 0).all(dim=1))
    # self collision (pair list from the q = 0 filter) and the composite
    selfc = PairWiseSelfCollFreeConst(colkin, jaxon, only_closest_feature=True)
    comp = IneqCompositeConst([env, selfc])
    cv = comp.is_valid_batch(qs[sl])
    vs, _ = selfc.evaluate(qs[sl], False)
    assert torch.equal(cv, ~((vals < 0).any(dim=1)) & ~((vs < 0).any(dim=1)))
    # pose rows: 4 end effectors x XYZW
    efkin = conf.get_endeffector_kin(rot_type=RotationType.XYZW)
    pose = PoseConstraint([np.zeros(7)] * 4, efkin, jaxon)
    pv, pj = pose.evaluate(qs[sl][:50_000], True)
    assert pv.shape == (50_000, 28) and pj.shape == (50_000, 28, 37)
    # oracle spot checks
    pick = np.arange(0, 200_000, 1000)
    q64 = qs[sl][pick].cpu().numpy().astype(np.float64)
    tree, feat, jl, base = oracle_args(colkin)
    rv, rj, kink = oracle.collfree(tree, q64, feat, jl, oracle.Sdf(UnionSDF([ground, table]).primitives()), colkin.get_radius_list(),
                                   base_type=base, with_kink=True)
    assert (np.abs(vals[pick].cpu().numpy() - rv) / np.maximum(1, np.abs(rv))).max() < 2e-5
    ok = kink > 2e-5
    assert (np.abs(jac[pick].cpu().numpy() - rj) / np.maximum(1, np.abs(rj)))[ok].max() < 5e-5
    tree, feat, jl, base = oracle_args(efkin)
    rp, rpj = oracle.solve_fk(tree, q64[:50], feat, jl, base, ROT[RotationType.XYZW])
    got = pv[pick[:50]].cpu().numpy() if False else None
    pv2, pj2 = pose.evaluate(qs[sl][pick[:50]], True)
    assert (np.abs(pv2.cpu().numpy() - rp.reshape(50, -1))).max() < 5e-5
    assert (np.abs(pj2.cpu().numpy() - rpj.reshape(50, 28, 37))).max() < 2e-4

"""GPU tests of the batched callers and of the sharded entry points on one rank:
results must equal the reference's one-configuration-at-a-time loops run through the same constraints, and
the oracle."""
import numpy as np
import pytest

from helpers import oracle_args, sample_q
from oracle import oracle

pytestmark = pytest.mark.gpu


def _scene():
    from skmp_b200.constraint import CollFreeConst
    from skmp_b200.robot.pr2 import PR2Config
    from skmp_b200.robot_model import PR2
    from skmp_b200.sdf import Box

    pr2 = PR2(); pr2.reset_manip_pose()
    conf = PR2Config()
    colkin = conf.get_collision_kin()
    box = Box([0.7, 0.5, 1.2], pos=[0.85, -0.2, 0.9])
    return pr2, conf, colkin, box, CollFreeConst(colkin, box.sdf, pr2)


def test_edge_validity_matches_reference_loop():
    """motion_step_box.py:44-57 semantics, decided by the ORACLE's values per sample."""
    from skmp_b200.batched import edge_samples, is_valid_motion_step_batch

    pr2, conf, colkin, box, const = _scene()
    rng = np.random.default_rng(20)
    q1s, q2s = sample_q(colkin, 400, rng), sample_q(colkin, 400, rng)
    q2s[:50] = q1s[:50] + rng.normal(size=(50, 7)) * 0.05
    step_box = conf.get_default_motion_step_box()
    got = is_valid_motion_step_batch(step_box, q1s, q2s, const, dtype=np.float64)
    samples, offsets = edge_samples(step_box, q1s, q2s)
    tree, feat, jl, base = oracle_args(colkin)
    rv, _ = oracle.collfree(tree, samples, feat, jl, oracle.Sdf(box.primitives()), colkin.get_radius_list(), with_jacobian=False)
    ok = (rv > 0).all(axis=1)
    margin = np.abs(rv).min(axis=1)
    for e in range(len(q1s)):
        sl = slice(offsets[e], offsets[e + 1])
        if (margin[sl] < 1e-9).any():
            continue
        assert got[e] == bool(ok[sl].all()), e
    assert 0.02 < got.mean() < 0.98
    got32 = is_valid_motion_step_batch(step_box, q1s, q2s, const, dtype=np.float32)
    assert (got32 == got).mean() > 0.99


def test_trajectory_stack_with_gpu_constraints():
    """Collision inequality on every waypoint + pose equality at the goal: batched stacking == the reference's
    per-waypoint evaluate_single assembly (trajectory_constraint.py:143-168) through the same GPU constraints."""
    from skmp_b200.batched import TrajectoryConstraintStack
    from skmp_b200.constraint import PoseConstraint
    from skmp_b200.tinyfk import RotationType

    pr2, conf, colkin, box, ineq = _scene()
    efkin = conf.get_endeffector_kin(RotationType.RPY)
    eq = PoseConstraint([np.array([0.7, -0.6, 1.0, 0.0, 0.0, 0.0])], efkin, pr2)
    n_wp, n_dof = 20, 7
    rng = np.random.default_rng(21)
    traj = sample_q(colkin, n_wp, rng)
    for table in ({i: ineq for i in range(n_wp)}, {n_wp - 1: eq}):
        stack = TrajectoryConstraintStack(n_dof, n_wp, table)
        v, J = stack.evaluate(traj.reshape(-1))
        dense = np.zeros((len(v), n_wp * n_dof))
        head, ref_v = 0, []
        for i, c in table.items():
            vi, ji = c.evaluate_single(traj[i], True)
            vi, ji = np.asarray(vi), np.asarray(ji)
            dense[head: head + len(vi), i * n_dof: (i + 1) * n_dof] = ji
            head += len(vi)
            ref_v.append(vi)
        assert np.allclose(v, np.hstack(ref_v), rtol=0, atol=1e-12)      # batch rows vs N=1 launches: same arithmetic
        assert np.allclose(J.toarray(), dense, rtol=0, atol=1e-12)


def test_sharded_entry_points_single_rank():
    from skmp_b200.sharded import sharded_evaluate, sharded_is_valid

    pr2, conf, colkin, box, const = _scene()
    rng = np.random.default_rng(22)
    qs = sample_q(colkin, 1000, rng).astype(np.float32)
    vals, jac, (s, e) = sharded_evaluate(const, qs, True)
    assert (s, e) == (0, 1000) and vals.is_cuda and vals.shape == (1000, 76) and jac.shape == (1000, 76, 7)
    valid = sharded_is_valid(const, qs)
    assert valid.shape == (1000,) and (valid.cpu() == (vals > 0).all(dim=1).cpu()).all()


@pytest.mark.parametrize("include_q1", [True, False])
def test_device_edge_discretisation_is_bit_identical_to_the_reference_loop(include_q1):
    """csrc/skb_edges.cu against the host restatement of motion_step_box.py:8-41 (itself compared with the reference
    loop line by line in test_batched_host.py): same counts, same fp64 samples bit for bit, fp32 = rounded fp64."""
    import torch

    from skmp_b200.batched import edge_samples, edge_samples_device

    rng = np.random.default_rng(0)
    box = np.array([0.05, 0.05, 0.1, 0.2, 0.03, 0.1, 0.5])
    q1s = rng.uniform(-2, 2, size=(3000, 7))
    q2s = q1s + rng.normal(size=(3000, 7)) * rng.choice([1e-8, 0.01, 0.3, 2.0], size=(3000, 1))
    q2s[5] = q1s[5]                                                   # identical points -> no sample
    q2s[6] = q1s[6] + np.array([0.1, 0, 0, 0, 0, 0, 0])               # exactly two steps on the active axis
    q2s[7] = q1s[7] + np.array([0.05 * 7, 0, 0, 0, 0, 0, 0])
    if not include_q1:
        # the reference raises IndexError for a step ratio of exactly 1.0 without q1 (empty list indexed); keep off it
        q2s[6] = q1s[6] + np.array([0.11, 0, 0, 0, 0, 0, 0])
    ref, off = edge_samples(box, q1s, q2s, include_q1)
    got, goff = edge_samples_device(box, q1s, q2s, include_q1, dtype=np.float64)
    assert np.array_equal(goff.cpu().numpy(), off)
    assert np.array_equal(got.cpu().numpy(), ref)
    got32, _ = edge_samples_device(box, torch.as_tensor(q1s, device="cuda"), torch.as_tensor(q2s, device="cuda"), include_q1)
    assert got32.dtype == torch.float32 and np.array_equal(got32.cpu().numpy(), ref.astype(np.float32))
    e0, o0 = edge_samples_device(box, q1s[:0], q2s[:0], include_q1)
    assert e0.shape == (0, 7) and o0.tolist() == [0]


def test_edge_validity_cuda_inputs_return_cuda_bools():
    import torch

    from skmp_b200.batched import is_valid_motion_step_batch

    pr2, conf, colkin, box, const = _scene()
    rng = np.random.default_rng(21)
    q1s, q2s = sample_q(colkin, 2000, rng), sample_q(colkin, 2000, rng)
    step_box = conf.get_default_motion_step_box()
    a = is_valid_motion_step_batch(step_box, q1s, q2s, const)
    b = is_valid_motion_step_batch(step_box, torch.as_tensor(q1s, device="cuda"), torch.as_tensor(q2s, device="cuda"), const)
    assert isinstance(a, np.ndarray) and b.is_cuda and b.dtype == torch.bool
    assert np.array_equal(a, b.cpu().numpy())


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_collfree_csc_blocks_are_the_transposed_jacobians(dtype):
    """``skb_collfree_csc`` / ``CollFreeConst.evaluate_csc``: the transposed (N, D, S) blocks written by the kernel equal the
    (N, S, D) Jacobians of ``evaluate`` bit for bit -- box, union and voxel grid SDFs, D = 7 (odd), D = 14, ragged batch
    sizes; D = 37 (JAXON) goes through the transpose fallback."""
    import torch

    import bench
    from skmp_b200.constraint import CollFreeConst
    from skmp_b200.robot.jaxon import JaxonConfig
    from skmp_b200.robot.pr2 import PR2Config
    from skmp_b200.robot_model import Jaxon
    from skmp_b200.sdf import Box, Cylinder, UnionSDF

    pr2, conf, colkin, box, const = _scene()
    pr2_d, colkin_d, grid = bench.make_problem()
    union = UnionSDF([box, Cylinder(0.15, 0.8, pos=[0.4, -0.6, 1.0])])
    jaxon = Jaxon(); jaxon.reset_manip_pose()
    jkin = JaxonConfig().get_collision_kin()
    cases = [(const, colkin), (CollFreeConst(colkin, union, pr2, distance_margin=0.02), colkin),
             (CollFreeConst(colkin_d, grid, pr2_d), colkin_d), (CollFreeConst(colkin_d, union, pr2_d), colkin_d),
             (CollFreeConst(jkin, UnionSDF([Box([4.0, 4.0, 0.2], pos=[0.0, 0.0, -0.1])]), jaxon), jkin)]
    rng = np.random.default_rng(70)
    for c, kin in cases:
        for n in (1, 31, 1003):
            q = torch.as_tensor(sample_q(kin, n, rng).astype(dtype), device="cuda")
            v, j = c.evaluate(q, True)
            vc, d = c.evaluate_csc(q)
            assert d.shape == (n, kin.dim_cspace, kin.n_feature) and d.is_contiguous()
            assert torch.equal(vc, v) and torch.equal(d, j.transpose(1, 2))
    # numpy in -> numpy out (host path: transpose of the host result)
    qn = sample_q(colkin, 17, rng).astype(dtype)
    vn, dn = const.evaluate_csc(qn)
    v, j = const.evaluate(qn, True)
    assert isinstance(dn, np.ndarray) and np.array_equal(dn, np.swapaxes(np.asarray(j), 1, 2)) and np.array_equal(vn, v)


def test_trajectory_batch_stack_csc_equals_reference_dense_assembly():
    """``TrajectoryBatchStack.evaluate_batch_csc``: for every trajectory of the batch, (data, indices, indptr) is the CSC
    form of the dense matrix the reference assembles waypoint by waypoint (trajectory_constraint.py:128-195, restated in
    tests/test_batched_host.py::ref_trajectory_evaluate) -- homogeneous collision table (the zero-copy path), and a mixed
    table in non-monotone insertion order (pose goal, collision on some waypoints, a configuration point)."""
    import torch
    from scipy.sparse import csc_matrix

    from skmp_b200.batched import TrajectoryBatchStack, TrajectoryConstraintStack
    from skmp_b200.constraint import ConfigPointConst, PoseConstraint
    from skmp_b200.tinyfk import RotationType
    from test_batched_host import ref_trajectory_evaluate

    pr2, conf, colkin, box, ineq = _scene()
    eq = PoseConstraint([np.array([0.7, -0.6, 1.0, 0.0, 0.0, 0.0])], conf.get_endeffector_kin(RotationType.RPY), pr2)
    cp = ConfigPointConst(np.linspace(-0.3, 0.3, 7))
    n_dof, n_wp, B = 7, 6, 5
    rng = np.random.default_rng(71)
    trajs = sample_q(colkin, B * n_wp, rng).reshape(B, n_wp * n_dof)
    tables = [{i: ineq for i in range(n_wp)}, {n_wp - 1: eq, 0: cp, 2: ineq, 4: ineq, 1: eq}]
    for table in tables:
        stack = TrajectoryBatchStack(n_dof, n_wp, table)
        values, data = stack.evaluate_batch_csc(torch.as_tensor(trajs, device="cuda"))
        assert values.shape == (B, stack.shape[0]) and data.shape == (B, len(stack.indices))
        vh, dh = stack.evaluate_batch_csc(trajs)                                 # numpy in -> numpy out
        assert np.array_equal(vh, values.cpu().numpy()) and np.array_equal(dh, data.cpu().numpy())
        for b in range(B):
            rv, rJ, dense = ref_trajectory_evaluate(n_dof, n_wp, table, trajs[b])
            J = csc_matrix((dh[b], stack.indices, stack.indptr), shape=stack.shape)
            assert J.shape == dense.shape
            assert np.abs(vh[b] - rv).max() < 1e-12 and np.abs(J.toarray() - dense).max() < 1e-12
            assert np.array_equal(stack.csc_matrix(data[b]).toarray(), J.toarray())
        # the single-trajectory front end (the reference's call shape) rides on the same path
        v1, J1 = TrajectoryConstraintStack(n_dof, n_wp, table).evaluate(trajs[0])
        rv, rJ, dense = ref_trajectory_evaluate(n_dof, n_wp, table, trajs[0])
        assert np.abs(v1 - rv).max() < 1e-12 and np.abs(J1.toarray() - dense).max() < 1e-12
        assert np.array_equal(J1.indices, rJ.indices) or J1.nnz >= rJ.nnz       # structural zeros are stored, like determine_sparse_pattern's fill
    # fp32 batch: same layout, values within fp32 of the fp64 run
    stack = TrajectoryBatchStack(n_dof, n_wp, tables[0])
    v32, d32 = stack.evaluate_batch_csc(torch.as_tensor(trajs.astype(np.float32), device="cuda"))
    v64, d64 = stack.evaluate_batch_csc(torch.as_tensor(trajs.astype(np.float32).astype(np.float64), device="cuda"))
    assert v32.dtype == torch.float32 and float((v32.double() - v64).abs().max()) < 1e-5

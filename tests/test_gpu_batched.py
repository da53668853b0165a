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

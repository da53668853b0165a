"""Host logic of the batched callers, CPU only: the discretisation and the CSC assembly
are compared with literal restatements of the reference loops (skmp/solver/motion_step_box.py:8-57,
skmp/solver/nlp_solver/trajectory_constraint.py:128-195) on constraints that need no GPU."""
import numpy as np
import pytest
from scipy.sparse import csc_matrix

from skmp_b200.batched import (TrajectoryConstraintStack, edge_samples, interpolate_fractions_batch,
                               is_valid_motion_step_batch)
from skmp_b200.constraint import BoxConst, ConfigPointConst, IneqCompositeConst


def ref_interpolate_fractions(motion_step_box, q1, q2, include_q1):
    """motion_step_box.py:8-41, restated line by line."""
    diff = q2 - q1
    active_idx = np.argmax(np.abs(diff) / motion_step_box)
    diff_active_axis = diff[active_idx]
    if abs(diff_active_axis) < 1e-6:
        return []
    step_ratio = motion_step_box[active_idx] / abs(diff_active_axis)
    if step_ratio > 1.0:
        return [1.0]
    travel_rate, fractions = 0.0, []
    if include_q1:
        fractions.append(travel_rate)
    while travel_rate + step_ratio < 1.0:
        travel_rate += step_ratio
        fractions.append(travel_rate)
    if abs(fractions[-1] - 1) > 1e-6:
        fractions.append(1.0)
    return fractions


@pytest.mark.parametrize("include_q1", [True, False])
def test_interpolate_fractions_match_reference_loop(include_q1):
    rng = np.random.default_rng(0)
    box = np.array([0.05, 0.05, 0.1, 0.2, 0.03, 0.1, 0.5])
    q1s = rng.uniform(-2, 2, size=(300, 7))
    q2s = q1s + rng.normal(size=(300, 7)) * rng.choice([1e-8, 0.01, 0.3, 2.0], size=(300, 1))
    q2s[5] = q1s[5]                                                   # identical points -> no sample
    q2s[6] = q1s[6] + np.array([0.1, 0, 0, 0, 0, 0, 0])               # exactly two steps on the active axis
    q2s[7] = q1s[7] + np.array([0.05 * 7, 0, 0, 0, 0, 0, 0])
    fractions, offsets = interpolate_fractions_batch(box, q1s, q2s, include_q1)
    assert offsets[0] == 0 and offsets[-1] == len(fractions)
    for e in range(len(q1s)):
        ref = ref_interpolate_fractions(box, q1s[e], q2s[e], include_q1) if include_q1 or True else None
        got = fractions[offsets[e]: offsets[e + 1]]
        assert len(got) == len(ref), e
        assert np.array_equal(got, np.array(ref)), e                  # bit-identical (sequential running sum)
    samples, off2 = edge_samples(box, q1s, q2s, include_q1)
    assert np.array_equal(off2, offsets) and samples.shape == (len(fractions), 7)


def test_edge_validity_matches_reference_loop_on_box_constraint():
    rng = np.random.default_rng(1)
    box = np.full(4, 0.1)
    const = BoxConst(-np.ones(4), np.ones(4))
    q1s = rng.uniform(-0.9, 0.9, size=(200, 4))
    q2s = rng.uniform(-1.4, 1.4, size=(200, 4))
    q2s[:3] = q1s[:3]
    got = is_valid_motion_step_batch(box, q1s, q2s, const, dtype=np.float64)
    for e in range(len(q1s)):                                          # motion_step_box.py:44-57
        ok = True
        for frac in ref_interpolate_fractions(box, q1s[e], q2s[e], True):
            fs, _ = const.evaluate_single(q1s[e] + (q2s[e] - q1s[e]) * frac, False)
            if not np.all(fs > 0.0):
                ok = False
                break
        assert got[e] == ok, e
    assert got[:3].all() and 0.1 < got.mean() < 0.9


def ref_trajectory_evaluate(n_dof, n_wp, table, traj_vector):
    """trajectory_constraint.py:128-168 + CachedCscMatrix, restated (local constraints only)."""
    traj = traj_vector.reshape(-1, n_dof)
    values, jac = [], {}
    for i in table.keys():
        v, j = table[i].evaluate_single(traj[i], True)
        values.append(v)
        jac[i] = j
    total = np.hstack(values)
    dense = np.zeros((len(total), len(traj_vector)))
    head = 0
    for i in table.keys():
        dense[head: head + jac[i].shape[0], i * n_dof: (i + 1) * n_dof] = jac[i]
        head += jac[i].shape[0]
    return total, csc_matrix(dense), dense


def test_trajectory_stack_equals_reference_dense_assembly():
    n_dof, n_wp = 5, 12
    rng = np.random.default_rng(2)
    ineq = BoxConst(-np.ones(n_dof), np.ones(n_dof) * 2)
    table = {i: ineq for i in range(n_wp)}                             # homogeneous inequality over all waypoints
    traj = rng.normal(size=n_wp * n_dof)
    stack = TrajectoryConstraintStack(n_dof, n_wp, table)
    v, J = stack.evaluate(traj)
    rv, rJ, dense = ref_trajectory_evaluate(n_dof, n_wp, table, traj)
    assert np.array_equal(v, rv) and J.shape == rJ.shape and np.array_equal(J.toarray(), dense)
    assert J.nnz == n_wp * 2 * n_dof * n_dof                           # full blocks, like determine_sparse_pattern's 1.0 fill
    v2, J2 = stack.evaluate(traj * 0.5)                                # pattern reused
    assert np.array_equal(J2.indices, J.indices) and np.array_equal(J2.indptr, J.indptr)
    # start/goal equality constraints in NON-monotone insertion order: rows follow the table, columns the waypoints
    c0, cg = ConfigPointConst(np.zeros(n_dof)), ConfigPointConst(np.ones(n_dof))
    table2 = {n_wp - 1: cg, 0: c0, 3: cg}
    stack2 = TrajectoryConstraintStack(n_dof, n_wp, table2)
    v, J = stack2.evaluate(traj)
    rv, rJ, dense = ref_trajectory_evaluate(n_dof, n_wp, table2, traj)
    assert np.array_equal(v, rv) and np.array_equal(J.toarray(), dense)
    with pytest.raises(ValueError):
        TrajectoryConstraintStack(n_dof, n_wp, {n_wp: c0})


def test_trajectory_stack_batch_blocks():
    n_dof, n_wp, B = 4, 6, 9
    rng = np.random.default_rng(3)
    comp = IneqCompositeConst([BoxConst(-np.ones(n_dof), np.ones(n_dof))])
    stack = TrajectoryConstraintStack(n_dof, n_wp, {i: comp for i in range(n_wp)})
    trajs = rng.normal(size=(B, n_wp * n_dof))
    blocks = stack.evaluate_batch(trajs)
    for b in range(0, B, 4):
        v, J = stack.evaluate(trajs[b])
        dense = J.toarray()
        for i in range(n_wp):
            vi, ji = blocks[i]
            m = vi.shape[1]
            assert np.array_equal(vi[b], v[i * m: (i + 1) * m])
            assert np.array_equal(ji[b], dense[i * m: (i + 1) * m, i * n_dof: (i + 1) * n_dof])

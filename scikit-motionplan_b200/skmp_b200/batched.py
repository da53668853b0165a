"""Batched callers of the hot path (SURVEY.md 8f "next" rows 1 and 2).

The reference feeds the constraint API one configuration at a time:
  * edge validity  -- ``is_valid_motion_step`` evaluates every interpolated sample of an edge with
    ``evaluate_single`` in a Python loop                       (skmp/solver/motion_step_box.py:8-57)
  * SQP stacking   -- ``TrajectoryConstraint.evaluate`` calls ``evaluate_single`` per waypoint, fills a dense
    (M_total x n_wp*D) matrix and converts it to CSC          (skmp/solver/nlp_solver/trajectory_constraint.py:128-195)
Here both become ONE kernel launch per distinct constraint, with identical results:
  * ``interpolate_fractions_batch`` / ``is_valid_motion_step_batch``: all samples of all edges in one
    ``is_valid_batch`` call + a per-edge AND reduction (BASELINE config 5: RRT edge validity);
  * ``TrajectoryConstraintStack``: per-waypoint blocks from one batched ``evaluate`` and the block-diagonal CSC
    ``data`` array written directly (BASELINE config 4: 4096 trajectories x 100 waypoints).
"""
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from ._array import _is_torch, torch


# ------------------------------------------------------------------ edge discretisation
def interpolate_fractions_batch(motion_step_box: np.ndarray, q1s: np.ndarray, q2s: np.ndarray,
                                include_q1: bool = True) -> Tuple[np.ndarray, np.ndarray]:
    """Vectorised ``interpolate_fractions`` (motion_step_box.py:8-41) for E edges.

    Returns (fractions (n_samples,), offsets (E+1,)): edge e owns ``fractions[offsets[e]:offsets[e+1]]``.
    Per edge this reproduces the reference list exactly, including its quirks: an edge shorter than 1e-6 on
    its active axis yields no sample; an edge shorter than one step yields only [1.0] (q1 is dropped even
    when ``include_q1``); the running sum ``travel_rate += step_ratio`` is accumulated sequentially."""
    box = np.asarray(motion_step_box, dtype=np.float64)
    q1s, q2s = np.atleast_2d(np.asarray(q1s, dtype=np.float64)), np.atleast_2d(np.asarray(q2s, dtype=np.float64))
    assert q1s.shape == q2s.shape and q1s.shape[1] == len(box)
    E = len(q1s)
    diff = q2s - q1s
    active = np.argmax(np.abs(diff) / box, axis=1)
    d_act = np.abs(diff[np.arange(E), active])
    too_close = d_act < 1e-6
    with np.errstate(divide="ignore"):
        step = box[active] / d_act
    single = (~too_close) & (step > 1.0)
    multi = (~too_close) & ~single
    counts = np.zeros(E, dtype=np.int64)
    counts[single] = 1
    frac_rows: Dict[int, np.ndarray] = {}
    if multi.any():
        idx = np.nonzero(multi)[0]
        st = step[idx]
        kmax = int(np.ceil(1.0 / st.min())) + 2
        # travel[:, k] = value of travel_rate after k sequential additions (cumsum adds left to right)
        travel = np.concatenate([np.zeros((len(idx), 1)), np.cumsum(np.repeat(st[:, None], kmax, axis=1), axis=1)], axis=1)
        # the while loop appends travel[k+1] as long as travel[k] + step < 1.0
        take = (travel[:, :-1] + st[:, None]) < 1.0
        n_steps = np.where(take.all(axis=1), kmax, np.argmin(take, axis=1))      # first False = number of appended steps
        for row, e in enumerate(idx):
            fr = travel[row, 1: n_steps[row] + 1]
            if include_q1:
                fr = np.concatenate([[0.0], fr])
            if len(fr) == 0 or abs(fr[-1] - 1.0) > 1e-6:
                fr = np.concatenate([fr, [1.0]])
            frac_rows[int(e)] = fr
            counts[e] = len(fr)
    offsets = np.concatenate([[0], np.cumsum(counts)])
    fractions = np.empty(offsets[-1])
    for e in np.nonzero(single)[0]:
        fractions[offsets[e]] = 1.0
    for e, fr in frac_rows.items():
        fractions[offsets[e]: offsets[e + 1]] = fr
    return fractions, offsets


def edge_samples(motion_step_box, q1s, q2s, include_q1: bool = True):
    """All interpolated configurations of all edges -> (samples (n_samples, D), offsets (E+1,))."""
    q1s, q2s = np.atleast_2d(np.asarray(q1s, dtype=np.float64)), np.atleast_2d(np.asarray(q2s, dtype=np.float64))
    fractions, offsets = interpolate_fractions_batch(motion_step_box, q1s, q2s, include_q1)
    edge_of = np.repeat(np.arange(len(q1s)), np.diff(offsets))
    samples = q1s[edge_of] + (q2s[edge_of] - q1s[edge_of]) * fractions[:, None]      # q1 + (q2 - q1) * frac, as the reference
    return samples, offsets


def edge_samples_device(motion_step_box, q1s, q2s, include_q1: bool = True, dtype=np.float32, device=None):
    """``edge_samples`` on the GPU (csrc/skb_edges.cu): -> (samples (n_samples, D) CUDA tensor of ``dtype``, offsets
    (E+1,) int64 CUDA tensor).  fp64 sequential fraction arithmetic, bit-identical to the host version; q1s / q2s may
    be numpy arrays or CUDA tensors (fp64 is used for the interpolation either way)."""
    import ctypes as C

    from . import _lib

    if torch is None or not torch.cuda.is_available():
        raise RuntimeError("edge_samples_device needs a CUDA device (no CPU fallback)")
    dev = (q1s.device if _is_torch(q1s) and q1s.is_cuda else torch.device("cuda", torch.cuda.current_device())) if device is None else device
    q1 = torch.as_tensor(q1s, device=dev).to(torch.float64).reshape(-1, len(motion_step_box)).contiguous()
    q2 = torch.as_tensor(q2s, device=dev).to(torch.float64).reshape(-1, len(motion_step_box)).contiguous()
    assert q1.shape == q2.shape
    box = torch.as_tensor(np.asarray(motion_step_box, dtype=np.float64), device=dev)
    E, D = q1.shape
    L = _lib.lib()
    p = lambda t: C.c_void_p(t.data_ptr())
    with torch.cuda.device(dev):
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        counts = torch.zeros(E, dtype=torch.int64, device=dev)
        _lib.check(L.skb_edge_counts(p(q1), p(q2), p(box), E, D, int(include_q1), p(counts), st))
        offsets = torch.zeros(E + 1, dtype=torch.int64, device=dev)
        torch.cumsum(counts, dim=0, out=offsets[1:])
        n = int(offsets[-1].item()) if E else 0                      # the one scalar read-back: output allocation
        tdt = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32
        samples = torch.empty((n, D), dtype=tdt, device=dev)
        if n:
            _lib.check(L.skb_edge_samples(p(q1), p(q2), p(box), E, D, int(include_q1), p(offsets),
                                          _lib.F64 if tdt == torch.float64 else _lib.F32, p(samples), st))
    return samples, offsets


def is_valid_motion_step_batch(motion_step_box, q1s, q2s, ineq_const, dtype=np.float32, device=None):
    """``is_valid_motion_step`` (motion_step_box.py:44-57) for E edges at once -> (E,) bool.

    An edge is valid iff every sample satisfies ``all(values > 0)``; an edge with no sample is valid.  With a CUDA
    device everything stays on the GPU: edge discretisation (``skb_edge_counts`` / ``skb_edge_samples``), ONE
    ``is_valid_batch`` launch per constraint over all samples, per-edge AND (``skb_segment_all``); the result is a
    numpy array for numpy inputs and a CUDA bool tensor for CUDA tensor inputs."""
    if torch is not None and torch.cuda.is_available():
        import ctypes as C

        from . import _lib

        samples, offsets = edge_samples_device(motion_step_box, q1s, q2s, True, dtype, device)
        E = offsets.numel() - 1
        out = torch.ones(E, dtype=torch.uint8, device=samples.device)
        if samples.shape[0]:
            ok = ineq_const.is_valid_batch(samples).to(torch.uint8).contiguous()
            with torch.cuda.device(samples.device):
                _lib.check(_lib.lib().skb_segment_all(C.c_void_p(ok.data_ptr()), C.c_void_p(offsets.data_ptr()), E,
                                                      C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        res = out.bool()
        return res if (_is_torch(q1s) and q1s.is_cuda) else res.cpu().numpy()
    samples, offsets = edge_samples(motion_step_box, q1s, q2s, include_q1=True)
    E = len(offsets) - 1
    if len(samples) == 0:
        return np.ones(E, dtype=bool)
    ok = np.asarray(ineq_const.is_valid_batch(samples.astype(dtype)))      # host constraints (BoxConst ...) only
    out = np.ones(E, dtype=bool)
    nz = np.diff(offsets) > 0
    out[nz] = np.minimum.reduceat(ok.astype(np.uint8), offsets[:-1][nz]).astype(bool)
    return out


# ------------------------------------------------------------------ trajectory-constraint stacking for SQP
class TrajectoryBatchStack:
    """Batched ``TrajectoryConstraint.evaluate`` (skmp/solver/nlp_solver/trajectory_constraint.py:128-195) for the
    local-constraint table, for B trajectories at once, entirely on the GPU.

    ``local_constraint_table`` maps waypoint index -> constraint (insertion order = row order, like the reference's
    ``local_table.keys()`` loop).  The stacked Jacobian of one trajectory is block diagonal, (M_total, n_wp * n_dof): the
    block of waypoint i sits in rows ``head_i .. head_i + M_i`` and columns ``i * n_dof .. (i + 1) * n_dof``.  Its CSC
    pattern is fixed by the table (the reference's ``determine_sparse_pattern``, :34-47): ``indices`` / ``indptr`` are built
    once; ``evaluate_batch_csc`` returns ``values (B, M_total)`` and ``data (B, nnz)`` with ``data[b]`` the CSC ``data``
    array of trajectory b -- for waypoints in ascending order the (n_dof, M_i) transpose of the waypoint's Jacobian.

    One launch per distinct constraint over all B x (its waypoints) rows.  A ``CollFreeConst`` writes its transposed blocks
    straight into ``data`` (``evaluate_csc``); when one constraint covers every waypoint -- the SQP solver's global
    inequality constraint -- nothing is gathered, copied or transposed at all."""

    def __init__(self, n_dof: int, n_wp: int, local_constraint_table: Dict[int, object]):
        self.n_dof, self.n_wp = n_dof, n_wp
        self.table = dict(local_constraint_table)
        for i in self.table:
            if not (0 <= i < n_wp):
                raise ValueError("index {} exceeds the waypoint number {}".format(i, n_wp))
        self._layout = None
        self._dev_index = {}

    # groups: constraint object -> waypoint indices, in table order
    def _groups(self):
        groups: Dict[int, Tuple[object, List[int]]] = {}
        for i, const in self.table.items():
            groups.setdefault(id(const), (const, []))[1].append(i)
        return list(groups.values())

    @staticmethod
    def _rows_of(const, n_dof: int) -> int:
        if hasattr(const, "colkin") and hasattr(const, "sdf"):                    # CollFreeConst
            return 1 if const.only_closest_feature else const.colkin.n_feature
        if hasattr(const, "check_sphere_id_pairs"):                               # PairWiseSelfCollFreeConst
            return 1 if const.only_closest_feature else len(const.check_sphere_id_pairs)
        if hasattr(const, "efkin") and hasattr(const, "desired_poses"):           # PoseConstraint
            return const.efkin.n_feature * const.efkin.dim_tspace
        v, _ = const.evaluate(np.zeros((1, n_dof)), False)                       # anything else: ask it once
        return int(v.shape[1])

    def layout(self):
        """(heads {waypoint: first row}, m_total, indices (nnz,), indptr (n_wp * n_dof + 1,), data offsets {waypoint: offset})."""
        if self._layout is None:
            dims = {i: self._rows_of(c, self.n_dof) for i, c in self.table.items()}
            heads, head = {}, 0
            for i in self.table:                       # row order = table order
                heads[i] = head
                head += dims[i]
            indptr = np.zeros(self.n_dof * self.n_wp + 1, dtype=np.int64)
            indices, doff, off = [], {}, 0
            for i in range(self.n_wp):                 # column order = waypoint order
                m = dims.get(i, 0)
                if m:
                    doff[i] = off
                    off += m * self.n_dof
                for d in range(self.n_dof):
                    indptr[i * self.n_dof + d + 1] = indptr[i * self.n_dof + d] + m
                    if m:
                        indices.append(heads[i] + np.arange(m))
            self._layout = (heads, head, np.concatenate(indices) if indices else np.zeros(0, dtype=np.int64), indptr, doff, dims)
        return self._layout

    @property
    def indices(self) -> np.ndarray:
        return self.layout()[2]

    @property
    def indptr(self) -> np.ndarray:
        return self.layout()[3]

    @property
    def shape(self) -> Tuple[int, int]:
        return self.layout()[1], self.n_dof * self.n_wp

    def evaluate_batch_csc(self, traj_vectors):
        """``traj_vectors`` (B, n_wp * n_dof), numpy or CUDA torch (fp32 / fp64) -> (values (B, M_total), data (B, nnz)) of the
        same array kind; with ``indices`` / ``indptr`` each row is one trajectory's CSC Jacobian."""
        if torch is None or not torch.cuda.is_available():
            raise RuntimeError("TrajectoryBatchStack needs a CUDA device (no CPU fallback)")
        is_np = not _is_torch(traj_vectors)
        t = torch.as_tensor(np.ascontiguousarray(traj_vectors)) if is_np else traj_vectors
        if t.dtype not in (torch.float32, torch.float64):
            t = t.to(torch.float64)
        if not t.is_cuda:
            t = t.cuda()
        B = t.shape[0]
        trajs = t.reshape(B, self.n_wp, self.n_dof)
        heads, m_total, indices, _, doff, dims = self.layout()
        nnz = len(indices)
        groups = self._groups()
        values = data = None
        if len(groups) == 1 and list(self.table) == sorted(self.table):
            # one constraint on waypoints in ascending order: its outputs ARE the stacked arrays (rows follow the table,
            # columns the waypoints -- the same order here); every waypoint: not even a gather of the rows
            const, idxs = groups[0]
            if len(idxs) == self.n_wp:
                rows = trajs.reshape(B * self.n_wp, self.n_dof)
            elif len(idxs) == 1:
                rows = trajs[:, idxs[0], :].contiguous()
            else:
                rows = trajs[:, idxs, :].reshape(B * len(idxs), self.n_dof)
            v, d = self._block(const, rows)
            values, data = v.reshape(B, m_total), d.reshape(B, nnz)
        else:
            values = torch.empty((B, m_total), dtype=t.dtype, device=t.device)
            data = torch.empty((B, nnz), dtype=t.dtype, device=t.device)
            for const, idxs in groups:
                order = sorted(idxs)
                rows = trajs[:, order, :].reshape(B * len(order), self.n_dof)
                v, d = self._block(const, rows)
                m = dims[order[0]]
                vcols, dcols = self._positions(const, order, heads, doff, m, t.device)
                values.index_copy_(1, vcols, v.reshape(B, len(order) * m))
                data.index_copy_(1, dcols, d.reshape(B, len(order) * m * self.n_dof))
        if is_np:
            # device -> pinned host, both copies in flight together, one synchronisation (a pageable destination would be
            # staged through the driver's bounce buffer at a few GB/s)
            hv = torch.empty(values.shape, dtype=values.dtype, pin_memory=True)
            hd = torch.empty(data.shape, dtype=data.dtype, pin_memory=True)
            hv.copy_(values, non_blocking=True)
            hd.copy_(data, non_blocking=True)
            torch.cuda.current_stream(values.device).synchronize()
            return hv.numpy(), hd.numpy()
        return values, data

    def _block(self, const, rows):
        """values (n, M), transposed Jacobian blocks (n, n_dof, M)."""
        if hasattr(const, "evaluate_csc"):
            return const.evaluate_csc(rows)
        v, j = const.evaluate(rows, True)
        return v, j.transpose(1, 2).contiguous()

    def _positions(self, const, order, heads, doff, m, device):
        key = (id(const), str(device))
        pos = self._dev_index.get(key)
        if pos is None:
            vcols = np.concatenate([heads[i] + np.arange(m) for i in order])
            dcols = np.concatenate([doff[i] + np.arange(m * self.n_dof) for i in order])
            pos = self._dev_index[key] = (torch.as_tensor(vcols, device=device), torch.as_tensor(dcols, device=device))
        return pos

    def csc_matrix(self, data_row):
        """scipy CSC matrix of one trajectory from its ``data`` row (host side, for a QP solver that wants scipy)."""
        from scipy.sparse import csc_matrix

        d = data_row.cpu().numpy() if _is_torch(data_row) else np.asarray(data_row)
        return csc_matrix((d.astype(np.float64), self.indices, self.indptr), shape=self.shape)


class TrajectoryConstraintStack:
    """Batched ``TrajectoryConstraint.evaluate`` (trajectory_constraint.py:128-195) for the local-constraint table.

    ``local_constraint_table`` maps waypoint index -> constraint (insertion order = row order, like the
    reference's ``local_table.keys()`` loop).  Waypoints sharing one constraint object are evaluated in ONE
    batched call.  The Jacobian is block diagonal; its CSC pattern is fixed by the table, so ``indices`` /
    ``indptr`` are built once (the reference's ``determine_sparse_pattern``) and ``evaluate`` only writes ``data``.
    """

    def __init__(self, n_dof: int, n_wp: int, local_constraint_table: Dict[int, object]):
        self.n_dof, self.n_wp = n_dof, n_wp
        self.table = dict(local_constraint_table)
        for i in self.table:
            if not (0 <= i < n_wp):
                raise ValueError("index {} exceeds the waypoint number {}".format(i, n_wp))
        self._pattern = None
        self._gpu_stack = None

    # groups: constraint object -> waypoint indices, in table order
    def _groups(self):
        groups: Dict[int, Tuple[object, List[int]]] = {}
        for i, const in self.table.items():
            groups.setdefault(id(const), (const, []))[1].append(i)
        return list(groups.values())

    def _blocks(self, traj: np.ndarray):
        """waypoint index -> (values (M_i,), jac (M_i, n_dof)) as numpy fp64, one launch per distinct constraint."""
        out = {}
        for const, idxs in self._groups():
            vals, jac = const.evaluate(traj[idxs], True)
            if _is_torch(vals):
                vals, jac = vals.cpu().numpy(), jac.cpu().numpy()
            for k, i in enumerate(idxs):
                out[i] = (np.asarray(vals[k], dtype=np.float64), np.asarray(jac[k], dtype=np.float64))
        return out

    def _build_pattern(self, dims: Dict[int, int]):
        heads, head = {}, 0
        for i in self.table:                       # row order = table order
            heads[i] = head
            head += dims[i]
        indptr = np.zeros(self.n_dof * self.n_wp + 1, dtype=np.int64)
        indices = []
        for i in range(self.n_wp):                 # column order = waypoint order
            m = dims.get(i, 0)
            for d in range(self.n_dof):
                indptr[i * self.n_dof + d + 1] = indptr[i * self.n_dof + d] + m
                if m:
                    indices.append(heads[i] + np.arange(m))
        self._pattern = (heads, head, np.concatenate(indices) if indices else np.zeros(0, dtype=np.int64), indptr, dict(dims))

    def evaluate(self, traj_vector: np.ndarray):
        """-> (values (M_total,), scipy.sparse.csc_matrix (M_total, n_wp * n_dof)), equal to the reference's.  With a CUDA
        device the blocks and the CSC ``data`` array come from ``TrajectoryBatchStack`` (B = 1): one launch per distinct
        constraint, nothing is looped over or transposed on the host."""
        from scipy.sparse import csc_matrix

        traj = np.asarray(traj_vector, dtype=np.float64).reshape(-1, self.n_dof)
        assert len(traj) == self.n_wp
        if torch is not None and torch.cuda.is_available():
            if self._gpu_stack is None:
                self._gpu_stack = TrajectoryBatchStack(self.n_dof, self.n_wp, self.table)
            st = self._gpu_stack
            values, data = st.evaluate_batch_csc(traj.reshape(1, -1))
            return values[0], csc_matrix((data[0], st.indices, st.indptr), shape=st.shape)
        blocks = self._blocks(traj)
        dims = {i: len(blocks[i][0]) for i in self.table}
        if self._pattern is None or self._pattern[4] != dims:
            self._build_pattern(dims)
        heads, m_total, indices, indptr, _ = self._pattern
        values = np.concatenate([blocks[i][0] for i in self.table]) if self.table else np.zeros(0)
        # column (i, d) holds jac_i[:, d]: per waypoint the (n_dof, M_i) transpose, waypoints in ascending order
        data = [blocks[i][1].T.reshape(-1) for i in sorted(self.table)]
        data = np.concatenate(data) if data else np.zeros(0)
        return values, csc_matrix((data, indices, indptr), shape=(m_total, self.n_dof * self.n_wp))

    def evaluate_batch(self, traj_vectors):
        """B trajectories at once (BASELINE config 4).  ``traj_vectors`` (B, n_wp * n_dof), numpy or CUDA torch.
        -> dict waypoint index -> (values (B, M_i), jac (B, M_i, n_dof)): the diagonal blocks of every trajectory's
        stacked Jacobian, produced by one launch per distinct constraint over all B x (its waypoints) rows.
        Arrays stay on the device when the input is a CUDA tensor."""
        B = traj_vectors.shape[0]
        trajs = traj_vectors.reshape(B, self.n_wp, self.n_dof)
        out = {}
        for const, idxs in self._groups():
            rows = trajs[:, idxs, :].reshape(B * len(idxs), self.n_dof)
            vals, jac = const.evaluate(rows, True)
            vals = vals.reshape(B, len(idxs), -1)
            jac = jac.reshape(B, len(idxs), -1, self.n_dof)
            for k, i in enumerate(idxs):
                out[i] = (vals[:, k], jac[:, k])
        return out

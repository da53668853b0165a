"""Batched multi-seed constraint satisfaction (IK / projection) -- skmp/satisfy.py:45-143 for many seeds at once.

The reference solves ONE seed with SLSQP and restarts from random seeds on failure (up to 20 / 30 / 100 sequential
trials: skmp/satisfy.py:119-143, solver/ompl_solver.py:146-154, solver/myrrt_solver.py:152-160).  Here all trials
run together: every iteration is one batched ``evaluate(Q (B, D))`` of the equality and inequality constraints on
the GPU (values + Jacobians from the CUDA kernels) followed by a batched projected Levenberg-Marquardt step

    minimise  |e(q)|^2 + w^2 |min(g(q) - m, 0)|^2      subject to  lb <= q <= ub

with a per-seed damping factor; violated inequality rows enter as hinge residuals, box bounds by projection.
The success test is the reference's (:108-112): |e|^2 < acceptable_error, inequality rows valid, inside the box.
Same public names: SatisfactionConfig, SatisfactionResult, satisfy_by_optimization(_with_budget); the batch entry
point is satisfy_by_optimization_batch.  The step itself (normal equations, Cholesky, bound projection) is the CUDA
kernel ``skb_lm_step`` (csrc/skb_lm.cu), one warp per seed; a whole iteration is captured in a CUDA graph and replayed.
"""
import time
from dataclasses import dataclass
from typing import Optional

import numpy as np

from ._array import torch
from .constraint import AbstractEqConst, AbstractIneqConst, BoxConst, ConfigPointConst


@dataclass
class SatisfactionConfig:
    ftol: float = 1e-6
    disp: bool = False
    n_max_eval: int = 200
    acceptable_error: float = 1e-6
    # batch-only knobs
    n_seeds: int = 64
    ineq_weight: float = 30.0
    margin_numerical: float = 1e-6
    use_cuda_graph: bool = True


@dataclass
class SatisfactionResult:
    q: np.ndarray
    elapsed_time: float
    success: bool


@dataclass
class BatchSatisfactionResult:
    qs: np.ndarray          # (B, D) final iterates
    success: np.ndarray     # (B,) bool
    eq_error: np.ndarray    # (B,) |e|^2 (0 when there is no equality constraint)
    ineq_min: np.ndarray    # (B,) min inequality value (+inf when there is none)
    n_iter: int
    n_eval: int
    elapsed_time: float

    def best(self) -> SatisfactionResult:
        """First successful seed (seed 0 is the caller's), else the one with the smallest equality error."""
        ok = np.flatnonzero(self.success)
        i = int(ok[0]) if len(ok) else int(np.argmin(self.eq_error + np.maximum(0.0, -self.ineq_min)))
        return SatisfactionResult(self.qs[i].copy(), self.elapsed_time, bool(self.success[i]))


def _device():
    if torch is None or not torch.cuda.is_available():
        raise RuntimeError("skmp_b200.satisfy needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _lm_step(J, r, lam, Q, lb, ub):
    """skb_lm_step: q_new = clamp(q + dq), (J^T J + lam (I + diag J^T J)) dq = -J^T r  for every seed."""
    import ctypes as C

    from . import _lib

    J, r = J.contiguous(), r.contiguous()
    B, M, D = J.shape
    out = torch.empty_like(Q)
    p = lambda t: C.c_void_p(t.data_ptr())
    with torch.cuda.device(Q.device):
        _lib.check(_lib.lib().skb_lm_step(_lib.F64, p(J), p(r), p(lam), p(Q), p(lb), p(ub), p(out), B, M, D,
                                          C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out


def satisfy_by_optimization_batch(eq_const: Optional[AbstractEqConst], box_const: BoxConst,
                                  ineq_const: Optional[AbstractIneqConst], q_seeds=None,
                                  config: Optional[SatisfactionConfig] = None, rng=None) -> BatchSatisfactionResult:
    """Run every row of ``q_seeds`` (B, D) -- or ``config.n_seeds`` uniform samples of the box -- to convergence."""
    ts = time.time()
    config = config or SatisfactionConfig()
    dev = _device()
    f64 = torch.float64
    lb = torch.as_tensor(np.asarray(box_const.lb, dtype=np.float64), device=dev)
    ub = torch.as_tensor(np.asarray(box_const.ub, dtype=np.float64), device=dev)
    D = lb.numel()
    if q_seeds is None:
        rng = rng or np.random.default_rng()
        q_seeds = rng.uniform(size=(config.n_seeds, D)) * (box_const.ub - box_const.lb) + box_const.lb
    Q = torch.as_tensor(np.atleast_2d(np.asarray(q_seeds, dtype=np.float64)), device=dev).clone()
    B = Q.shape[0]
    if isinstance(eq_const, ConfigPointConst):
        q = np.tile(np.asarray(eq_const.desired_angles, dtype=np.float64), (B, 1))
        return BatchSatisfactionResult(q, np.ones(B, bool), np.zeros(B), np.full(B, np.inf), 0, 0, time.time() - ts)
    Q = torch.minimum(torch.maximum(Q, lb), ub)
    Q0 = Q.clone()
    w, m = config.ineq_weight, config.margin_numerical
    eye = torch.eye(D, dtype=f64, device=dev)

    def residuals(Qc):
        """-> stacked residual (B, M), Jacobian (B, M, D), |e|^2 (B,), min inequality value (B,)"""
        if eq_const is not None:
            e, Je = eq_const.evaluate(Qc, True)
        else:                                   # projection: stay close to the seed (ConfigPointConst(q_seed), :61-64)
            e, Je = Qc - Q0, eye.expand(B, D, D)
        eq_err = (e * e).sum(dim=1)
        if ineq_const is None:
            return e, Je, eq_err, torch.full((B,), float("inf"), dtype=f64, device=dev)
        g, Jg = ineq_const.evaluate(Qc, True)
        # a small safety band keeps accepted iterates strictly inside (values are metres or squared metres)
        viol = g < (m + 1e-4)
        h = torch.where(viol, g - (m + 2e-4), torch.zeros_like(g)) * w
        Jh = Jg * (viol.to(f64) * w)[:, :, None]
        return torch.cat([e, h], dim=1), torch.cat([Je, Jh], dim=1), eq_err, g.min(dim=1).values

    lam = torch.full((B,), 1e-2, dtype=f64, device=dev)
    r, J, eq_err, gmin = residuals(Q)
    r, J, eq_err, gmin = r.clone(), J.contiguous().clone(), eq_err.clone(), gmin.clone()
    cost = (r * r).sum(dim=1)
    n_eval, n_iter = 1, 0
    done = torch.zeros(B, dtype=torch.bool, device=dev)
    stalled = torch.zeros(B, dtype=torch.bool, device=dev)
    need_eq = eq_const is not None

    def finished():
        if need_eq:
            return (eq_err < config.acceptable_error) & (gmin > m)
        hinge_off = (r[:, D:] == 0).all(dim=1) if r.shape[1] > D else torch.ones_like(done)
        return (gmin > m) & hinge_off            # projection: feasible with every hinge term inactive

    def iteration():
        """One LM iteration on the persistent state, no host synchronisation (so it can be captured in a CUDA graph)."""
        done.copy_(finished())
        Qn = _lm_step(J, r, lam, Q, lb, ub)      # CUDA: normal equations + Cholesky + bound projection, warp per seed
        Qn = torch.where(done[:, None], Q, Qn)
        rn, Jn, eq_n, gmin_n = residuals(Qn)
        cost_n = (rn * rn).sum(dim=1)
        better = (cost_n < cost) & ~done
        stalled.copy_(~better & ~done & (lam >= 1e6))
        lam.copy_(torch.where(better, lam / 3.0, lam * 4.0).clamp(1e-9, 1e6))
        Q.copy_(torch.where(better[:, None], Qn, Q))
        r.copy_(torch.where(better[:, None], rn, r))
        J.copy_(torch.where(better[:, None, None], Jn, J))
        eq_err.copy_(torch.where(better, eq_n, eq_err))
        gmin.copy_(torch.where(better, gmin_n, gmin))
        cost.copy_(torch.where(better, cost_n, cost))

    # The iteration is launch-bound (two constraint kernels, the LM step and a dozen element-wise updates on B x D data),
    # so it is captured once in a CUDA graph and replayed; convergence is polled every CHECK replays.
    CHECK = 8
    graph = None
    if config.use_cuda_graph:
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                iteration()                       # warm-up: lazy plan uploads, allocator pools
                n_eval += 1
            torch.cuda.current_stream(dev).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                iteration()
        except Exception:
            graph = None                          # capture not possible (e.g. a host-side constraint): eager replays
            torch.cuda.synchronize(dev)
    n_iter = 1 if graph is not None else 0
    while n_iter < config.n_max_eval - 1:
        for _ in range(min(CHECK, config.n_max_eval - 1 - n_iter)):
            graph.replay() if graph is not None else iteration()
            n_iter += 1
            n_eval += 1
        if bool((finished() | stalled).all()):
            break
    ok_eq = eq_err < config.acceptable_error if need_eq else torch.ones_like(done)
    inside = ((Q >= lb) & (Q <= ub)).all(dim=1)
    success = ok_eq & (gmin > 0) & inside
    return BatchSatisfactionResult(Q.cpu().numpy(), success.cpu().numpy(), (eq_err if need_eq else torch.zeros_like(cost)).cpu().numpy(),
                                   gmin.cpu().numpy(), n_iter, n_eval, time.time() - ts)


def satisfy_by_optimization(eq_const: Optional[AbstractEqConst], box_const: BoxConst,
                            ineq_const: Optional[AbstractIneqConst], q_seed: Optional[np.ndarray],
                            config: Optional[SatisfactionConfig] = None) -> SatisfactionResult:
    """Drop-in for skmp/satisfy.py:45-116.  The caller's seed (when given) is row 0 of the batch; for an IK problem
    the remaining ``config.n_seeds - 1`` rows are random restarts evaluated in the same launches, for a projection
    (``eq_const is None``) only the caller's seed is meaningful."""
    config = config or SatisfactionConfig()
    if isinstance(eq_const, ConfigPointConst):
        return SatisfactionResult(eq_const.desired_angles, 0.0, True)
    lbn, ubn = np.asarray(box_const.lb, dtype=np.float64), np.asarray(box_const.ub, dtype=np.float64)
    if eq_const is None:
        assert q_seed is not None
        seeds = np.asarray(q_seed, dtype=np.float64)[None]
    else:
        n_rand = config.n_seeds - (0 if q_seed is None else 1)
        seeds = np.random.uniform(size=(max(0, n_rand), len(lbn))) * (ubn - lbn) + lbn
        if q_seed is not None:
            seeds = np.vstack([np.asarray(q_seed, dtype=np.float64)[None], seeds])
    return satisfy_by_optimization_batch(eq_const, box_const, ineq_const, seeds, config).best()


def satisfy_by_optimization_with_budget(eq_const: AbstractEqConst, box_const: BoxConst,
                                        ineq_const: Optional[AbstractIneqConst], q_seed: Optional[np.ndarray],
                                        config: Optional[SatisfactionConfig] = None,
                                        n_trial_budget: int = 20) -> SatisfactionResult:
    """skmp/satisfy.py:119-143: the budget of sequential restarts becomes the batch size of one call."""
    config = config or SatisfactionConfig()
    cfg = SatisfactionConfig(**{**config.__dict__, "n_seeds": max(1, n_trial_budget)})
    return satisfy_by_optimization(eq_const, box_const, ineq_const, q_seed, cfg)

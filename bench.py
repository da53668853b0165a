#!/usr/bin/env python
"""Benchmark of the hot path on the BASELINE.json configurations: configs/sec, with the roofline of the dominant kernel and
the reference's CPU path timed beside it.

  python bench.py --gpus N --steps K --warmup W [--workload cfg3]     # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ... [--workload ..]   # the reference's CPU algorithm (oracle port)

Workloads (``--workload``; the default, cfg3, is the configuration BASELINE.json's metric is quoted on):
  cfg3  PR2 dual-arm 14-DoF, 152 spheres, CollFreeConst vs a 128^3 voxel-grid SDF, values (N,152) + Jacobians (N,152,14).
        The named 10 M configurations are sharded over 8 GPUs: every rank processes 1.25 M (weak scaling, --gpus 8 is
        exactly the named config; --configs-per-gpu 10000000 puts all of it on one GPU).
  cfg2  Fetch 8-DoF PairWiseSelfCollFreeConst, all pairs, values (N,P) + Jacobians (N,P,8), 1 M configurations per GPU.
  cfg4  SQP trajectory batch: 4096 PR2 right-arm trajectories x 100 waypoints per GPU -- CollFreeConst rows of every
        waypoint + PoseConstraint (RPY) rows of every goal waypoint, and the block-diagonal CSC ``data`` arrays of the
        stacked Jacobians written on the GPU (trajectory_constraint.py:128-195).
  cfg5  JAXON 37-DoF floating base, RRT edge validity: environment + self-collision validity bytes (the composite
        short-circuits in member order like the reference) and PoseConstraint (4 x XYZW) rows, 1.25 M configurations per
        GPU (10 M / 8); under torchrun the validity bytes are all-gathered over NCCL (``gather_ms``, reported beside
        the step time, not inside it).

One JSON line on stdout (rank 0).  ``value`` times the kernels with inputs resident in HBM (CUDA events, max over ranks);
``e2e`` goes through the public API with numpy host buffers (H2D + D2H inside the timed region); ``roofline`` is the
dominant kernel of the step against the measured HBM peak; ``cpu_baseline`` is the oracle's C port of the reference
algorithm on all host threads.  See DESIGN.md "Measurement".
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
for p in (ROOT, ROOT / "scikit-motionplan_b200"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

S, D = 152, 14
BYTES_PER_CONFIG = 4 * (D + S + S * D)          # cfg3 algorithmic bytes: q row in, values + Jacobian out
WORKLOAD = ("cfg3: PR2 dual-arm 14-DoF CollFreeConst vs 128^3 voxel-grid SDF, values (N,152) + Jacobians (N,152,14), "
            "fp32, {n} configs per GPU (10M / 8 GPUs)")


# ------------------------------------------------------------------ scenes
def scene_primitives():
    """Union of 4 boxes inside the workspace AABB [-0.5,1.5]x[-1.5,1.5]x[0,2] (seed 0)."""
    from skmp_b200.rotmath import rotation_matrix
    from skmp_b200.sdf import Box, UnionSDF

    rng = np.random.default_rng(0)
    boxes = []
    table = Box([0.7, 0.5, 1.2]); table.translate([0.85, -0.2, 0.9]); boxes.append(table)     # the reference's test obstacle
    for _ in range(3):
        b = Box(rng.uniform(0.2, 0.6, 3))
        b.translate(rng.uniform([0.2, -1.0, 0.3], [1.2, 1.0, 1.7]))
        b.rotate_with_matrix(rotation_matrix(rng.uniform(-0.6, 0.6), rng.normal(size=3)))
        boxes.append(b)
    return UnionSDF(boxes)


GRID_LB, GRID_UB, GRID_DIMS = np.array([-0.5, -1.5, 0.0]), np.array([1.5, 1.5, 2.0]), (128, 128, 128)


def grid_values_cpu(union) -> np.ndarray:
    """128^3 fp32 samples of the union of boxes, in plain numpy (scene generation only: neither the CUDA engine nor the
    oracle is involved).  Box SDF: |max(|p| - h, 0)| + min(max_i(|p_i| - h_i), 0) in the box frame; union = min."""
    spacing = (GRID_UB - GRID_LB) / (np.array(GRID_DIMS) - 1)
    axes = [GRID_LB[i] + spacing[i] * np.arange(GRID_DIMS[i]) for i in range(3)]
    pts = np.stack(np.meshgrid(*axes, indexing="ij"), axis=-1).reshape(-1, 3)
    best = np.full(len(pts), np.inf)
    for kind, pos, rot, half in union.primitives():
        assert kind == "box"
        local = (pts - np.asarray(pos)) @ np.asarray(rot)            # R^T (p - pos), row-vector form
        s = np.abs(local) - np.asarray(half)
        d = np.linalg.norm(np.maximum(s, 0.0), axis=1) + np.minimum(s.max(axis=1), 0.0)
        best = np.minimum(best, d)
    return best.reshape(GRID_DIMS).astype(np.float32)


def make_problem():
    from skmp_b200.robot.pr2 import PR2Config
    from skmp_b200.robot_model import PR2
    from skmp_b200.sdf import GridSDF

    pr2 = PR2()
    pr2.reset_manip_pose()
    colkin = PR2Config(control_arm="dual").get_collision_kin()
    colkin.reflect_skrobot_model(pr2)
    union = scene_primitives()
    spacing = (GRID_UB - GRID_LB) / (np.array(GRID_DIMS) - 1)
    grid = GridSDF(grid_values_cpu(union), GRID_LB, spacing, fill_value=1.0e3)
    return pr2, colkin, grid


def joint_bounds(colkin, base_lo=-1.0, base_hi=1.0, scale=1.0):
    """Uniform sampling box: URDF joint limits (continuous joints +-2 pi), base dofs in [base_lo, base_hi]."""
    jm = colkin.fksolver.urdf.joint_map
    lo = [-2 * np.pi if jm[n].limit.lower is None else jm[n].limit.lower for n in colkin.control_joint_names]
    hi = [2 * np.pi if jm[n].limit.upper is None else jm[n].limit.upper for n in colkin.control_joint_names]
    nb = colkin.dim_cspace - len(lo)
    return np.array(lo + [base_lo] * nb) * scale, np.array(hi + [base_hi] * nb) * scale


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).  NVML through
    pynvml (~1 ms per sample) so that even a 30 ms timed region gets tens of samples; falls back to the
    nvidia-smi query line of the recipe when pynvml is missing."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self._stop_evt = index, threading.Event()
        self.sm, self.mx, self.reasons, self.how = [], [], set(), "nvml"
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self.nv, self.how = None, "nvidia-smi"

    def _nvml_once(self):
        nv = self.nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
        self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)))
        try:
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for name, bit in (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40)):
            if r & bit:
                self.reasons.add(name)

    def _smi_once(self):
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        if out:
            r = [c.strip() for c in out.split(",")]
            self.sm.append(float(r[0])); self.mx.append(float(r[1]))
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)

    def run(self):
        while not self._stop_evt.is_set():
            try:
                self._nvml_once() if self.nv is not None else self._smi_once()
            except Exception:
                pass
            self._stop_evt.wait(0.002 if self.nv is not None else 0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "how": self.how}


# ------------------------------------------------------------------ CPU reference (oracle port), threaded
def _threaded_rate(work, qs, threads: int, block: int = 2048):
    """Rows/s of ``work(rows)`` (a ctypes call into the oracle: releases the GIL) over ``threads`` host threads, each walking
    its contiguous shard in blocks and dropping the outputs (timing only)."""
    shards = np.array_split(qs, max(1, threads))

    def run(shard):
        for i in range(0, len(shard), block):
            work(shard[i:i + block])
        return len(shard)

    t0 = time.perf_counter()
    if threads <= 1:
        run(qs)
    else:
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(run, shards))
    return time.perf_counter() - t0


def cpu_rate(work, sample, seconds: float, threads: int):
    """Time ``work`` on a bounded sample sized for about ``seconds`` of wall clock -> (rows/s, rows, seconds)."""
    probe = sample(512)
    t0 = time.perf_counter()
    work(probe)
    per_row = (time.perf_counter() - t0) / len(probe)
    n = int(max(threads * 256, min(4_000_000, seconds * threads / per_row)))
    qs = sample(n)
    dt = _threaded_rate(work, qs, threads)
    return n / dt, n, dt


# ------------------------------------------------------------------ workloads
class Workload:
    """One BASELINE.json configuration: device step, algorithmic bytes, e2e call, CPU reference."""

    name = ""
    metric = "configs/sec"
    dtype = "f32"
    default_n = 0

    def describe(self, n):               # config.workload
        raise NotImplementedError

    def setup(self, dev, rank, n):       # build constraints + resident inputs / outputs; returns nothing
        raise NotImplementedError

    def tune(self):                      # explicit launch tuning (skb_plan_tune) during warm-up
        pass

    def step(self):                      # one pass of the hot path over the resident batch
        raise NotImplementedError

    def kernels(self):                   # [(label, algorithmic bytes per launch, event pair index)] for the roofline
        raise NotImplementedError

    def e2e(self, n):                    # -> (callable doing one public-API call on host buffers, n, h2d bytes, d2h bytes, api string)
        raise NotImplementedError

    def cpu(self):                       # -> (work(rows), sample(n), description)
        raise NotImplementedError

    def extra_config(self):
        return {}


def _rand_rows(lo, hi, seed):
    def sample(n):
        rng = np.random.default_rng(seed)
        return rng.uniform(size=(n, len(lo))) * (hi - lo) + lo
    return sample


def _device_uniform(lo, hi, n, dev, seed):
    import torch

    gen = torch.Generator(device=dev).manual_seed(seed)
    lo_t, hi_t = torch.tensor(lo, dtype=torch.float32, device=dev), torch.tensor(hi, dtype=torch.float32, device=dev)
    return torch.rand((n, len(lo)), generator=gen, device=dev, dtype=torch.float32) * (hi_t - lo_t) + lo_t


class Cfg3(Workload):
    name = "cfg3"
    metric = "configs/sec (CollFreeConst value+Jacobian)"
    default_n = 1_250_000

    def describe(self, n):
        return WORKLOAD.format(n=n)

    def setup(self, dev, rank, n):
        import torch

        from skmp_b200 import _lib
        from skmp_b200.constraint import CollFreeConst

        self.pr2, self.colkin, self.grid = make_problem()
        self.const = CollFreeConst(self.colkin, self.grid, self.pr2)
        self.lo, self.hi = joint_bounds(self.colkin)
        self.n, self.dev = n, dev
        self.qs = _device_uniform(self.lo, self.hi, n, dev, 1000 + rank)      # each rank owns a different shard
        self.plan = self.const._plan()
        self.L = _lib.lib()
        self.vals = torch.empty((n, S), dtype=torch.float32, device=dev)
        self.jac = torch.empty((n, S, D), dtype=torch.float32, device=dev)
        self.stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        self._lib = _lib

    def tune(self):
        self.const.tune(self.qs)

    def step(self):
        _lib = self._lib
        _lib.check(self.L.skb_collfree(self.plan.handle, self.grid.native(), _lib.F32, C.c_void_p(self.qs.data_ptr()), self.n, 0.0,
                                       _lib.MODE_ALL, C.c_void_p(self.vals.data_ptr()), C.c_void_p(self.jac.data_ptr()), None, self.stream))

    def kernels(self):
        return [("skb::s2_kernel<float, COLLFREE, grid-fast, VW2, u32, all, jac, NP<=8> (skb_sphere2.cu)", BYTES_PER_CONFIG * self.n)]

    def traffic(self):
        tp = ROOT / "profiles" / "traffic_collfree_cfg3.json"
        if tp.exists():
            t = json.loads(tp.read_text())
            # per-launch DRAM bytes of the ncu --set full capture, scaled to this launch's configuration count
            return t.get("dram_bytes_per_launch") * self.n / t.get("configs_per_launch", 1_250_000), \
                "profiles/traffic_collfree_cfg3.json (static: one ncu --set full capture, not measured in this run)"
        return None, None

    def e2e(self, n):
        import torch

        n = min(self.n, n)
        q_host = torch.empty((n, D), dtype=torch.float32, pin_memory=True)
        q_host.copy_(self.qs[:n])
        q_np = q_host.numpy()
        return (lambda: self.const.evaluate(q_np, True)), n, n * D * 4, n * (S + S * D) * 4, \
            "CollFreeConst.evaluate(numpy (N,14) float32, with_jacobian=True) -> numpy"

    def cpu(self):
        from oracle import oracle

        s = self.colkin.fksolver
        tree = oracle.Tree.from_solver(s)
        feat = np.array(self.colkin.tinyfk_feature_ids)
        jl = s.joint_child_links(self.colkin.tinyfk_joint_ids)
        sdf = oracle.Sdf(self.grid.primitives())
        radii = self.colkin.get_radius_list()
        return (lambda rows: oracle.collfree(tree, rows, feat, jl, sdf, radii, grad_mode=oracle.GRAD_FD)), _rand_rows(self.lo, self.hi, 0), \
            "C port of the reference algorithm (tree-walk FK + Jacobian, forward-difference SDF gradient, fp64)"

    def extra_config(self):
        valid_frac = float((self.vals.min(dim=1).values > 0).float().mean())
        return {"spheres": S, "dof": D, "grid": "128^3 fp32",
                "l2": "outputs (%.1f GB/step) and inputs exceed the 126 MB L2; no flush needed" % (self.n * (S + S * D) * 4 / 1e9),
                "valid_fraction": valid_frac}


class Cfg2(Workload):
    name = "cfg2"
    metric = "configs/sec (PairWiseSelfCollFreeConst value+Jacobian, all pairs)"
    default_n = 1_000_000

    def describe(self, n):
        return (f"cfg2: Fetch 8-DoF arm+torso PairWiseSelfCollFreeConst, all {self.P} sphere pairs (q=0 filter), values (N,{self.P}) + "
                f"Jacobians (N,{self.P},8), fp32, {n} configs per GPU")

    def setup(self, dev, rank, n):
        from skmp_b200.robot.fetch import FetchConfig
        from skmp_b200.robot_model import Fetch

        fetch = Fetch(); fetch.reset_pose()
        self.const = FetchConfig().get_pairwise_selcol_consts(fetch)
        self.kin = self.const.colkin
        self.P, self.Dd = len(self.const.check_sphere_id_pairs), self.kin.dim_cspace
        self.lo, self.hi = joint_bounds(self.kin)
        self.n = n
        self.qs = _device_uniform(self.lo, self.hi, n, dev, 2000 + rank)
        self.out = None

    def tune(self):
        self.const.tune(self.qs)

    def step(self):
        self.out = None                                # free the previous step's 27 GB before allocating the next
        self.out = self.const.evaluate(self.qs, True)

    def kernels(self):
        return [("skb::s2_kernel<float, PAIRWISE, all, jac, NP<=4> (skb_sphere2.cu)", 4 * (self.Dd + self.P + self.P * self.Dd) * self.n)]

    def e2e(self, n):
        import torch

        n = min(self.n, n // 4)
        q_host = torch.empty((n, self.Dd), dtype=torch.float32, pin_memory=True)
        q_host.copy_(self.qs[:n])
        q_np = q_host.numpy()
        return (lambda: self.const.evaluate(q_np, True)), n, n * self.Dd * 4, n * (self.P + self.P * self.Dd) * 4, \
            "PairWiseSelfCollFreeConst.evaluate(numpy (N,8) float32, with_jacobian=True) -> numpy"

    def cpu(self):
        from oracle import oracle

        s = self.kin.fksolver
        tree = oracle.Tree.from_solver(s)
        jl = s.joint_child_links(self.kin.tinyfk_joint_ids)
        pairs = np.array(self.const.check_sphere_id_pairs)
        return (lambda rows: oracle.inter_link_sqdists(tree, rows, pairs, jl)), _rand_rows(self.lo, self.hi, 0), \
            "C port of tinyfk.compute_inter_link_sqdists (tree-walk FK + Jacobians of both spheres of every pair, fp64)"

    def extra_config(self):
        return {"pairs": self.P, "dof": self.Dd, "l2": "outputs (%.1f GB/step) exceed the 126 MB L2; no flush needed" % (self.n * (self.P + self.P * self.Dd) * 4 / 1e9)}


class Cfg4(Workload):
    name = "cfg4"
    metric = "configs/sec (SQP batch: stacked collision + pose constraint Jacobians, CSC data)"
    default_n = 409_600
    N_WP = 100

    def describe(self, n):
        return (f"cfg4: SQP trajectory batch, {n // self.N_WP} PR2 right-arm trajectories x {self.N_WP} waypoints per GPU: CollFreeConst "
                f"(76 spheres vs box) on every waypoint + PoseConstraint (RPY, goal (0.7,-0.6,1.0)) on the last, values + block-diagonal "
                f"CSC data of the stacked Jacobians, fp32")

    def setup(self, dev, rank, n):
        import torch

        from skmp_b200.batched import TrajectoryBatchStack
        from skmp_b200.constraint import CollFreeConst, PoseConstraint
        from skmp_b200.robot.pr2 import PR2Config
        from skmp_b200.robot_model import PR2
        from skmp_b200.sdf import Box
        from skmp_b200.tinyfk import RotationType

        pr2 = PR2(); pr2.reset_manip_pose()
        conf = PR2Config()
        self.colkin = conf.get_collision_kin()
        self.box = Box([0.7, 0.5, 1.2], pos=[0.85, -0.2, 0.9])
        self.ineq = CollFreeConst(self.colkin, self.box.sdf, pr2)
        self.efkin = conf.get_endeffector_kin(RotationType.RPY)
        self.eq = PoseConstraint([np.array([0.7, -0.6, 1.0, 0.0, 0.0, 0.0])], self.efkin, pr2)
        self.lo, self.hi = joint_bounds(self.colkin)
        self.B, self.n = n // self.N_WP, (n // self.N_WP) * self.N_WP
        a = _device_uniform(self.lo, self.hi, self.B, dev, 4000 + rank)
        b = _device_uniform(self.lo, self.hi, self.B, dev, 5000 + rank)
        t = torch.linspace(0, 1, self.N_WP, device=dev)[None, :, None]
        self.trajs = (a[:, None, :] * (1 - t) + b[:, None, :] * t).reshape(self.B, self.N_WP * 7).contiguous()   # straight lines (trajectory.py:166-169)
        self.stack_ineq = TrajectoryBatchStack(7, self.N_WP, {i: self.ineq for i in range(self.N_WP)})
        self.stack_eq = TrajectoryBatchStack(7, self.N_WP, {self.N_WP - 1: self.eq})
        self.out = None

    def tune(self):
        self.ineq.tune(self.trajs.reshape(-1, 7), csc=True)

    def step(self):
        self.out = None
        self.out = (self.stack_ineq.evaluate_batch_csc(self.trajs), self.stack_eq.evaluate_batch_csc(self.trajs))

    def kernels(self):
        col = 4 * (7 + 76 + 76 * 7) * self.n
        return [("skb::s2_kernel<float, COLLFREE, box, all, jac, CSC rows> + pose_kernel<RPY> (goal waypoints)", col + 4 * (7 + 6 + 42) * self.B)]

    def e2e(self, n):
        b = max(1, min(self.B, n // self.N_WP // 4))
        t_np = self.trajs[:b].cpu().numpy()
        si, se = self.stack_ineq, self.stack_eq
        return (lambda: (si.evaluate_batch_csc(t_np), se.evaluate_batch_csc(t_np))), b * self.N_WP, t_np.nbytes * 2, \
            b * (self.N_WP * (76 + 76 * 7) + 6 + 42) * 4, \
            "TrajectoryBatchStack.evaluate_batch_csc(numpy (B, 700) float32) -> numpy values + CSC data (collision and goal pose stacks)"

    def cpu(self):
        from oracle import oracle

        s = self.colkin.fksolver
        tree = oracle.Tree.from_solver(s)
        feat = np.array(self.colkin.tinyfk_feature_ids)
        jl = s.joint_child_links(self.colkin.tinyfk_joint_ids)
        sdf = oracle.Sdf(self.box.primitives())
        radii = self.colkin.get_radius_list()
        s2 = self.efkin.fksolver
        tree2, feat2, jl2 = oracle.Tree.from_solver(s2), np.array(self.efkin.tinyfk_feature_ids), s2.joint_child_links(self.efkin.tinyfk_joint_ids)

        def work(rows):                       # every waypoint: collision rows; one waypoint in N_WP: the goal pose rows
            oracle.collfree(tree, rows, feat, jl, sdf, radii, grad_mode=oracle.GRAD_FD)
            oracle.solve_fk(tree2, rows[:: self.N_WP], feat2, jl2, 0, oracle.ROT_RPY)
        return work, _rand_rows(self.lo, self.hi, 0), \
            "C port of the reference algorithm per waypoint (FK + Jacobian, forward-difference box SDF gradient; goal pose rows for 1 waypoint in 100), fp64; the reference's dense stacking and CSC conversion are NOT included"

    def extra_config(self):
        return {"trajectories": self.B, "waypoints": self.N_WP, "dof": 7, "spheres": 76,
                "l2": "outputs (%.2f GB/step) exceed the 126 MB L2; no flush needed" % (self.n * (76 + 76 * 7) * 4 / 1e9)}


class Cfg5(Workload):
    name = "cfg5"
    metric = "configs/sec (JAXON env + self collision validity + PoseConstraint rows)"
    default_n = 1_250_000

    def describe(self, n):
        return (f"cfg5: JAXON 37-DoF floating base, RRT edge validity: env CollFreeConst (85 spheres vs ground + table) and "
                f"PairWiseSelfCollFreeConst ({self.P} pairs) validity bytes (composite, short-circuit in member order) + PoseConstraint "
                f"4 x XYZW values (N,28) + Jacobians (N,28,37), fp32, {n} configs per GPU (10M / 8 GPUs)")

    def setup(self, dev, rank, n):
        from skmp_b200.constraint import CollFreeConst, IneqCompositeConst, PairWiseSelfCollFreeConst, PoseConstraint
        from skmp_b200.robot.jaxon import JaxonConfig
        from skmp_b200.robot_model import Jaxon
        from skmp_b200.sdf import Box, UnionSDF
        from skmp_b200.tinyfk import RotationType

        jaxon = Jaxon(); jaxon.reset_manip_pose()
        conf = JaxonConfig()
        self.colkin = conf.get_collision_kin()
        self.env_sdf = UnionSDF([Box([4.0, 4.0, 0.2], pos=[0.0, 0.0, -0.1]), Box([0.6, 1.0, 0.8], pos=[0.9, 0.0, 0.4])])
        self.env = CollFreeConst(self.colkin, self.env_sdf, jaxon)
        self.selfc = PairWiseSelfCollFreeConst(self.colkin, jaxon, only_closest_feature=True)
        self.comp = IneqCompositeConst([self.env, self.selfc])
        self.efkin = conf.get_endeffector_kin(rot_type=RotationType.XYZW)
        self.pose = PoseConstraint([np.zeros(7)] * 4, self.efkin, jaxon)
        self.P = len(self.selfc.check_sphere_id_pairs)
        self.lo, self.hi = joint_bounds(self.colkin, scale=0.5)
        self.lo[33] += 1.0; self.hi[33] += 1.0                  # lift the base: a mix of valid / invalid (as tests/test_gpu_scale.py)
        self.n = n
        self.qs = _device_uniform(self.lo, self.hi, n, dev, 6000 + rank)
        self.valid, self.out = None, None

    def tune(self):
        self.env.tune(self.qs, with_jacobian=False, values=True, validity=False)       # the composite asks for the row minimum
        self.selfc.tune(self.qs, with_jacobian=False, values=True, validity=False)

    def step(self):
        self.out = None
        self.valid = self.comp.is_valid_batch(self.qs)
        self.out = self.pose.evaluate(self.qs, True)

    def kernels(self):
        return [("skb::pose_kernel<float, XYZW> (skb_pose.cu) + s2_kernel<COLLFREE / PAIRWISE, closest value> validity", (4 * (37 + 28 + 28 * 37) + 2 * 4 * 37 + 2) * self.n)]

    def e2e(self, n):
        n = min(self.n, n)
        q_np = self.qs[:n].cpu().numpy()
        return (lambda: (self.comp.is_valid_batch(q_np), self.pose.evaluate(q_np, True))), n, 2 * q_np.nbytes, n * (1 + (28 + 28 * 37) * 4), \
            "IneqCompositeConst([env, self]).is_valid_batch(numpy) + PoseConstraint.evaluate(numpy, True) -> numpy"

    def cpu(self):
        from oracle import oracle

        s = self.colkin.fksolver
        tree = oracle.Tree.from_solver(s)
        feat = np.array(self.colkin.tinyfk_feature_ids)
        jl = s.joint_child_links(self.colkin.tinyfk_joint_ids)
        sdf = oracle.Sdf(self.env_sdf.primitives())
        radii = self.colkin.get_radius_list()
        pairs = np.array(self.selfc.check_sphere_id_pairs)
        s2 = self.efkin.fksolver
        tree2, feat2, jl2 = oracle.Tree.from_solver(s2), np.array(self.efkin.tinyfk_feature_ids), s2.joint_child_links(self.efkin.tinyfk_joint_ids)

        def work(rows):
            v, _ = oracle.collfree(tree, rows, feat, jl, sdf, radii, base_type=oracle.BASE_FLOATING, with_jacobian=False)
            alive = rows[(v >= 0).all(axis=1)]                  # the reference's composite short-circuits: constraint.py:167-176
            if len(alive):
                oracle.inter_link_sqdists(tree, alive, pairs, jl, oracle.BASE_FLOATING, with_jacobian=False)
            oracle.solve_fk(tree2, rows, feat2, jl2, oracle.BASE_FLOATING, oracle.ROT_XYZW)
        return work, _rand_rows(self.lo, self.hi, 0), \
            "C port of the reference algorithm (env collision values, self-collision squared distances of the survivors, 4 XYZW pose rows + Jacobians), fp64"

    def extra_config(self):
        return {"dof": 37, "spheres": 85, "pairs": self.P, "valid_fraction": float(self.valid.float().mean()) if self.valid is not None else None,
                "l2": "pose outputs (%.1f GB/step) exceed the 126 MB L2; no flush needed" % (self.n * (28 + 28 * 37) * 4 / 1e9)}


WORKLOADS = {"cfg2": Cfg2, "cfg3": Cfg3, "cfg4": Cfg4, "cfg5": Cfg5}


# ------------------------------------------------------------------ reference arm
def run_reference(args):
    """--impl reference: rank 0 times the reference CPU algorithm (oracle port) on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]()
    n = args.configs_per_gpu or w.default_n
    work, sample, what, desc = reference_problem(args.workload, n)
    threads = os.cpu_count() or 1
    per_step_seconds = max(1.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    rates, n_last = [], 0
    for i in range(args.warmup + args.steps):
        rate, n_last, dt = cpu_rate(work, sample, per_step_seconds, threads)
        if i >= args.warmup:
            rates.append((n_last, dt))
    total_n, total_t = sum(r[0] for r in rates), sum(r[1] for r in rates)
    value = total_n / total_t
    smp = f"{n_last} configs per step of the same workload (uniform within joint limits, seed 0): {what}, {threads} threads"
    line = {
        "impl": "reference", "metric": w.metric, "value": value, "unit": "configs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "reference_sample_configs_per_step": n_last},
        "cpu_baseline": {"value": value, "unit": "configs/s", "cores": threads, "kind": "port", "sample": smp},
        "e2e": {"value": value, "unit": "configs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def reference_problem(workload: str, n: int):
    """The CPU arm needs the same scene / robot tables but no GPU: build the workload's host-side objects only.
    -> (work(rows), sample(n), description of the port, config.workload string)."""
    w = WORKLOADS[workload]()
    if workload == "cfg3":
        w.pr2, w.colkin, w.grid = make_problem()
        w.lo, w.hi = joint_bounds(w.colkin)
        work, sample, what = w.cpu()
        return work, sample, what, w.describe(n)
    # the other workloads build their pair lists / plans through GPU calls in the product API (q = 0 filter): restate the
    # host-side tables directly
    from oracle import oracle
    if workload == "cfg2":
        from skmp_b200.robot.fetch import FetchConfig
        from skmp_b200.robot_model import Fetch

        fetch = Fetch(); fetch.reset_pose()
        kin = FetchConfig().get_collision_kin()
        kin.reflect_skrobot_model(fetch)
        pairs, _ = reference_pair_list(kin)
        w.kin, w.P, w.Dd = kin, len(pairs), kin.dim_cspace
        w.lo, w.hi = joint_bounds(kin)
        w.const = type("C", (), {"check_sphere_id_pairs": [tuple(p) for p in pairs]})()
        work, sample, what = w.cpu()
        return work, sample, what, w.describe(n)
    if workload == "cfg4":
        from skmp_b200.robot.pr2 import PR2Config
        from skmp_b200.robot_model import PR2
        from skmp_b200.sdf import Box
        from skmp_b200.tinyfk import RotationType

        pr2 = PR2(); pr2.reset_manip_pose()
        conf = PR2Config()
        w.colkin = conf.get_collision_kin(); w.colkin.reflect_skrobot_model(pr2)
        w.efkin = conf.get_endeffector_kin(RotationType.RPY); w.efkin.reflect_skrobot_model(pr2)
        w.box = Box([0.7, 0.5, 1.2], pos=[0.85, -0.2, 0.9])
        w.lo, w.hi = joint_bounds(w.colkin)
        work, sample, what = w.cpu()
        return work, sample, what, w.describe(n)
    from skmp_b200.robot.jaxon import JaxonConfig
    from skmp_b200.robot_model import Jaxon
    from skmp_b200.sdf import Box, UnionSDF
    from skmp_b200.tinyfk import RotationType

    jaxon = Jaxon(); jaxon.reset_manip_pose()
    conf = JaxonConfig()
    w.colkin = conf.get_collision_kin(); w.colkin.reflect_skrobot_model(jaxon)
    w.efkin = conf.get_endeffector_kin(rot_type=RotationType.XYZW); w.efkin.reflect_skrobot_model(jaxon)
    w.env_sdf = UnionSDF([Box([4.0, 4.0, 0.2], pos=[0.0, 0.0, -0.1]), Box([0.6, 1.0, 0.8], pos=[0.9, 0.0, 0.4])])
    pairs, _ = reference_pair_list(w.colkin)
    w.P = len(pairs)
    w.selfc = type("C", (), {"check_sphere_id_pairs": [tuple(p) for p in pairs]})()
    w.lo, w.hi = joint_bounds(w.colkin, scale=0.5)
    w.lo[33] += 1.0; w.hi[33] += 1.0
    work, sample, what = w.cpu()
    return work, sample, what, w.describe(n)


def reference_pair_list(kin):
    """PairWiseSelfCollFreeConst.__init__'s pair list (skmp/constraint.py:709-744) from the oracle: every pair whose distance
    at q = 0 is at least 3 (r_i + r_j)."""
    import itertools

    from oracle import oracle

    s = kin.fksolver
    tree = oracle.Tree.from_solver(s)
    ids = list(kin.tinyfk_feature_ids)
    rad = dict(zip(ids, kin.get_radius_list()))
    all_pairs = np.array(list(itertools.combinations(ids, 2)))
    base = {0: oracle.BASE_FIXED, 1: oracle.BASE_PLANER, 2: oracle.BASE_FLOATING}[kin.base_type.value]
    sq, _ = oracle.inter_link_sqdists(tree, np.zeros((1, kin.dim_cspace)), all_pairs, s.joint_child_links(kin.tinyfk_joint_ids), base, with_jacobian=False)
    rs = np.array([rad[a] + rad[b] for a, b in all_pairs])
    keep = np.sqrt(sq[0]) - 3.0 * rs >= 0
    return all_pairs[keep], rs[keep]


def bind_to_gpu_numa_node(local: int):
    """Pin this rank's host threads (and therefore its first-touch pinned buffers) to the CPUs next to its GPU, so that the
    D2H streams of several ranks do not all land in one socket's memory.  /sys first (PCI device -> NUMA node -> cpulist), the
    "CPU Affinity" column of `nvidia-smi topo -m` when the PCI device reports no node.  Never narrows the affinity to fewer
    than 4 CPUs (a single-node box reports every CPU for every GPU: nothing to do).  Best effort; returns a description."""
    try:
        import torch

        pr = torch.cuda.get_device_properties(local)
        bus = "%04x:%02x:%02x.0" % (getattr(pr, "pci_domain_id", 0), pr.pci_bus_id, pr.pci_device_id)
        node = -1
        try:
            node = int(Path(f"/sys/bus/pci/devices/{bus}/numa_node").read_text().strip())
        except Exception:
            pass

        def parse(cpulist):
            cpus = []
            for part in cpulist.strip().split(","):
                a, _, b = part.strip().partition("-")
                if a.isdigit():
                    cpus.extend(range(int(a), int(b or a) + 1))
            return cpus

        cpus, how = [], "sysfs"
        if node >= 0:
            cpus = parse(Path(f"/sys/devices/system/node/node{node}/cpulist").read_text())
        else:
            how = "nvidia-smi topo -m"
            out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=10).stdout
            import re
            for line in out.splitlines():
                clean = re.sub(r"\x1b\[[0-9;]*m", "", line)
                if clean.startswith(f"GPU{local}\t") or clean.startswith(f"GPU{local} "):
                    for field in clean.split("\t"):                       # the affinity column looks like 0-23 or 0-15,32-47
                        if re.fullmatch(r"\s*\d+(-\d+)?(,\d+(-\d+)?)*\s*", field) and ("-" in field or "," in field):
                            cpus = parse(field)
                            break
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if len(allowed) >= 4 and len(allowed) < len(os.sched_getaffinity(0)):
            os.sched_setaffinity(0, allowed)
            return f"gpu {bus} -> {len(allowed)} cpus ({how}, node {node})"
        return f"gpu {bus}: not bound ({how}: {len(allowed)} of {len(os.sched_getaffinity(0))} cpus, node {node})"
    except Exception as e:      # pragma: no cover
        return f"not bound ({type(e).__name__})"


# ------------------------------------------------------------------ CUDA arm
def run_b200(args):
    import torch
    import torch.distributed as dist

    from skmp_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else "single rank: not bound"
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    w = WORKLOADS[args.workload]()
    n = args.configs_per_gpu or w.default_n
    w.setup(dev, rank, n)
    n = w.n
    L = _lib.lib()
    w.step()                                     # first call: plans compiled, tables uploaded
    torch.cuda.synchronize()
    w.tune()                                     # explicit launch tuning (skb_plan_tune) -- never inside the timed region
    for _ in range(max(3, args.warmup)):
        w.step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = L.skb_launch_count()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    evs[0].record()
    torch.cuda.nvtx.range_push("timed")          # lets `ncu --nvtx --nvtx-include "timed/"` pick the timed launches
    for i in range(args.steps):
        w.step()
        evs[i + 1].record()
    torch.cuda.nvtx.range_pop()
    barrier()
    launches = L.skb_launch_count() - launches0
    per_step_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    total_ms = evs[0].elapsed_time(evs[-1])
    clocks = sampler.stop()
    t = torch.tensor([total_ms, float(np.mean(per_step_ms))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, kernel_ms = float(t[0]), float(t[1])
    value = world * n * args.steps / (total_ms * 1e-3)

    # ---- the path's one collective: all-gather of the validity bytes (cfg5 under torchrun), timed beside the step
    gather = None
    if args.workload == "cfg5":
        from skmp_b200.sharded import gather_validity

        v8 = w.valid.to(torch.uint8)
        if world > 1:
            for _ in range(2):
                gather_validity(v8, world * n)
            barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            allv = gather_validity(v8, world * n)
            g1.record()
            torch.cuda.synchronize()
            tg = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=dev)
            dist.all_reduce(tg, op=dist.ReduceOp.MAX)
            gather = {"gather_ms": float(tg[0]), "bytes_gathered": int(allv.numel()), "backend": "nccl all_gather_into_tensor",
                      "all_valid_fraction": float(allv.float().mean())}
        else:
            gather = {"gather_ms": 0.0, "bytes_gathered": int(v8.numel()), "backend": "single rank: nothing to gather"}

    # ---- end to end through the public API with host buffers (numpy in -> numpy out; H2D + D2H inside the timed region).
    # Throughput does not depend on the batch once it spans many pipeline chunks, so the e2e batch is capped: the full
    # 1.25 M-configuration cfg3 result is 10.6 GB of host memory per rank, 85 GB on an 8-GPU box for no extra information.
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    call, n_e2e, h2d, d2h, api = w.e2e(args.e2e_configs)
    for _ in range(2):
        res = call()
    del res
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res = call()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    del res
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * n_e2e * e2e_steps / float(te[0])

    if rank == 0:
        peaks_path = ROOT / "MEASURED_PEAKS.json"
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        if peaks_path.exists():
            peak, peak_src = float(json.loads(peaks_path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        (kernel_name, kernel_bytes), = w.kernels()
        achieved = kernel_bytes / (kernel_ms * 1e-3) / 1e9
        traffic, traffic_src = (w.traffic() if hasattr(w, "traffic") else (None, None))
        cpu = {"value": None}
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            work, sample, what = w.cpu()
            rate, n_cpu, dt = cpu_rate(work, sample, args.cpu_seconds, threads)
            cpu = {"value": rate, "unit": "configs/s", "cores": threads, "kind": "port",
                   "sample": f"{n_cpu} configs of the same workload in {dt:.1f} s: {what}"}
        cfg = {"workload": w.describe(n), "configs_per_gpu": n}
        cfg.update(w.extra_config())
        cfg["timed_call"] = ("the device-timed loop calls the C ABI (skb_collfree) directly on resident buffers" if args.workload == "cfg3"
                             else "the device-timed loop calls the Python API on resident CUDA tensors (outputs allocated per step from torch's caching allocator)")
        cfg["sphere_evals_per_s" if args.workload in ("cfg3", "cfg4") else "rows_per_s"] = value * {"cfg3": S, "cfg4": 76, "cfg2": getattr(w, "P", 0), "cfg5": 85 + 28}.get(args.workload, 1)
        # the e2e leg is a D2H stream: compare its per-rank rate with the box's raw pinned-copy ceiling at this rank count
        # (profiles/r2_d2h.json, measured once with tools/d2h_bench.py -- a static reference, not measured in this run)
        pcie = None
        dp = ROOT / "profiles" / "r2_d2h.json"
        if dp.exists():
            per_n = json.loads(dp.read_text()).get("per_n", {})
            if str(world) in per_n:
                raw = per_n[str(world)]["d2h_gbs_per_rank_3stream"]
                pcie = {"raw_d2h_gbs_per_rank": raw, "frac_of_pcie": (d2h * e2e_steps / float(te[0]) / 1e9) / raw,
                        "source": "profiles/r2_d2h.json (static)"}
        line = {
            "metric": w.metric, "value": value, "unit": "configs/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "bytes_per_config": kernel_bytes / n, "kernel_ms": kernel_ms, "kernel": kernel_name},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "configs/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps, "configs_per_step": n_e2e, "host_affinity_rank0": numa, "api": api,
                    "d2h_gbs_per_rank": d2h * e2e_steps / float(te[0]) / 1e9, "pcie": pcie},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if gather is not None:
            line["gather"] = gather
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--configs-per-gpu", type=int, default=0, help="0 = the workload's default (cfg3 / cfg5: 1.25 M, cfg2: 1 M, cfg4: 409 600)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-configs", type=int, default=524_288)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

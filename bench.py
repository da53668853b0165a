#!/usr/bin/env python
"""Headline benchmark: configs/sec for CollFreeConst value + Jacobian (BASELINE.json metric).

Workload (config.workload): BASELINE config 3 -- PR2 dual-arm 14-DoF, 152 collision spheres,
CollFreeConst vs a 128^3 voxel-grid SDF, fp32 values (N,152) + Jacobians (N,152,14).  The 10 M
configurations of the named config are sharded over 8 GPUs, i.e. 1.25 M per GPU; that per-GPU shard
is what every rank processes (weak scaling: --gpus 8 is exactly the named config).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)

One JSON line on stdout (rank 0).  Keys follow the driver contract; see DESIGN.md "Measurement".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
for p in (ROOT, ROOT / "scikit-motionplan_b200"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

S, D = 152, 14
BYTES_PER_CONFIG = 4 * (D + S + S * D)          # algorithmic bytes: q row in, values + Jacobian out
WORKLOAD = ("cfg3: PR2 dual-arm 14-DoF CollFreeConst vs 128^3 voxel-grid SDF, values (N,152) + Jacobians (N,152,14), "
            "fp32, {n} configs per GPU (10M / 8 GPUs)")


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


def joint_bounds(colkin):
    jm = colkin.fksolver.urdf.joint_map
    lo = np.array([-2 * np.pi if jm[n].limit.lower is None else jm[n].limit.lower for n in colkin.control_joint_names])
    hi = np.array([2 * np.pi if jm[n].limit.upper is None else jm[n].limit.upper for n in colkin.control_joint_names])
    return lo, hi


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


def cpu_reference_rate(colkin, grid, lo, hi, seconds: float, threads: int):
    """The reference's CPU algorithm for this path, restated in C (oracle port): per-config tree-walk FK +
    geometric Jacobian, 4 SDF sweeps for the forward-difference gradient (eps=1e-7), chain-rule contraction
    (skmp/constraint.py:335-374), fp64, `threads` host threads over contiguous shards."""
    from oracle import oracle

    s = colkin.fksolver
    tree = oracle.Tree.from_solver(s)
    feat = np.array(colkin.tinyfk_feature_ids)
    jl = s.joint_child_links(colkin.tinyfk_joint_ids)
    sdf = oracle.Sdf(grid.primitives())
    radii = colkin.get_radius_list()
    rng = np.random.default_rng(0)
    probe = rng.uniform(size=(512, D)) * (hi - lo) + lo
    t0 = time.perf_counter()
    oracle.collfree(tree, probe, feat, jl, sdf, radii, grad_mode=oracle.GRAD_FD)
    per_cfg = (time.perf_counter() - t0) / len(probe)
    n = int(max(threads * 256, min(4_000_000, seconds * threads / per_cfg)))
    qs = rng.uniform(size=(n, D)) * (hi - lo) + lo
    t0 = time.perf_counter()
    oracle.collfree_threaded(threads, tree, qs, feat, jl, sdf, radii, grad_mode=oracle.GRAD_FD, keep=False)
    dt = time.perf_counter() - t0
    return n / dt, n, dt


def run_reference(args):
    """--impl reference: rank 0 times the reference CPU algorithm (oracle port) on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    _, colkin, grid = make_problem()
    lo, hi = joint_bounds(colkin)
    threads = os.cpu_count() or 1
    per_step_seconds = max(1.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    rates, n_last = [], 0
    for i in range(args.warmup + args.steps):
        rate, n_last, dt = cpu_reference_rate(colkin, grid, lo, hi, per_step_seconds, threads)
        if i >= args.warmup:
            rates.append((n_last, dt))
    total_n, total_t = sum(r[0] for r in rates), sum(r[1] for r in rates)
    value = total_n / total_t
    sample = (f"{n_last} configs per step of the same workload (uniform within joint limits, seed 0), C port of the reference "
              f"algorithm incl. forward-difference SDF gradient, fp64, {threads} threads")
    line = {
        "impl": "reference", "metric": "configs/sec (CollFreeConst value+Jacobian)", "value": value, "unit": "configs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD.format(n=args.configs_per_gpu), "reference_sample_configs_per_step": n_last},
        "cpu_baseline": {"value": value, "unit": "configs/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "configs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(local: int):
    """Pin this rank's host threads (and therefore its first-touch pinned buffers) to the NUMA node its GPU hangs on, so
    that the D2H streams of several ranks do not all land in one socket's memory.  Best effort; returns a description."""
    try:
        import torch

        pr = torch.cuda.get_device_properties(local)
        bus = "%04x:%02x:%02x.0" % (getattr(pr, "pci_domain_id", 0), pr.pci_bus_id, pr.pci_device_id)
        node = int(Path(f"/sys/bus/pci/devices/{bus}/numa_node").read_text().strip())
        if node < 0:
            return f"gpu {bus}: no NUMA information"
        cpus = []
        for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
            a, _, b = part.partition("-")
            cpus.extend(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return f"gpu {bus} -> NUMA node {node}, {len(allowed)} cpus"
    except Exception as e:      # pragma: no cover
        return f"not bound ({type(e).__name__})"


def run_b200(args):
    import torch
    import torch.distributed as dist

    from skmp_b200 import _lib
    from skmp_b200.constraint import CollFreeConst

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

    pr2, colkin, grid = make_problem()
    const = CollFreeConst(colkin, grid, pr2)
    lo, hi = joint_bounds(colkin)
    n = args.configs_per_gpu
    gen = torch.Generator(device=dev).manual_seed(1000 + rank)      # each rank owns a different shard
    lo_t, hi_t = torch.tensor(lo, dtype=torch.float32, device=dev), torch.tensor(hi, dtype=torch.float32, device=dev)
    qs = torch.rand((n, D), generator=gen, device=dev, dtype=torch.float32) * (hi_t - lo_t) + lo_t
    plan = const._plan()
    L = _lib.lib()
    import ctypes as C
    vals = torch.empty((n, S), dtype=torch.float32, device=dev)
    jac = torch.empty((n, S, D), dtype=torch.float32, device=dev)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def step():
        _lib.check(L.skb_collfree(plan.handle, grid.native(), _lib.F32, C.c_void_p(qs.data_ptr()), n, 0.0, _lib.MODE_ALL,
                                  C.c_void_p(vals.data_ptr()), C.c_void_p(jac.data_ptr()), None, stream))

    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = L.skb_launch_count()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    evs[0].record()
    torch.cuda.nvtx.range_push("timed")          # lets `ncu --nvtx --nvtx-include "timed/"` pick the timed launches
    for i in range(args.steps):
        step()
        evs[i + 1].record()
    torch.cuda.nvtx.range_pop()
    barrier()
    launches = L.skb_launch_count() - launches0
    per_launch_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    total_ms = evs[0].elapsed_time(evs[-1])
    clocks = sampler.stop()
    t = torch.tensor([total_ms, float(np.mean(per_launch_ms))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, kernel_ms = float(t[0]), float(t[1])
    value = world * n * args.steps / (total_ms * 1e-3)

    # ---- end to end through the public API with host buffers (numpy in -> numpy out; H2D + D2H inside the timed region).
    # Throughput does not depend on the batch once it spans many pipeline chunks, so the e2e batch is capped: the full
    # 1.25 M-configuration result is 10.6 GB of host memory per rank, 85 GB on an 8-GPU box for no extra information.
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    n_e2e = min(n, args.e2e_configs)
    q_host = torch.empty((n_e2e, D), dtype=torch.float32, pin_memory=True)
    q_host.copy_(qs[:n_e2e])
    q_np = q_host.numpy()
    for _ in range(2):
        v_np, j_np = const.evaluate(q_np, True)
    del v_np, j_np
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        v_np, j_np = const.evaluate(q_np, True)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    chk = float(v_np[0, 0])
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * n_e2e * e2e_steps / float(te[0])
    valid_frac = float((vals.min(dim=1).values > 0).float().mean())

    if rank == 0:
        peaks_path = ROOT / "MEASURED_PEAKS.json"
        peak, peak_src = 6650.0, "fallback"
        if peaks_path.exists():
            peak, peak_src = float(json.loads(peaks_path.read_text())["hbm_gbs"]), "measured"
        achieved = BYTES_PER_CONFIG * n / (kernel_ms * 1e-3) / 1e9
        traffic = None
        tp = ROOT / "profiles" / "traffic_collfree_cfg3.json"
        if tp.exists():
            traffic = json.loads(tp.read_text()).get("dram_bytes_per_launch")
        cpu = {"value": None}
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            rate, n_cpu, dt = cpu_reference_rate(colkin, grid, lo, hi, args.cpu_seconds, threads)
            cpu = {"value": rate, "unit": "configs/s", "cores": threads, "kind": "port",
                   "sample": f"{n_cpu} configs of the same workload in {dt:.1f} s: C port of the reference algorithm "
                             f"(tree-walk FK + Jacobian, forward-difference SDF gradient, fp64)"}
        line = {
            "metric": "configs/sec (CollFreeConst value+Jacobian)", "value": value, "unit": "configs/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD.format(n=n), "configs_per_gpu": n, "spheres": S, "dof": D, "grid": "128^3 fp32",
                       "l2": "outputs (11.5 GB/step) and inputs (70 MB) exceed the 126 MB L2; no flush needed",
                       "valid_fraction": valid_frac, "sphere_evals_per_s": value * S},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "bytes_per_config": BYTES_PER_CONFIG,
                         "kernel_ms": kernel_ms, "kernel": "skb::s2_kernel<COLLFREE, grid-fast, VW2, u32, all, jac, NP<=8> (skb_sphere2.cu)"},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "configs/s", "h2d_bytes_per_step": n_e2e * D * 4, "d2h_bytes_per_step": n_e2e * (S + S * D) * 4,
                    "steps": e2e_steps, "configs_per_step": n_e2e, "host_affinity_rank0": numa, "api": "CollFreeConst.evaluate(numpy (N,14) float32, with_jacobian=True) -> numpy", "check": chk},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--configs-per-gpu", type=int, default=1_250_000)
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

"""Kernel-only timings (CUDA events, inputs resident) of every BASELINE.json configuration, one JSON line each:
algorithmic bytes per configuration, achieved GB/s and the fraction of the measured HBM peak.
bench.py stays the contract benchmark (cfg3); this script feeds DESIGN.md / profiles/ with the other rows.

  python scikit-motionplan_b200/tools/bench_kernels.py [--only cfg2_all,cfg3] [--reps 10]
"""
import argparse
import json
import os
import sys
from pathlib import Path

# This script measures kernels: the first (warm-up) call of every workload may time its launch candidates inside the call,
# whatever the batch size -- what a product caller does explicitly with `constraint.tune(qs)` / skb_plan_tune.
os.environ.setdefault("SKB_S2_AUTOTUNE_MIN", "32768")

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
for p in (ROOT, ROOT / "scikit-motionplan_b200"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

import torch  # noqa: E402


def device_samples(kin, n, seed, scale=1.0):
    jm = kin.fksolver.urdf.joint_map
    lo = [(-2 * np.pi if jm[n_].limit.lower is None else jm[n_].limit.lower) for n_ in kin.control_joint_names]
    hi = [(2 * np.pi if jm[n_].limit.upper is None else jm[n_].limit.upper) for n_ in kin.control_joint_names]
    nb = kin.dim_cspace - len(lo)
    lo, hi = np.array(lo + [-1.0] * nb) * scale, np.array(hi + [1.0] * nb) * scale
    g = torch.Generator(device="cuda").manual_seed(seed)
    lo_t, hi_t = torch.tensor(lo, dtype=torch.float32, device="cuda"), torch.tensor(hi, dtype=torch.float32, device="cuda")
    return torch.rand((n, kin.dim_cspace), generator=g, device="cuda", dtype=torch.float32) * (hi_t - lo_t) + lo_t


def timed(fn, reps, warmup=3):
    for _ in range(warmup):
        out = fn()
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    evs[0].record()
    torch.cuda.nvtx.range_push("timed")          # lets `ncu --nvtx --nvtx-include "timed/"` pick the timed launches
    for i in range(reps):
        out = fn()
        evs[i + 1].record()
    torch.cuda.nvtx.range_pop()
    torch.cuda.synchronize()
    ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(reps)]
    return float(np.mean(ms)), float(np.min(ms)), out


def workloads():
    import bench
    from skmp_b200.constraint import CollFreeConst, PairWiseSelfCollFreeConst, PoseConstraint
    from skmp_b200.robot.fetch import FetchConfig
    from skmp_b200.robot.jaxon import JaxonConfig
    from skmp_b200.robot.pr2 import PR2Config
    from skmp_b200.robot_model import PR2, Fetch, Jaxon
    from skmp_b200.sdf import Box, UnionSDF
    from skmp_b200.tinyfk import RotationType

    W = {}

    def cfg1(n):
        pr2 = PR2(); pr2.reset_manip_pose()
        colkin = PR2Config().get_collision_kin()
        const = CollFreeConst(colkin, Box([0.7, 0.5, 1.2], pos=[0.85, -0.2, 0.9]).sdf, pr2)
        qs = device_samples(colkin, n, 1)
        S, D = colkin.n_feature, colkin.dim_cspace
        return (lambda: const.evaluate(qs, True)), n, 4 * (D + S + S * D), f"PR2 right arm CollFreeConst vs box, all + Jacobian, N={n}"
    W["cfg1_1k"] = lambda: cfg1(1000)
    W["cfg4_409600"] = lambda: cfg1(409_600)

    def cfg2(closest, n):
        fetch = Fetch(); fetch.reset_pose()
        const = FetchConfig().get_pairwise_selcol_consts(fetch, only_closest_feature=closest)
        kin = const.colkin
        qs = device_samples(kin, n, 8)
        P, D = len(const.check_sphere_id_pairs), kin.dim_cspace
        M = 1 if closest else P
        S = kin.n_feature
        flops = (126 * D + 18 * S + 8 * P + 17 * D) if closest else None
        return (lambda: const.evaluate(qs, True)), n, 4 * (D + M + M * D), \
            f"Fetch PairWiseSelfCollFreeConst P={P} D={D}, {'closest' if closest else 'all pairs'} + Jacobian, N={n}", flops
    W["cfg2_all"] = lambda: cfg2(False, 1_000_000)
    W["cfg2_closest"] = lambda: cfg2(True, 1_000_000)

    def cfg3(mode):
        pr2, colkin, grid = bench.make_problem()
        n = 1_250_000
        qs = device_samples(colkin, n, 3)
        S, D = colkin.n_feature, colkin.dim_cspace
        if mode == "all":
            const = CollFreeConst(colkin, grid, pr2)
            return (lambda: const.evaluate(qs, True)), n, 4 * (D + S + S * D), f"PR2 dual arm vs 128^3 grid, all + Jacobian, N={n}"
        if mode == "closest":
            const = CollFreeConst(colkin, grid, pr2, only_closest_feature=True)
            return (lambda: const.evaluate(qs, True)), n, 4 * (D + 1 + D), f"PR2 dual arm vs 128^3 grid, closest + Jacobian, N={n}", \
                126 * D + (18 + 60) * S + 17 * D
        const = CollFreeConst(colkin, grid, pr2)
        return (lambda: const.is_valid_batch(qs)), n, 4 * D + 1, f"PR2 dual arm vs 128^3 grid, validity bytes, N={n}", \
            126 * D + (18 + 60) * S
    def cfg3_f64():
        pr2, colkin, grid = bench.make_problem()
        n = 400_000
        qs = device_samples(colkin, n, 3).double()
        S, D = colkin.n_feature, colkin.dim_cspace
        const = CollFreeConst(colkin, grid, pr2)
        return (lambda: const.evaluate(qs, True)), n, 8 * (D + S + S * D), f"PR2 dual arm vs 128^3 grid, all + Jacobian, fp64, N={n}"
    W["cfg3_f64"] = cfg3_f64

    def cfg2_f64():
        fetch = Fetch(); fetch.reset_pose()
        const = FetchConfig().get_pairwise_selcol_consts(fetch)
        n = 200_000
        qs = device_samples(const.colkin, n, 8).double()
        P, D = len(const.check_sphere_id_pairs), const.colkin.dim_cspace
        return (lambda: const.evaluate(qs, True)), n, 8 * (D + P + P * D), f"Fetch pairwise all pairs + Jacobian, fp64, N={n}"
    W["cfg2_f64"] = cfg2_f64

    def com(n=1_000_000):
        jaxon = Jaxon(); jaxon.reset_manip_pose()
        const = JaxonConfig().get_com_stability_const(jaxon, Box([0.2, 0.6, 5.0]), ["RARM_LINK7", "LARM_LINK7"], [5.0, 5.0])
        qs = (torch.rand((n, 37), device="cuda", generator=torch.Generator("cuda").manual_seed(2)) - 0.5) * 0.6
        return (lambda: const.evaluate(qs, True)), n, 4 * (37 + 1 + 37), f"JAXON COMStabilityConst (40 weighted points) + Jacobian, N={n}"
    W["cfg5_com"] = com
    def cfg3_map():
        pr2, colkin, grid = bench.make_problem()
        n = 400_000
        qs = device_samples(colkin, n, 3)
        S, D = colkin.n_feature, colkin.dim_cspace
        return (lambda: colkin.map(qs)), n, 4 * (D + 3 * S + 3 * S * D), f"PR2 dual arm KinematicsMap.map: 152 points + (152,3,14) Jacobians, N={n}"
    W["cfg3_map"] = cfg3_map
    W["cfg3"] = lambda: cfg3("all")
    W["cfg3_closest"] = lambda: cfg3("closest")
    W["cfg3_valid"] = lambda: cfg3("valid")

    def cfg5(mode):
        jaxon = Jaxon(); jaxon.reset_manip_pose()
        conf = JaxonConfig()
        colkin = conf.get_collision_kin()
        env = CollFreeConst(colkin, UnionSDF([Box([4.0, 4.0, 0.2], pos=[0.0, 0.0, -0.1]), Box([0.6, 1.0, 0.8], pos=[0.9, 0.0, 0.4])]), jaxon)
        S, D = colkin.n_feature, colkin.dim_cspace
        if mode == "valid":
            n = 10_000_000
            qs = device_samples(colkin, n, 11, scale=0.5); qs[:, 33] += 1.0
            return (lambda: env.is_valid_batch(qs)), n, 4 * D + 1, f"JAXON 37-DoF env CollFreeConst validity bytes, N={n}", \
                126 * D + (18 + 2 * 40) * S
        n = 1_000_000
        qs = device_samples(colkin, n, 11, scale=0.5); qs[:, 33] += 1.0
        if mode == "all":
            return (lambda: env.evaluate(qs, True)), n, 4 * (D + S + S * D), f"JAXON env CollFreeConst S={S} all + Jacobian, N={n}"
        if mode == "self_valid":
            selfc = PairWiseSelfCollFreeConst(colkin, jaxon, only_closest_feature=True)
            P = len(selfc.check_sphere_id_pairs)
            return (lambda: selfc.is_valid_batch(qs)), n, 4 * D + 1, f"JAXON self collision P={P} validity bytes, N={n}", \
                126 * D + 18 * S + 8 * P
        if mode == "self":
            selfc = PairWiseSelfCollFreeConst(colkin, jaxon, only_closest_feature=True)
            P = len(selfc.check_sphere_id_pairs)
            return (lambda: selfc.evaluate(qs, True)), n, 4 * (D + 1 + D), f"JAXON self collision P={P} closest + Jacobian, N={n}", \
                126 * D + 18 * S + 8 * P + 17 * D
        if mode == "edges":
            # RRT edge validity (BASELINE config 5): env + self collision over every interpolated sample of every edge
            from skmp_b200.batched import edge_samples_device, is_valid_motion_step_batch
            from skmp_b200.constraint import IneqCompositeConst

            selfc = PairWiseSelfCollFreeConst(colkin, jaxon, only_closest_feature=True)
            comp = IneqCompositeConst([env, selfc])
            step_box = conf.get_motion_step_box()
            E = 200_000
            q1 = qs[:E].double()
            g = torch.Generator(device="cuda").manual_seed(5)
            q2 = q1 + torch.randn((E, D), generator=g, device="cuda", dtype=torch.float64) * 0.25
            _, off = edge_samples_device(step_box, q1, q2)
            ns = int(off[-1])
            return (lambda: is_valid_motion_step_batch(step_box, q1, q2, comp)), ns, 4 * D + 1, \
                f"JAXON RRT edge validity: {E} edges -> {ns} samples, env + self collision, discretisation + validity + per-edge AND on the GPU"
        efkin = conf.get_endeffector_kin(rot_type=RotationType.XYZW)
        pose = PoseConstraint([np.zeros(7)] * 4, efkin, jaxon)
        return (lambda: pose.evaluate(qs, True)), n, 4 * (D + 28 + 28 * D), f"JAXON PoseConstraint 4 x XYZW + Jacobian, N={n}"
    W["cfg5_valid"] = lambda: cfg5("valid")
    W["cfg5_env_all"] = lambda: cfg5("all")
    W["cfg5_self_closest"] = lambda: cfg5("self")
    W["cfg5_self_valid"] = lambda: cfg5("self_valid")
    W["cfg5_pose"] = lambda: cfg5("pose")
    W["cfg5_edges"] = lambda: cfg5("edges")
    return W


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    peak = 6650.0
    pp = ROOT / "MEASURED_PEAKS.json"
    if pp.exists():
        peak = float(json.loads(pp.read_text())["hbm_gbs"])
    W = workloads()
    names = [s for s in args.only.split(",") if s] or list(W)
    # FMA-pipe peaks of this device (the denominator of the compute-bound modes), best of three
    import ctypes

    from skmp_b200 import _lib
    torch.cuda.init()
    fma = {}
    for kind, key in ((0, "fp32_ffma"), (1, "fp32_ffma2"), (2, "fp64_dfma")):
        t = ctypes.c_double(0.0)
        best = 0.0
        for _ in range(3):
            _lib.check(_lib.lib().skb_fma_peak(kind, 4096 if kind < 2 else 256, ctypes.byref(t)))
            best = max(best, t.value)
        fma[key] = best
    print(json.dumps({"workload": "fma_peak", "desc": "FMA-pipe micro-benchmark (skb_fma_peak), TFLOP/s", **fma}), flush=True)
    # write-only stream ceiling (skb_store_peak): plain streaming stores, and the step kernel's shared memory -> TMA bulk
    # store path at its own block sizes (cfg3: 3 584-byte steps; cfg2: 2 048; cfg4: 1 792)
    st = {}
    for kind, blk, key in ((0, 0, "stg_v4"), (1, 3584, "bulk_3584"), (1, 2048, "bulk_2048"), (1, 1792, "bulk_1792"), (1, 8192, "bulk_8192")):
        t = ctypes.c_double(0.0)
        _lib.check(_lib.lib().skb_store_peak(kind, 8 << 30, blk, ctypes.byref(t)))
        st[key] = t.value
    print(json.dumps({"workload": "store_peak", "desc": "write-only stream of 8 GiB (skb_store_peak), GB/s", "hbm_copy_peak_gbs": peak, **st}), flush=True)
    for name in names:
        fn, n, bpc, desc, *rest = W[name]()
        mean_ms, min_ms, _ = timed(fn, args.reps)
        gbs = bpc * n / (mean_ms * 1e-3) / 1e9
        rec = {"workload": name, "desc": desc, "configs": n, "ms": mean_ms, "ms_min": min_ms,
               "configs_per_s": n / (mean_ms * 1e-3), "bytes_per_config": bpc, "achieved_gbs": gbs,
               "hbm_peak_gbs": peak, "frac_of_hbm": gbs / peak}
        if rest and rest[0]:
            # compute-bound mode: algorithmic flop per configuration by SURVEY.md 8(d)'s count (126 per joint, 18 per sphere
            # transform, 60 per trilinear / 40 per box SDF value + gradient, 8 per pair, 17 per column of the one Jacobian
            # row); an upper bound on the work done, the validity modes stop at the first colliding step of a configuration
            tf = rest[0] * n / (mean_ms * 1e-3) / 1e12
            rec.update({"flop_per_config": rest[0], "achieved_tflops": tf, "fp32_peak_tflops": fma["fp32_ffma"],
                        "frac_of_fp32": tf / fma["fp32_ffma"]})
        print(json.dumps(rec), flush=True)
        del fn
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()

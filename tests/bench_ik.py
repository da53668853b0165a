"""IK throughput on the reference's reaching problem (tests/test_satisfy.py:22-53: PR2 right arm, pose target
(0.7, -0.6, 1.0), 0.7 x 0.5 x 1.2 box obstacle): the batched GPU solver (skmp_b200.satisfy) against the reference's
algorithm restated on the CPU -- scipy SLSQP with the objective |e|^2 and the inequality g - 1e-6 >= 0 exactly as
skmp/satisfy.py:66-106, constraints evaluated by the CPU oracle (fp64, forward-difference SDF gradient).

Lives under tests/ because it executes the oracle (test infrastructure) as the CPU side of the comparison.

  python tests/bench_ik.py [--seeds 4096] [--cpu-seeds 24]
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "scikit-motionplan_b200", ROOT / "tests"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=4096)
    ap.add_argument("--cpu-seeds", type=int, default=24)
    args = ap.parse_args()
    import torch
    from scipy.optimize import Bounds, minimize

    from helpers import ROT, oracle_args
    from oracle import oracle
    from skmp_b200.constraint import CollFreeConst, PoseConstraint
    from skmp_b200.robot.pr2 import PR2Config
    from skmp_b200.robot_model import PR2, Coordinates
    from skmp_b200.satisfy import SatisfactionConfig, satisfy_by_optimization_batch
    from skmp_b200.sdf import Box
    from skmp_b200.tinyfk import BaseType

    pr2 = PR2(); pr2.reset_manip_pose()
    conf = PR2Config(base_type=BaseType.FIXED)
    colkin, efkin = conf.get_collision_kin(), conf.get_endeffector_kin()
    box = Box([0.7, 0.5, 1.2]); box.translate([0.85, -0.2, 0.9])
    ineq = CollFreeConst(colkin, box.sdf, pr2)
    eq = PoseConstraint.from_skrobot_coords([Coordinates(pos=[0.7, -0.6, 1.0])], efkin, pr2)
    box_const = conf.get_box_const()
    rng = np.random.default_rng(0)
    seeds = rng.uniform(size=(args.seeds, 7)) * (box_const.ub - box_const.lb) + box_const.lb

    satisfy_by_optimization_batch(eq, box_const, ineq, seeds[:64])          # warm-up: plans, lazy uploads, module load
    torch.cuda.synchronize()
    times = []
    for _ in range(3):                                                       # every call captures its own CUDA graph
        t0 = time.perf_counter()
        res = satisfy_by_optimization_batch(eq, box_const, ineq, seeds, SatisfactionConfig())
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    gpu_s = float(np.median(times))

    # ---- the reference algorithm on the CPU (oracle constraints + scipy SLSQP), sequential like the reference
    tree_c, feat_c, jl_c, base_c = oracle_args(colkin)
    tree_e, feat_e, jl_e, base_e = oracle_args(efkin)
    sdf = oracle.Sdf(box.primitives())
    radii = colkin.get_radius_list()
    target = np.array([0.7, -0.6, 1.0, 0.0, 0.0, 0.0])

    def objective(q):
        pts, jac = oracle.solve_fk(tree_e, q[None], feat_e, jl_e, base_e, ROT[efkin.rot_type])
        e = pts.reshape(-1) - target
        return float(e @ e), 2.0 * e @ jac.reshape(6, 7)

    def ineq_fun(q):
        v, j = oracle.collfree(tree_c, q[None], feat_c, jl_c, sdf, radii, grad_mode=oracle.GRAD_FD)
        return v[0] - 1e-6, j[0]

    n_cpu, ok_cpu = min(args.cpu_seeds, args.seeds), 0
    t0 = time.perf_counter()
    for q0 in seeds[:n_cpu]:
        cache = {}

        def f(x):
            cache["f"] = objective(x)
            return cache["f"][0]

        def g(x):
            cache["g"] = ineq_fun(x)
            return cache["g"][0]

        out = minimize(f, q0, method="SLSQP", jac=lambda x: cache["f"][1], bounds=Bounds(box_const.lb, box_const.ub, keep_feasible=True),
                       constraints=[{"type": "ineq", "fun": g, "jac": lambda x: cache["g"][1]}],
                       options={"ftol": 1e-6, "disp": False, "maxiter": 199})
        ok_cpu += bool(out.success and out.fun < 1e-6)
    cpu_s = time.perf_counter() - t0
    print(json.dumps({
        "problem": "PR2 right-arm pose IK with box obstacle (tests/test_satisfy.py:22-53)",
        "gpu": {"seeds": args.seeds, "seconds": gpu_s, "seeds_per_s": args.seeds / gpu_s, "success_rate": float(res.success.mean()),
                "solutions_per_s": float(res.success.sum()) / gpu_s, "iterations": res.n_iter, "constraint_evals": res.n_eval, "seconds_each_of_3_calls": times},
        "cpu_reference_algorithm": {"seeds": n_cpu, "seconds": cpu_s, "seeds_per_s": n_cpu / cpu_s, "success_rate": ok_cpu / n_cpu,
                                    "solutions_per_s": ok_cpu / cpu_s, "threads": 1,
                                    "what": "scipy SLSQP over the oracle's constraint values / Jacobians (fp64, FD gradient), sequential"},
    }))


if __name__ == "__main__":
    main()

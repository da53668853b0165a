"""Latency of the planners' call pattern: ONE configuration per call (skmp/solver/ompl_solver.py:74-80,
motion_step_box.py:50-56, satisfy.py:67-82 all call evaluate_single / is_valid on a single q).

    python tests/latency_report.py --out gpurun_out/r2_latency.json

Per case: microseconds per call, host wall clock around the blocking call (numpy in -> numpy / bool out, the reference's
call shape) or around launch + stream synchronisation (CUDA tensors), median of 5 x 2000 calls after warm-up.  The oracle's
C port of the reference algorithm on one core is timed beside it (which is why this script lives under tests/: only test
infrastructure may load oracle/).
"""
import argparse
import ctypes as C
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "scikit-motionplan_b200"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from skmp_b200 import _lib  # noqa: E402
from skmp_b200.constraint import CollFreeConst, IneqCompositeConst, PairWiseSelfCollFreeConst  # noqa: E402


def per_call_us(fn, sync=False, reps=2000, rounds=5):
    for _ in range(200):
        fn()
    torch.cuda.synchronize()
    res = []
    for _ in range(rounds):
        t = time.perf_counter()
        for _ in range(reps):
            fn()
            if sync:
                torch.cuda.current_stream().synchronize()
        res.append((time.perf_counter() - t) / reps * 1e6)
    return float(np.median(res))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/r2_latency.json")
    args = ap.parse_args()
    out = {"unit": "microseconds per call, N = 1 configuration", "cases": {}}
    L = _lib.lib()

    def scene(name, const, kin, sdf_native):
        rng = np.random.default_rng(0)
        lo, hi = bench.joint_bounds(kin, scale=0.5)
        q64 = rng.uniform(size=kin.dim_cspace) * (hi - lo) + lo
        q32 = q64.astype(np.float32)
        c = out["cases"].setdefault(name, {})
        c["is_valid(q) numpy fp32"] = per_call_us(lambda: const.is_valid(q32))
        c["is_valid(q) numpy fp64"] = per_call_us(lambda: const.is_valid(q64))
        c["evaluate_single(q, False) numpy fp32"] = per_call_us(lambda: const.evaluate_single(q32, False))
        c["evaluate_single(q, True) numpy fp32"] = per_call_us(lambda: const.evaluate_single(q32, True))
        c["evaluate_single(q, True) numpy fp64"] = per_call_us(lambda: const.evaluate_single(q64, True))
        qd = torch.as_tensor(q32[None], device="cuda")
        c["is_valid_batch(cuda (1,D)) + sync"] = per_call_us(lambda: const.is_valid_batch(qd), sync=True)
        c["evaluate(cuda (1,D), True) + sync"] = per_call_us(lambda: const.evaluate(qd, True), sync=True)
        c["evaluate(cuda (1,D), True) launch only (asynchronous)"] = per_call_us(lambda: const.evaluate(qd, True))
        if sdf_native is not None:                      # raw C ABI, outputs preallocated
            plan = const._plan()
            m, d = (1 if const.only_closest_feature else kin.n_feature), kin.dim_cspace
            vals = np.empty((1, m), dtype=np.float32); jac = np.empty((1, m, d), dtype=np.float32); valid = np.empty(1, dtype=np.uint8)
            qp, vp, jp, bp = (C.c_void_p(a.ctypes.data) for a in (q32, vals, jac, valid))
            mode = _lib.MODE_CLOSEST if const.only_closest_feature else _lib.MODE_ALL
            c["C ABI skb_collfree_host validity byte"] = per_call_us(lambda: L.skb_collfree_host(plan.handle, sdf_native, _lib.F32, qp, 1, 0.0, mode, None, None, bp))
            c["C ABI skb_collfree_host values + jacobian"] = per_call_us(lambda: L.skb_collfree_host(plan.handle, sdf_native, _lib.F32, qp, 1, 0.0, mode, vp, jp, None))
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            dv = torch.empty((1, m), device="cuda"); dj = torch.empty((1, m, d), device="cuda")
            c["C ABI skb_collfree (device pointers) launch only"] = per_call_us(
                lambda: L.skb_collfree(plan.handle, sdf_native, _lib.F32, C.c_void_p(qd.data_ptr()), 1, 0.0, mode, C.c_void_p(dv.data_ptr()), C.c_void_p(dj.data_ptr()), None, st))
        return q64

    # cfg1: PR2 right arm vs the reference's box (the standard planning problem of the reference's tests)
    from skmp_b200.robot.pr2 import PR2Config
    from skmp_b200.robot_model import PR2
    from skmp_b200.sdf import Box

    pr2 = PR2(); pr2.reset_manip_pose()
    colkin = PR2Config().get_collision_kin()
    box = Box([0.7, 0.5, 1.2], pos=[0.85, -0.2, 0.9])
    const1 = CollFreeConst(colkin, box.sdf, pr2)
    q1 = scene("cfg1 PR2 right arm CollFreeConst vs box (S=76, D=7)", const1, colkin, box.sdf.native())
    # cfg3: PR2 dual arm vs voxel grid
    pr2d, colkin3, grid = bench.make_problem()
    const3 = CollFreeConst(colkin3, grid, pr2d)
    q3 = scene("cfg3 PR2 dual arm CollFreeConst vs 128^3 grid (S=152, D=14)", const3, colkin3, grid.native())
    # cfg5: JAXON env + self collision composite
    from skmp_b200.robot.jaxon import JaxonConfig
    from skmp_b200.robot_model import Jaxon
    from skmp_b200.sdf import UnionSDF

    jaxon = Jaxon(); jaxon.reset_manip_pose()
    jkin = JaxonConfig().get_collision_kin()
    env_sdf = UnionSDF([Box([4.0, 4.0, 0.2], pos=[0.0, 0.0, -0.1]), Box([0.6, 1.0, 0.8], pos=[0.9, 0.0, 0.4])])
    env = CollFreeConst(jkin, env_sdf, jaxon)
    comp = IneqCompositeConst([env, PairWiseSelfCollFreeConst(jkin, jaxon, only_closest_feature=True)])
    q5 = scene("cfg5 JAXON IneqCompositeConst([env, self collision]) (D=37)", comp, jkin, None)

    # the reference algorithm restated on the CPU, one configuration per call, one core
    from oracle import oracle
    for name, kin, sdf, q in (("cfg1 PR2 right arm CollFreeConst vs box (S=76, D=7)", colkin, box.sdf, q1),
                              ("cfg3 PR2 dual arm CollFreeConst vs 128^3 grid (S=152, D=14)", colkin3, grid, q3)):
        s = kin.fksolver
        tree, feat, jl = oracle.Tree.from_solver(s), np.array(kin.tinyfk_feature_ids), s.joint_child_links(kin.tinyfk_joint_ids)
        osdf, radii = oracle.Sdf(sdf.primitives()), kin.get_radius_list()
        qq = q[None]
        out["cases"][name]["CPU port of the reference (oracle, fp64, FD gradient): values + jacobian"] = per_call_us(
            lambda: oracle.collfree(tree, qq, feat, jl, osdf, radii, grad_mode=oracle.GRAD_FD), reps=500)
        out["cases"][name]["CPU port of the reference (oracle, fp64): values only"] = per_call_us(
            lambda: oracle.collfree(tree, qq, feat, jl, osdf, radii, with_jacobian=False), reps=500)
    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    Path(args.out).write_text(json.dumps(out, indent=1))
    for name, c in out["cases"].items():
        print(name)
        for k, v in c.items():
            print(f"   {k:75s} {v:8.1f} us")


if __name__ == "__main__":
    main()

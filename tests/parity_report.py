"""Measured parity of the CUDA path against the CPU oracle, for every BASELINE.json configuration, fp32 and fp64,
against BOTH gradient semantics of the oracle:

  * GRAD_ANALYTIC  the analytic SDF gradient (what the CUDA engine computes), and
  * GRAD_FD        the reference's own forward difference, eps = 1e-7 (skmp/constraint.py:320-331,362-369).

Test infrastructure (it loads oracle/): run on the GPU box as

    python tests/parity_report.py --out gpurun_out/parity_r2.json

and copy the result to profiles/.  tests/test_gpu_parity_report.py runs the same measurement at a reduced size and
asserts the north_star tolerances on it, so the numbers in profiles/ cannot drift away from the code.

Error measure (the one every parity test uses): |got - ref| / max(1, |ref|), maximum over all entries.  Jacobian
entries of samples closer than `kink_band` to a gradient discontinuity of the SDF (box medial axis, voxel cell faces,
union switch-over: measure-zero sets where fp32 and fp64 may legitimately pick different sides) are excluded from the
Jacobian comparison and counted in `kink_excluded_frac`; values are always compared.
"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "scikit-motionplan_b200", ROOT / "tests"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

from helpers import FK_POS_ERR, KINK_BAND, ROT, RPY_ROT_ERR, TOL, collfree_jac_excess, oracle_args, pose_excess, sample_q  # noqa: E402
from oracle import oracle  # noqa: E402

KINK_BAND_FD = 4e-7                                 # FD mode: the forward step (1e-7) must not straddle a kink either
SMOOTH_CURV = 1.0                                   # 1/m: rows whose gradient-conditioning term is below 2e-6 ("smooth")


def _err(got, ref):
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    e = np.abs(got - ref) / np.maximum(1.0, np.abs(ref))
    return e


def _stats(e):
    if e.size == 0:
        return {"max": 0.0, "p999": 0.0, "mean": 0.0}
    return {"max": float(e.max()), "p999": float(np.quantile(e, 0.999)), "mean": float(e.mean())}


def _gpu(const, qs, dtype, with_jac=True):
    import torch

    q = torch.as_tensor(qs.astype(dtype), device="cuda")
    v, j = const.evaluate(q, with_jac)
    torch.cuda.synchronize()
    return v.cpu().numpy(), (j.cpu().numpy() if with_jac else None)


def collfree_entry(const, colkin, sdf, qs, dtype, margin=0.0):
    """CollFreeConst (all spheres) vs oracle.collfree in both gradient modes."""
    v, j = _gpu(const, qs, dtype)
    tree, feat, jl, base = oracle_args(colkin)
    osdf = oracle.Sdf(sdf.primitives())
    dt = np.dtype(dtype).type
    out = {}
    for mode, name in ((oracle.GRAD_ANALYTIC, "analytic"), (oracle.GRAD_FD, "fd")):
        rv, rj, kink, curv = oracle.collfree(tree, qs, feat, jl, osdf, colkin.get_radius_list(), margin=margin, base_type=base,
                                             grad_mode=mode, with_kink=True, with_curv=True)
        band = max(KINK_BAND[dt], KINK_BAND_FD if mode == oracle.GRAD_FD else 0.0)
        ok = kink > band
        smooth = ok & (curv <= SMOOTH_CURV)
        # FD mode: the reference's own truncation error (eps / 2 x second derivative) is part of what is compared; fp64 rows
        # also carry the FD's round-off 2^-52 |sdf| / eps ~ 5e-9
        excess, _ = collfree_jac_excess(j, rj, np.where(ok, kink, 0.0), curv, dt, fd_eps=1e-7 if mode == oracle.GRAD_FD else 0.0)
        if mode == oracle.GRAD_FD and dt == np.float64:
            excess = float(np.max((_err(j, rj) / (1e-8 + (2e-15 + 1e-7) * np.minimum(curv, 1e12))[..., None])[ok])) if ok.any() else 0.0
        out[name] = {"values": _stats(_err(v, rv)), "jacobians": _stats(_err(j[ok], rj[ok])),
                     "jacobians_smooth_rows": _stats(_err(j[smooth], rj[smooth])), "smooth_row_frac": float(smooth.mean()),
                     "jacobian_err_over_tol_max": excess,
                     "kink_band_m": band, "kink_excluded_frac": float(1.0 - ok.mean())}
    out["validity_mismatch_outside_1e-6_band"] = int((((v > 0) != (rv > 0)) & (np.abs(rv) > 1e-6)).sum())
    return out


def pairwise_entry(const, qs, dtype):
    v, j = _gpu(const, qs, dtype)
    kin = const.colkin
    tree, _, jl, base = oracle_args(kin)
    sq, gr = oracle.inter_link_sqdists(tree, qs, np.array(const.check_sphere_id_pairs), jl, base)
    rv = sq - const.check_sphere_pair_sqdists
    if const.only_closest_feature:
        idx = rv.argmin(axis=1)
        n = np.arange(len(rv))
        # exact ties between pairs are resolved by index on both sides; compare the value, and the row of the same pair
        out = {"values": _stats(_err(v[:, 0], rv[n, idx])), "jacobians": _stats(_err(j[:, 0], gr[n, idx]))}
    else:
        out = {"values": _stats(_err(v, rv)), "jacobians": _stats(_err(j, gr))}
    out["validity_mismatch_outside_1e-6_band"] = int((((v > 0).all(axis=1) != (rv > 0).all(axis=1)) & (np.abs(rv).min(axis=1) > 1e-6)).sum())
    return out


def pose_entry(const, qs, dtype):
    v, j = _gpu(const, qs, dtype)
    kin = const.efkin
    tree, feat, jl, base = oracle_args(kin)
    rp, rj = oracle.solve_fk(tree, qs, feat, jl, base, ROT[kin.rot_type])
    n = len(qs)
    target = np.hstack(const.desired_poses)
    rv = rp.reshape(n, -1) - target
    ev, ej = pose_excess(v + target, j, rp, rj, kin.rot_type, np.dtype(dtype).type)
    out = {"values": _stats(_err(v, rv)), "jacobians": _stats(_err(j, rj.reshape(n, rv.shape[1], -1))),
           "value_err_over_tol_max": ev, "jacobian_err_over_tol_max": ej}
    if rp.shape[2] == 6:       # RPY: the rate map is singular at pitch = +-pi/2; report the well-conditioned part separately
        good = (np.abs(np.cos(rp[:, :, 4])) > 0.3).all(axis=1)
        out["jacobians_cos_pitch_gt_0.3"] = _stats(_err(j[good], rj.reshape(n, rv.shape[1], -1)[good]))
        out["cos_pitch_gt_0.3_frac"] = float(good.mean())
        out["min_cos_pitch"] = float(np.abs(np.cos(rp[:, :, 4])).min())
    return out


def measure(n_scale: float = 1.0, dtypes=(np.float32, np.float64)):
    import bench
    from skmp_b200.constraint import CollFreeConst, PairWiseSelfCollFreeConst, PoseConstraint
    from skmp_b200.robot.fetch import FetchConfig
    from skmp_b200.robot.jaxon import JaxonConfig
    from skmp_b200.robot.pr2 import PR2Config
    from skmp_b200.robot_model import PR2, Fetch, Jaxon
    from skmp_b200.sdf import Box, UnionSDF
    from skmp_b200.tinyfk import RotationType

    def N(n):
        return max(32, int(n * n_scale))

    report = {"error_measure": "|got - ref| / max(1, |ref|), maximum over all entries",
              "tolerance_model": {"values": "TOL", "collfree_jacobian_row": "TOL + FK_POS_ERR * curv (curv = the oracle's Lipschitz bound of the SDF gradient at the sphere centre, 1/m)",
                                  "rpy_rows": "angles TOL + RPY_ROT_ERR / cos(pitch), their Jacobian rows TOL + RPY_ROT_ERR / cos(pitch)^2",
                                  "TOL": {k.__name__: v for k, v in TOL.items()}, "FK_POS_ERR_m": {k.__name__: v for k, v in FK_POS_ERR.items()},
                                  "RPY_ROT_ERR": {k.__name__: v for k, v in RPY_ROT_ERR.items()},
                                  "err_over_tol_max": "<= 1 passes; tests/helpers.py"},
              "reference_gradient": "fd = forward difference eps 1e-7 in fp64 (skmp/constraint.py:320-331,362-369), analytic = exact",
              "north_star_tolerance": {"float32": 1e-5, "float64": 1e-10}, "configs": {}}
    pr2 = PR2(); pr2.reset_manip_pose()
    # cfg1 / cfg4 collision part: PR2 right arm vs the reference's test box
    colkin1 = PR2Config().get_collision_kin()
    box = Box([0.7, 0.5, 1.2], pos=[0.85, -0.2, 0.9])
    const1 = CollFreeConst(colkin1, box.sdf, pr2)
    # cfg2: Fetch pairwise (all pairs, and closest)
    fetch = Fetch(); fetch.reset_pose()
    pw_all = FetchConfig().get_pairwise_selcol_consts(fetch)
    # cfg3: PR2 dual arm vs the 128^3 voxel grid of bench.py
    pr2_3, colkin3, grid3 = bench.make_problem()
    const3 = CollFreeConst(colkin3, grid3, pr2_3)
    # cfg4 pose part: right-arm end effector, RPY rows, the reference's goal
    pose4 = PoseConstraint([np.array([0.7, -0.6, 1.0, 0.0, 0.0, 0.0])], PR2Config().get_endeffector_kin(RotationType.RPY), pr2)
    # cfg5: JAXON floating base: environment + self collision + 4 end effectors XYZW
    jaxon = Jaxon(); jaxon.reset_manip_pose()
    jconf = JaxonConfig()
    colkin5 = jconf.get_collision_kin()
    env_sdf = UnionSDF([Box([4.0, 4.0, 0.2], pos=[0.0, 0.0, -0.1]), Box([0.6, 1.0, 0.8], pos=[0.9, 0.0, 0.4])])
    env5 = CollFreeConst(colkin5, env_sdf, jaxon)
    self5 = PairWiseSelfCollFreeConst(colkin5, jaxon, only_closest_feature=True)
    pose5 = PoseConstraint([np.zeros(7)] * 4, jconf.get_endeffector_kin(rot_type=RotationType.XYZW), jaxon)

    for dtype in dtypes:
        name = np.dtype(dtype).name
        rng = np.random.default_rng(20)

        def q(kin, n, scale=1.0, lift=None):
            x = sample_q(kin, n, rng) * scale
            if lift is not None:
                x[:, lift] += 1.0
            return x.astype(dtype).astype(np.float64)

        c = report["configs"]
        c.setdefault("cfg1 PR2 right arm CollFreeConst vs box (S=76, D=7)", {})[name] = collfree_entry(const1, colkin1, box.sdf, q(colkin1, N(1000)), dtype)
        c.setdefault("cfg2 Fetch PairWiseSelfCollFreeConst, all pairs (values in m^2)", {})[name] = pairwise_entry(pw_all, q(pw_all.colkin, N(2000)), dtype)
        c.setdefault("cfg3 PR2 dual arm CollFreeConst vs 128^3 voxel grid (S=152, D=14)", {})[name] = collfree_entry(const3, colkin3, grid3, q(colkin3, N(2000)), dtype)
        c.setdefault("cfg4 SQP batch: PR2 right arm collision rows (as cfg1, other samples)", {})[name] = collfree_entry(const1, colkin1, box.sdf, q(colkin1, N(2000)), dtype)
        c.setdefault("cfg4 SQP batch: PoseConstraint RPY at the goal waypoint", {})[name] = pose_entry(pose4, q(pose4.efkin, N(2000)), dtype)
        c.setdefault("cfg5 JAXON env CollFreeConst vs ground + table (S=85, D=37)", {})[name] = collfree_entry(env5, colkin5, env_sdf, q(colkin5, N(1000), 0.5, 33), dtype)
        c.setdefault("cfg5 JAXON PairWiseSelfCollFreeConst, closest pair (values in m^2)", {})[name] = pairwise_entry(self5, q(colkin5, N(1000), 0.5), dtype)
        c.setdefault("cfg5 JAXON PoseConstraint 4 x XYZW", {})[name] = pose_entry(pose5, q(pose5.efkin, N(1000), 0.5), dtype)
    return report


def worst(report, dtype_name: str):
    """(max value error, max Jacobian error vs analytic, max Jacobian error vs fd) over all configs of one dtype."""
    ev = ej = ef = 0.0
    for entry in report["configs"].values():
        e = entry[dtype_name]
        if "analytic" in e:
            ev = max(ev, e["analytic"]["values"]["max"])
            ej = max(ej, e["analytic"]["jacobians"]["max"])
            ef = max(ef, e["fd"]["jacobians"]["max"])
        else:
            ev = max(ev, e["values"]["max"])
            ej = max(ej, e["jacobians"]["max"])
    return ev, ej, ef


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/parity_r2.json")
    ap.add_argument("--scale", type=float, default=1.0)
    args = ap.parse_args()
    rep = measure(args.scale)
    for dt in ("float32", "float64"):
        rep[f"worst_{dt}"] = dict(zip(("values", "jacobians_vs_analytic", "jacobians_vs_fd"), worst(rep, dt)))
    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    Path(args.out).write_text(json.dumps(rep, indent=1))
    print(json.dumps({k: rep[k] for k in rep if k.startswith("worst")}))
    for cfg, entry in rep["configs"].items():
        for dt, e in entry.items():
            if "analytic" in e:
                print(f"{cfg:75s} {dt}: val {e['analytic']['values']['max']:.2e}  jac(analytic) {e['analytic']['jacobians']['max']:.2e} "
                      f"[smooth rows {e['analytic']['jacobians_smooth_rows']['max']:.2e}, err/tol {e['analytic']['jacobian_err_over_tol_max']:.2f}] "
                      f"jac(fd) {e['fd']['jacobians']['max']:.2e} [smooth {e['fd']['jacobians_smooth_rows']['max']:.2e}, err/tol {e['fd']['jacobian_err_over_tol_max']:.2f}]  "
                      f"kink-excluded {e['analytic']['kink_excluded_frac']:.4f}/{e['fd']['kink_excluded_frac']:.4f}")
            else:
                extra = f" [err/tol val {e['value_err_over_tol_max']:.2f} jac {e['jacobian_err_over_tol_max']:.2f}]" if "jacobian_err_over_tol_max" in e else ""
                print(f"{cfg:75s} {dt}: val {e['values']['max']:.2e}  jac {e['jacobians']['max']:.2e}{extra}")

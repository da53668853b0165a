"""GPU parity against the COMMITTED golden fixtures (tests/golden/*.npz, written by make_golden.py from
the CPU oracle): the CUDA path must reproduce them through the public API without the oracle running.
Tolerances as in test_gpu_parity.py (fp32 1e-5 relative, fp64 1e-10; Jacobians excluded within the kink band)."""
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"
from helpers import KINK_BAND, TOL, collfree_jac_excess  # noqa: E402


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return (np.abs(a - b) / np.maximum(1.0, np.abs(b))).max()


def dev(x, dtype):
    import torch

    return torch.as_tensor(np.asarray(x).astype(dtype), device="cuda")


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_cfg1_golden(dtype):
    from golden.make_golden import _pr2, box_scene
    from skmp_b200.constraint import CollFreeConst
    from skmp_b200.robot.pr2 import PR2Config

    g = np.load(GOLDEN / "cfg1_pr2_rarm_box.npz")
    pr2, kin = _pr2(), PR2Config().get_collision_kin()
    tol = TOL[dtype]
    v, j = CollFreeConst(kin, box_scene().sdf, pr2, distance_margin=0.05).evaluate(dev(g["qs"], dtype), True)
    ok = g["kink"] > KINK_BAND[dtype]
    assert relerr(v.cpu(), g["values"]) < tol and collfree_jac_excess(j.cpu().numpy(), g["jac"], g["kink"], g["curv"], dtype)[0] <= 1.0
    vc, jc = CollFreeConst(kin, box_scene().sdf, pr2, only_closest_feature=True, distance_margin=0.05).evaluate(dev(g["qs"], dtype), True)
    assert relerr(vc.cpu(), g["closest_values"]) < tol
    idx = g["values"].argmin(axis=1)
    okc = ok[np.arange(len(idx)), idx] & (np.diff(np.sort(g["values"], axis=1)[:, :2], axis=1)[:, 0] > 1e-4)
    n = np.arange(len(idx))
    assert collfree_jac_excess(jc.cpu().numpy()[okc], g["closest_jac"][okc], g["kink"][n, idx][okc, None], g["curv"][n, idx][okc, None], dtype)[0] <= 1.0
    pts, pj = kin.map(dev(g["qs"][:8], dtype))
    assert relerr(pts.cpu(), g["points"]) < tol and relerr(pj.cpu(), g["points_jac"]) < tol


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_cfg2_golden(dtype):
    from skmp_b200.constraint import PairWiseSelfCollFreeConst
    from skmp_b200.robot.fetch import FetchConfig
    from skmp_b200.robot_model import Fetch

    g = np.load(GOLDEN / "cfg2_fetch_pairwise.npz")
    fetch = Fetch(); fetch.reset_pose()
    kin = FetchConfig().get_collision_kin()
    ids = [(kin.tinyfk_feature_ids[a], kin.tinyfk_feature_ids[b]) for a, b in g["pairs_local"]]
    const = PairWiseSelfCollFreeConst(kin, fetch, id_pairs=ids)
    assert np.allclose(const.check_sphere_pair_sqdists, g["thresholds"], atol=1e-15)
    v, j = const.evaluate(dev(g["qs"], dtype), True)
    assert relerr(v.cpu(), g["values"]) < TOL[dtype] and relerr(j.cpu(), g["jac"]) < TOL[dtype]
    # the constructor's own q = 0 filter yields the same pair list (itertools.combinations order)
    auto = PairWiseSelfCollFreeConst(kin, fetch)
    assert auto.check_sphere_id_pairs == ids


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_cfg3_golden(dtype):
    from golden.make_golden import _pr2, grid_scene
    from skmp_b200.constraint import CollFreeConst
    from skmp_b200.robot.pr2 import PR2Config

    g = np.load(GOLDEN / "cfg3_pr2_dual_grid.npz")
    const = CollFreeConst(PR2Config(control_arm="dual").get_collision_kin(), grid_scene(), _pr2())
    v, j = const.evaluate(dev(g["qs"], dtype), True)
    assert relerr(v.cpu(), g["values"]) < TOL[dtype]
    assert collfree_jac_excess(j.cpu().numpy(), g["jac"], g["kink"], g["curv"], dtype)[0] <= 1.0
    valid = const.is_valid_batch(dev(g["qs"], dtype)).cpu().numpy()
    decisive = np.abs(g["values"]).min(axis=1) > 1e-6
    assert (valid == (g["values"] > 0).all(axis=1))[decisive].all()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_cfg5_pose_golden(dtype):
    from skmp_b200.robot.jaxon import JaxonConfig
    from skmp_b200.robot_model import Jaxon
    from skmp_b200.tinyfk import RotationType

    g = np.load(GOLDEN / "cfg5_jaxon_pose_xyzw.npz")
    jaxon = Jaxon(); jaxon.reset_manip_pose()
    kin = JaxonConfig().get_endeffector_kin(rot_type=RotationType.XYZW)
    kin.reflect_skrobot_model(jaxon)
    pts, jac = kin.map(dev(g["qs"], dtype))
    assert relerr(pts.cpu(), g["points"]) < TOL[dtype] and relerr(jac.cpu(), g["jac"]) < TOL[dtype]      # XYZW rows: no singularity

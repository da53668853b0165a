"""Tests that need more than one GPU in the box (skipped otherwise): several devices in ONE process (plans and SDF handles
are per device), and the path's one collective on real NCCL (the CPU suite covers it with gloo, tests/test_sharding.py)."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _n_gpus():
    import torch

    return torch.cuda.device_count()


@pytest.mark.skipif("_n_gpus() < 2")
def test_two_devices_in_one_process():
    """The same constraint objects evaluated with tensors on cuda:0 and on cuda:1: each device gets its own plan and its own
    SDF handle (tables and voxel grids live in that device's memory), results are identical; handing the C ABI a plan of
    another device is an error, not an illegal address."""
    import ctypes as C

    import torch

    import bench
    from skmp_b200 import _lib
    from skmp_b200.constraint import CollFreeConst, PairWiseSelfCollFreeConst
    from skmp_b200.robot.fetch import FetchConfig
    from skmp_b200.robot_model import Fetch

    pr2, colkin, grid = bench.make_problem()
    const = CollFreeConst(colkin, grid, pr2)
    fetch = Fetch(); fetch.reset_pose()
    pw = FetchConfig().get_pairwise_selcol_consts(fetch, only_closest_feature=True)
    rng = np.random.default_rng(5)
    q = rng.uniform(-1.0, 1.0, size=(777, 14)).astype(np.float32)
    q8 = rng.uniform(-1.0, 1.0, size=(555, 8)).astype(np.float32)
    res = []
    for d in (0, 1, 0):
        qd = torch.as_tensor(q, device=f"cuda:{d}")
        v, j = const.evaluate(qd, True)
        ok = const.is_valid_batch(qd)
        pv, pj = pw.evaluate(torch.as_tensor(q8, device=f"cuda:{d}"), True)
        assert v.device.index == d and j.device.index == d and ok.device.index == d and pv.device.index == d
        pts = grid(torch.as_tensor(q[:, :3], device=f"cuda:{d}"))
        res.append([t.cpu() for t in (v, j, ok, pv, pj, pts)])
    for a, b in zip(res[0], res[1]):
        assert torch.equal(a, b)
    for a, b in zip(res[0], res[2]):
        assert torch.equal(a, b)
    # C ABI: a plan created on cuda:0 used while cuda:1 is current
    plan0 = const._plan(0)
    with torch.cuda.device(1):
        qd = torch.as_tensor(q, device="cuda:1")
        v = torch.empty((len(q), 152), device="cuda:1")
        rc = _lib.lib().skb_collfree(plan0.handle, grid.native(1), _lib.F32, C.c_void_p(qd.data_ptr()), len(q), 0.0, _lib.MODE_ALL,
                                     C.c_void_p(v.data_ptr()), None, None, None)
        assert rc != 0 and b"created on cuda:0" in _lib.lib().skb_last_error()
        rc = _lib.lib().skb_collfree(const._plan(1).handle, grid.native(0), _lib.F32, C.c_void_p(qd.data_ptr()), len(q), 0.0, _lib.MODE_ALL,
                                     C.c_void_p(v.data_ptr()), None, None, None)
        assert rc != 0 and b"SDF lives on cuda:0" in _lib.lib().skb_last_error()


@pytest.mark.skipif("_n_gpus() < 2")
def test_nccl_gather_validity(tmp_path):
    """``sharded_is_valid`` on two ranks with the NCCL backend == the single-GPU answer for the whole batch."""
    import torch

    from skmp_b200.constraint import CollFreeConst
    from skmp_b200.robot.pr2 import PR2Config
    from skmp_b200.robot_model import PR2
    from skmp_b200.sdf import Box

    out = tmp_path / "valid.npy"
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29617", str(ROOT / "tests" / "nccl_gather_worker.py"), str(out)],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    got = np.load(out)
    pr2 = PR2(); pr2.reset_manip_pose()
    colkin = PR2Config(control_arm="dual").get_collision_kin()
    const = CollFreeConst(colkin, Box([0.7, 0.5, 1.2], pos=[0.85, -0.2, 0.9]).sdf, pr2)
    qs = np.random.default_rng(123).uniform(-1.5, 1.5, size=(100_003, 14)).astype(np.float32)
    ref = const.is_valid_batch(torch.as_tensor(qs, device="cuda:0")).cpu().numpy()
    assert got.shape == ref.shape and np.array_equal(got, ref) and 0.05 < ref.mean() < 0.95

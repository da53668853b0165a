"""Worker of tests/test_gpu_multi.py::test_nccl_gather_validity (one process per GPU, launched by torch.distributed.run):
``sharded_is_valid`` = this rank's row block on its own GPU + the path's ONE collective, an NCCL all-gather of validity bytes."""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "scikit-motionplan_b200", ROOT / "tests"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    out = sys.argv[1]
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from skmp_b200.constraint import CollFreeConst
    from skmp_b200.robot.pr2 import PR2Config
    from skmp_b200.robot_model import PR2
    from skmp_b200.sdf import Box
    from skmp_b200.sharded import shard_bounds, sharded_evaluate, sharded_is_valid

    pr2 = PR2(); pr2.reset_manip_pose()
    colkin = PR2Config(control_arm="dual").get_collision_kin()
    const = CollFreeConst(colkin, Box([0.7, 0.5, 1.2], pos=[0.85, -0.2, 0.9]).sdf, pr2)
    rng = np.random.default_rng(123)                         # every rank builds the same full batch
    qs = rng.uniform(-1.5, 1.5, size=(100_003, 14)).astype(np.float32)
    valid = sharded_is_valid(const, qs)                      # (N,) bool on this rank's GPU, gathered from all ranks
    vals, jac, (a, b) = sharded_evaluate(const, qs, True)    # stays sharded
    assert valid.is_cuda and valid.device.index == local and valid.shape == (len(qs),)
    assert (a, b) == shard_bounds(len(qs), dist.get_world_size(), dist.get_rank()) and vals.shape == (b - a, 152)
    assert torch.equal(valid[a:b], (vals > 0).all(dim=1))
    if dist.get_rank() == 0:
        np.save(out, valid.cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

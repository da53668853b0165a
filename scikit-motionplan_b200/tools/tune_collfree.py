"""Kernel-only timing sweep of the cfg3 workload (PR2 dual arm vs 128^3 grid) over kernel variants and
tuning knobs, in one process.  Development tool (run on the GPU box through gpurun):

    python scikit-motionplan_b200/tools/tune_collfree.py [n_configs] [variant ...]

A variant is  kernel:cells:G:warps:ctas  e.g.  sphere:1:8:8:2  tree:0:0:0:0 ; 0 = auto.
Prints one line per variant: ms per launch, configs/s, achieved GB/s (algorithmic bytes), frac of the measured peak.
"""
import ctypes as C
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
for p in (ROOT, ROOT / "scikit-motionplan_b200"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

import numpy as np
import torch

import bench
from skmp_b200 import _lib
from skmp_b200.constraint import CollFreeConst
from skmp_b200.sdf import GridSDF


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
    variants = sys.argv[2:] or ["tree:0:0:0:0", "sphere:0:0:0:0", "sphere:1:0:0:0", "sphere:1:4:8:2", "sphere:1:8:8:2", "sphere:1:16:8:1",
                                "sphere:1:8:4:4", "sphere:1:8:16:1", "sphere:1:4:8:3"]
    dev = torch.device("cuda", 0)
    pr2, colkin, grid0 = bench.make_problem()
    lo, hi = bench.joint_bounds(colkin)
    gen = torch.Generator(device=dev).manual_seed(1000)
    lo_t, hi_t = torch.tensor(lo, dtype=torch.float32, device=dev), torch.tensor(hi, dtype=torch.float32, device=dev)
    qs = torch.rand((n, bench.D), generator=gen, device=dev, dtype=torch.float32) * (hi_t - lo_t) + lo_t
    vals = torch.empty((n, bench.S), dtype=torch.float32, device=dev)
    jac = torch.empty((n, bench.S, bench.D), dtype=torch.float32, device=dev)
    peak = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]) if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
    L = _lib.lib()
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ref = None
    grids = {}
    for var in variants:
        kernel, cells, G, warps, ctas = var.split(":")
        os.environ["SKB_KERNEL"] = kernel
        os.environ["SKB_GRID_CELLS"] = cells
        if cells not in grids:
            grids[cells] = GridSDF(grid0.values, grid0.lb, grid0.spacing, fill_value=grid0.fill_value)
            grids[cells].native()
        grid = grids[cells]
        const = CollFreeConst(colkin, grid, pr2)
        plan = const._plan()
        _lib.check(L.skb_plan_set_tuning(plan.handle, int(G), 0, int(warps), int(ctas)))

        def step():
            _lib.check(L.skb_collfree(plan.handle, grid.native(), _lib.F32, C.c_void_p(qs.data_ptr()), n, 0.0, _lib.MODE_ALL,
                                      C.c_void_p(vals.data_ptr()), C.c_void_p(jac.data_ptr()), None, stream))

        vals.zero_(); jac.fill_(7.0)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        gbs = bench.BYTES_PER_CONFIG * n / (ms * 1e-3) / 1e9
        chk = ""
        if ref is None:
            ref = (vals.clone(), jac.clone())
        else:
            dv = (vals - ref[0]).abs().max().item()
            dj = (jac - ref[1]).abs()
            chk = f"  max|dv|={dv:.2e} max|dJ|={dj.max().item():.2e} q999|dJ|={torch.quantile(dj.flatten()[:10_000_000], 0.999).item():.2e}"
        print(f"{var:24s} {ms:8.3f} ms  {n / ms * 1e3:12.4g} cfg/s  {gbs:8.1f} GB/s  frac {gbs / peak:5.3f}{chk}", flush=True)
        _lib.check(L.skb_plan_set_tuning(plan.handle, 0, 0, 0, 0))


if __name__ == "__main__":
    main()

"""Raw pinned device -> host copy bandwidth at 1 ... N concurrent ranks of one box: the platform ceiling of the e2e leg.

cfg3 returns 9.1 KB per configuration to host memory, so `CollFreeConst.evaluate(numpy) -> numpy` is a D2H stream (bench.py's
`e2e`).  This tool measures what the box can do with NOTHING of ours in the way -- plain `cudaMemcpyAsync` (through torch) of
48 MB chunks from device memory into pinned host memory, the chunk size of the host pipeline (csrc/skb_api.cpp) -- so that
e2e can be read as a fraction of it:

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scikit-motionplan_b200/tools/d2h_bench.py [--bind]

Every rank copies concurrently (barrier before, max time over ranks); rank 0 prints one JSON line with the per-rank and
aggregate GB/s, for one stream and for three streams, with and without binding the rank to its GPU's NUMA node.
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
for p in (ROOT, ROOT / "scikit-motionplan_b200"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bind", action="store_true", help="bind the rank to its GPU's NUMA node first (bench.bind_to_gpu_numa_node)")
    ap.add_argument("--total-mb", type=int, default=4096)
    ap.add_argument("--chunk-mb", type=int, default=48)
    args = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    bound = "not bound"
    if args.bind:
        import bench
        bound = bench.bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    chunk = args.chunk_mb << 20
    n_chunks = max(1, (args.total_mb << 20) // chunk)
    src = [torch.empty(chunk, dtype=torch.uint8, device=dev) for _ in range(3)]
    dst = torch.empty(n_chunks * chunk, dtype=torch.uint8, pin_memory=True)
    dst.fill_(1)                                          # first touch by this rank's (possibly bound) thread
    h2d_src = torch.empty(chunk, dtype=torch.uint8, pin_memory=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(n_streams, direction):
        streams = [torch.cuda.Stream() for _ in range(n_streams)]
        barrier()
        t0 = time.perf_counter()
        for i in range(n_chunks):
            with torch.cuda.stream(streams[i % n_streams]):
                if direction == "d2h":
                    dst[i * chunk:(i + 1) * chunk].copy_(src[i % 3], non_blocking=True)
                else:
                    src[i % 3].copy_(h2d_src, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return n_chunks * chunk / float(t[0]) / 1e9

    out = {"tool": "d2h_bench", "n_ranks": world, "chunk_mb": args.chunk_mb, "total_mb_per_rank": n_chunks * args.chunk_mb, "host_affinity_rank0": bound}
    for direction in ("d2h", "h2d"):
        for ns in (1, 3):
            run(ns, direction)                            # warm-up
            per_rank = max(run(ns, direction) for _ in range(3))
            out[f"{direction}_gbs_per_rank_{ns}stream"] = per_rank
            out[f"{direction}_gbs_aggregate_{ns}stream"] = per_rank * world
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

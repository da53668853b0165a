"""Multi-GPU sharding of configuration batches: one process per GPU (torch.distributed), contiguous row
blocks, replicated robot constants / SDF, outputs stay sharded, NO collective on the compute path.

Every row of ``qs`` is independent in map / CollFreeConst / PairWise / PoseConstraint
(skmp/constraint.py:287-374,749-778,451-472 have no cross-row term), so rank g of G evaluates
``qs[shard_bounds(N, G, g)]`` on its own device.  The only exchange offered is a final *gather* of the
per-configuration validity bytes (N bytes in total) -- what an RRT edge checker or ``Problem.is_satisfied``
(skmp/solver/interface.py:72-92) consumes.  Values / Jacobians are never gathered (cfg3: 92 GB).
"""
import os
from typing import Optional, Tuple

import numpy as np

try:
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch, dist = None, None


def world_info(group=None) -> Tuple[int, int]:
    """(rank, world_size) of the default / given process group; (0, 1) outside torch.distributed."""
    if dist is not None and dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row block of rank ``rank``: sizes differ by at most one, the first ``n % world`` ranks get the
    extra row, blocks are ordered by rank and cover [0, n) exactly."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(int(n), int(world))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_rows(qs, group=None):
    """This rank's row block of a batch that every rank holds in full -> (local view, (start, stop))."""
    rank, world = world_info(group)
    start, stop = shard_bounds(qs.shape[0], world, rank)
    return qs[start:stop], (start, stop)


def local_device(group=None):
    """cuda:LOCAL_RANK (one process per GPU); the caller is expected to have set it current."""
    if torch is None or not torch.cuda.is_available():
        raise RuntimeError("skmp_b200 needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", int(os.environ.get("LOCAL_RANK", torch.cuda.current_device())))


def gather_validity(valid_local, n_total: int, group=None):
    """all_gather of the per-rank validity bytes -> (n_total,) uint8 tensor on every rank (same device as the
    input; CPU tensors work with the gloo backend).  Shards are padded to the largest block for the
    collective and trimmed afterwards; this is the ONLY collective of the path."""
    rank, world = world_info(group)
    v = torch.as_tensor(valid_local).to(torch.uint8).reshape(-1)
    start, stop = shard_bounds(n_total, world, rank)
    if v.numel() != stop - start:
        raise ValueError(f"rank {rank} holds {v.numel()} validity bytes, its shard has {stop - start} rows")
    if world == 1:
        return v
    cap = shard_bounds(n_total, world, 0)[1]        # rank 0 always owns a largest block
    padded = torch.zeros(cap, dtype=torch.uint8, device=v.device)
    padded[: v.numel()] = v
    out = torch.empty(world * cap, dtype=torch.uint8, device=v.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    parts = []
    for r in range(world):
        s, e = shard_bounds(n_total, world, r)
        parts.append(out[r * cap: r * cap + (e - s)])
    return torch.cat(parts)


def sharded_evaluate(const, qs, with_jacobian: bool = True, group=None):
    """Evaluate this rank's row block on its GPU.  ``qs`` is the FULL (N, D) batch (numpy or CPU torch; every rank
    passes the same array) -> (values, jacobians, (start, stop)); results are CUDA tensors and stay sharded."""
    local, (start, stop) = shard_rows(qs, group)
    dev = local_device(group)
    q = torch.as_tensor(np.ascontiguousarray(local) if isinstance(local, np.ndarray) else local).to(dev, non_blocking=True)
    vals, jac = const.evaluate(q, with_jacobian)
    return vals, jac, (start, stop)


def sharded_is_valid(const, qs, group=None, validity_fn=None):
    """Validity of every row of the FULL batch ``qs``, computed shard-wise, gathered to every rank -> (N,) bool.
    ``validity_fn(local_rows) -> (n_local,) bool/uint8`` defaults to ``const.is_valid_batch`` on this rank's GPU."""
    local, (start, stop) = shard_rows(qs, group)
    if validity_fn is None:
        dev = local_device(group)
        q = torch.as_tensor(np.ascontiguousarray(local) if isinstance(local, np.ndarray) else local).to(dev)
        v = const.is_valid_batch(q)
    else:
        v = validity_fn(local)
    v = torch.as_tensor(np.asarray(v) if not (torch is not None and isinstance(v, torch.Tensor)) else v)
    rank, world = world_info(group)
    if world > 1 and dist.get_backend(group) == "gloo":
        v = v.cpu()
    return gather_validity(v, qs.shape[0], group).bool()

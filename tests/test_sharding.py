"""Multi-GPU host logic on CPU: world_size-2 gloo process group.  The compute itself needs a GPU,
so the per-shard validity function is injected; what is tested is the partition and the single collective."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from skmp_b200.sharded import gather_validity, shard_bounds, shard_rows, sharded_is_valid


def test_shard_bounds_partition_properties():
    for n in (0, 1, 7, 8, 9, 1000, 1_250_000, 10_000_001):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard_bounds(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))       # contiguous, ordered
            sizes = [e - s for s, e in blocks]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)   # balanced
    assert shard_bounds(10_000_000, 8, 3) == (3_750_000, 5_000_000)                        # BASELINE cfg3 / cfg5
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def test_single_process_is_identity():
    qs = np.arange(30.0).reshape(10, 3)
    local, (s, e) = shard_rows(qs)
    assert (s, e) == (0, 10) and local is not None and np.array_equal(local, qs)
    v = gather_validity(torch.tensor([1, 0, 1], dtype=torch.uint8), 3)
    assert v.tolist() == [1, 0, 1]
    out = sharded_is_valid(None, qs, validity_fn=lambda rows: rows[:, 0] > 10)
    assert out.tolist() == (qs[:, 0] > 10).tolist()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)                       # every rank builds the same full batch
        qs = rng.normal(size=(n, 5))
        local, (s, e) = shard_rows(qs)
        assert (s, e) == shard_bounds(n, world, rank) and np.array_equal(local, qs[s:e])
        seen = []

        def fake_validity(rows):                             # stands in for const.is_valid_batch on this rank's GPU
            seen.append(len(rows))
            return (rows.sum(axis=1) > 0).astype(np.uint8)

        valid = sharded_is_valid(None, qs, validity_fn=fake_validity)
        assert seen == [e - s]                               # each rank touched only its own rows
        q.put((rank, valid.numpy().tolist()))
        with pytest.raises(ValueError):                      # a shard of the wrong size is rejected before the collective
            gather_validity(torch.zeros(e - s + 1, dtype=torch.uint8), n)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [11, 64, 1])
def test_world_size_2_gloo_gather(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    rng = np.random.default_rng(0)
    expect = (rng.normal(size=(n, 5)).sum(axis=1) > 0).tolist()
    assert results[0] == expect and results[1] == expect     # every rank ends with the full, ordered result

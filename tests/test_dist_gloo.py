"""world_size-2 gloo tests (CPU) of the data-parallel host logic: sharding, metric reduction, mask gathering and the
gradient all-reduce + identical-update property the trainer relies on."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ucod_dpl_b200 import dist as ud
        n = 11
        mine = list(ud.shard_indices(n))
        # every item exactly once across ranks, interleaved
        assert mine == list(range(rank, n, world))
        # metric reduction: per-image values i and i^2, global sums must equal the serial sums
        vals = torch.tensor([[float(i), float(i * i)] for i in mine], dtype=torch.float64)
        sums, count = ud.reduce_metric_sums(vals.sum(0), len(mine))
        assert count == n
        assert torch.allclose(sums, torch.tensor([sum(range(n)), sum(i * i for i in range(n))], dtype=torch.float64))
        # mask gather restores item order on rank 0
        local = torch.stack([torch.full((4, 4), i, dtype=torch.uint8) for i in mine])
        full = ud.gather_sharded_masks(local, n)
        if rank == 0:
            assert full.shape == (n, 4, 4) and all(int(full[i, 0, 0]) == i for i in range(n))
        else:
            assert full is None
        # gradient all-reduce: mean-of-grads over ranks == gradient of the global batch (what DDP would compute)
        g = torch.Generator().manual_seed(0)
        w = torch.randn(16, generator=g)
        x = torch.randn(8, 16, generator=g)
        y = torch.randn(8, generator=g)
        xs, ys = x[rank::world], y[rank::world]
        grad_local = 2 * xs.t() @ (xs @ w - ys) / xs.shape[0]
        flat = grad_local.clone()
        dist.all_reduce(flat)
        flat /= world
        grad_global = 2 * x.t() @ (x @ w - y) / x.shape[0]
        assert torch.allclose(flat, grad_global, atol=1e-5)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_two_rank_host_logic():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_single_process_passthrough():
    from ucod_dpl_b200 import dist as ud
    assert list(ud.shard_indices(5)) == [0, 1, 2, 3, 4]
    s, c = ud.reduce_metric_sums(torch.tensor([1.0, 2.0]), 3)
    assert c == 3 and s.tolist() == [1.0, 2.0]
    m = torch.zeros(3, 2, 2, dtype=torch.uint8)
    assert ud.gather_sharded_masks(m, 3) is m

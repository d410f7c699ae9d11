"""N > 1 host path on CPU: two gloo ranks shard a session list, each takes its own part, and the union is exact
(the only cross-rank traffic in the product is this kind of host-side bookkeeping + the bench's max-over-ranks)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from llmvox_b200.sharding import owner, shard, shard_indices, unshard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_sessions, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_indices(n_sessions, world, rank)
    # every rank reports how many sessions it owns and a checksum of their ids
    t = torch.tensor([len(mine), sum(mine)], dtype=torch.int64)
    gathered = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, t)
    # bench.py's timing reduction: max over ranks of the elapsed time
    el = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(el, op=dist.ReduceOp.MAX)
    out_q.put((rank, mine, [g.tolist() for g in gathered], float(el)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_sessions", [64, 4097])
def test_two_ranks_partition_sessions_exactly(n_sessions):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_sessions, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    parts = [r[1] for r in res]
    assert sorted(parts[0] + parts[1]) == list(range(n_sessions))
    assert not set(parts[0]) & set(parts[1])
    assert abs(len(parts[0]) - len(parts[1])) <= 1
    for rank, mine, gathered, el in res:
        assert gathered == [[len(parts[0]), sum(parts[0])], [len(parts[1]), sum(parts[1])]]
        assert el == 11.0
        assert all(owner(i, world) == rank for i in mine)
    assert unshard(parts) == list(range(n_sessions))


def test_shard_helpers():
    items = list("abcdefg")
    assert shard(items, 3, 0) == ["a", "d", "g"] and shard(items, 3, 2) == ["c", "f"]
    assert unshard([shard(items, 3, r) for r in range(3)]) == items
    with pytest.raises(ValueError):
        shard_indices(4, 2, 2)

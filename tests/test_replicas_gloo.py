"""N > 1 host logic on CPU: two processes, gloo backend (the GPU path uses the same code on nccl)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from itermvs_b200 import replicas


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert replicas.world() == (rank, world)
        units = replicas.shard_units(7, rank, world)
        replicas.barrier()
        # rank r pretends to need (r + 1) * 10 ms for its units
        mx = replicas.max_over_ranks([10.0 * (rank + 1), float(len(units))])
        tp = replicas.aggregate_throughput(len(units), 10.0 * (rank + 1))
        # replicas are independent: a deterministic per-unit "result" must not depend on who computed it
        res = {u: float(torch.manual_seed(u).initial_seed() % 97) for u in units}
        q.put((rank, units, mx, tp, res))
    finally:
        dist.destroy_process_group()


def test_two_rank_replicas_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    out.sort()
    all_units = sorted(u for _, units, _, _, _ in out for u in units)
    assert all_units == list(range(7))                       # every unit exactly once
    assert out[0][1] == [0, 2, 4, 6] and out[1][1] == [1, 3, 5]
    for _, _, mx, tp, _ in out:
        assert mx == [20.0, 4.0]                             # MAX over ranks, identical on both
        assert abs(tp - 7 / 0.020) < 1e-6                    # all units / slowest rank
    merged = {}
    for _, _, _, _, res in out:
        merged.update(res)
    assert merged == {u: float(u % 97) for u in range(7)}


def test_single_process_defaults():
    assert replicas.world() == (0, 1)
    assert replicas.shard_units(5, 0, 1) == [0, 1, 2, 3, 4]
    assert replicas.max_over_ranks([3.0]) == [3.0]
    assert replicas.aggregate_throughput(4, 8.0) == 500.0
    with pytest.raises(ValueError):
        replicas.shard_units(5, 2, 2)

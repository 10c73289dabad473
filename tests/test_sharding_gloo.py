"""CPU, world_size 2 over gloo: the N>1 host logic of bench.py — partition placement is a disjoint cover, routing by
crc64(hashkey) agrees between ranks, and the timing reduction is the max over ranks."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from incubator_pegasus_b200 import sharding
    mine = sharding.partitions_of_rank(256, rank, world)
    t = torch.zeros(256, dtype=torch.int64)
    t[mine] = 1
    dist.all_reduce(t)  # every partition owned exactly once
    keys = [b"user%d" % i for i in range(1000)]
    routed = torch.tensor([sharding.partition_of(k, 256) for k in keys])
    other = routed.clone()
    dist.broadcast(other, src=0)
    ms = torch.tensor([10.0 + 5.0 * rank], dtype=torch.float64)

    def reduce_max(x):
        v = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        return float(v)

    rate = sharding.whole_job_rate(100.0, float(ms) / 1e3, world, reduce_max)
    local = sum(1 for k in keys if sharding.owner_rank(sharding.partition_of(k, 256), world) == rank)
    lt = torch.tensor([local])
    dist.all_reduce(lt)
    q.put((rank, bool((t == 1).all()), bool((routed == other).all()), rate, int(lt)))
    dist.destroy_process_group()


def test_two_rank_sharding_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    for rank, cover, same, rate, total in res:
        assert cover and same
        assert rate == pytest.approx(2 * 100.0 / 0.015)  # slowest rank (15 ms) sets the job time
        assert total == 1000

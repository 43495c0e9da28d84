"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: unique-id broadcast, input sharding,
file ownership (balanced, and slot j of all ranks = a contiguous file range, which the pass planner
relies on), metric reductions."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import meryl_b200 as mb
from meryl_b200 import dist as mbd


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        uid = mbd.broadcast_unique_id(lambda: bytes(range(128)))
        first, cnt = mbd.shard_range(1001, rank, world)
        total = mbd.all_ranks_sum(cnt)
        mx = mbd.all_ranks_max(float(rank + 1))
        q.put((rank, uid == bytes(range(128)), first, cnt, total, mx, mbd.owned_files(rank, world)))
    finally:
        dist.destroy_process_group()


def test_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, ok0, f0, c0, t0, m0, o0), (r1, ok1, f1, c1, t1, m1, o1) = res
    assert ok0 and ok1
    assert (f0, c0, f1, c1) == (0, 501, 501, 500) and t0 == t1 == 1001 and m0 == m1 == 2.0
    assert sorted(o0 + o1) == list(range(64)) and len(o0) == len(o1) == 32


@pytest.mark.parametrize("world", [1, 2, 4, 8, 16, 32])
def test_ownership_map(world):
    owners = [mb.file_owner(f, world) for f in range(64)]
    assert owners == [min(f, 63 - f) % world for f in range(64)]
    per = 64 // world
    for r in range(world):
        mine = [f for f in range(64) if owners[f] == r]
        assert mine == mbd.owned_files(r, world) and len(mine) == per
        # heavy (low) files are paired with light (high) ones: the linear canonical skew cancels
        weight = sum(2.0 * (63.5 - f) / 64 for f in mine)
        assert abs(weight - per) < 1e-9
    # slot j of every rank together = files [j*world, (j+1)*world)
    for j in range(per):
        slot = sorted(mbd.owned_files(r, world)[j] for r in range(world))
        assert slot == list(range(j * world, (j + 1) * world))

"""Multi-GPU host logic on CPU: sharding plan + the all-to-all exchanges with world_size 2
over gloo ('all-to-all round-trip = identity permutation')."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from recommend_tf2_b200.sharded import (ShardLayout, exchange_ids, exchange_rows_bwd,
                                        exchange_rows_fwd, plan_table_owners)

CRITEO = [1460, 583, 10000000, 2202608, 305, 24, 12517, 633, 3, 93145, 5683, 8351593, 3194, 27,
          14992, 5461306, 10, 5652, 2173, 4, 7046547, 18, 15, 286181, 105, 142572]


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_plan_balances_lookups_then_bytes(world):
    owners = plan_table_owners(CRITEO, [128] * 26, world)
    assert owners == plan_table_owners(CRITEO, [128] * 26, world)          # deterministic
    counts = [owners.count(g) for g in range(world)]
    assert max(counts) - min(counts) <= 1
    nbytes = [sum(CRITEO[t] for t in range(26) if owners[t] == g) for g in range(world)]
    if world > 1:   # the four multi-million-row tables land on different ranks
        big = sorted(range(26), key=lambda t: -CRITEO[t])[:min(4, world)]
        assert len({owners[t] for t in big}) == len(big)
        assert max(nbytes) < 0.75 * sum(nbytes)
    lay = ShardLayout(CRITEO, [128] * 26, world, 0, owners)
    assert sorted(t for s in lay.slots for t in s) == list(range(26))
    offs, total = lay.block_offsets(16, 128)
    assert total == 16 * 26 * 128
    seen = set()
    for t in range(26):
        off, st = lay.row_location(t, 16, 128)
        seen.add(off)
        assert 0 <= off < total and st % 128 == 0
    assert len(seen) == 26


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, F, D, B_local):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rows = [100 * (t + 1) for t in range(F)]
        lay = ShardLayout(rows, [D] * F, world, rank)
        B = B_local * world
        # ids exchange: rank-major concatenation
        ids_local = (torch.arange(B_local * F).view(B_local, F) + 1000 * rank).to(torch.int32)
        ids = exchange_ids(ids_local, world)
        for r in range(world):
            assert torch.equal(ids[r * B_local:(r + 1) * B_local] - 1000 * r, ids_local - 1000 * rank)
        # forward exchange: value encodes (table, global sample, d)
        mine = lay.slots[rank]
        b = torch.arange(B).view(B, 1, 1)
        tt = torch.tensor(mine).view(1, -1, 1)
        d = torch.arange(D).view(1, 1, D)
        local_out = (10000.0 * tt + 10.0 * b + d).to(torch.float32).reshape(B, len(mine) * D)
        recv, _ = exchange_rows_fwd(local_out, lay, B_local, D)
        for t in range(F):
            off, st = lay.row_location(t, B_local, D)
            for bl in range(B_local):
                row = recv[off + bl * st: off + bl * st + D]
                want = 10000.0 * t + 10.0 * (rank * B_local + bl) + torch.arange(D)
                assert torch.equal(row, want.float()), (rank, t, bl)
        # reverse exchange is the inverse permutation
        back, _ = exchange_rows_bwd(recv * 2, lay, B_local, D)
        assert torch.equal(back, local_out * 2)
        # async handles work too
        recv2, work = exchange_rows_fwd(local_out, lay, B_local, D, async_op=True)
        work.wait()
        assert torch.equal(recv2, recv)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,F", [(2, 5), (2, 26), (3, 7)])
def test_exchange_round_trip_gloo(world, F):
    mp.spawn(_worker, args=(world, _free_port(), F, 4, 3), nprocs=world, join=True)

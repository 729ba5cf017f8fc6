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


@pytest.mark.parametrize("world,F", [(2, 5), (2, 26), (3, 7), (4, 26)])
def test_exchange_round_trip_gloo(world, F):
    mp.spawn(_worker, args=(world, _free_port(), F, 4, 3), nprocs=world, join=True)


# ---- table-wise + row-wise placement of the peer-memory path (host logic only) ---------------
from recommend_tf2_b200.sharded import PeerLayout, local_shard_ids  # noqa: E402


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_peer_layout_row_wise_and_table_wise(world):
    lay = PeerLayout(CRITEO, [128] * 26, world)
    big = [t for t in range(26) if CRITEO[t] >= 5_000_000]
    assert len(big) == 4
    assert [t for t in range(26) if lay.row_wise[t]] == (big if world > 1 else [])
    assert lay.rw_mask == sum(1 << t for t in big) * (world > 1)
    # every table-wise table has exactly one owner, every row-wise table a shard on every rank
    for t in range(26):
        holders = [g for g in range(world) if t in lay.fields[g]]
        assert holders == (list(range(world)) if lay.row_wise[t] else [lay.owners[t]])
        assert sum(lay.local_rows(g, t) for g in holders) == CRITEO[t]
    counts = [len(f) for f in lay.fields]
    assert max(counts) - min(counts) <= 1
    # shard bytes are balanced now that the multi-million-row tables are split
    sizes = [lay.shard_offsets(g)[1] for g in range(world)]
    assert lay.buffer_elems() == max(sizes)
    if world > 1:
        assert max(sizes) < 1.5 * (sum(sizes) / world)    # the 2.2 M-row table stays whole
    # pointer tables: one distinct, 16-byte aligned address per shard / gradient column
    tab_ptrs = [(g + 1) << 40 for g in range(world)]
    grad_ptrs = [(g + 101) << 40 for g in range(world)]
    tab, gptr, gstr = lay.peer_pointer_tables(tab_ptrs, grad_ptrs, 128)
    seen = set()
    for t in range(26):
        n = world if lay.row_wise[t] else 1
        for e in range(n):
            g = e if lay.row_wise[t] else lay.owners[t]
            assert tab[t][e] >> 40 == g + 1 and tab[t][e] % 16 == 0
            assert gptr[t][e] >> 40 == g + 101 and gptr[t][e] % 16 == 0
            assert gstr[t][e] == len(lay.fields[g]) * 128
            seen.add(tab[t][e])
            seen.add(gptr[t][e])
    assert len(seen) == 2 * sum(world if lay.row_wise[t] else 1 for t in range(26))


@pytest.mark.parametrize("world", [2, 3, 8])
def test_local_shard_ids_partition_every_lookup_once(world):
    rows = [7, 1000, 13, 4097, 5]
    lay = PeerLayout(rows, [8] * 5, world, row_wise_min_rows=1000)
    assert lay.row_wise == [False, True, False, True, False]
    g = torch.Generator().manual_seed(world)
    B = 257
    ids = torch.stack([torch.randint(0, r, (B,), generator=g) for r in rows], 1).to(torch.int32)
    ids[3, 1] = -1            # invalid ids stay invalid on every rank
    ids[5, 3] = 4097
    hits = torch.zeros(B, 5, dtype=torch.int32)
    for rank in range(world):
        loc = local_shard_ids(ids, lay, rank)
        assert loc.dtype == torch.int32 and loc.shape == (B, len(lay.fields[rank]))
        for j, t in enumerate(lay.fields[rank]):
            col = loc[:, j]
            ok = (col >= 0) & (col < lay.local_rows(rank, t))
            hits[:, t] += ok.int()
            if lay.row_wise[t]:      # local row maps back to the global row, on this rank
                back = col[ok].long() * world + rank
                assert torch.equal(back, ids[ok, t].long())
                for b in torch.nonzero(ok).flatten()[:20].tolist():
                    assert lay.holder(t, int(ids[b, t])) == (rank, int(col[b]))
            else:
                assert torch.equal(col, ids[:, t])
    want = torch.ones(B, 5, dtype=torch.int32)
    want[3, 1] = 0
    want[5, 3] = 0
    assert torch.equal(hits, want)


@pytest.mark.parametrize("world", [2, 8])
def test_peer_layout_replicates_small_tables(world):
    lay = PeerLayout(CRITEO, [128] * 26, world, replicate_max_rows=16384)
    rep = [t for t in range(26) if CRITEO[t] <= 16384]
    assert lay.rep_fields == rep and len(rep) == 18
    assert sum(CRITEO[t] for t in rep) == 47398          # 24 MB of rows to all-reduce per step
    for g in range(world):
        f = lay.fields[g]
        assert f[len(lay.shard_fields[g]):] == rep        # replicated tables come last, on every rank
        assert not set(lay.shard_fields[g]) & set(rep)
        # the replicated shards are contiguous at the end of the rank's table buffer
        off, total = lay.shard_offsets(g)
        o = off[rep[0]]
        for t in rep:
            assert off[t] == o
            o += CRITEO[t] * 128
        assert o == total
    shard_tables = sorted({t for g in range(world) for t in lay.shard_fields[g]})
    assert shard_tables == [t for t in range(26) if t not in rep]
    assert [t for t in range(26) if lay.row_wise[t]] == [t for t in range(26) if CRITEO[t] >= 5_000_000]
    # pointer tables: a replicated field points at the asking rank's own copy / buffer
    tab_ptrs = [(g + 1) << 40 for g in range(world)]
    grad_ptrs = [(g + 101) << 40 for g in range(world)]
    for me in (0, world - 1):
        tab, gptr, gstr = lay.peer_pointer_tables(tab_ptrs, grad_ptrs, 128, rank=me)
        for t in rep:
            assert tab[t][0] >> 40 == me + 1 and gptr[t][0] >> 40 == me + 101
            assert gstr[t][0] == len(lay.fields[me]) * 128
    # ids handed to the holders cover the sharded tables only
    ids = torch.stack([torch.randint(0, r, (64,)) for r in CRITEO], 1).to(torch.int32)
    loc = local_shard_ids(ids, lay, 0)
    assert loc.shape == (64, len(lay.shard_fields[0]))


# ---- replicated (data-parallel) small tables: dense gradient combine + touched-row update -----
from recommend_tf2_b200.embedding import SparseOptimizer  # noqa: E402
from recommend_tf2_b200.sharded import apply_touched_rows, scatter_unique_rows  # noqa: E402
from oracle import embedding as OE  # noqa: E402


def _rep_worker(rank, world, port, kind):
    """Every rank reduces its own lookups (what K2 emits: sorted unique keys + summed rows),
    scatters them into the dense block, all-reduces it and updates the rows touched anywhere;
    the replicas must stay identical and equal the single-process update over the whole batch."""
    import numpy as np
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rows, D, Ts, B_local = [5, 40, 3], 8, 2, 64          # tables 2..4 replicated; 0..1 sharded
        R = sum(rows)
        row_bits = 13                                         # as if a sharded table had 8192 rows
        rep_off = torch.tensor([0, 5, 45, 48])
        g = torch.Generator().manual_seed(7)
        W0 = torch.randn(R, D, generator=g)
        ids_all = torch.stack([torch.randint(0, r, (B_local * world,), generator=g) for r in rows], 1)
        grad_all = torch.randn(B_local * world, len(rows), D, generator=g)
        opt = SparseOptimizer(kind, lr=1e-2, l2=1e-3)
        # --- this rank: unique (table,row) keys + sums over ITS samples, padded like K2's output
        sl = slice(rank * B_local, (rank + 1) * B_local)
        dense = torch.zeros(R, D)
        for j in range(len(rows)):
            dense.index_add_(0, rep_off[j] + ids_all[sl, j], grad_all[sl, j])
        keys = torch.cat([((Ts + j) << row_bits) | torch.unique(ids_all[sl, j]) for j in range(len(rows))])
        idx = torch.cat([rep_off[j] + torch.unique(ids_all[sl, j]) for j in range(len(rows))])
        cap = R + 1
        uk = torch.full((cap,), 0xFFFFFFFF, dtype=torch.int64)
        uk[: keys.numel()] = keys
        sums = torch.zeros(cap, D)
        sums[: keys.numel()] = dense[idx]
        G = scatter_unique_rows(uk, sums, row_bits, Ts, rep_off, R, D)
        assert torch.equal(G[:R, :D], dense) and torch.equal(G[:R, D] > 0, dense.abs().sum(1) > 0)
        dist.all_reduce(G)
        W, m, v = W0.clone(), torch.zeros(R, D), torch.zeros(R, D)
        lr_t = opt.struct_for_step(1).lr
        apply_touched_rows(opt, lr_t, W, m, v, G)
        # --- reference: the numpy oracle's row update over the union of all ranks' lookups
        Wn = [W0[rep_off[j]: rep_off[j + 1]].numpy().copy() for j in range(len(rows))]
        s1 = [np.zeros_like(w) for w in Wn]
        s2 = [np.zeros_like(w) for w in Wn]
        full = torch.zeros(R, D)
        for j in range(len(rows)):
            full.index_add_(0, rep_off[j] + ids_all[:, j], grad_all[:, j])
        for j in range(len(rows)):
            for r in torch.unique(ids_all[:, j]).tolist():
                OE.sparse_optimizer_step(kind, Wn[j], s1[j], s2[j], r,
                                         full[rep_off[j] + r].numpy(), lr_t, eps=opt.eps, l2=opt.l2)
        want = torch.from_numpy(np.concatenate(Wn, 0))
        torch.testing.assert_close(W, want, rtol=2e-5, atol=1e-6)
        untouched = torch.ones(R, dtype=torch.bool)
        for j in range(len(rows)):
            untouched[rep_off[j] + torch.unique(ids_all[:, j])] = False
        assert torch.equal(W[untouched], W0[untouched]) and not bool(m[untouched].any())
        gathered = [torch.empty_like(W) for _ in range(world)]
        dist.all_gather(gathered, W)
        assert all(torch.equal(gathered[0], t) for t in gathered)        # replicas stay identical
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,kind", [(2, "adam"), (2, "adagrad"), (2, "sgd"), (4, "adam")])
def test_replicated_tables_combine_and_update_gloo(world, kind):
    mp.spawn(_rep_worker, args=(world, _free_port(), kind), nprocs=world, join=True)


# ---- property test: any table list / world size gives a consistent placement -------------------
from hypothesis import given, settings  # noqa: E402
from hypothesis import strategies as st  # noqa: E402


@settings(max_examples=60, deadline=None)
@given(st.lists(st.integers(1, 3_000_000), min_size=1, max_size=40), st.sampled_from([1, 2, 3, 4, 8]),
       st.integers(0, 50_000), st.integers(100_000, 2_000_000))
def test_peer_layout_is_consistent_for_any_tables(rows, world, rep_max, rw_min):
    D = 16
    lay = PeerLayout(rows, [D] * len(rows), world, row_wise_min_rows=rw_min, replicate_max_rows=rep_max)
    n = len(rows)
    for t in range(n):
        kinds = [lay.replicated[t], lay.row_wise[t], lay.owners[t] >= 0]
        assert sum(kinds) == 1                                   # exactly one placement class
        holders = [g for g in range(world) if t in lay.fields[g]]
        if lay.replicated[t] or lay.row_wise[t]:
            assert holders == list(range(world))
        else:
            assert holders == [lay.owners[t]]
        if not lay.replicated[t]:                                # sharded rows add up to the table
            assert sum(lay.local_rows(g, t) for g in holders) == rows[t]
    if world == 1:
        assert not any(lay.replicated) and not any(lay.row_wise)
    for g in range(world):
        assert lay.fields[g] == lay.shard_fields[g] + lay.rep_fields
        off, total = lay.shard_offsets(g)
        spans = sorted((off[t], off[t] + lay.local_rows(g, t) * D) for t in lay.fields[g])
        if not spans:                                            # more ranks than tables
            assert total == 0
            continue
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))     # shards tile the buffer
        assert lay.buffer_elems() >= total
    # a few random rows: the holder map agrees with the shard sizes
    for t in range(min(n, 5)):
        if lay.replicated[t]:
            continue
        for r in {0, rows[t] - 1, rows[t] // 2}:
            g, local = lay.holder(t, r)
            assert 0 <= g < world and 0 <= local < lay.local_rows(g, t)

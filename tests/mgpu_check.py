"""Run under torchrun (>= 2 GPUs): sharded DLRM == single-GPU DLRM on the same global batch.
Checks the forward predictions and, after one training step, every table and MLP weight."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import recommend_tf2_b200 as pkg  # noqa: E402
from recommend_tf2_b200.sharded import PeerShardedDLRM, ShardedDLRM, ShardedDLRMTrainer  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    torch.backends.cuda.matmul.allow_tf32 = False
    F, D, B_local = 26, 32, 48
    rows = [37 + 11 * t for t in range(F)]
    fc = [[{"feat": f"I{i}"} for i in range(13)],
          [{"feat": f"C{t}", "feat_num": rows[t], "embed_dim": D} for t in range(F)]]
    kw = dict(bot_dnn_hidden_units=(64, D), top_dnn_hidden_units=(128, 64), input_bn=False)
    mode = os.environ.get("RTF_EXCHANGE", "peer")
    single = pkg.DLRM(fc, seed=5, **kw)
    if mode == "peer":   # tables with >= 200 rows are split row-wise, the rest placed table-wise
        # rows >= 200: row-wise; rows <= RTF_REPLICATE (default 70: 4 tables): replicated; else table-wise
        sharded = PeerShardedDLRM(fc, seed=5, row_wise_min_rows=200,
                                  gather=os.environ.get("RTF_PEER_GATHER", "owner"),
                                  replicate_max_rows=int(os.environ.get("RTF_REPLICATE", "70")), **kw)
        lay = sharded.layout
        assert any(lay.row_wise) and not all(lay.row_wise)
        assert sum(lay.replicated) == sum(r <= int(os.environ.get("RTF_REPLICATE", "70")) for r in rows)
        mine = lay.fields[rank]

        def shard_of(w, t):
            return w[rank::world] if lay.row_wise[t] else w
    else:
        sharded = ShardedDLRM(fc, seed=5, exchange=mode, **kw)
        mine = sharded.layout.slots[rank]

        def shard_of(w, t):
            return w
    g = torch.Generator(device="cuda").manual_seed(99)
    B = B_local * world
    dense = torch.rand(B, 13, device="cuda", generator=g)
    sparse = torch.stack([torch.randint(0, r, (B,), device="cuda", generator=g) for r in rows], 1).to(torch.int32)
    y = (torch.rand(B, 1, device="cuda", generator=g) < 0.3).float()
    sl = slice(rank * B_local, (rank + 1) * B_local)

    # build both, then copy weights single -> sharded
    with torch.no_grad():
        single([dense, sparse])
        sharded([dense[sl], sparse[sl]])
        for j, t in enumerate(mine):
            sharded.embed_layers.weights[j].copy_(shard_of(single.embed_layers.weights[t], t))
        for ps, pd in zip(single.dense_parameters(), sharded.dense_parameters()):
            pd.copy_(ps)
        p1 = single([dense, sparse])
        p2 = sharded([dense[sl], sparse[sl]])
    torch.testing.assert_close(p2, p1[sl], rtol=1e-5, atol=1e-6)

    t1 = pkg.DLRMTrainer(single, lr=1e-2)
    t2 = ShardedDLRMTrainer(sharded, lr=1e-2)
    for _ in range(3):
        l1 = t1.step(dense, sparse, y)
        l2 = t2.step(dense[sl], sparse[sl], y[sl])
    lsum = l2.clone()
    dist.all_reduce(lsum)
    torch.testing.assert_close(lsum / world, l1, rtol=1e-5, atol=1e-6)
    for j, t in enumerate(mine):
        torch.testing.assert_close(sharded.embed_layers.weights[j],
                                   shard_of(single.embed_layers.weights[t], t), rtol=1e-4, atol=2e-6)
        torch.testing.assert_close(sharded.embed_layers.state1[j],
                                   shard_of(single.embed_layers.state1[t], t), rtol=1e-4, atol=1e-7)
    for ps, pd in zip(single.dense_parameters(), sharded.dense_parameters()):
        torch.testing.assert_close(pd, ps, rtol=1e-4, atol=2e-6)
    sharded.embed_layers.check_ids()
    dist.barrier()
    if rank == 0:
        print(f"mgpu_check ok: world={world} exchange={mode}"
              + (f"/{sharded.gather}" if mode == "peer" else "") + f" owners={sharded.layout.owners}"
              + (f" row_wise={[t for t in range(F) if sharded.layout.row_wise[t]]}"
                 f" replicated={sharded.layout.rep_fields}" if mode == "peer" else ""))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

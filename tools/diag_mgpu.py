#!/usr/bin/env python
"""Where does a sharded training step spend its time?  Run under torchrun (>= 2 GPUs).  Rank 0
writes one JSON object (stdout, and --out FILE): the real (overlapped) step time, the device time
of every kernel of one step on rank 0 (CUPTI), grouped into phases, and the NVLink bytes the
exchange moves per step against the 770 GB/s/direction reference.
    torchrun --nproc-per-node 8 tools/diag_mgpu.py --out gpurun_out/r2_mgpu_phases_n8.json"""
import argparse
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench
import recommend_tf2_b200 as pkg
from recommend_tf2_b200.sharded import PeerShardedDLRM, ShardedDLRM, ShardedDLRMTrainer

ap = argparse.ArgumentParser()
ap.add_argument("--exchange", default="peer")
ap.add_argument("--peer-gather", default="owner")
ap.add_argument("--replicate-max-rows", type=int, default=-1)
ap.add_argument("--batch", type=int, default=65536)
ap.add_argument("--steps", type=int, default=8)
ap.add_argument("--out", default="")
a = ap.parse_args()
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
torch.backends.cuda.matmul.allow_tf32 = False
if a.replicate_max_rows < 0:
    a.replicate_max_rows = 16384 if world >= 4 else 0
fc = pkg.criteo_feature_columns(bench.EMBED_DIM, rows=bench.CRITEO_ROWS)
if a.exchange == "peer":
    m = PeerShardedDLRM(fc, bench.BOT_MLP, bench.TOP_MLP, seed=1, pad_to=8, gather=a.peer_gather,
                        replicate_max_rows=a.replicate_max_rows)
else:
    m = ShardedDLRM(fc, bench.BOT_MLP, bench.TOP_MLP, seed=1, pad_to=8, exchange=a.exchange)
tr = ShardedDLRMTrainer(m, lr=1e-3)
host = bench.make_batches(3 + a.steps, a.batch, bench.CRITEO_ROWS, "uniform", seed=100 + rank)
dev = [tuple(t.cuda() for t in b) for b in host]
for i in range(3):
    tr.step(*dev[i])
torch.cuda.synchronize()
dist.barrier()
times = []
for i in range(3, 3 + a.steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tr.step(*dev[i])
    e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
t = torch.tensor([statistics.median(times)], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
step_ms = float(t[0])

from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.step(*dev[3])
    torch.cuda.synchronize()
if rank == 0:
    PH = [("dense GEMMs (tcgen05 bf16x6 + cuBLAS first/last layer)", ("FastF32", "gemm", "Gemm", "gemv")),
          ("holder-side K1 gather (global batch)", ("embed_fwd",)),
          ("NVLink row pull (peer_pull_kernel)", ("peer_pull",)),
          ("K4 interaction forward", ("dot_fwd",)),
          ("K4 backward + NVLink gradient push", ("dot_bwd",)),
          ("K2 keys / sort / segments (side stream)", ("make_keys", "sort_", "scan_")),
          ("K2 apply: segment reduce + sparse Adam", ("seg_apply",)),
          ("cross-GPU barriers (symmetric memory)", ("barrier", "Barrier", "signal")),
          ("NCCL collectives (ids all-gather, MLP / replicated-row all-reduce)", ("nccl", "Nccl")),
          ("ReLU mask + bias gradient", ("relu_bwd", "colsum")),
          ("dense Adam + replicated-row update", ("dense_adam", "rows_apply"))]
    agg, kern = {n: 0.0 for n, _ in PH}, {}
    agg["other (BatchNorm, loss, elementwise, memsets)"] = 0.0
    for ev in prof.events():
        us = ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
        k = kern.setdefault(ev.name[:100], [0.0, 0])
        k[0] += us
        k[1] += 1
        for n, pats in PH:
            if any(p in ev.name for p in pats):
                agg[n] += us
                break
        else:
            agg["other (BatchNorm, loss, elementwise, memsets)"] += us
    lay = getattr(m, "layout", None)
    B, D, F = a.batch, bench.EMBED_DIM, len(bench.CRITEO_ROWS)
    remote = None
    if a.exchange == "peer":
        n_rep = len(lay.rep_fields)
        n_rw = sum(lay.row_wise)
        n_tw_remote = sum(1 for t in range(F) if not lay.row_wise[t] and not lay.replicated[t]
                          and lay.owners[t] != 0)
        remote_rows = n_tw_remote + n_rw * (world - 1) / world
        remote = {"fields_replicated": n_rep, "fields_row_wise": n_rw,
                  "remote_rows_per_sample": round(remote_rows, 2),
                  "bytes_pulled_per_step": int(B * remote_rows * D * 4),
                  "bytes_pushed_per_step": int(B * remote_rows * D * 4),
                  "ms_at_770GBs_each_way": round(B * remote_rows * D * 4 / 770e9 * 1e3, 3)}
    res = {"world": world, "exchange": a.exchange, "gather": a.peer_gather,
           "replicate_max_rows": a.replicate_max_rows, "batch_per_gpu": a.batch,
           "step_ms_overlapped_max_over_ranks": round(step_ms, 3),
           "kernel_ms_sum_rank0": round(sum(v[0] for v in kern.values()) / 1e3, 3),
           "phases_ms_rank0": {n: round(v / 1e3, 3) for n, v in agg.items()},
           "nvlink": remote,
           "top_kernels_rank0": [{"us": round(us, 1), "n": n, "name": nm}
                                 for nm, (us, n) in sorted(kern.items(), key=lambda kv: -kv[1][0])[:14]],
           "note": "phase times are device times of rank 0's kernels in ONE step (CUPTI); phases on the "
                   "exchange / side streams overlap the dense GEMMs, so their sum exceeds the step time"}
    txt = json.dumps(res, indent=1)
    print(txt)
    if a.out:
        with open(a.out, "w") as f:
            f.write(txt + "\n")
dist.destroy_process_group()

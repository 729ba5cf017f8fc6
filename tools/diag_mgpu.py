#!/usr/bin/env python
"""Where does a sharded training step spend its time?  Run under torchrun (>= 2 GPUs): per-phase
CUDA-event times and the host-side enqueue time of each phase, rank 0 prints medians.
    torchrun --nproc-per-node 2 tools/diag_mgpu.py [--exchange peer|p2p|nccl]"""
import argparse
import os
import statistics
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench
import recommend_tf2_b200 as pkg
from recommend_tf2_b200.core import binary_crossentropy
from recommend_tf2_b200.sharded import PeerShardedDLRM, ShardedDLRM, ShardedDLRMTrainer

ap = argparse.ArgumentParser()
ap.add_argument("--exchange", default="peer")
ap.add_argument("--peer-gather", default="owner")
ap.add_argument("--replicate-max-rows", type=int, default=16384)
ap.add_argument("--batch", type=int, default=65536)
ap.add_argument("--steps", type=int, default=8)
a = ap.parse_args()
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
torch.backends.cuda.matmul.allow_tf32 = False
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

names = ["forward", "loss+backward", "finish_backward(K2)", "allreduce+copy", "dense_adam"]
gpu = {n: [] for n in names}
cpu = {n: [] for n in names}
tot_gpu, tot_cpu = [], []
for i in range(3, 3 + a.steps):
    d, s, y = dev[i]
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    ts = []
    evs[0].record(); ts.append(time.perf_counter())
    m.embed_layers.begin_step()
    pred = m([d, s])
    evs[1].record(); ts.append(time.perf_counter())
    loss = binary_crossentropy(y, pred)
    tr.dense_opt.zero_grad(set_to_none=True)
    (loss / world).backward()
    evs[2].record(); ts.append(time.perf_counter())
    m.finish_backward()
    evs[3].record(); ts.append(time.perf_counter())
    grads = [p.grad for p in m.dense_parameters() if p.grad is not None]
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat)
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()
    evs[4].record(); ts.append(time.perf_counter())
    tr.dense_opt.step()
    evs[5].record(); ts.append(time.perf_counter())
    torch.cuda.synchronize()
    for k, n in enumerate(names):
        gpu[n].append(evs[k].elapsed_time(evs[k + 1]))
        cpu[n].append((ts[k + 1] - ts[k]) * 1e3)
    tot_gpu.append(evs[0].elapsed_time(evs[-1]))
    tot_cpu.append((ts[-1] - ts[0]) * 1e3)
if rank == 0:
    print(f"exchange={a.exchange} world={world} (each step synchronised: no CPU run-ahead)")
    for n in names:
        print(f"  {n:22s} gpu {statistics.median(gpu[n]):7.3f} ms   host enqueue {statistics.median(cpu[n]):7.3f} ms")
    print(f"  {'total':22s} gpu {statistics.median(tot_gpu):7.3f} ms   host enqueue {statistics.median(tot_cpu):7.3f} ms")
# one more step under the profiler: device time per kernel on this rank
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.step(*dev[3])
    torch.cuda.synchronize()
if rank == 0:
    agg = {}
    for ev in prof.events():
        nm = ev.name[:90]
        t = agg.setdefault(nm, [0.0, 0])
        t[0] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
        t[1] += 1
    tot = sum(v[0] for v in agg.values())
    print(f"  profiled step: {tot / 1e3:.3f} ms of kernels")
    for nm, (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:16]:
        print(f"    {us:9.1f} us  n={n:3d}  {nm}")
dist.destroy_process_group()

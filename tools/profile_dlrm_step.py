#!/usr/bin/env python
"""Device time per kernel of ONE single-GPU DLRM training step at the bench configuration
(torch.profiler / CUPTI; the launch-list companion of profiles/r2_launches_bench_step.csv that
needs no ncu).   python tools/profile_dlrm_step.py [--top 40] [--batch 65536]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

import bench
import recommend_tf2_b200 as pkg
from recommend_tf2_b200 import core
from recommend_tf2_b200.data import synthetic_criteo_batch

ap = argparse.ArgumentParser()
ap.add_argument("--top", type=int, default=40)
ap.add_argument("--batch", type=int, default=65536)
a = ap.parse_args()
torch.backends.cuda.matmul.allow_tf32 = False
core.set_dense_gemm("bf16x6")
fc = pkg.criteo_feature_columns(bench.EMBED_DIM, rows=bench.CRITEO_ROWS)
model = pkg.DLRM(fc, bench.BOT_MLP, bench.TOP_MLP, interaction="dot", seed=1234, pad_to=8)
tr = pkg.DLRMTrainer(model, lr=1e-3)
rng = np.random.default_rng(0)
dev = [tuple(torch.from_numpy(np.ascontiguousarray(x)).cuda()
             for x in synthetic_criteo_batch(rng, a.batch, bench.CRITEO_ROWS, "uniform")) for _ in range(6)]
for b in dev[:5]:
    tr.step(*b)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for b in dev[:5]:
    tr.step(*b)
e1.record()
torch.cuda.synchronize()
print(f"dlrm: {e0.elapsed_time(e1) / 5:.3f} ms per step (5 steps, CUDA events)")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.step(*dev[5])
    torch.cuda.synchronize()
agg = {}
for ev in prof.events():
    t = agg.setdefault(ev.name[:120], [0.0, 0])
    t[0] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
    t[1] += 1
tot = sum(v[0] for v in agg.values())
print(f"dlrm: {tot / 1e3:.3f} ms of kernels in one step, {sum(v[1] for v in agg.values())} launches")
for nm, (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:a.top]:
    print(f"  {us:9.1f} us  n={n:3d}  {nm}")

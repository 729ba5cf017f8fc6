#!/usr/bin/env python
"""Device time per kernel of ONE training step of a bench workload (torch.profiler / CUPTI).
    python tools/profile_step.py --workload sasrec [--top 25]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

import bench_workloads as BW
import recommend_tf2_b200 as pkg

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="sasrec")
ap.add_argument("--top", type=int, default=25)
a = ap.parse_args()
torch.backends.cuda.matmul.allow_tf32 = False
wl = BW.WORKLOADS[a.workload]()
st = wl.build(pkg)
rng = np.random.default_rng(0)
dev = [tuple(torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in wl.host_batch(rng, wl.batch)) for _ in range(5)]
for b in dev[:4]:
    st.step(*b)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    st.step(*dev[4])
    torch.cuda.synchronize()
agg = {}
for ev in prof.events():
    t = agg.setdefault(ev.name[:110], [0.0, 0])
    t[0] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
    t[1] += 1
tot = sum(v[0] for v in agg.values())
print(f"{a.workload}: {tot / 1e3:.3f} ms of kernels in one step, {sum(v[1] for v in agg.values())} launches")
for nm, (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:a.top]:
    print(f"  {us:9.1f} us  n={n:3d}  {nm}")

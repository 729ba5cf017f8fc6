#!/usr/bin/env python
"""Host-side (Python) cost of a DLRM training step: cProfile over N steps without device syncs
inside, so the numbers are enqueue time.    python tools/profile_host.py [--steps 20]"""
import argparse
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import recommend_tf2_b200 as pkg

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--batch", type=int, default=65536)
a = ap.parse_args()
torch.backends.cuda.matmul.allow_tf32 = False
fc = pkg.criteo_feature_columns(bench.EMBED_DIM, rows=bench.CRITEO_ROWS)
model = pkg.DLRM(fc, bench.BOT_MLP, bench.TOP_MLP, interaction="dot", seed=1, pad_to=8)
tr = pkg.DLRMTrainer(model, lr=1e-3)
host = bench.make_batches(4 + a.steps, a.batch, bench.CRITEO_ROWS, "uniform", seed=3)
dev = [tuple(t.cuda() for t in b) for b in host]
for b in dev[:4]:
    tr.step(*b)
torch.cuda.synchronize()
t0 = time.perf_counter()
pr = cProfile.Profile()
pr.enable()
for b in dev[4:]:
    tr.step(*b)
pr.disable()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3 * (t1 - t0) / a.steps:.3f} ms/step (under cProfile), drain {1e3 * (t2 - t1):.1f} ms")
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
st.sort_stats("cumulative").print_stats(36)

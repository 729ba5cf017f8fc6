#!/usr/bin/env python
"""CUDA-event timing of the fused gather + dot-interaction forward on the bench's table set
(B = 65 536, 26 Criteo tables, D = 128).  Set RTF_DOT_FWD_V2=1 to time variant 2."""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import recommend_tf2_b200 as pkg

rows, D, B = bench.CRITEO_ROWS, bench.EMBED_DIM, 65536
ts = pkg.EmbeddingTables(rows, [D] * len(rows), seed=1)
ids = [b[1].cuda() for b in bench.make_batches(6, B, rows, "uniform", seed=7)]
dense = torch.randn(B, D, device="cuda")
with torch.no_grad():
    for i in range(3):
        pkg.embed_dot(ts, ids[i], dense, pad_to=8)
    torch.cuda.synchronize()
    ev = []
    for i in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(600_000)
        a.record()
        pkg.embed_dot(ts, ids[i], dense, pad_to=8)
        b.record()
        ev.append((a, b))
    torch.cuda.synchronize()
print("RTF_DOT_FWD_V2=%s  dot_fwd median %.4f ms" % (os.environ.get("RTF_DOT_FWD_V2"),
                                                   statistics.median(a.elapsed_time(b) for a, b in ev)))

#!/usr/bin/env python
"""ncu driver: a few forward + backward passes of the fused AutoInt interacting layer (K6) at the
BASELINE shape (B = 4096, 39 fields, 32 -> 2 heads x 16, residual)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import recommend_tf2_b200 as pkg

torch.manual_seed(0)
layer = pkg.layers.ctr.MultiHeadAttention(16, 2, use_res=True)
x = torch.randn(4096, 39, 32, device="cuda", requires_grad=True)
g = torch.randn(4096, 39, 32, device="cuda")
for _ in range(3):
    out = layer(x)
    out.backward(g)
torch.cuda.synchronize()
print("prof_autoint done")

#!/usr/bin/env python
"""One launch of each dense-GEMM layout on the biggest bench layer (M = 65 536, 1024 x 1024), for
`ncu -k regex:device_kernel`: tensor-pipe utilisation and DRAM traffic of the tcgen05 GEMM."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from recommend_tf2_b200 import core

B, K, N = 65536, 1024, 1024
x = torch.randn(B, K, device="cuda")
w = torch.randn(K, N, device="cuda") * 0.05
b = torch.randn(N, device="cuda")
g = torch.randn(B, N, device="cuda")
core.dense_gemm("nn", x, w, b, True)
core.dense_gemm("nt", g, w)
core.dense_gemm("tn", x, g, splits=core._wgrad_splits(B, K, N))
torch.cuda.synchronize()
print("prof_gemm done")

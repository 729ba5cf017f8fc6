#!/usr/bin/env python
"""Times the DLRM bench's dense-layer GEMMs (B = 65 536): librtf_b200's 6-product bf16 split on
tcgen05 vs the framework path (cuBLAS 12.9 BF16x9 emulation when preloaded, else SGEMM), with
the error of each against fp64.   python tools/probe_dense_gemm.py [--lib-emulation]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

emu = "--lib-emulation" in sys.argv
desc = bench.preload_cublas_fp32_emulation() if emu else "native SGEMM"
import torch  # noqa: E402
from recommend_tf2_b200 import core  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
B = 65536
LAYERS = [(512, 256), (256, 128), (480, 1024), (1024, 1024), (1024, 512), (512, 256)]


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def rel(out, ref):
    return float((out.double() - ref).abs().max() / ref.abs().max())


for K, N in LAYERS:
    x = torch.randn(B, K, device="cuda")
    w = torch.randn(K, N, device="cuda") * 0.05
    b = torch.randn(N, device="cuda")
    g = torch.randn(B, N, device="cuda")
    flops = 2.0 * B * K * N
    sub = slice(0, 4096)
    ref_f = (x[sub].double() @ w.double() + b.double()).clamp_min(0)
    ref_d = g[sub].double() @ w.double().t()
    ref_w = x.double().t() @ g.double()
    row = {"K": K, "N": N, "library": desc}
    for name, ours, lib, ref, cut in [
        ("fwd", lambda: core.dense_gemm("nn", x, w, b, True),
         lambda: torch._addmm_activation(b, x, w, use_gelu=False), ref_f, True),
        ("dgrad", lambda: core.dense_gemm("nt", g, w), lambda: g @ w.t(), ref_d, True),
        ("wgrad", lambda: core.dense_gemm("tn", x, g, splits=core._wgrad_splits(B, K, N)),
         lambda: torch.bmm(x.view(16, B // 16, -1).transpose(1, 2), g.view(16, B // 16, -1)).sum(0),
         ref_w, False)]:
        t_o, t_l = timed(ours), timed(lib)
        o, l = ours(), lib()
        if cut:
            o, l = o[sub], l[sub]
        row[name] = {"ours_ms": round(t_o, 4), "lib_ms": round(t_l, 4),
                     "ours_tflops_fp32": round(flops / t_o / 1e9, 1),
                     "lib_tflops_fp32": round(flops / t_l / 1e9, 1),
                     "ours_err": rel(o, ref), "lib_err": rel(l, ref)}
    print(json.dumps(row), flush=True)

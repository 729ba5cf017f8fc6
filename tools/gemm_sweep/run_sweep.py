#!/usr/bin/env python
"""Times the FastF32Gemm configuration variants of tools/gemm_sweep (built by `make` there) on the
bench's forward shapes and prints one JSON line per (variant, shape)."""
import ctypes as C
import json
import os
import subprocess
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
lib = C.CDLL(os.path.join(HERE, "libsweep.so"))
names = [l.split()[-1][len("sweep_"):] for l in subprocess.run(
    ["nm", "-D", os.path.join(HERE, "libsweep.so")], capture_output=True, text=True).stdout.splitlines()
    if " T sweep_" in l]
only = sys.argv[1:]
B = 65536
ws = torch.empty(1 << 24, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for K, N in [(1024, 1024), (480, 1024), (512, 256)]:
    x = torch.randn(B, K, device="cuda")
    w = torch.randn(K, N, device="cuda") * 0.05
    b = torch.randn(N, device="cuda")
    ref = (x[:2048].double() @ w.double() + b.double()).clamp_min(0)
    for nm in sorted(names):
        if only and not any(o in nm for o in only):
            continue
        fn = getattr(lib, "sweep_" + nm)
        fn.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p,
                       C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
        out = torch.zeros(B, N, device="cuda")

        def run():
            return fn(x.data_ptr(), K, w.data_ptr(), N, b.data_ptr(), 1, out.data_ptr(), N, B, N, K,
                      ws.data_ptr(), ws.numel(), st)
        rc = run()
        torch.cuda.synchronize()
        if rc != 0:
            print(json.dumps({"variant": nm, "K": K, "N": N, "rc": rc}), flush=True)
            continue
        err = float((out[:2048].double() - ref).abs().max() / ref.abs().max())
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            run()
        e.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(e) / 10
        print(json.dumps({"variant": nm, "K": K, "N": N, "ms": round(ms, 4),
                          "tflops_fp32": round(2.0 * B * K * N / ms / 1e9, 1), "err": err}), flush=True)

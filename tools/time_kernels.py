#!/usr/bin/env python
"""CUDA-event timing of the hot-path kernels on the bench's DLRM table set (Criteo cardinalities,
D = 128, B = 65 536, uniform ids), one JSON line per kernel.  Distinct batches per iteration
(tables >> L2), the stream kept busy while the host enqueues so launch latency is excluded.

    python tools/time_kernels.py [--which k2apply,k2,fwd,bwd,k1] [--tag NAME]
"""
import argparse
import ctypes as C
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import bench
import recommend_tf2_b200 as pkg
from recommend_tf2_b200 import _lib as L
from recommend_tf2_b200.embedding import _ptr_array


def timed(fn, n, warm=3, pre=None):
    for i in range(warm):
        if pre:
            pre(i % n)
        fn(i % n)
    torch.cuda.synchronize()
    evs = []
    for i in range(n):
        if pre:
            pre(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(600_000)
        a.record()
        fn(i)
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = [a.elapsed_time(b) for a, b in evs]
    return statistics.median(ts), min(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--iters", type=int, default=8)
    ap.add_argument("--which", default="k2apply,k2,fwd,bwd,k1")
    ap.add_argument("--ids", default="uniform")
    ap.add_argument("--tag", default="")
    a = ap.parse_args()
    which = a.which.split(",")
    rows = bench.CRITEO_ROWS
    D, F, B = bench.EMBED_DIM, len(rows), a.batch
    peak = bench.load_peaks()[0]["hbm_gbs"]
    ts = pkg.EmbeddingTables(rows, [D] * F, seed=1, optimizer=pkg.SparseOptimizer("adam", l2=1e-4))
    ts.begin_step()
    n = a.iters
    batches = [b[1].cuda() for b in bench.make_batches(n, B, rows, a.ids, seed=7)]
    cols = pkg.dot_out_cols(F + 1, D, 8)
    dense = torch.randn(B, D, device="cuda")
    gout = torch.randn(B, cols, device="cuda")
    gemb = torch.randn(B, F * D, device="cuda")
    gdense = torch.empty(B, D, device="cuda")
    out = torch.empty(B, F * D, device="cuda")
    tables = list(ts.weights)
    rows_arr = L.host_array(C.c_int64, rows)
    uniq = []
    for ids in batches:
        keys = ids.long() + (torch.arange(F, device="cuda").view(1, F) << 32)
        uniq.append(int(torch.unique(keys).numel()))
    nu = statistics.mean(uniq)
    res = {}

    def emit(name, ms, best, nbytes):
        gbs = nbytes / (ms * 1e-3) / 1e9
        res[name] = {"ms": round(ms, 4), "best_ms": round(best, 4), "algorithmic_bytes": int(nbytes),
                     "gbs": round(gbs, 1), "frac_hbm": round(gbs / peak, 4)}
        print(json.dumps({"tag": a.tag, "kernel": name, **res[name]}), flush=True)

    if "k2apply" in which:
        prepared = {}

        def pre(i):
            prepared[i] = ts.prepare_backward(batches[i], list(range(F)))
            torch.cuda.current_stream().wait_event(prepared[i]["ev"])

        ms, best = timed(lambda i: ts.apply_prepared(prepared[i], gemb), n, pre=pre)
        emit("K2 apply (seg_apply: segment reduce + Adam)", ms, best, B * F * D * 4 + nu * 6 * D * 4)
    if "k2" in which:
        ms, best = timed(lambda i: ts.apply_sparse_grad(batches[i], list(range(F)), gemb), n)
        emit("K2 pipeline (keys+sort+segments+apply)", ms, best, B * F * (4 + D * 4) + nu * 6 * D * 4)
    if "fwd" in which:
        with torch.no_grad():
            ms, best = timed(lambda i: pkg.embed_dot(ts, batches[i], dense, pad_to=8), n)
        emit("K1+K4 fwd (dot_fwd_kernel)", ms, best, B * (F * (4 + D * 4) + D * 4 + cols * 4))
    if "bwd" in which:
        def bwd(i):
            ids = batches[i]
            L.check(L.lib().rtf_embed_dot_bwd(_ptr_array(tables), rows_arr, F, D, ids.data_ptr(), 0, B,
                                              ids.stride(0), ids.stride(1), dense.data_ptr(), D,
                                              gout.data_ptr(), cols, gdense.data_ptr(), D,
                                              gemb.data_ptr(), F * D, L.current_stream_ptr()), "bwd")
        ms, best = timed(bwd, n)
        emit("K4 bwd (dot_bwd_kernel)", ms, best,
             B * (F * (4 + D * 4) + D * 4 + cols * 4 + D * 4 + F * D * 4))
    if "k1" in which:
        ms, best = timed(lambda i: pkg.embed_fwd(tables, batches[i], "BF", None, out=out), n)
        emit("K1 (embed_fwd_vec)", ms, best, B * F * (4 + D * 4 + D * 4))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Small driver for ncu: runs each hot-path kernel a few times on the bench's DLRM table set
(Criteo cardinalities, D=128, uniform ids) so that `ncu -k regex:...` can pick launches.

    python tools/prof_kernels.py [--batch 65536] [--iters 3] [--which k1,fwd,bwd,k2]
"""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import bench
import recommend_tf2_b200 as pkg
from recommend_tf2_b200 import _lib as L
from recommend_tf2_b200.embedding import _ptr_array


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--which", default="k1,fwd,bwd,k2")
    ap.add_argument("--ids", default="uniform")
    ap.add_argument("--row-cap", type=int, default=10_000_000,
                    help="cap table rows (ncu --set full saves/restores every buffer a kernel writes)")
    a = ap.parse_args()
    which = set(a.which.split(","))
    rows = [min(r, a.row_cap) for r in bench.CRITEO_ROWS]
    D, F, B = bench.EMBED_DIM, len(rows), a.batch
    ts = pkg.EmbeddingTables(rows, [D] * F, seed=1, optimizer=pkg.SparseOptimizer("adam"))
    ts.begin_step()
    batches = [b[1].cuda() for b in bench.make_batches(a.iters, B, rows, a.ids, seed=7)]
    cols = pkg.dot_out_cols(F + 1, D)
    dense = torch.randn(B, D, device="cuda")
    gout = torch.randn(B, cols, device="cuda")
    gemb = torch.randn(B, F * D, device="cuda")
    gdense = torch.empty(B, D, device="cuda")
    out = torch.empty(B, F * D, device="cuda")
    tables = list(ts.weights)
    rows_arr = L.host_array(C.c_int64, rows)
    torch.cuda.synchronize()
    for ids in batches:
        if "k1" in which:
            pkg.embed_fwd(tables, ids, "BF", None, out=out)
        if "fwd" in which:
            with torch.no_grad():
                pkg.embed_dot(ts, ids, dense)
        if "bwd" in which:
            L.check(L.lib().rtf_embed_dot_bwd(_ptr_array(tables), rows_arr, F, D, ids.data_ptr(), 0, B,
                                              ids.stride(0), ids.stride(1), dense.data_ptr(), D,
                                              gout.data_ptr(), cols, gdense.data_ptr(), D,
                                              gemb.data_ptr(), F * D, L.current_stream_ptr()), "bwd")
        if "k2" in which:
            ts.apply_sparse_grad(ids, list(range(F)), gemb)   # keys, sort, segments, work items, seg_apply
    torch.cuda.synchronize()
    print("prof_kernels done")


if __name__ == "__main__":
    main()

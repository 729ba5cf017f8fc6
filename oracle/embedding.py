"""Oracle (test infrastructure): embedding lookup, pooling, gradient and optimizers.

Follows, per function, the reference lines named in the docstring; TF semantics from
SURVEY.md Appendix A (A1 gather, A2 out-of-range, A3 l2, A12 Adam).  fp32 throughout, with
the summation orders the CUDA kernels fix, so integer work and sums compare bit for bit.
"""
from __future__ import annotations

import numpy as np

SEG_CHUNK = 64  # == RTF_SEG_CHUNK in include/rtf_b200.h
SEG_GROUP = 64  # == RTF_SEG_GROUP: chunk partials are added per group of 64 chunks, then the groups


def embed_lookup_concat(tables, ids_bfl, pool=None):
    """src/ctr/dlrm/model.py:45-46 (and deep_fm/model.py:53-54, autoint/model.py:46-47,
    din/model.py:62-74, match/sasrec/model.py:75-79): one gather per field (A1), concat on
    the last axis; optional pooling over the length axis as src/match/fm/model.py:73,77
    (`reduce_sum(axis=1)`), summed in ascending l.  An out-of-range id raises like TF's CPU
    gather (A2).
      tables : list of (N_f, D_f) fp32, one per field (a table may repeat)
      ids_bfl: (B, F, L) integer
    returns (B, L, sumD) for pool None, else (B, sumD)."""
    ids = np.asarray(ids_bfl)
    B, F, L = ids.shape
    outs = []
    for f in range(F):
        W = np.asarray(tables[f], dtype=np.float32)
        idx = ids[:, f, :].astype(np.int64)
        if idx.size and (idx.min() < 0 or idx.max() >= W.shape[0]):
            raise IndexError(f"indices out of range [0, {W.shape[0]}) for field {f}")
        e = W[idx]                                   # (B, L, D_f)  — gather, bit-exact copy
        if pool is None:
            outs.append(e)
        else:
            acc = np.zeros((B, W.shape[1]), np.float32)
            for l in range(L):                       # fixed ascending-l fp32 order
                acc = acc + e[:, l, :]
            if pool == "mean":
                acc = acc / np.float32(L)
            elif pool != "sum":
                raise ValueError(pool)
            outs.append(acc)
    return np.concatenate(outs, axis=-1)


def make_keys(ids_bfl, field_table, rows):
    """(table, id) sort keys in lookup-position order p = (b*L + l)*F + f, as K2 builds them."""
    ids = np.asarray(ids_bfl).astype(np.int64)
    B, F, L = ids.shape
    n_tables = len(rows)
    row_bits = 1
    while (1 << row_bits) < max(rows):
        row_bits += 1
    ft = np.asarray(field_table, np.int64)
    ids_blf = np.transpose(ids, (0, 2, 1)).reshape(-1)           # position order (b, l, f)
    tab = np.tile(ft, B * L)
    bad = (ids_blf < 0) | (ids_blf >= np.asarray(rows, np.int64)[tab])
    keys = (tab << row_bits) | np.where(bad, 0, ids_blf)
    keys = np.where(bad, n_tables << row_bits, keys)
    return keys.astype(np.int64), row_bits


def embed_grad_unique(ids_bfl, field_table, rows, dims, grad, pool=None, chunk=SEG_CHUNK):
    """The IndexedSlices -> UnsortedSegmentSum reduction implied by the backward of the
    gathers, with the order fixed: for every touched (table,row), gradient
    rows are added in ascending lookup position; segments longer than `chunk` are summed as
    chunk partials; the partials of every `SEG_GROUP` consecutive chunks are added in chunk order,
    the group sums in group order (one group = the plain chunk-order sum).
      grad: (B, L, sumD) for pool None, (B, sumD) for pooled lookups.
    returns (keys ascending int64, (n_unique, dim_max) fp32 sums, row_bits)."""
    ids = np.asarray(ids_bfl)
    B, F, L = ids.shape
    grad = np.asarray(grad, np.float32)
    keys, row_bits = make_keys(ids, field_table, rows)
    n = keys.shape[0]
    dim_max = max(dims)
    fdim = [dims[t] for t in field_table]
    off = np.concatenate([[0], np.cumsum(fdim)])[:-1]
    # gradient row of every lookup position, padded to dim_max
    g = np.zeros((n, dim_max), np.float32)
    p = np.arange(n)
    b, r = p // (L * F), p % (L * F)
    l, f = r // F, r % F
    scale = np.float32(1.0) / np.float32(L)
    for ff in range(F):
        sel = f == ff
        d = fdim[ff]
        if pool is None:
            rows_g = grad.reshape(B, L, -1)[b[sel], l[sel], off[ff]:off[ff] + d]
        else:
            rows_g = grad[b[sel], off[ff]:off[ff] + d]
            if pool == "mean":
                rows_g = rows_g * scale
        g[sel, :d] = rows_g
    order = np.argsort(keys, kind="stable")
    sk = keys[order]
    valid = (sk >> row_bits) < len(rows)
    sk, order = sk[valid], order[valid]
    if sk.size == 0:
        return sk, np.zeros((0, dim_max), np.float32), row_bits
    head = np.ones(sk.size, bool)
    head[1:] = sk[1:] != sk[:-1]
    seg = np.cumsum(head) - 1
    seg_start = np.flatnonzero(head)
    within = np.arange(sk.size) - seg_start[seg]
    ch = within // chunk
    chunk_head = head | (within % chunk == 0)
    cuid = np.cumsum(chunk_head) - 1                       # unique (segment, chunk) id, ascending
    partial = np.zeros((cuid[-1] + 1, dim_max), np.float32)
    np.add.at(partial, cuid, g[order])                     # sequential fp32 adds, ascending p
    # group sums: partials in chunk order inside each group of SEG_GROUP chunks ...
    pseg, pch = seg[chunk_head], ch[chunk_head]
    group_head = (pch % SEG_GROUP) == 0
    guid = np.cumsum(group_head) - 1
    gsum = np.zeros((guid[-1] + 1, dim_max), np.float32)
    np.add.at(gsum, guid, partial)
    # ... then the group sums in group order
    total = np.zeros((seg[-1] + 1, dim_max), np.float32)
    np.add.at(total, pseg[group_head], gsum)
    return sk[head], total, row_bits


def sparse_optimizer_step(kind, W, s1, s2, row, g, lr, beta1=0.9, beta2=0.999, eps=1e-7, l2=0.0):
    """Row update K2 applies (Keras formulas; `lr` already bias-corrected for
    Adam).  Operates in place on the fp32 arrays; every op is a separately rounded fp32 op in
    the order the kernel uses."""
    f = np.float32
    w = W[row]
    g = g.astype(np.float32)
    if l2 > 0:
        g = g + (f(2.0) * f(l2)) * w
    if kind == "sgd":
        W[row] = w - f(lr) * g
    elif kind == "adagrad":
        a = s1[row] + g * g
        s1[row] = a
        W[row] = w - (f(lr) * g) / (np.sqrt(a) + f(eps))
    elif kind == "adam":
        m = f(beta1) * s1[row] + (f(1.0) - f(beta1)) * g
        v = f(beta2) * s2[row] + (f(1.0) - f(beta2)) * (g * g)
        s1[row], s2[row] = m, v
        W[row] = w - (f(lr) * m) / (np.sqrt(v) + f(eps))
    else:
        raise ValueError(kind)


def adam_lr_t(lr, beta1, beta2, step):
    """Keras folds the bias corrections into the step size (A12)."""
    import math
    return lr * math.sqrt(1.0 - beta2 ** step) / (1.0 - beta1 ** step)


def embed_bwd_apply(kind, weights, state1, state2, ids_bfl, field_table, grad, pool=None,
                    lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-7, l2=0.0):
    """Gradient reduction + in-place sparse optimizer over all touched rows."""
    rows = [w.shape[0] for w in weights]
    dims = [w.shape[1] for w in weights]
    keys, total, row_bits = embed_grad_unique(ids_bfl, field_table, rows, dims, grad, pool)
    tab = keys >> row_bits
    row = keys & ((1 << row_bits) - 1)
    for t in range(len(weights)):
        sel = tab == t
        if not sel.any():
            continue
        d = dims[t]
        sparse_optimizer_step(kind, weights[t], None if state1 is None else state1[t],
                              None if state2 is None else state2[t], row[sel], total[sel, :d],
                              lr, beta1, beta2, eps, l2)
    return keys, total, row_bits


def dense_reference_grad(tables_shapes, ids_bfl, field_table, grad, pool=None):
    """Independent check in fp64: dense dL/dW_t by scatter-add (no ordering games)."""
    ids = np.asarray(ids_bfl)
    B, F, L = ids.shape
    out = [np.zeros(s, np.float64) for s in tables_shapes]
    fdim = [tables_shapes[t][1] for t in field_table]
    off = np.concatenate([[0], np.cumsum(fdim)])
    g = np.asarray(grad, np.float64)
    for f in range(F):
        t = field_table[f]
        for l in range(L):
            if pool is None:
                rows_g = g.reshape(B, L, -1)[:, l, off[f]:off[f + 1]]
            else:
                rows_g = g[:, off[f]:off[f + 1]] * (1.0 / L if pool == "mean" else 1.0)
            np.add.at(out[t], ids[:, f, l], rows_g)
    return out

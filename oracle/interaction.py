"""Oracle (test infrastructure): feature-interaction layers, restated in numpy.

  dot_interact      — DLRM pairwise dot.  The reference's DLRM.call only concatenates
                      (src/ctr/dlrm/model.py:48) and cites arXiv 1906.00091 at :7; the op is
                      defined from that paper.
  dlrm_concat       — the reference's literal `tf.concat([sparse_embed, dense_fea], -1)` (:48).
  fm_layer          — ctr.layers.modules.FM.call, src/ctr/layers/modules.py:57-72 (quirks kept).
  fm_model_onehot   — ctr.fm.model.FM.call one-hot form, src/ctr/fm/model.py:34-53.
"""
from __future__ import annotations

import numpy as np


def dot_interact(x, dtype=np.float64):
    """x (B, F1, D) -> (B, D + F1(F1-1)/2): [x[:,0,:], <x_i, x_j> for i>j row-major]."""
    x = np.asarray(x, dtype)
    B, F1, D = x.shape
    z = np.einsum("bid,bjd->bij", x, x)
    ii, jj = np.tril_indices(F1, -1)
    return np.concatenate([x[:, 0, :], z[:, ii, jj]], axis=1)


def dot_interact_bwd(x, gout, dtype=np.float64):
    """d(out)/d(x) contracted with gout."""
    x = np.asarray(x, dtype)
    g = np.asarray(gout, dtype)
    B, F1, D = x.shape
    ii, jj = np.tril_indices(F1, -1)
    S = np.zeros((B, F1, F1), dtype)
    S[:, ii, jj] = g[:, D:D + ii.size]
    S = S + np.transpose(S, (0, 2, 1))
    gx = np.einsum("bij,bjd->bid", S, x)
    gx[:, 0, :] += g[:, :D]
    return gx


def fm_layer(first_inputs, second_inputs, w, dtype=np.float64):
    """ctr.layers.modules.FM.call, src/ctr/layers/modules.py:63-72, line for line:
        first_order  = reduce_sum(matmul(first_inputs, w))                  # scalar, whole batch
        square_sum   = square(reduce_sum(second_inputs, axis=1, keepdims=True))
        sum_square   = reduce_sum(square(second_inputs), axis=1, keepdims=True)
        second_order = 0.5 * reduce_sum(square_sum - sum_square, axis=1, keepdims=False)
        output       = reshape(first_order + second_order, (-1, 1))
    second_inputs may be 2-D (B,M) — what DeepFM passes (deep_fm/model.py:58-59) — or 3-D."""
    a = np.asarray(first_inputs, dtype)
    x = np.asarray(second_inputs, dtype)
    first_order = np.sum(a @ np.asarray(w, dtype).reshape(-1, 1))
    square_sum = np.square(np.sum(x, axis=1, keepdims=True))
    sum_square = np.sum(np.square(x), axis=1, keepdims=True)
    second_order = 0.5 * np.sum(square_sum - sum_square, axis=1, keepdims=False)
    return np.reshape(first_order + second_order, (-1, 1))


def fm_layer_paper(first_inputs, second_inputs_bfd, w, dtype=np.float64):
    """Paper-correct variant: per-sample first order + per-dimension field
    cross summed over D -> (B,1).  Identity: 0.5((sum x)^2 - sum x^2) = sum_{i<j} x_i x_j."""
    a = np.asarray(first_inputs, dtype)
    x = np.asarray(second_inputs_bfd, dtype)
    first = a @ np.asarray(w, dtype).reshape(-1, 1)
    sec = 0.5 * (np.square(x.sum(1)) - np.square(x).sum(1)).sum(-1, keepdims=True)
    return first + sec


def fm_model_onehot(dense_inputs, sparse_inputs, feat_nums, w0, w, V, dtype=np.float64):
    """ctr.fm.model.FM.call, src/ctr/fm/model.py:34-53, one-hot form exactly as written:
        stack  = concat([dense_inputs] + [one_hot(sparse[:, i], depth=N_i)], -1)    (B, M)
        first  = w0 + stack @ w
        second = 0.5 * reduce_sum((stack @ V^T)^2 - (stack^2) @ (V^T)^2, 1, keepdims=True)
        out    = sigmoid(first + second)
    V is (k, M).  tf.one_hot of an out-of-range id is an all-zero row (A15)."""
    dense = np.asarray(dense_inputs, dtype)
    sp = np.asarray(sparse_inputs)
    B = dense.shape[0]
    hots = []
    for i, n in enumerate(feat_nums):
        h = np.zeros((B, n), dtype)
        ok = (sp[:, i] >= 0) & (sp[:, i] < n)
        h[np.flatnonzero(ok), sp[ok, i]] = 1
        hots.append(h)
    stack = np.concatenate([dense] + hots, axis=-1)
    Vt = np.asarray(V, dtype).T
    first = np.asarray(w0, dtype).reshape(1) + stack @ np.asarray(w, dtype).reshape(-1, 1)
    second = 0.5 * np.sum(np.square(stack @ Vt) - np.square(stack) @ np.square(Vt), axis=1,
                          keepdims=True)
    z = first + second
    return 1.0 / (1.0 + np.exp(-z))


def dlrm_concat(sparse_embed, dense_fea):
    """src/ctr/dlrm/model.py:48 exactly: x = tf.concat([sparse_embed, dense_fea], axis=-1)."""
    return np.concatenate([sparse_embed, dense_fea], axis=-1)

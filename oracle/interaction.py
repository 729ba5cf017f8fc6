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


def dlrm_concat(sparse_embed, dense_fea):
    """src/ctr/dlrm/model.py:48 exactly: x = tf.concat([sparse_embed, dense_fea], axis=-1)."""
    return np.concatenate([sparse_embed, dense_fea], axis=-1)

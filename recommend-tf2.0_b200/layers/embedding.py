"""Drop-ins for `tensorflow.keras.layers.Embedding` as the reference uses it, plus the fused
multi-table form of its per-field lookup + concat (src/ctr/dlrm/model.py:30-37,45-46)."""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from ..embedding import EmbeddingTables, SparseOptimizer
from .core import Layer, l2

_INIT = {"uniform": "random_uniform", "random_uniform": "random_uniform",
         "random_normal": "random_normal", "normal": "random_normal", "zeros": "zeros"}


def _as_int_ids(ids: torch.Tensor) -> torch.Tensor:
    """Keras casts non-integer ids to int32 (truncation; exact for float ids < 2**24, A1) —
    the reference declares many sparse inputs float32 (src/ctr/din/model.py:97-100)."""
    if ids.dtype in (torch.int32, torch.int64):
        return ids
    return ids.to(torch.int32)


class Embedding(Layer):
    """Embedding(input_dim, output_dim, embeddings_initializer, embeddings_regularizer,
    input_length): output shape = ids.shape + (output_dim,)."""

    def __init__(self, input_dim: int, output_dim: int, embeddings_initializer="uniform",
                 embeddings_regularizer: Optional[l2] = None, input_length=None,
                 sparse_optimizer: Optional[SparseOptimizer] = None, **kwargs):
        super().__init__(**kwargs)
        self.input_dim, self.output_dim, self.input_length = input_dim, output_dim, input_length
        self.embeddings_regularizer = embeddings_regularizer
        self.tables = EmbeddingTables([input_dim], [output_dim],
                                      _INIT.get(embeddings_initializer, embeddings_initializer),
                                      optimizer=sparse_optimizer)

    @property
    def embeddings(self):
        return self.tables.weights[0]

    def call(self, ids, **kwargs):
        ids = _as_int_ids(ids)
        flat = ids.reshape(-1, 1)
        out = self.tables.lookup(flat, (0,), "BF")
        return out.reshape(tuple(ids.shape) + (self.output_dim,))

    def regularization_loss(self):
        if self.embeddings_regularizer is None:
            return 0.0
        return self.embeddings_regularizer(self.embeddings)


class MultiTableEmbedding(Layer):
    """All sparse fields of a model in one launch: call(sparse_inputs (B, F)) -> (B, sumD), the
    same tensor as tf.concat([embed_i(sparse_inputs[:, i]) for i in range(F)], axis=-1).
    `sparse_feature_columns` is the reference's list of sparseFeature dicts
    ({'feat', 'feat_num', 'embed_dim'}, src/ctr/utils/data_process.py:13-21)."""

    def __init__(self, sparse_feature_columns: Sequence[dict],
                 embeddings_initializer="random_uniform", embed_reg: float = 0.0,
                 sparse_optimizer: Optional[SparseOptimizer] = None, seed=None, **kwargs):
        super().__init__(**kwargs)
        self.columns = list(sparse_feature_columns)
        self.embed_reg = embed_reg
        self.tables = EmbeddingTables([c["feat_num"] for c in self.columns],
                                      [c["embed_dim"] for c in self.columns],
                                      _INIT.get(embeddings_initializer, embeddings_initializer),
                                      optimizer=sparse_optimizer, seed=seed)

    def call(self, sparse_inputs, pool=None, layout="BF", **kwargs):
        return self.tables.lookup(_as_int_ids(sparse_inputs), None, layout, pool)

    def regularization_loss(self):
        if not self.embed_reg:
            return 0.0
        return sum(self.embed_reg * (w * w).sum() for w in self.tables.weights)

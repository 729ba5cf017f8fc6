"""Layer classes with the reference's names, constructor arguments and call signatures
(SURVEY.md §8b).  `layers.ctr` mirrors src/ctr/layers/modules.py, `layers.match` mirrors
src/match/layers/modules.py (both define a MultiHeadAttention and a DNN, as the reference
does, so they are reached through their submodule)."""
from . import ctr, match
from .core import (BatchNormalization, DNN, Dense, Dropout, Layer, binary_crossentropy,
                   get_activation, l2)
from .ctr import FM, AttentionLayer, Dice
from .embedding import Embedding, MultiTableEmbedding
from .match import (FFN, LayerNormalization, PoolingLayer, SampledSoftmaxLayer,
                    TransformerEncoder, sampled_softmax_loss, sampledsoftmaxloss)

__all__ = ["ctr", "match", "Layer", "Dense", "DNN", "BatchNormalization", "Dropout", "l2",
           "get_activation", "binary_crossentropy", "Embedding", "MultiTableEmbedding", "FM",
           "AttentionLayer", "Dice", "FFN", "LayerNormalization", "PoolingLayer",
           "SampledSoftmaxLayer", "TransformerEncoder", "sampled_softmax_loss",
           "sampledsoftmaxloss"]

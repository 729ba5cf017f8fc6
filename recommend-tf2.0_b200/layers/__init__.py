"""Layer classes with the reference's names, constructor arguments and call signatures
(SURVEY.md §8b).  ctr-side layers mirror src/ctr/layers/modules.py, match-side layers mirror
src/match/layers/modules.py."""
from .core import (BatchNormalization, DNN, Dense, Dropout, Layer, binary_crossentropy,
                   get_activation, l2)
from .embedding import Embedding, MultiTableEmbedding

__all__ = ["Layer", "Dense", "DNN", "BatchNormalization", "Dropout", "l2", "get_activation",
           "binary_crossentropy", "Embedding", "MultiTableEmbedding"]

"""Keras-protocol glue (Layer, Dense, DNN, ...) — defined in ../core.py, re-exported here so
that `layers.core` keeps working as an import path."""
from ..core import (BatchNormalization, DNN, Dense, Dropout, Layer, binary_crossentropy,  # noqa: F401
                    get_activation, l2, _default_device, _init_, _shape_of)

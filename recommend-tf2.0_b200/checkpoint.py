"""Weight import / export with the variable names Keras gives the reference's models
(`ModelCheckpoint(check_path, save_weights_only=True)` at
src/ctr/fm/train.py:52-55 and every other script).

`weights_dict(model)` walks the model in layer-creation order and names every variable the way
tf.keras names it: the layer name is the snake_case class name with Keras' per-class counter
(`embedding`, `embedding_1`, ..., `dense`, `dense_1`, `batch_normalization`, ...), the weight
name is the `add_weight` name (`embeddings`, `kernel`, `bias`, `gamma`, `beta`, `moving_mean`,
`moving_variance`, or the literal name a layer passes: `w0`, `w`, `V` in src/ctr/fm/model.py:22-32,
`w` in the FM layer, `alpha` in Dice) -> `"<layer>/<weight>:0"`, i.e. `[v.name for v in
model.weights]` of the reference model.  An `EmbeddingTables` set stands for the per-field
`Embedding` layers the reference creates in a loop (src/ctr/dlrm/model.py:30-37): one
`embedding[_i]/embeddings:0` per table, in field order.  `FMModel` exports the reference's
(w0, w, V) layout (one-hot column order), not its internal gather rows.

`save_weights` / `load_weights` store that mapping in a NumPy `.npz` container (TensorFlow's
own `.ckpt` / HDF5 writers are not available in this image; a maintainer with TF loads the same
arrays by name with `model.get_layer(name).set_weights(...)`).  [TF-knowledge]: the auto-naming
rule is Keras' `backend.unique_object_name`; nested name scopes (`build_graph()` wrappers) only
add a prefix, which `prefix=` reproduces.
"""
from __future__ import annotations

import re
from collections import OrderedDict, defaultdict
from typing import Dict

import numpy as np
import torch

from .core import BatchNormalization, Dense, Layer
from .embedding import EmbeddingTables

_WEIGHT_ALIASES = {"att_dense_kernel": ("dense", "kernel"), "att_dense_bias": ("dense", "bias")}


def _snake(name: str) -> str:
    s = re.sub(r"(.)([A-Z][a-z0-9]+)", r"\1_\2", name)
    return re.sub(r"([a-z0-9])([A-Z])", r"\1_\2", s).lower()


class _Namer:
    def __init__(self):
        self.count = defaultdict(int)

    def __call__(self, base: str) -> str:
        n = self.count[base]
        self.count[base] += 1
        return base if n == 0 else f"{base}_{n}"


def _entries(model: torch.nn.Module):
    """-> list of (key, getter, setter) in Keras variable order."""
    from .fm import FMModel
    namer = _Namer()
    out = []

    def add(key, tensor_fn, set_fn):
        out.append((key, tensor_fn, set_fn))

    def tensor_entry(layer_name, wname, t):
        def setter(a, t=t):
            a = torch.as_tensor(np.asarray(a), dtype=t.dtype).reshape(t.shape)
            with torch.no_grad():
                t.copy_(a.to(t.device))
        add(f"{layer_name}/{wname}:0", lambda t=t: t.detach().cpu().numpy().copy(), setter)

    def visit(m):
        if isinstance(m, FMModel):            # the reference layout: w0 (1,), w (M,1), V (k,M)
            holder = {}

            def get(i, m=m):
                return m.reference_weights()[i].detach().cpu().numpy().copy()

            def setter_for(i, m=m):
                def setter(a):
                    holder[i] = np.asarray(a)
                    if len(holder) == 3:
                        m.load_reference_weights(torch.as_tensor(holder[0]), torch.as_tensor(holder[1]),
                                                 torch.as_tensor(holder[2]))
                        holder.clear()
                return setter
            for i, nm in enumerate(("w0", "w", "V")):
                add(f"{nm}:0", lambda i=i: get(i), setter_for(i))
            return
        if isinstance(m, EmbeddingTables):
            for w in m.weights:
                tensor_entry(namer("embedding"), "embeddings", w)
            return
        if isinstance(m, Dense):
            if m._built:
                nm = namer("dense")
                tensor_entry(nm, "kernel", m.kernel)
                if m.bias is not None:
                    tensor_entry(nm, "bias", m.bias)
            return
        if isinstance(m, BatchNormalization):
            if m._built:
                nm = namer("batch_normalization")
                for wn in ("gamma", "beta"):
                    if getattr(m, wn, None) is not None:
                        tensor_entry(nm, wn, getattr(m, wn))
                tensor_entry(nm, "moving_mean", m.moving_mean)
                tensor_entry(nm, "moving_variance", m.moving_variance)
            return
        # generic layer: its own add_weight variables first (Keras: layer variables before the
        # sub-layers' only for weights created in build; names are the add_weight names), then
        # the children in creation order
        own = [(n, p) for n, p in m._parameters.items() if p is not None]
        if own and isinstance(m, Layer):
            lname = namer(_snake(type(m).__name__))
            for n, p in own:
                if n in _WEIGHT_ALIASES:      # a Dense the reference layer holds as a sub-layer
                    sub, wn = _WEIGHT_ALIASES[n]
                    tensor_entry(f"{lname}/{sub}", wn, p)
                else:
                    tensor_entry(lname, n, p)
        elif own:
            for n, p in own:
                tensor_entry(_snake(type(m).__name__), n, p)
        for child in m.children():
            visit(child)

    visit(model)
    return out


def weights_dict(model: torch.nn.Module, prefix: str = "") -> "OrderedDict[str, np.ndarray]":
    """{Keras variable name: ndarray} in `model.weights` order (lazily built layers must have
    been called once, exactly as Keras cannot save an unbuilt model)."""
    return OrderedDict((prefix + k, get()) for k, get, _ in _entries(model))


def set_weights_dict(model: torch.nn.Module, mapping: Dict[str, np.ndarray], prefix: str = "",
                     strict: bool = True) -> None:
    entries = _entries(model)
    keys = {prefix + k for k, _, _ in entries}
    if strict:
        missing, extra = keys - set(mapping), set(mapping) - keys
        if missing or extra:
            raise KeyError(f"weight names do not match: missing {sorted(missing)[:5]}, "
                           f"unexpected {sorted(extra)[:5]}")
    for k, _, setter in entries:
        if prefix + k in mapping:
            setter(mapping[prefix + k])


def save_weights(model: torch.nn.Module, path: str, prefix: str = "") -> None:
    """model.save_weights(path) of the reference scripts: every variable under its Keras name."""
    d = weights_dict(model, prefix)
    np.savez(path, __order__=np.array(list(d.keys())), **{k.replace("/", "|"): v for k, v in d.items()})


def load_weights(model: torch.nn.Module, path: str, prefix: str = "", strict: bool = True) -> None:
    """model.load_weights(path): by name, shapes checked, in place (tables stay in HBM)."""
    if not path.endswith(".npz"):
        path = path + ".npz"
    with np.load(path, allow_pickle=False) as z:
        mapping = {k.replace("|", "/"): z[k] for k in z.files if k != "__order__"}
    set_weights_dict(model, mapping, prefix, strict)

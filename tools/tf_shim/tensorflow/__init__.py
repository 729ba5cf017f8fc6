"""A numpy stand-in for the handful of TensorFlow / Keras symbols the reference's hot-path files
use — TEST TOOLING ONLY (tools/make_golden.py).

TensorFlow is not installable in the build image, so the reference's own layer source
(/root/reference/src/{ctr,match}/layers/modules.py, ctr/fm/model.py, match/sasrec/model.py)
is executed over this shim to produce golden vectors: the control flow, reshapes, tilings,
mask handling and op order are then the reference's own code; only the tf.* op semantics are
restated here.  Everything is eager, float64 by default for accuracy
(pass float32 arrays to keep float32), tensors are plain numpy arrays.
"""
import sys
import types

import numpy as np

_rng = np.random.default_rng(20211021)
Tensor = np.ndarray
float32, float64, int32, int64 = np.float32, np.float64, np.int32, np.int64
FDT = np.float64  # working dtype of created weights


def _seed(s):
    global _rng
    _rng = np.random.default_rng(s)


def _a(x):
    return x if isinstance(x, np.ndarray) else np.asarray(x)


# ---------------------------------------------------------------- ops
def matmul(a, b, transpose_a=False, transpose_b=False, name=None):
    a, b = _a(a), _a(b)
    if transpose_a:
        a = np.swapaxes(a, -1, -2)
    if transpose_b:
        b = np.swapaxes(b, -1, -2)
    return a @ b


def reduce_sum(x, axis=None, keepdims=False, name=None):
    return np.sum(_a(x), axis=axis, keepdims=keepdims)


def reduce_mean(x, axis=None, keepdims=False, name=None):
    return np.mean(_a(x), axis=axis, keepdims=keepdims)


def reduce_max(x, axis=None, keepdims=False, name=None):
    return np.max(_a(x), axis=axis, keepdims=keepdims)


def square(x):
    return np.square(_a(x))


def pow(x, y):  # noqa: A001
    return np.power(_a(x), y)


def sqrt(x):
    return np.sqrt(_a(x))


def reshape(x, shape):
    return np.reshape(_a(x), tuple(int(s) for s in shape))


def transpose(x, perm=None):
    return np.transpose(_a(x), perm)


def concat(values, axis):
    return np.concatenate([_a(v) for v in values], axis=axis)


def expand_dims(x, axis):
    return np.expand_dims(_a(x), axis)


def squeeze(x, axis=None):
    return np.squeeze(_a(x), axis=axis)


def tile(x, multiples):
    return np.tile(_a(x), tuple(multiples))


def ones_like(x):
    return np.ones_like(_a(x))


def zeros_like(x):
    return np.zeros_like(_a(x))


def where(cond, a, b):
    return np.where(cond, a, b)      # SelectV2: broadcasting (App. A7)


def equal(a, b):
    return np.equal(a, b)


def not_equal(a, b):
    return np.not_equal(a, b)


def cast(x, dtype):
    x = _a(x)
    if dtype in (np.float32, np.float64) and x.dtype == np.bool_:
        return x.astype(FDT)
    if dtype in (np.float32, np.float64):
        return x.astype(FDT if x.dtype != np.float32 else np.float32)
    return x.astype(dtype)


def multiply(a, b):
    return _a(a) * _a(b)


def add(a, b):
    return _a(a) + _a(b)


def maximum(a, b):
    return np.maximum(a, b)


def minimum(a, b):
    return np.minimum(a, b)


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-_a(x)))


def one_hot(indices, depth):
    idx = _a(indices).astype(np.int64)
    out = np.zeros(idx.shape + (depth,), FDT)
    ok = (idx >= 0) & (idx < depth)                                   # App. A15
    np.put_along_axis(out, np.where(ok, idx, 0)[..., None], ok[..., None].astype(FDT), axis=-1)
    return out


def shape(x):
    return np.asarray(_a(x).shape)


def tensordot(a, b, axes):
    return np.tensordot(_a(a), _a(b), axes=axes)


class random_normal_initializer:
    def __init__(self, mean=0.0, stddev=0.05, seed=None):
        self.mean, self.stddev = mean, stddev

    def __call__(self, shape, dtype=None):
        return _rng.normal(self.mean, self.stddev, size=tuple(shape)).astype(FDT)


class zeros_initializer:
    def __call__(self, shape, dtype=None):
        return np.zeros(tuple(shape), FDT)


def _softmax(logits, axis=-1, name=None):
    x = _a(logits)
    m = np.max(x, axis=axis, keepdims=True)
    e = np.exp(x - m)
    return e / np.sum(e, axis=axis, keepdims=True)


LAST_SAMPLED_VALUES = None


def _log_uniform_candidate_sampler(num_sampled, range_max):
    """unique=True log-uniform sampler (App. A14) on the shim's own RNG."""
    seen, out, tries = set(), [], 0
    lr = np.log(range_max + 1.0)
    while len(out) < num_sampled:
        tries += 1
        c = int(np.exp(_rng.random() * lr)) - 1
        c = min(max(c, 0), range_max - 1)
        if c not in seen:
            seen.add(c)
            out.append(c)
    return np.asarray(out, np.int64), tries


def _expected(c, range_max, tries):
    p = (np.log(c + 2.0) - np.log(c + 1.0)) / np.log(range_max + 1.0)
    return -np.expm1(tries * np.log1p(-p))


def _sampled_softmax_loss(weights, biases, labels, inputs, num_sampled, num_classes, num_true=1,
                          sampled_values=None, remove_accidental_hits=True, **kw):
    """tf.nn.sampled_softmax_loss (App. A13)."""
    global LAST_SAMPLED_VALUES
    W, b, x = _a(weights), _a(biases), _a(inputs)
    lab = _a(labels).astype(np.int64).reshape(-1)
    if sampled_values is None:
        s, tries = _log_uniform_candidate_sampler(num_sampled, num_classes)
        sampled_values = (s, _expected(lab.astype(np.float64), num_classes, tries),
                          _expected(s.astype(np.float64), num_classes, tries))
    LAST_SAMPLED_VALUES = sampled_values
    s, te, se = sampled_values
    true_logits = np.sum(x * W[lab], axis=1) + b[lab]
    sampled_logits = x @ W[s].T + b[s]
    if remove_accidental_hits:
        sampled_logits = sampled_logits + np.where(lab[:, None] == s[None, :],
                                                   -np.finfo(np.float32).max, 0.0)
    true_logits = true_logits - np.log(te)
    sampled_logits = sampled_logits - np.log(se)[None, :]
    logits = np.concatenate([true_logits[:, None], sampled_logits], 1)
    m = logits.max(1, keepdims=True)
    return (m[:, 0] + np.log(np.exp(logits - m).sum(1))) - logits[:, 0]


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


nn = _mod("tensorflow.nn", softmax=_softmax, relu=lambda x: np.maximum(_a(x), 0), sigmoid=sigmoid,
          sampled_softmax_loss=_sampled_softmax_loss)
math = _mod("tensorflow.math", log=lambda x: np.log(_a(x)))
linalg = _mod("tensorflow.linalg",
              matmul=lambda a, b, transpose_a=False, transpose_b=False: matmul(a, b, transpose_a, transpose_b))


# ---------------------------------------------------------------- keras
class _l2:
    def __init__(self, l2=0.01):
        self.l2 = l2

    def __call__(self, w):
        return self.l2 * np.sum(np.square(w))


def _init(initializer, shape):
    shape = tuple(int(s) for s in shape)
    if initializer is None or initializer == "glorot_uniform":
        if len(shape) >= 2:
            fi, fo = shape[-2], shape[-1]
            if len(shape) > 2:                      # conv kernels: receptive field * channels
                rf = int(np.prod(shape[:-2]))
                fi, fo = fi * rf, fo * rf
        else:
            fi = fo = max(1, int(np.prod(shape)) if shape else 1)
        lim = np.sqrt(6.0 / (fi + fo))
        return _rng.uniform(-lim, lim, size=shape).astype(FDT)
    if isinstance(initializer, type):
        initializer = initializer()
    if callable(initializer):
        return _a(initializer(shape)).astype(FDT)
    if initializer in ("random_normal", "normal"):
        return _rng.normal(0, 0.05, size=shape).astype(FDT)
    if initializer in ("random_uniform", "uniform"):
        return _rng.uniform(-0.05, 0.05, size=shape).astype(FDT)
    if initializer == "zeros":
        return np.zeros(shape, FDT)
    if initializer == "ones":
        return np.ones(shape, FDT)
    raise ValueError(f"Unknown initializer: {initializer}")


def _activation(act):
    if act is None or act == "linear":
        return lambda z: z
    if callable(act):
        return act
    table = {"relu": lambda z: np.maximum(z, 0), "sigmoid": sigmoid, "tanh": np.tanh,
             "softmax": _softmax}
    if act not in table:
        raise ValueError(f"Unknown activation function: {act}")   # e.g. 'prelu'
    return table[act]


def _shape_of(x):
    if isinstance(x, (list, tuple)):
        return [_shape_of(t) for t in x]
    if isinstance(x, dict):
        return {k: _shape_of(v) for k, v in x.items()}
    return tuple(_a(x).shape)


CREATED_LAYERS = []   # every layer instance in creation order (the ctr MultiHeadAttention
                      # builds its Dense layers inside call(): this is how their weights are found)


class Layer:
    def __init__(self, name=None, **kwargs):
        self._built = False
        self.weights_dict = {}
        self._losses = []
        CREATED_LAYERS.append(self)

    def add_weight(self, name=None, shape=(), initializer=None, regularizer=None, trainable=True,
                   dtype=None, **kw):
        w = _init(initializer, shape)
        self.weights_dict[name or f"w{len(self.weights_dict)}"] = w
        return w

    def add_loss(self, loss):
        self._losses.append(loss)

    def build(self, input_shape):
        self._built = True

    def __call__(self, inputs, *args, **kwargs):
        if not self._built:
            self.build(_shape_of(inputs))
            self._built = True
        return self.call(inputs, *args, **kwargs)


class Model(Layer):
    def __init__(self, *args, **kwargs):
        Layer.__init__(self)


class Dense(Layer):
    def __init__(self, units, activation=None, use_bias=True, kernel_regularizer=None, **kw):
        super().__init__()
        self.units, self.use_bias = units, use_bias
        self.activation = _activation(activation)

    def build(self, input_shape):
        self.kernel = self.add_weight("kernel", (input_shape[-1], self.units), "glorot_uniform")
        self.bias = self.add_weight("bias", (self.units,), "zeros") if self.use_bias else None

    def call(self, x, **kw):
        y = _a(x) @ self.kernel
        if self.bias is not None:
            y = y + self.bias
        return self.activation(y)


class Conv1D(Dense):
    """kernel_size == 1 only: a per-position Dense (App. A4)."""

    def __init__(self, filters, kernel_size, activation=None, use_bias=True, **kw):
        assert kernel_size == 1
        super().__init__(filters, activation, use_bias)


class Dropout(Layer):
    def __init__(self, rate=0.0, **kw):
        super().__init__()

    def call(self, x, **kw):
        return x


class ReLU(Layer):
    def call(self, x, **kw):
        return np.maximum(_a(x), 0)


class LayerNormalization(Layer):
    def __init__(self, epsilon=1e-3, **kw):
        super().__init__()
        self.epsilon = epsilon

    def build(self, input_shape):
        self.gamma = self.add_weight("gamma", (input_shape[-1],), "ones")
        self.beta = self.add_weight("beta", (input_shape[-1],), "zeros")

    def call(self, x, **kw):
        x = _a(x)
        mu = x.mean(-1, keepdims=True)
        var = ((x - mu) ** 2).mean(-1, keepdims=True)
        return self.gamma * (x - mu) / np.sqrt(var + self.epsilon) + self.beta


class BatchNormalization(Layer):
    """Inference-mode call (moving mean 0 / variance 1 at init), as an eager Keras layer does
    outside fit; `training=True` uses batch statistics (App. A9)."""

    def __init__(self, center=True, scale=True, epsilon=1e-3, **kw):
        super().__init__()
        self.center, self.scale, self.epsilon = center, scale, epsilon

    def call(self, x, training=False, **kw):
        x = _a(x)
        if training:
            mu, var = x.mean(0), x.var(0)
        else:
            mu, var = 0.0, 1.0
        return (x - mu) / np.sqrt(var + self.epsilon)


class Embedding(Layer):
    def __init__(self, input_dim, output_dim, embeddings_initializer="uniform",
                 embeddings_regularizer=None, input_length=None, **kw):
        super().__init__()
        self.embeddings = self.add_weight("embeddings", (input_dim, output_dim), embeddings_initializer)

    def call(self, ids, **kw):
        ids = _a(ids)
        if ids.dtype.kind == "f":
            ids = ids.astype(np.int32)                                 # App. A1
        if ids.size and (ids.min() < 0 or ids.max() >= self.embeddings.shape[0]):
            raise IndexError("indices out of range")                  # App. A2 (CPU)
        return self.embeddings[ids]


class Concatenate(Layer):
    def __init__(self, axis=-1, **kw):
        super().__init__()
        self.axis = axis

    def call(self, inputs, **kw):
        return np.concatenate([_a(t) for t in inputs], axis=self.axis)


class Lambda(Layer):
    def __init__(self, fn, **kw):
        super().__init__()
        self.fn = fn

    def call(self, x, **kw):
        return self.fn(x)


class PReLU(Layer):
    def build(self, input_shape):
        self.alpha = self.add_weight("alpha", input_shape[1:], "zeros")

    def call(self, x, **kw):
        x = _a(x)
        return np.maximum(x, 0) + self.alpha * np.minimum(x, 0)


def Input(shape=None, dtype=None, **kw):
    raise RuntimeError("symbolic Input is not part of the shim: call the layers eagerly")


class _NotUsed(Layer):
    pass


_layers = _mod("tensorflow.keras.layers", Layer=Layer, Dense=Dense, Dropout=Dropout, ReLU=ReLU,
               BatchNormalization=BatchNormalization, LayerNormalization=LayerNormalization,
               Conv1D=Conv1D, Embedding=Embedding, Concatenate=Concatenate, Input=Input, PReLU=PReLU,
               Lambda=Lambda, GlobalAveragePooling1D=_NotUsed, GlobalMaxPooling1D=_NotUsed,
               Activation=_NotUsed, GlobalAveragePooling2D=_NotUsed, Reshape=_NotUsed)
_regs = _mod("tensorflow.keras.regularizers", l2=_l2)
_inits = _mod("tensorflow.keras.initializers", RandomNormal=random_normal_initializer,
              Zeros=zeros_initializer)
_models = _mod("tensorflow.keras.models", Model=Model)
_backend = _mod("tensorflow.python.keras.backend", mean=lambda x: np.mean(_a(x)))
keras = _mod("tensorflow.keras", layers=_layers, regularizers=_regs, initializers=_inits,
             models=_models, Model=Model, Input=Input, backend=_backend)
_pk_layers = _mod("tensorflow.python.keras.layers", Lambda=Lambda)
_pk = _mod("tensorflow.python.keras", backend=_backend, layers=_pk_layers)
python = _mod("tensorflow.python", keras=_pk)

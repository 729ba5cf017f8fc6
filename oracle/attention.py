"""Oracle (test infrastructure): attention-family layers restated in numpy, quirks kept.

  din_attention_layer  — ctr AttentionLayer.call, src/ctr/layers/modules.py:149-175
  ctr_mha              — ctr MultiHeadAttention.call, src/ctr/layers/modules.py:285-325
  match_sdpa / match_mha / transformer_encoder — src/match/layers/modules.py:76-96,115-131,173-185
  sasrec_scores_loss   — src/match/sasrec/model.py:88-96
  sampled_softmax_loss — tf.nn.sampled_softmax_loss as called at src/match/layers/modules.py:54-60
                         (TF semantics: SURVEY.md App. A13/A14)
"""
from __future__ import annotations

import numpy as np

PAD = float(-2 ** 32 + 1)   # the source's padding constant; rounds to -4294967296.0 in fp32


def softmax(x, axis=-1):
    """tf.nn.softmax: exp(x - max) / sum (A6)."""
    m = np.max(x, axis=axis, keepdims=True)
    e = np.exp(x - m)
    return e / np.sum(e, axis=axis, keepdims=True)


def _act(name):
    return {None: lambda z: z, "linear": lambda z: z, "relu": lambda z: np.maximum(z, 0),
            "sigmoid": lambda z: 1.0 / (1.0 + np.exp(-z)), "tanh": np.tanh}[name]


def din_attention_layer(q, k, v, mask, W, bias, activation="sigmoid", dtype=np.float64):
    """AttentionLayer.call, line for line (modules.py:149-175); W (4d,1), bias (1,) are the
    Dense(hidden_unit=1) parameters.  `mask` None (= "not a tf.Tensor") pads every score."""
    q, k, v = (np.asarray(t, dtype) for t in (q, k, v))
    B, L, d = k.shape
    qt = np.tile(q, (1, L)).reshape(-1, L, d)                        # :150-151
    info = np.concatenate([qt, k, qt - k, qt * k], axis=-1)          # :154
    outputs = _act(activation)(info @ np.asarray(W, dtype).reshape(4 * d, 1) +
                               np.asarray(bias, dtype).reshape(1))   # :157
    outputs = outputs.reshape(-1, L)                                 # :159
    paddings = np.ones_like(outputs) * dtype(np.float32(PAD))        # :161
    if mask is not None:
        outputs = np.where(np.asarray(mask).reshape(B, L) == 0, paddings, outputs)   # :163
    else:
        outputs = paddings                                           # :165
    outputs = softmax(outputs)[:, None, :]                           # :169-170
    return (outputs @ v)[:, 0, :]                                    # :172-173


def ctr_mha(xq, xk, xv, Wq, Wk, Wv, head_num, head_size, activation="relu", W0=None,
            scale="reference", dtype=np.float64):
    """ctr MultiHeadAttention.call (modules.py:285-325): Dense without bias + activation on
    q, k AND v; product / head_size**-0.5 (:235-237, i.e. times sqrt(head_size)); softmax over
    the last axis, no mask; use_res <=> W0 given: relu(out + act(ori_v @ W0)) (:316-323)."""
    act = _act(activation)
    q = act(np.asarray(xq, dtype) @ np.asarray(Wq, dtype))
    k = act(np.asarray(xk, dtype) @ np.asarray(Wk, dtype))
    v = act(np.asarray(xv, dtype) @ np.asarray(Wv, dtype))
    B = q.shape[0]

    def split(t):
        return np.transpose(t.reshape(B, t.shape[1], head_num, head_size), (0, 2, 1, 3))

    qh, kh, vh = split(q), split(k), split(v)
    div = head_size ** -0.5 if scale == "reference" else head_size ** 0.5
    product = (qh @ np.transpose(kh, (0, 1, 3, 2))) / div
    out = softmax(product) @ vh
    out = np.transpose(out, (0, 2, 1, 3)).reshape(B, out.shape[2], head_num * head_size)
    if W0 is not None:
        return np.maximum(out + act(np.asarray(xv, dtype) @ np.asarray(W0, dtype)), 0)
    return out


def match_sdpa(q, k, v, mask):
    """match scaled_dot_product_attention (modules.py:76-96).  mask broadcastable (...,L,1):
    `where(mask == 0, paddings, logits)` blanks whole query ROWS (A7)."""
    mat_qk = q @ np.swapaxes(k, -1, -2)
    dk = np.float32(k.shape[-1]).astype(q.dtype)
    scaled = mat_qk / np.sqrt(dk)
    paddings = np.ones_like(scaled) * q.dtype.type(np.float32(PAD))
    outputs = np.where(np.equal(mask, np.zeros_like(mask)), paddings, scaled)
    return softmax(outputs) @ v


def match_mha(q, k, v, mask, Wq, bq, Wk, bk, Wv, bv, num_heads, dtype=np.float64):
    """match MultiHeadAttention.call (modules.py:115-131); mask (B,L,1)."""
    q = np.asarray(q, dtype) @ np.asarray(Wq, dtype) + np.asarray(bq, dtype)
    k = np.asarray(k, dtype) @ np.asarray(Wk, dtype) + np.asarray(bk, dtype)
    v = np.asarray(v, dtype) @ np.asarray(Wv, dtype) + np.asarray(bv, dtype)
    B, L, dm = q.shape
    depth = dm // num_heads

    def split(t):
        return np.transpose(t.reshape(B, L, num_heads, depth), (0, 2, 1, 3))

    m = np.tile(np.asarray(mask, dtype)[:, None, :, :], (1, num_heads, 1, 1))
    att = match_sdpa(split(q), split(k), split(v), m)
    return np.transpose(att, (0, 2, 1, 3)).reshape(B, L, dm)


def layer_norm(x, gamma, beta, eps):
    mu = x.mean(-1, keepdims=True)
    var = ((x - mu) ** 2).mean(-1, keepdims=True)
    return gamma * (x - mu) / np.sqrt(var + eps) + beta


def transformer_encoder(x, mask, p, num_heads=1, eps=1e-6, dtype=np.float64):
    """TransformerEncoder.call (modules.py:173-185), dropout 0.  p: dict of parameters."""
    x = np.asarray(x, dtype)
    g = lambda n: np.asarray(p[n], dtype)
    att = match_mha(x, x, x, mask, g("wq"), g("bq"), g("wk"), g("bk"), g("wv"), g("bv"), num_heads, dtype)
    out1 = layer_norm(x + att, g("ln1_g"), g("ln1_b"), eps)
    ffn = np.maximum(out1 @ g("w1") + g("b1"), 0) @ g("w2") + g("b2")
    return layer_norm(out1 + ffn, g("ln2_g"), g("ln2_b"), eps)


def sasrec_scores_loss(att_outputs, pos_embed, neg_embed):
    """SASRec.call tail (sasrec/model.py:88-96): last position, dot scores, log loss."""
    seq_info = att_outputs[:, -1:, :]                                   # (B,1,D)
    pos = np.sum(seq_info * pos_embed, axis=-1)                         # (B,1)
    neg = np.sum(seq_info * neg_embed, axis=-1)                         # (B,neg_len)
    sig = lambda z: 1.0 / (1.0 + np.exp(-z))
    loss = np.mean(-np.log(sig(pos)) - np.log(1 - sig(neg))) / 2
    return np.concatenate([pos, neg], axis=-1), loss


def log_uniform_prob(c, range_max):
    c = np.asarray(c, np.float64)
    return (np.log(c + 2.0) - np.log(c + 1.0)) / np.log(range_max + 1.0)


def log_uniform_expected(c, range_max, num_tries):
    """unique=True expected counts: -expm1(num_tries * log1p(-P(c)))  (A14)."""
    return -np.expm1(num_tries * np.log1p(-log_uniform_prob(c, range_max)))


def sampled_softmax_loss(weights, biases, labels, inputs, sampled, true_exp, samp_exp,
                         remove_accidental_hits=True, dtype=np.float64):
    """tf.nn.sampled_softmax_loss with injected sampled_values (A13), num_true = 1."""
    W = np.asarray(weights, dtype)
    x = np.asarray(inputs, dtype)
    b = np.zeros(W.shape[0], dtype) if biases is None else np.asarray(biases, dtype)
    labels = np.asarray(labels).reshape(-1)
    sampled = np.asarray(sampled).reshape(-1)
    true_logits = np.sum(x * W[labels], axis=1) + b[labels]
    sampled_logits = x @ W[sampled].T + b[sampled]
    if remove_accidental_hits:
        hits = labels[:, None] == sampled[None, :]
        sampled_logits = sampled_logits + np.where(hits, -np.finfo(np.float32).max, 0.0)
    true_logits = true_logits - np.log(np.asarray(true_exp, dtype).reshape(-1))
    sampled_logits = sampled_logits - np.log(np.asarray(samp_exp, dtype).reshape(-1))[None, :]
    logits = np.concatenate([true_logits[:, None], sampled_logits], axis=1)
    m = logits.max(1, keepdims=True)
    lse = m[:, 0] + np.log(np.exp(logits - m).sum(1))
    return lse - logits[:, 0]

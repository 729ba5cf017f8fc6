"""match-side layers, same names / constructor arguments / call structure as
src/match/layers/modules.py: DNN (:8-26), SampledSoftmaxLayer (:28-61), MultiHeadAttention
(:98-131), FFN (:134-149), TransformerEncoder (:152-185), PoolingLayer (:187-211)."""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import torch
import torch.nn.functional as F

from .. import _lib as L
from ..attention import attention
from ..embedding import embed_bwd
from .core import DNN as _CoreDNN
from .core import Dense, Dropout, Layer


class DNN(_CoreDNN):
    """match DNN: Dense stack + dropout, no BatchNormalization (src/match/layers/modules.py:8-26)."""

    def __init__(self, hidden_units, activation="relu", dnn_dropout=0.0, **kwargs):
        super().__init__(hidden_units, activation, dnn_dropout, input_bn=False, **kwargs)


class _LayerNormFn(torch.autograd.Function):
    """rtf_layernorm_fwd / rtf_layernorm_bwd: a warp per row, dgamma / dbeta as deterministic
    two-stage column sums (the framework kernel runs one CTA per 64-float row: 0.3 TB/s)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        from ..core import _scratch
        L.require_cuda(x, "LayerNormalization(x)")
        Cc = x.shape[-1]
        x2 = x.reshape(-1, Cc)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        rows = x2.shape[0]
        y = torch.empty_like(x2)
        stats = torch.empty((2, rows), dtype=torch.float32, device=x.device)
        L.check(L.lib().rtf_layernorm_fwd(x2.data_ptr(), rows, Cc, gamma.data_ptr(), beta.data_ptr(), eps,
                                          y.data_ptr(), stats[0].data_ptr(), stats[1].data_ptr(),
                                          L.current_stream_ptr()), "rtf_layernorm_fwd")
        ctx.save_for_backward(x2, stats, gamma)
        ctx.shape = x.shape
        ctx._scratch = _scratch
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, stats, gamma = ctx.saved_tensors
        rows, Cc = x2.shape
        dy2 = dy.reshape(rows, Cc)
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        nb = C.c_size_t(0)
        L.check(L.lib().rtf_layernorm_workspace(rows, Cc, C.byref(nb)), "rtf_layernorm_workspace")
        ws = ctx._scratch(nb.value, dy.device)
        dx = torch.empty_like(x2) if ctx.needs_input_grad[0] else None
        dg = torch.empty(Cc, dtype=torch.float32, device=dy.device)
        db = torch.empty(Cc, dtype=torch.float32, device=dy.device)
        L.check(L.lib().rtf_layernorm_bwd(dy2.data_ptr(), x2.data_ptr(), stats[0].data_ptr(),
                                          stats[1].data_ptr(), gamma.data_ptr(), rows, Cc,
                                          None if dx is None else dx.data_ptr(), dg.data_ptr(),
                                          db.data_ptr(), ws.data_ptr(), ws.numel(),
                                          L.current_stream_ptr()), "rtf_layernorm_bwd")
        return (None if dx is None else dx.view(ctx.shape)), dg, db, None


class LayerNormalization(Layer):
    """Keras LayerNormalization(epsilon): last axis, biased variance, gamma=1 / beta=0 (A8)."""

    def __init__(self, epsilon: float = 1e-3, **kwargs):
        super().__init__(**kwargs)
        self.epsilon = epsilon

    def build(self, input_shape):
        c = input_shape[-1]
        self.gamma = self.add_weight("gamma", (c,), "ones")
        self.beta = self.add_weight("beta", (c,), "zeros")

    def call(self, x, **kwargs):
        if x.is_cuda and x.dtype == torch.float32 and x.shape[-1] <= 256 and x.numel() > 0:
            return _LayerNormFn.apply(x, self.gamma, self.beta, self.epsilon)
        # wider rows than the warp-per-row kernel takes (or CPU tensors in the layer-building tests)
        return F.layer_norm(x, (x.shape[-1],), self.gamma, self.beta, self.epsilon)


class MultiHeadAttention(Layer):
    """MultiHeadAttention(d_model, num_heads); call(q, k, v, mask) with mask (B, L, 1).

    wq/wk/wv are Dense(d_model) WITH bias and no activation (:110-112); logits are divided by
    sqrt(d_model/num_heads) (:85-88); the (B,L,1) mask is tiled over heads and broadcasts over
    the KEY axis (:126,90-91), so it blanks whole query rows (uniform attention on padded
    queries), keys are not masked and there is no causal mask; no output projection (:130)."""

    def __init__(self, d_model, num_heads):
        super().__init__()
        self.d_model, self.num_heads = d_model, num_heads
        self.wq = Dense(d_model, activation=None)
        self.wk = Dense(d_model, activation=None)
        self.wv = Dense(d_model, activation=None)

    def forward(self, q, k, v, mask):          # the reference calls it with four positionals
        q, k, v = self.wq(q), self.wk(k), self.wv(v)
        depth = self.d_model // self.num_heads
        return attention(q, k, v, self.num_heads, 1.0 / math.sqrt(depth), row_mask=mask)


class FFN(Layer):
    """FFN(hidden_unit, d_model): Conv1D(k=1, relu) -> Conv1D(k=1) == per-position Dense (A4)."""

    def __init__(self, hidden_unit, d_model):
        super().__init__()
        self.conv1 = Dense(hidden_unit, activation="relu", use_bias=True)
        self.conv2 = Dense(d_model, activation=None, use_bias=True)

    def call(self, inputs, **kwargs):
        return self.conv2(self.conv1(inputs))


class TransformerEncoder(Layer):
    """TransformerEncoder(d_model, num_heads=1, ffn_hidden_unit=128, dropout=0.,
    layer_norm_eps=1e-6); call([x, mask]) — src/match/layers/modules.py:173-185."""

    def __init__(self, d_model, num_heads=1, ffn_hidden_unit=128, dropout=0.0, layer_norm_eps=1e-6):
        super().__init__()
        self.mha = MultiHeadAttention(d_model, num_heads)
        self.ffn = FFN(ffn_hidden_unit, d_model)
        self.layernorm1 = LayerNormalization(epsilon=layer_norm_eps)
        self.layernorm2 = LayerNormalization(epsilon=layer_norm_eps)
        self.dropout1 = Dropout(dropout)
        self.dropout2 = Dropout(dropout)

    def call(self, inputs, **kwargs):
        x, mask = inputs
        att_out = self.dropout1(self.mha(x, x, x, mask))
        out1 = self.layernorm1(x + att_out)
        ffn_out = self.dropout2(self.ffn(out1))
        return self.layernorm2(out1 + ffn_out)


class PoolingLayer(Layer):
    """PoolingLayer(mode in {'mean','max','sum'}); call(list of same-shape tensors): one tensor
    is returned unchanged, several are stacked on a NEW last axis and reduced over it — i.e. an
    element-wise mean/sum/max ACROSS tensors; `mask` is ignored (:199-209)."""

    def __init__(self, mode="mean", **kwargs):
        if mode not in ["mean", "max", "sum"]:
            raise ValueError("mode must be max or mean")
        self.mode = mode
        super().__init__(**kwargs)

    def call(self, inputs, mask=None, **kwargs):
        if not isinstance(inputs, (list, tuple)):
            inputs = [inputs]
        if len(inputs) == 1:
            return inputs[0]
        a = torch.stack(list(inputs), dim=-1)
        if self.mode == "mean":
            return a.mean(-1)
        if self.mode == "sum":
            return a.sum(-1)
        return a.max(-1).values


# ------------------------------------------------------------------------------ sampled softmax
def log_uniform_candidate_sampler(num_sampled: int, range_max: int, seed, device="cuda"):
    """Device log-uniform (Zipfian) sampler with unique=True (App. A14).  Returns
    (sampled int64 (S), num_tries int32 (1)); expected counts via `log_uniform_expected`.
    `seed`: a Python int, or a 1-element int64 CUDA tensor read at execution time (a step replayed
    from a CUDA graph bumps it between replays)."""
    nb = C.c_size_t(0)
    L.check(L.lib().rtf_log_uniform_workspace(num_sampled, C.byref(nb)), "rtf_log_uniform_workspace")
    ws = torch.empty(nb.value, dtype=torch.uint8, device=device)
    sampled = torch.empty(num_sampled, dtype=torch.int64, device=device)
    tries = torch.zeros(1, dtype=torch.int32, device=device)
    if isinstance(seed, torch.Tensor):
        L.require_cuda(seed, "log_uniform_candidate_sampler(seed)")
        if seed.dtype != torch.int64 or seed.numel() != 1:
            raise TypeError("device seed must be one int64")
        L.check(L.lib().rtf_log_uniform_sample_dseed(seed.data_ptr(), num_sampled, range_max,
                                                     sampled.data_ptr(), tries.data_ptr(),
                                                     ws.data_ptr(), L.current_stream_ptr()),
                "rtf_log_uniform_sample_dseed")
        return sampled, tries
    L.check(L.lib().rtf_log_uniform_sample(seed & 0xFFFFFFFFFFFFFFFF, num_sampled, range_max,
                                           sampled.data_ptr(), tries.data_ptr(), ws.data_ptr(),
                                           L.current_stream_ptr()), "rtf_log_uniform_sample")
    return sampled, tries


def log_uniform_expected(ids: torch.Tensor, range_max: int, num_tries: torch.Tensor) -> torch.Tensor:
    ids = ids.reshape(-1).to(torch.int64).contiguous()
    out = torch.empty(ids.numel(), dtype=torch.float32, device=ids.device)
    L.check(L.lib().rtf_log_uniform_expected(ids.data_ptr(), ids.numel(), range_max,
                                             num_tries.data_ptr(), out.data_ptr(),
                                             L.current_stream_ptr()), "rtf_log_uniform_expected")
    return out


def _ssm_ws(S, D, device):
    nb = C.c_size_t(0)
    L.check(L.lib().rtf_sampled_softmax_workspace(S, D, C.byref(nb)), "rtf_sampled_softmax_workspace")
    return torch.empty(nb.value, dtype=torch.uint8, device=device)


class _SampledSoftmaxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weights, biases, labels, inputs, sampled, true_exp, samp_exp, remove_hits, err,
                table=None):
        L.require_cuda(inputs, "sampled_softmax_loss(inputs)")
        weights, inputs = weights.contiguous(), inputs.contiguous()
        B, D = inputs.shape
        N, S = weights.shape[0], sampled.numel()
        loss = torch.empty(B, dtype=torch.float32, device=inputs.device)
        lse = torch.empty_like(loss)
        ws = _ssm_ws(S, D, inputs.device)
        rc = L.lib().rtf_sampled_softmax_fwd(
            inputs.data_ptr(), inputs.stride(0), weights.data_ptr(),
            None if biases is None else biases.data_ptr(), labels.data_ptr(), sampled.data_ptr(),
            true_exp.data_ptr(), samp_exp.data_ptr(), B, N, S, D, int(remove_hits), loss.data_ptr(),
            lse.data_ptr(), ws.data_ptr(), None if err is None else err.data_ptr(),
            L.current_stream_ptr())
        L.check(rc, "rtf_sampled_softmax_fwd")
        ctx.save_for_backward(weights, inputs, labels, sampled, true_exp, samp_exp, lse)
        ctx.biases, ctx.remove_hits, ctx.table = biases, remove_hits, table
        return loss

    @staticmethod
    def backward(ctx, gloss):
        weights, inputs, labels, sampled, true_exp, samp_exp, lse = ctx.saved_tensors
        biases = ctx.biases
        B, D = inputs.shape
        N, S = weights.shape[0], sampled.numel()
        gloss = gloss.contiguous()
        gx = torch.empty_like(inputs)
        G = torch.empty((B, S + 1), dtype=torch.float32, device=inputs.device)
        ws = _ssm_ws(S, D, inputs.device)
        rc = L.lib().rtf_sampled_softmax_bwd(
            inputs.data_ptr(), inputs.stride(0), weights.data_ptr(),
            None if biases is None else biases.data_ptr(), labels.data_ptr(), sampled.data_ptr(),
            true_exp.data_ptr(), samp_exp.data_ptr(), B, N, S, D, int(ctx.remove_hits),
            lse.data_ptr(), gloss.data_ptr(), gx.data_ptr(), gx.stride(0), G.data_ptr(),
            ws.data_ptr(), L.current_stream_ptr())
        L.check(rc, "rtf_sampled_softmax_bwd")
        gw = None
        if ctx.table is not None and ctx.table[0].optimizer is not None:
            # the class-weight matrix is an embedding table with a fused sparse optimizer: its
            # touched rows ([labels | sampled]) are reduced and updated in place by K2 — no dense
            # (N, D) gradient is ever materialised
            tset, t = ctx.table
            rows = torch.cat([G[:, :1] * inputs, G[:, 1:].t() @ inputs], 0)
            ids = torch.cat([labels, sampled]).reshape(-1, 1)
            tset.apply_sparse_grad(ids, [t], rows)
        elif ctx.needs_input_grad[0]:
            # weight-row gradients: true rows g0*x (B,D), sampled rows G[:,1:]^T x (S,D), reduced
            # per touched row by K2 (deterministic), then placed into a dense (N,D) gradient
            rows = torch.cat([G[:, :1] * inputs, G[:, 1:].t() @ inputs], 0)
            ids = torch.cat([labels, sampled]).reshape(-1, 1)
            keys, tot, _ = embed_bwd([weights], [0], ids, rows, "BF", None, want_unique=True)
            gw = torch.zeros_like(weights)
            gw[keys] = tot[:, :D]
        return gw, None, None, gx, None, None, None, None, None, None


class _SampledSoftmaxGemmFn(torch.autograd.Function):
    """K8 for large S: the (B, S) sampled logits are one tensor-core GEMM x . Ws^T
    (rtf_dense_gemm_nt, fp32-accurate), the per-sample work (true logit, hit removal, log-sum-exp)
    a streaming epilogue; backward = epilogue (g0, G1) + two GEMMs.  Same semantics as
    _SampledSoftmaxFn (rtf_sampled_softmax_*), which re-streams the S rows per sample."""

    @staticmethod
    def forward(ctx, weights, biases, labels, inputs, sampled, true_exp, samp_exp, remove_hits, err,
                table=None):
        from ..core import dense_gemm
        L.require_cuda(inputs, "sampled_softmax_loss(inputs)")
        weights, inputs = weights.contiguous(), inputs.contiguous()
        B, D = inputs.shape
        N, S = weights.shape[0], sampled.numel()
        dev = inputs.device
        Ws = torch.empty((S, D), dtype=torch.float32, device=dev)
        cs = torch.empty(S, dtype=torch.float32, device=dev)
        bptr = None if biases is None else biases.data_ptr()
        eptr = None if err is None else err.data_ptr()
        L.check(L.lib().rtf_ssm_gather(weights.data_ptr(), bptr, sampled.data_ptr(), samp_exp.data_ptr(),
                                       N, S, D, Ws.data_ptr(), cs.data_ptr(), eptr,
                                       L.current_stream_ptr()), "rtf_ssm_gather")
        logits = dense_gemm("nt", inputs, Ws)                        # (B, S)
        loss = torch.empty(B, dtype=torch.float32, device=dev)
        lse = torch.empty_like(loss)
        rc = L.lib().rtf_ssm_logits_fwd(logits.data_ptr(), logits.stride(0), inputs.data_ptr(),
                                        inputs.stride(0), weights.data_ptr(), bptr, labels.data_ptr(),
                                        sampled.data_ptr(), true_exp.data_ptr(), cs.data_ptr(), B, N, S, D,
                                        int(remove_hits), loss.data_ptr(), lse.data_ptr(), eptr,
                                        L.current_stream_ptr())
        L.check(rc, "rtf_ssm_logits_fwd")
        ctx.save_for_backward(weights, inputs, labels, sampled, true_exp, logits, Ws, cs, lse)
        ctx.biases, ctx.remove_hits, ctx.table = biases, remove_hits, table
        return loss

    @staticmethod
    def backward(ctx, gloss):
        from ..core import dense_gemm
        weights, inputs, labels, sampled, true_exp, logits, Ws, cs, lse = ctx.saved_tensors
        biases = ctx.biases
        B, D = inputs.shape
        N, S = weights.shape[0], sampled.numel()
        gloss = gloss.contiguous()
        g0 = torch.empty(B, dtype=torch.float32, device=inputs.device)
        G1 = torch.empty((B, S), dtype=torch.float32, device=inputs.device)
        rc = L.lib().rtf_ssm_logits_bwd(logits.data_ptr(), logits.stride(0), inputs.data_ptr(),
                                        inputs.stride(0), weights.data_ptr(),
                                        None if biases is None else biases.data_ptr(), labels.data_ptr(),
                                        sampled.data_ptr(), true_exp.data_ptr(), cs.data_ptr(), B, N, S, D,
                                        int(ctx.remove_hits), lse.data_ptr(), gloss.data_ptr(), g0.data_ptr(),
                                        G1.data_ptr(), G1.stride(0), L.current_stream_ptr())
        L.check(rc, "rtf_ssm_logits_bwd")
        gx = dense_gemm("nn", G1, Ws)                                # sampled share of d loss / d x
        L.check(L.lib().rtf_ssm_true_gx(weights.data_ptr(), labels.data_ptr(), g0.data_ptr(), B, N, D,
                                        gx.data_ptr(), gx.stride(0), L.current_stream_ptr()),
                "rtf_ssm_true_gx")
        gw = None
        fused = ctx.table is not None and ctx.table[0].optimizer is not None
        if fused or ctx.needs_input_grad[0]:
            # weight-row gradients: true rows g0*x (B,D), sampled rows G1^T x (S,D); K2 reduces them
            # per touched row ([labels | sampled] may repeat), deterministically
            rows = torch.cat([g0.unsqueeze(1) * inputs, dense_gemm("tn", G1, inputs)], 0)
            ids = torch.cat([labels, sampled]).reshape(-1, 1)
            if fused:
                tset, t = ctx.table
                tset.apply_sparse_grad(ids, [t], rows)
            else:
                keys, tot, _ = embed_bwd([weights], [0], ids, rows, "BF", None, want_unique=True)
                gw = torch.zeros_like(weights)
                gw[keys] = tot[:, :D]
        return gw, None, None, gx, None, None, None, None, None, None


def _gemm_form_ok(B, S, D) -> bool:
    """The GEMM form pays from a few dozen sampled classes up (and needs 16-byte rows)."""
    return S >= 32 and S % 4 == 0 and D % 4 == 0 and B % 4 == 0 and B >= 64


def sampled_softmax_loss(weights, biases, labels, inputs, num_sampled, num_classes, num_true=1,
                         sampled_values=None, remove_accidental_hits=True, seed: int = 0,
                         err: Optional[torch.Tensor] = None, table=None):
    """tf.nn.sampled_softmax_loss (A13).  `sampled_values = (sampled, true_expected_count,
    sampled_expected_count)` may be injected (the only way to reproduce a TF run); otherwise
    they come from the device log-uniform sampler seeded with `seed`.  table=(EmbeddingTables, t):
    `weights` is table t of that set — with a fused sparse optimizer its touched rows are updated
    in place by K2 in the backward instead of returning a dense (N, D) gradient."""
    if num_true != 1:
        raise NotImplementedError("num_true != 1")
    labels = labels.reshape(-1).to(torch.int64).contiguous()
    if sampled_values is None:
        sampled, tries = log_uniform_candidate_sampler(num_sampled, num_classes, seed, inputs.device)
        true_exp = log_uniform_expected(labels, num_classes, tries)
        samp_exp = log_uniform_expected(sampled, num_classes, tries)
    else:
        sampled, true_exp, samp_exp = sampled_values
        sampled = sampled.reshape(-1).to(torch.int64).contiguous()
        true_exp = true_exp.reshape(-1).to(torch.float32).contiguous()
        samp_exp = samp_exp.reshape(-1).to(torch.float32).contiguous()
    fn = _SampledSoftmaxGemmFn if _gemm_form_ok(inputs.shape[0], sampled.numel(), inputs.shape[1]) \
        else _SampledSoftmaxFn
    return fn.apply(weights, biases, labels, inputs, sampled, true_exp, samp_exp,
                    remove_accidental_hits, err, table)


class SampledSoftmaxLayer(Layer):
    """SampledSoftmaxLayer(num_sampled=5); call([item_embeddings (B,1,n), user_embeddings
    (B,1,n), label_idx (B,1)]) -> (B,1).  As in the source (:34-60) the class-weight matrix is
    the squeezed item tensor itself, num_classes = its last dimension and the bias is a
    non-trainable zero vector of that size."""

    def __init__(self, num_sampled=5, seed: int = 0, **kwargs):
        super().__init__(**kwargs)
        self.num_sampled = num_sampled
        self.seed = seed
        self._calls = 0

    def build(self, input_shape):
        self.size = input_shape[0][2]
        self.zero_bias = self.add_weight("bias", (self.size,), "zeros", trainable=False)

    def call(self, inputs_with_label_idx, training=None, sampled_values=None, **kwargs):
        item_embeddings, user_embeddings, label_idx = inputs_with_label_idx
        item_embeddings = item_embeddings.squeeze(1)
        user_embeddings = user_embeddings.squeeze(1)
        self._calls += 1
        loss = sampled_softmax_loss(weights=item_embeddings, biases=self.zero_bias, labels=label_idx,
                                    inputs=user_embeddings, num_sampled=self.num_sampled,
                                    num_classes=self.size, sampled_values=sampled_values,
                                    seed=self.seed + self._calls)
        return loss.unsqueeze(1)


def sampledsoftmaxloss(y_true, y_pred):
    """src/match/utils/loss_util.py:11-13: K.mean(y_pred)."""
    return y_pred.mean()

"""ctr-side layers, same names / constructor arguments / call structure as
src/ctr/layers/modules.py: FM (:36-72), AttentionLayer (:137-175), MultiHeadAttention
(:177-325), Dice (:327-337).  DNN / Dense glue lives in core.py."""
from __future__ import annotations

import ctypes as C
import math

import torch

from .. import _lib as L
from ..attention import attention
from ..fm import FM, colsum  # noqa: F401  (FM re-exported under its reference name)
from .core import BatchNormalization, Dense, Layer, get_activation, l2

_ACT_CODE = {None: 0, "linear": 0, "relu": 1, "sigmoid": 2, "tanh": 3}


class _DinAttnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, mask, W, bias, act):
        L.require_cuda(k, "AttentionLayer(k)")
        q, k = q.contiguous(), k.contiguous()
        v = k if (v is k or v.data_ptr() == k.data_ptr()) else v.contiguous()
        B, Lk, d = k.shape
        out = torch.empty((B, d), dtype=torch.float32, device=k.device)
        Wf = W.reshape(-1).contiguous()
        rc = L.lib().rtf_din_attn_fwd(q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0),
                                      v.data_ptr(), v.stride(0),
                                      None if mask is None else mask.data_ptr(), Lk, Wf.data_ptr(),
                                      bias.data_ptr(), act, B, Lk, d, out.data_ptr(), out.stride(0),
                                      L.current_stream_ptr())
        L.check(rc, "rtf_din_attn_fwd")
        ctx.save_for_backward(q, k, v, Wf, bias)
        ctx.mask, ctx.act, ctx.wshape = mask, act, W.shape
        return out

    @staticmethod
    def backward(ctx, gout):
        q, k, v, Wf, bias = ctx.saved_tensors
        mask, act = ctx.mask, ctx.act
        gout = gout.contiguous()
        B, Lk, d = k.shape
        gq, gk, gv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(k)
        gw_rows = torch.empty((B, 4 * d + 1), dtype=torch.float32, device=k.device)
        rc = L.lib().rtf_din_attn_bwd(q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0),
                                      v.data_ptr(), v.stride(0),
                                      None if mask is None else mask.data_ptr(), Lk, Wf.data_ptr(),
                                      bias.data_ptr(), act, B, Lk, d, gout.data_ptr(),
                                      gout.stride(0), gq.data_ptr(), gq.stride(0), gk.data_ptr(),
                                      gk.stride(0), gv.data_ptr(), gv.stride(0), gw_rows.data_ptr(),
                                      L.current_stream_ptr())
        L.check(rc, "rtf_din_attn_bwd")
        gw = colsum(gw_rows)
        return gq, gk, gv, None, gw[: 4 * d].reshape(ctx.wshape), gw[4 * d:].reshape(1), None


class AttentionLayer(Layer):
    """DIN local activation unit: AttentionLayer(hidden_unit, activation='prelu');
    call([q (B,d), k (B,L,d), v (B,L,d), mask (B,L)]) -> (B,d).

    As in the source the score reshape (:159) only admits hidden_unit == 1, and the default
    activation string 'prelu' is not a Keras activation (the constructor raises, as Keras does;
    the reference's own script passes 'sigmoid', src/ctr/din/train.py:31).  If `mask` is not a
    tensor every score is replaced by the pad value -> uniform weights (:164-165)."""

    def __init__(self, hidden_unit, activation="prelu"):
        super().__init__()
        if activation not in _ACT_CODE:
            get_activation(activation)          # raises ValueError for unknown strings
        if hidden_unit != 1:
            raise ValueError("AttentionLayer: the reshape to (B, L) requires hidden_unit == 1")
        self.hidden_unit, self.activation = hidden_unit, activation

    def build(self, input_shape):
        d = input_shape[1][-1]
        self.att_dense_kernel = self.add_weight("att_dense_kernel", (4 * d, 1), "glorot_uniform")
        self.att_dense_bias = self.add_weight("att_dense_bias", (1,), "zeros")

    def call(self, inputs, **kwargs):
        q, k, v, mask = inputs
        m = None
        if isinstance(mask, torch.Tensor):
            m = mask.reshape(k.shape[0], k.shape[1]).to(torch.float32).contiguous()
        return _DinAttnFn.apply(q, k, v, m, self.att_dense_kernel, self.att_dense_bias,
                                _ACT_CODE[self.activation])


class _AutoIntLayerFn(torch.autograd.Function):
    """K6: the whole interacting layer (projections + activation, per-head softmax(QK^T s)V,
    residual relu) in one launch per direction (csrc/autoint_layer.cu)."""

    @staticmethod
    def forward(ctx, x, wq, wk, wv, w0, H, hs, act, scale):
        L.require_cuda(x, "MultiHeadAttention(x)")
        x = x.contiguous()
        ws_ = [w.contiguous() for w in (wq, wk, wv)] + [None if w0 is None else w0.contiguous()]
        B, F, dm = x.shape
        out = torch.empty((B, F, H * hs), dtype=torch.float32, device=x.device)
        rc = L.lib().rtf_autoint_layer_fwd(x.data_ptr(), B, F, dm, ws_[0].data_ptr(), ws_[1].data_ptr(),
                                           ws_[2].data_ptr(), None if w0 is None else ws_[3].data_ptr(),
                                           H, hs, act, scale, out.data_ptr(), L.current_stream_ptr())
        L.check(rc, "rtf_autoint_layer_fwd")
        ctx.save_for_backward(x, ws_[0], ws_[1], ws_[2], ws_[3], out)
        ctx.cfg = (H, hs, act, scale)
        return out

    @staticmethod
    def backward(ctx, gout):
        x, wq, wk, wv, w0, out = ctx.saved_tensors
        H, hs, act, scale = ctx.cfg
        gout = gout.contiguous()
        B, F, dm = x.shape
        HS = H * hs
        gx = torch.empty_like(x)
        gw = torch.empty((4, dm, HS), dtype=torch.float32, device=x.device)
        nb = C.c_size_t(0)
        L.check(L.lib().rtf_autoint_layer_workspace(B, dm, HS, C.byref(nb)), "rtf_autoint_layer_workspace")
        ws = torch.empty(max(nb.value, 16), dtype=torch.uint8, device=x.device)
        rc = L.lib().rtf_autoint_layer_bwd(x.data_ptr(), B, F, dm, wq.data_ptr(), wk.data_ptr(),
                                           wv.data_ptr(), None if w0 is None else w0.data_ptr(), H, hs,
                                           act, scale, out.data_ptr(), gout.data_ptr(), gx.data_ptr(),
                                           gw.data_ptr(), ws.data_ptr(), ws.numel(), L.current_stream_ptr())
        L.check(rc, "rtf_autoint_layer_bwd")
        return gx, gw[0], gw[1], gw[2], (None if w0 is None else gw[3]), None, None, None, None


class MultiHeadAttention(Layer):
    """AutoInt interacting layer: MultiHeadAttention(head_size, head_num=1, l2_reg=l2(1e-4),
    activation='relu', use_res=False, name=''); call(X) or call([q, k, v]) / call([x]).

    Q,K,V = act(X W) with no bias and the activation on all three (:255-270); the Dense layers
    are created on the first call (A5).  scale='reference' keeps the source's
    `product / head_size**-0.5` (= multiply by sqrt(head_size), :235-237); scale='paper' is the
    usual 1/sqrt(head_size).  No mask.  use_res adds relu(out + act(X W0)) (:316-323)."""

    def __init__(self, head_size, head_num=1, l2_reg=None, activation="relu", use_res=False,
                 name="", scale="reference"):
        super().__init__()
        self._head_num, self._head_size = head_num, head_size
        self._l2_reg = l2(1e-4) if l2_reg is None else l2_reg
        self._activation, self._use_res, self._scale = activation, use_res, scale
        self.fused = True       # False: always compose from the projection GEMMs + rtf_attn_*
        hs = head_num * head_size
        self.q_dense = Dense(hs, activation=activation, use_bias=False, kernel_regularizer=self._l2_reg)
        self.k_dense = Dense(hs, activation=activation, use_bias=False, kernel_regularizer=self._l2_reg)
        self.v_dense = Dense(hs, activation=activation, use_bias=False, kernel_regularizer=self._l2_reg)
        self.res_dense = (Dense(hs, activation=activation, use_bias=False,
                                kernel_regularizer=self._l2_reg) if use_res else None)

    def call(self, inputs, **kwargs):
        if isinstance(inputs, (list, tuple)):
            assert len(inputs) in (1, 3), \
                "If the input of multi_head_attention is a list, the length must be 1 or 3."
            ori_q, ori_k, ori_v = (inputs if len(inputs) == 3 else (inputs[0],) * 3)
        else:
            ori_q = ori_k = ori_v = inputs
        hs = self._head_size
        scale = math.sqrt(hs) if self._scale == "reference" else 1.0 / math.sqrt(hs)
        if (ori_q is ori_k and ori_k is ori_v and ori_q.is_cuda and ori_q.dim() == 3
                and ori_q.dtype == torch.float32 and self._activation in _ACT_CODE and self.fused
                and L.lib().rtf_autoint_layer_supported(ori_q.shape[1], ori_q.shape[2], self._head_num, hs)):
            # K6: self-attention input of a compiled shape -> the whole layer in one launch
            for dns in (self.q_dense, self.k_dense, self.v_dense, self.res_dense):
                if dns is not None and not dns._built:
                    dns.build(tuple(ori_q.shape))
                    dns._built = True
            return _AutoIntLayerFn.apply(ori_q, self.q_dense.kernel, self.k_dense.kernel, self.v_dense.kernel,
                                         self.res_dense.kernel if self._use_res else None, self._head_num,
                                         hs, _ACT_CODE[self._activation], scale)
        q, k, v = self.q_dense(ori_q), self.k_dense(ori_k), self.v_dense(ori_v)
        out = attention(q, k, v, self._head_num, scale)
        if self._use_res:
            return torch.relu(out + self.res_dense(ori_v))
        return out


class Dice(Layer):
    """Dice (:327-337): p = sigmoid(BN(x)) with BatchNormalization(center=False, scale=False);
    out = alpha*(1-p)*x + p*x.  alpha has no initializer in the source, i.e. Keras'
    glorot_uniform on a scalar = U(-sqrt(3), sqrt(3))."""

    def __init__(self):
        super().__init__()
        self.bn = BatchNormalization(center=False, scale=False)

    def build(self, input_shape):
        self.alpha = self.add_weight("alpha", (), "glorot_uniform")

    def call(self, x, **kwargs):
        x_p = torch.sigmoid(self.bn(x))
        return self.alpha * (1.0 - x_p) * x + x_p * x

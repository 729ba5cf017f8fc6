"""Host side of K3: the FM layer (src/ctr/layers/modules.py:36-72) and the FM model in gather
form (src/ctr/fm/model.py:5-59)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib as L
from .embedding import EmbeddingTables, SparseOptimizer, _ptr_array
from .core import Layer, l2


def colsum(x: torch.Tensor, rowscale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Deterministic out[c] = sum_b rowscale[b] * x[b, c] (fixed two-stage order)."""
    L.require_cuda(x, "colsum")
    x = x.contiguous()
    B, cols = x.shape
    nbytes = C.c_size_t(0)
    L.check(L.lib().rtf_colsum_workspace(B, cols, C.byref(nbytes)), "rtf_colsum_workspace")
    ws = torch.empty(max(nbytes.value, 16), dtype=torch.uint8, device=x.device)
    out = torch.empty(cols, dtype=torch.float32, device=x.device)
    L.check(L.lib().rtf_colsum(x.data_ptr(), x.stride(0),
                               None if rowscale is None else rowscale.contiguous().data_ptr(), B,
                               cols, out.data_ptr(), ws.data_ptr(), L.current_stream_ptr()),
            "rtf_colsum")
    return out


def _fm_ws(B, P1, device):
    nbytes = C.c_size_t(0)
    L.check(L.lib().rtf_fm_layer_workspace(B, P1, C.byref(nbytes)), "rtf_fm_layer_workspace")
    return torch.empty(max(nbytes.value, 16), dtype=torch.uint8, device=device)


class _FMLayerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, first, w, second, batch_scalar, sum_d):
        L.require_cuda(first, "FM(first_inputs)")
        first, second, w = first.contiguous(), second.contiguous(), w.contiguous()
        B, P1 = first.shape
        _, F, D = second.shape
        out = torch.empty(B if sum_d else B * D, dtype=torch.float32, device=first.device)
        ws = _fm_ws(B, P1, first.device)
        L.check(L.lib().rtf_fm_layer_fwd(first.data_ptr(), first.stride(0), w.data_ptr(), P1,
                                         second.data_ptr(), second.stride(0), F, D, B,
                                         int(batch_scalar), int(sum_d), out.data_ptr(),
                                         ws.data_ptr(), L.current_stream_ptr()), "rtf_fm_layer_fwd")
        ctx.save_for_backward(first, w, second)
        ctx.flags = (batch_scalar, sum_d)
        return out.view(-1, 1)

    @staticmethod
    def backward(ctx, gout):
        first, w, second = ctx.saved_tensors
        batch_scalar, sum_d = ctx.flags
        gout = gout.contiguous()
        B, P1 = first.shape
        _, F, D = second.shape
        gfirst, gsecond = torch.empty_like(first), torch.empty_like(second)
        gw = torch.empty(P1, dtype=torch.float32, device=first.device)
        ws = _fm_ws(B, P1, first.device)
        L.check(L.lib().rtf_fm_layer_bwd(first.data_ptr(), first.stride(0), w.data_ptr(), P1,
                                         second.data_ptr(), second.stride(0), F, D, B,
                                         int(batch_scalar), int(sum_d), gout.data_ptr(),
                                         gfirst.data_ptr(), gfirst.stride(0), gw.data_ptr(),
                                         gsecond.data_ptr(), gsecond.stride(0), ws.data_ptr(),
                                         L.current_stream_ptr()), "rtf_fm_layer_bwd")
        return gfirst, gw.view_as(w), gsecond, None, None


class FM(Layer):
    """ctr.layers.modules.FM(feature_length, w_reg=1e-6); call([first_inputs, second_inputs]).

    mode='reference' keeps the source's semantics: the first-order term is reduced over the
    whole batch to one scalar (:65); a 2-D second_inputs (what DeepFM passes) is crossed over
    all its F*D scalars -> (B,1); a 3-D (B,F,D) one gives (B*D,1) (:70-71).
    mode='paper' is the per-sample first order plus the second order summed over D -> (B,1)."""

    def __init__(self, feature_length: int, w_reg: float = 1e-6, mode: str = "reference"):
        super().__init__()
        if mode not in ("reference", "paper"):
            raise ValueError(mode)
        self.feature_length, self.w_reg, self.mode = feature_length, w_reg, mode

    def build(self, input_shape):
        self.w = self.add_weight("w", (self.feature_length, 1), "random_normal", l2(self.w_reg))

    def call(self, inputs, **kwargs):
        first_inputs, second_inputs = inputs
        if second_inputs.dim() == 2:
            second_inputs = second_inputs.unsqueeze(-1)      # (B, M) == (B, F=M, D=1)
        ref = self.mode == "reference"
        return _FMLayerFn.apply(first_inputs, self.w, second_inputs, ref, not ref)


class _FMGatherFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model: "FMModel", dense, ids, w0, dense_table, *weights):
        L.require_cuda(ids, "FM model (sparse_inputs)")
        tset = model.tables
        dense = dense.contiguous()
        B, F = ids.shape
        k, kp, nd = model.k, model.kp, dense.shape[1]
        out = torch.empty(B, dtype=torch.float32, device=ids.device)
        A = torch.empty((B, kp), dtype=torch.float32, device=ids.device)
        rows = L.host_array(C.c_int64, [int(t.shape[0]) for t in weights])
        rc = L.lib().rtf_fm_gather_fwd(_ptr_array(weights), rows, F, k, kp, dense_table.data_ptr(),
                                       nd, dense.data_ptr(), dense.stride(0), ids.data_ptr(),
                                       int(ids.dtype == torch.int64), B, ids.stride(0),
                                       ids.stride(1), w0.data_ptr(), out.data_ptr(), A.data_ptr(),
                                       tset.err.data_ptr(), L.current_stream_ptr())
        L.check(rc, "rtf_fm_gather_fwd")
        ctx.model, ctx.ids = model, ids
        ctx.save_for_backward(dense, out, A, dense_table)
        return out.view(B, 1)

    @staticmethod
    def backward(ctx, gout):
        dense, out, A, dense_table = ctx.saved_tensors
        model, ids = ctx.model, ctx.ids
        tset = model.tables
        weights = list(tset.weights)
        B, F = ids.shape
        k, kp, nd = model.k, model.kp, dense.shape[1]
        gout = gout.contiguous()
        gsparse = torch.empty((B, F * kp), dtype=torch.float32, device=ids.device)
        gdrows = torch.empty((B, nd * kp), dtype=torch.float32, device=ids.device)
        dz = torch.empty(B, dtype=torch.float32, device=ids.device)
        rows = L.host_array(C.c_int64, [int(t.shape[0]) for t in weights])
        rc = L.lib().rtf_fm_gather_bwd(_ptr_array(weights), rows, F, k, kp, dense_table.data_ptr(),
                                       nd, dense.data_ptr(), dense.stride(0), ids.data_ptr(),
                                       int(ids.dtype == torch.int64), B, ids.stride(0),
                                       ids.stride(1), out.data_ptr(), A.data_ptr(), gout.data_ptr(),
                                       gsparse.data_ptr(), gdrows.data_ptr(), dz.data_ptr(),
                                       L.current_stream_ptr())
        L.check(rc, "rtf_fm_gather_bwd")
        gw0 = colsum(dz.view(B, 1)).view(1)
        gdt = colsum(gdrows).view(nd, kp)
        wgrads = tset.grads_from_lookup_grad(ids, tuple(range(F)), gsparse, "BF", None)
        return (None, None, None, gw0, gdt) + wgrads


class FMModel(Layer):
    """ctr.fm.model.FM(feature_columns, k, w_reg=1e-4, v_reg=1e-4) — src/ctr/fm/model.py:5-32;
    call([dense_inputs (B,13) f32, sparse_inputs (B,26) i32]) -> sigmoid(first + second) (B,1).

    The reference multiplies a dense one-hot (B, M) matrix by w (M,1) and V^T; here every
    feature owns one HBM row [V[:, m] (k) | w[m] | 0-pad] (kp floats, 16-byte aligned) and a
    sample gathers its 13 + 26 rows.  `reference_weights()` / `load_reference_weights()` convert
    to and from the reference's (w0, w, V) layout (dense features first, then the fields in
    order, as `tf.concat([dense_inputs] + one_hots)` stacks them at :37-42)."""

    def __init__(self, feature_columns, k: int, w_reg: float = 1e-4, v_reg: float = 1e-4,
                 sparse_optimizer: Optional[SparseOptimizer] = None, seed: Optional[int] = None):
        super().__init__()
        self.dense_feature_columns, self.sparse_feature_columns = feature_columns
        self.k, self.w_reg, self.v_reg = k, w_reg, v_reg
        self.kp = (k + 1 + 3) // 4 * 4
        self.index_mapping = []
        self.feature_length = 0
        for feat in self.sparse_feature_columns:                 # model.py:18-21
            self.index_mapping.append(self.feature_length)
            self.feature_length += feat["feat_num"]
        rows = [f["feat_num"] for f in self.sparse_feature_columns]
        self.tables = EmbeddingTables(rows, [self.kp] * len(rows), "random_normal",
                                      optimizer=sparse_optimizer, seed=seed)
        with torch.no_grad():
            for w in self.tables.weights:
                w[:, k + 1:] = 0.0
        nd = len(self.dense_feature_columns)
        dev = self.tables.weights[0].device
        dt = torch.zeros((nd, self.kp), device=dev)
        dt[:, : k + 1].normal_(0.0, 0.05)
        self.dense_table = torch.nn.Parameter(dt)
        self.w0 = torch.nn.Parameter(torch.zeros(1, device=dev))  # model.py:23-25

    def call(self, inputs, **kwargs):
        dense_inputs, sparse_inputs = inputs
        if sparse_inputs.dtype not in (torch.int32, torch.int64):
            sparse_inputs = sparse_inputs.to(torch.int32)
        return _FMGatherFn.apply(self, dense_inputs, sparse_inputs, self.w0, self.dense_table,
                                 *self.tables.weights)

    # ---- reference layout <-> row layout
    def reference_weights(self):
        k = self.k
        rows = torch.cat([self.dense_table.detach()] + [w.detach() for w in self.tables.weights], 0)
        return self.w0.detach().clone(), rows[:, k:k + 1].clone(), rows[:, :k].t().contiguous()

    @torch.no_grad()
    def load_reference_weights(self, w0, w, V):
        """w0 (1,), w (M,1), V (k,M) with M = n_dense + sum(feat_num)."""
        k, nd = self.k, self.dense_table.shape[0]
        rows = torch.zeros((w.shape[0], self.kp), device=self.w0.device)
        rows[:, :k] = torch.as_tensor(V).t().to(rows)
        rows[:, k] = torch.as_tensor(w).reshape(-1).to(rows)
        self.w0.copy_(torch.as_tensor(w0).reshape(1))
        self.dense_table.copy_(rows[:nd])
        off = nd
        for t in self.tables.weights:
            t.copy_(rows[off:off + t.shape[0]])
            off += t.shape[0]

    def regularization_loss(self):
        k = self.k
        rows = [self.dense_table] + list(self.tables.weights)
        return sum(self.w_reg * (r[:, k] ** 2).sum() + self.v_reg * (r[:, :k] ** 2).sum() for r in rows)

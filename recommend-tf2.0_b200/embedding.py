"""Host side of K1/K2: multi-table embedding lookup (+pooling) and its deterministic
backward with in-place sparse optimizers.

Mirrors what the reference builds per model: a dict of `Embedding(input_dim, output_dim,
embeddings_initializer, embeddings_regularizer=l2(embed_reg))` and
`tf.concat([embed_i(sparse_inputs[:, i]) ...], axis=-1)` (src/ctr/dlrm/model.py:30-37,45-46).
All arithmetic happens in librtf_b200.so; torch only owns memory, streams and autograd glue.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib as L

_POOL = {None: L.POOL_NONE, "none": L.POOL_NONE, "sum": L.POOL_SUM, "mean": L.POOL_MEAN}
_OPT = {None: L.OPT_NONE, "none": L.OPT_NONE, "sgd": L.OPT_SGD, "adagrad": L.OPT_ADAGRAD,
        "adam": L.OPT_ADAM}


class SparseOptimizer:
    """Row-wise optimizer applied by K2 to the rows a batch touches.

    Defaults follow Keras `Adam(learning_rate=1e-3)` (beta 0.9/0.999, eps 1e-7), which is
    what every reference script compiles with (src/ctr/fm/train.py:49-50).  `l2` is the
    `embeddings_regularizer=l2(embed_reg)` coefficient: its gradient 2*l2*W is added on the
    touched rows only (documented divergence from TF's dense sweep)."""

    def __init__(self, kind: str = "adam", lr: float = 1e-3, beta1: float = 0.9,
                 beta2: float = 0.999, eps: float = 1e-7, l2: float = 0.0):
        if kind not in _OPT:
            raise ValueError(f"unknown sparse optimizer {kind!r}")
        self.kind, self.lr, self.beta1, self.beta2, self.eps, self.l2 = kind, lr, beta1, beta2, eps, l2
        self.step = 0
        self.lr_dev: Optional[torch.Tensor] = None

    def lr_for_step(self, step: int) -> float:
        if self.kind == "adam":  # Keras folds both bias corrections into the step size
            return self.lr * math.sqrt(1.0 - self.beta2 ** step) / (1.0 - self.beta1 ** step)
        return self.lr

    def struct_for_step(self, step: int) -> L.rtf_opt:
        return L.rtf_opt(_OPT[self.kind], self.lr_for_step(step), self.beta1, self.beta2, self.eps,
                         self.l2, None if self.lr_dev is None else self.lr_dev.data_ptr())

    def enable_device_lr(self, device):
        """Keep the step size in a device scalar (rtf_opt.lr_dev) refreshed by advance(): K2 then
        reads it from memory instead of its launch parameters, which is what lets a whole training
        step be replayed from a CUDA graph (core.StepGraph) while Adam's bias-corrected step size
        keeps changing.  The value is the same fp32 number the host path passes."""
        if self.lr_dev is None:
            self.lr_dev = torch.zeros(1, dtype=torch.float32, device=device)
            self.lr_dev.fill_(self.lr_for_step(max(self.step, 1)))

    def advance(self):
        """Next training step: bump the counter and refresh the device step size, if enabled."""
        self.step += 1
        if self.lr_dev is not None:
            self.lr_dev.fill_(self.lr_for_step(self.step))

    @property
    def n_states(self) -> int:
        return {"adam": 2, "adagrad": 1}.get(self.kind, 0)


def ids_strides(ids: torch.Tensor, layout: str) -> Tuple[int, int, int, int, int]:
    """-> (B, F, L, sb, sf, sl)-style description of an id tensor.
    layout 'BF' : (B, F);  'BFL' : (B, F, L);  'BLF' : (B, L, F);  'BL' : (B, L) one field."""
    if ids.dtype not in (torch.int32, torch.int64):
        raise TypeError("ids must be int32 or int64 (float ids are cast by the layer classes)")
    if layout == "BF":
        B, F = ids.shape
        return B, F, 1, ids.stride(0), ids.stride(1), 0
    if layout == "BFL":
        B, F, Lq = ids.shape
        return B, F, Lq, ids.stride(0), ids.stride(1), ids.stride(2)
    if layout == "BLF":
        B, Lq, F = ids.shape
        return B, F, Lq, ids.stride(0), ids.stride(2), ids.stride(1)
    if layout == "BL":
        B, Lq = ids.shape
        return B, 1, Lq, ids.stride(0), 0, ids.stride(1)
    raise ValueError(f"unknown id layout {layout!r}")


def _ptr_array(tensors: Sequence[Optional[torch.Tensor]]):
    return (C.c_void_p * len(tensors))(*[None if t is None else t.data_ptr() for t in tensors])


def embed_fwd(tables: Sequence[torch.Tensor], ids: torch.Tensor, layout: str = "BF",
              pool: Optional[str] = None, err: Optional[torch.Tensor] = None,
              out: Optional[torch.Tensor] = None, skip_invalid: bool = False) -> torch.Tensor:
    """K1.  tables[f] is the (rows, dim) fp32 table of lookup field f (a table may repeat).
    Returns (B, sumD) for L == 1 or pooled lookups, else (B, L, sumD).  skip_invalid: rows of
    out-of-range ids are left untouched in `out` instead of zero-filled and flagged."""
    lib = L.lib()
    L.require_cuda(ids, "embed_fwd(ids)")
    B, F, Lq, sb, sf, sl = ids_strides(ids, layout)
    if F != len(tables):
        raise ValueError(f"{len(tables)} tables for {F} id fields")
    for t in tables:
        L.require_cuda(t, "embed_fwd(table)")
        if t.dtype != torch.float32 or not t.is_contiguous() or t.dim() != 2:
            raise TypeError("tables must be contiguous fp32 (rows, dim)")
    dims = [int(t.shape[1]) for t in tables]
    rows = [int(t.shape[0]) for t in tables]
    sumD = sum(dims)
    pm = _POOL[pool]
    flags = 0x100 if skip_invalid else 0        # RTF_POOL_SKIP_INVALID
    if out is None:
        shape = (B, Lq, sumD) if (pm == L.POOL_NONE and layout != "BF") else (B, sumD)
        out = torch.empty(shape, dtype=torch.float32, device=ids.device)
    out_sb = out.stride(0) if B > 0 else (Lq * sumD)
    rc = lib.rtf_embed_fwd(_ptr_array(tables), L.host_array(C.c_int64, rows),
                           L.host_array(C.c_int32, dims), F, ids.data_ptr(),
                           int(ids.dtype == torch.int64), B, Lq, sb, sf, sl, pm | flags,
                           out.data_ptr(), out_sb, None if err is None else err.data_ptr(),
                           L.current_stream_ptr())
    L.check(rc, "rtf_embed_fwd")
    return out


def embed_bwd(weights: Sequence[torch.Tensor], field_table: Sequence[int], ids: torch.Tensor,
              grad: torch.Tensor, layout: str = "BF", pool: Optional[str] = None,
              opt: Optional[L.rtf_opt] = None, state1: Optional[Sequence[torch.Tensor]] = None,
              state2: Optional[Sequence[torch.Tensor]] = None, want_unique: bool = False,
              sync: bool = True):
    """K2.  Sorts the (table, id) keys of the batch, sums every touched row's gradient in
    ascending lookup position and applies `opt` in place.  With want_unique=True also returns
    (keys uint32 as int64 tensor, summed grads (n_unique, dim_max), row_bits); with sync=False
    the host is not synchronised: the arrays come back full-size (min(lookups, total rows)
    entries, unused keys = 0xFFFFFFFF) together with the device-side count."""
    lib = L.lib()
    L.require_cuda(ids, "embed_bwd(ids)")
    L.require_cuda(grad, "embed_bwd(grad)")
    B, F, Lq, sb, sf, sl = ids_strides(ids, layout)
    if F != len(field_table):
        raise ValueError("field_table must have one entry per id field")
    if grad.dtype != torch.float32 or grad.stride(-1) != 1:
        raise TypeError("grad must be fp32 with unit inner stride")
    nt = len(weights)
    dims = [int(w.shape[1]) for w in weights]
    rows = [int(w.shape[0]) for w in weights]
    dim_max = max(dims)
    n = B * F * Lq
    nbytes = C.c_size_t(0)
    L.check(lib.rtf_embed_bwd_workspace(n, dim_max, C.byref(nbytes)), "rtf_embed_bwd_workspace")
    ws = torch.empty(max(nbytes.value, 16), dtype=torch.uint8, device=ids.device)
    uk = ug = nu = None
    if want_unique:
        cap = max(n, 1) if sync else max(min(n, sum(rows)), 1)
        uk = torch.full((cap,), 0 if sync else -1, dtype=torch.int32, device=ids.device)
        ug = torch.zeros((cap, dim_max), dtype=torch.float32, device=ids.device)
        nu = torch.zeros(1, dtype=torch.int32, device=ids.device)
    if opt is None:
        opt = L.rtf_opt(L.OPT_NONE, 0.0, 0.0, 0.0, 0.0, 0.0, None)
    row_bits = C.c_int(0)
    gsb = grad.stride(0) if grad.dim() >= 2 and B > 0 else 0
    rc = lib.rtf_embed_bwd(_ptr_array(weights),
                           _ptr_array(state1 if state1 is not None else [None] * nt),
                           _ptr_array(state2 if state2 is not None else [None] * nt),
                           L.host_array(C.c_int64, rows), L.host_array(C.c_int32, dims), nt,
                           L.host_array(C.c_int32, list(field_table)), F, ids.data_ptr(),
                           int(ids.dtype == torch.int64), B, Lq, sb, sf, sl, _POOL[pool],
                           grad.data_ptr(), gsb, C.byref(opt),
                           None if uk is None else uk.data_ptr(),
                           None if ug is None else ug.data_ptr(),
                           None if nu is None else nu.data_ptr(), C.byref(row_bits),
                           ws.data_ptr(), ws.numel(), L.current_stream_ptr())
    L.check(rc, "rtf_embed_bwd")
    if want_unique and not sync:
        return uk.to(torch.int64) & 0xFFFFFFFF, ug, row_bits.value, nu
    if want_unique:
        k = int(nu.item())
        keys = uk[:k].to(torch.int64) & 0xFFFFFFFF
        return keys, ug[:k], row_bits.value
    return None


class _LookupFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tset: "EmbeddingTables", ids, field_table, layout, pool, *weights):
        ctx.tset, ctx.ids, ctx.field_table, ctx.layout, ctx.pool = tset, ids, field_table, layout, pool
        tables = [weights[t] for t in field_table]
        return embed_fwd(tables, ids, layout, pool, err=tset.err)

    @staticmethod
    def backward(ctx, grad):
        grads = ctx.tset.grads_from_lookup_grad(ctx.ids, ctx.field_table, grad.contiguous(),
                                                ctx.layout, ctx.pool)
        return (None,) * 5 + grads


class EmbeddingTables(torch.nn.Module):
    """A set of embedding tables resident in HBM plus (optionally) the fused sparse optimizer
    state.  `lookup` is the drop-in for the reference's per-field Embedding + concat."""

    def __init__(self, rows: Sequence[int], dims: Sequence[int], initializer: str = "random_uniform",
                 device=None, optimizer: Optional[SparseOptimizer] = None, seed: Optional[int] = None):
        super().__init__()
        device = torch.device("cuda" if device is None else device)
        gen = None
        if seed is not None:
            gen = torch.Generator(device=device).manual_seed(seed)
        self.weights = torch.nn.ParameterList()
        for n, d in zip(rows, dims):
            # initialised in place in HBM (a Criteo-sized table set is tens of GB)
            w = torch.empty((int(n), int(d)), dtype=torch.float32, device=device)
            if initializer == "random_uniform":      # Keras: U(-0.05, 0.05)
                w.uniform_(-0.05, 0.05, generator=gen)
            elif initializer == "random_normal":     # Keras: N(0, 0.05^2)
                w.normal_(0.0, 0.05, generator=gen)
            elif initializer == "zeros":
                w.zero_()
            else:
                raise ValueError(f"unknown initializer {initializer!r}")
            self.weights.append(torch.nn.Parameter(w))
        self.register_buffer("err", torch.zeros(1, dtype=torch.int32, device=device))
        self.optimizer = None
        self.state1: List[Optional[torch.Tensor]] = []
        self.state2: List[Optional[torch.Tensor]] = []
        if optimizer is not None:
            self.set_optimizer(optimizer)

    @classmethod
    def from_tensors(cls, weights: Sequence[torch.Tensor], optimizer: Optional[SparseOptimizer] = None):
        """Wrap existing (rows, dim) fp32 device tensors — e.g. shards living in NVLink
        peer-mapped (symmetric) memory — without allocating or initialising anything."""
        self = cls.__new__(cls)
        torch.nn.Module.__init__(self)
        self.weights = torch.nn.ParameterList([torch.nn.Parameter(w, requires_grad=False) for w in weights])
        self.register_buffer("err", torch.zeros(1, dtype=torch.int32, device=weights[0].device))
        self.optimizer = None
        self.state1, self.state2 = [], []
        if optimizer is not None:
            self.set_optimizer(optimizer)
        return self

    def wlist(self):
        """The tables as a plain Python list (ParameterList indexing costs ~2 us per access and
        the step touches the list hundreds of times)."""
        wl = self.__dict__.get("_wl")
        if wl is None or len(wl) != len(self.weights):
            wl = self.__dict__["_wl"] = list(self.weights)
            self.__dict__["_rows"] = [int(w.shape[0]) for w in wl]
            self.__dict__["_dims"] = [int(w.shape[1]) for w in wl]
            self.__dict__["_rows_arr"] = L.host_array(C.c_int64, self._rows)
            self.__dict__["_dims_arr"] = L.host_array(C.c_int32, self._dims)
        return wl

    def set_optimizer(self, optimizer: Optional[SparseOptimizer]):
        self.optimizer = optimizer
        n = 0 if optimizer is None else optimizer.n_states
        self.state1 = [torch.zeros_like(w) if n >= 1 else None for w in self.weights]
        self.state2 = [torch.zeros_like(w) if n >= 2 else None for w in self.weights]

    def begin_step(self):
        """Advance the optimizer's step counter (call once per training step before backward)."""
        if self.optimizer is not None:
            self.optimizer.advance()

    def lookup(self, ids: torch.Tensor, field_table: Optional[Sequence[int]] = None,
               layout: str = "BF", pool: Optional[str] = None) -> torch.Tensor:
        self.wait_pending()
        wl = self.wlist()
        if field_table is None:
            field_table = list(range(len(wl)))
        return _LookupFn.apply(self, ids, tuple(field_table), layout, pool, *wl)

    # ---- K2 split: id-only half on a side stream, gradient half on the critical path
    def prepare_backward(self, ids, field_table, layout="BF"):
        """Launch keys + sort + segments for this batch on a side stream (they depend on the ids
        only) so they overlap the dense forward/backward; returns a handle for apply_prepared."""
        lib = L.lib()
        B, F, Lq, sb, sf, sl = ids_strides(ids, layout)
        self.wlist()
        rows, dims = self._rows, self._dims
        n = B * F * Lq
        wsz = self.__dict__.setdefault("_ws_bytes", {})
        if (n, max(dims)) not in wsz:
            nbytes = C.c_size_t(0)
            L.check(lib.rtf_embed_bwd_workspace(n, max(dims), C.byref(nbytes)), "rtf_embed_bwd_workspace")
            wsz[(n, max(dims))] = max(nbytes.value, 16)
        ws = torch.empty(wsz[(n, max(dims))], dtype=torch.uint8, device=ids.device)
        self.wait_pending()
        cur = torch.cuda.current_stream()
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream()
        side = self._side
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            rc = lib.rtf_embed_bwd_prepare(self._rows_arr, self._dims_arr,
                                           len(rows), L.host_array(C.c_int32, list(field_table)), F,
                                           ids.data_ptr(), int(ids.dtype == torch.int64), B, Lq, sb, sf,
                                           sl, None, None, ws.data_ptr(), ws.numel(), side.cuda_stream)
            L.check(rc, "rtf_embed_bwd_prepare")
            ev = torch.cuda.Event()
            ev.record(side)
        ws.record_stream(side)
        ids.record_stream(side)
        return {"ws": ws, "ev": ev, "B": B, "F": F, "L": Lq, "field_table": tuple(field_table)}

    def apply_prepared(self, h, grad, pool=None, reduce_only=None):
        """Gradient half of K2 for a batch prepared by prepare_backward.  reduce_only=(keys int32
        (cap,), sums (cap, dim_max)): no optimizer — the touched rows' keys and summed gradients
        are written there instead (cap >= number of touched rows; unused keys keep their value)."""
        lib = L.lib()
        opt = self.optimizer
        if reduce_only is not None:
            st = L.rtf_opt(L.OPT_NONE, 0.0, 0.0, 0.0, 0.0, 0.0, None)
        else:
            st = opt.struct_for_step(max(opt.step, 1))
        wl = self.wlist()
        rows = self._rows
        cur = torch.cuda.current_stream()
        cur.wait_event(h["ev"])
        h["ws"].record_stream(cur)      # the work list may be consumed on a stream other than the
        #                                 one it was allocated / prepared on (exchange stream)
        rc = lib.rtf_embed_bwd_apply(_ptr_array(wl), _ptr_array(self.state1),
                                     _ptr_array(self.state2), self._rows_arr,
                                     self._dims_arr, len(rows),
                                     L.host_array(C.c_int32, list(h["field_table"])), h["F"], h["B"],
                                     h["L"], _POOL[pool], grad.data_ptr(), grad.stride(0),
                                     C.byref(st),
                                     None if reduce_only is None else reduce_only[0].data_ptr(),
                                     None if reduce_only is None else reduce_only[1].data_ptr(),
                                     h["ws"].data_ptr(), h["ws"].numel(), L.current_stream_ptr())
        L.check(rc, "rtf_embed_bwd_apply")

    def apply_prepared_async(self, h, grad, pool=None):
        """apply_prepared on the side stream: the gradient half of K2 (HBM-bound) then overlaps
        whatever the caller enqueues next on the current stream (the bottom MLP's backward GEMMs,
        tensor-core-bound).  The tables are consistent again after wait_pending()."""
        cur = torch.cuda.current_stream()
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream()
        side = self._side
        side.wait_stream(cur)                    # grad is complete
        with torch.cuda.stream(side):
            self.apply_prepared(h, grad, pool)
            ev = torch.cuda.Event()
            ev.record(side)
        grad.record_stream(side)
        self._pending_ev = ev

    def wait_pending(self):
        """Make the current stream wait for an asynchronous row update (apply_prepared_async)."""
        ev = getattr(self, "_pending_ev", None)
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)
            self._pending_ev = None

    def apply_sparse_grad(self, ids, field_table, grad, layout="BF", pool=None):
        opt = self.optimizer
        step = max(opt.step, 1)
        embed_bwd([w.data for w in self.weights], field_table, ids, grad, layout, pool,
                  opt=opt.struct_for_step(step), state1=self.state1, state2=self.state2)

    def grads_from_lookup_grad(self, ids, field_table, grad, layout="BF", pool=None):
        """Backward of a lookup given d(out): with a fused optimizer the touched rows are
        updated in place (returns Nones); otherwise returns one sparse COO gradient per table."""
        nw = len(self.weights)
        if self.optimizer is not None:
            self.apply_sparse_grad(ids, field_table, grad, layout, pool)
            return (None,) * nw
        keys, g, row_bits = embed_bwd(list(self.weights), field_table, ids, grad, layout, pool,
                                      want_unique=True)
        tab = keys >> row_bits
        row = keys & ((1 << row_bits) - 1)
        grads = []
        for t, w in enumerate(self.weights):
            sel = tab == t
            grads.append(torch.sparse_coo_tensor(row[sel].unsqueeze(0), g[sel][:, : w.shape[1]],
                                                 size=w.shape, check_invariants=False,
                                                 is_coalesced=True))   # keys are unique and sorted
        return tuple(grads)

    def check_ids(self):
        """Raise like TF's CPU gather (InvalidArgument) if any lookup since the last check used an
        out-of-range id.  Reads a device flag, i.e. synchronises."""
        self.wait_pending()
        if int(self.err.item()) != 0:
            self.err.zero_()
            raise IndexError("embedding lookup: id out of range [0, input_dim)")

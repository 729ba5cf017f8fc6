"""C1 — multi-GPU embedding model parallelism for the DLRM path.

The reference only knows tf.distribute.MirroredStrategy: every table is replicated and, being
l2-regularised, its dense gradient is all-reduced every step (src/ctr/fm/train.py:43-50).
Here the dense MLPs stay data-parallel and the tables are placed per size; one process per GPU,
torch.distributed (NCCL over NVLink) for the plumbing.  Two implementations:

PeerShardedDLRM (default) — three placements chosen per table:
  row-wise     (>= 5 M rows)   row r on rank r % G at local row r // G
  table-wise   (the middle)    one owner, greedy balance of lookup count then bytes
  replicated   (<= 16 K rows)  a copy on every rank; gradients combined by one dense all-reduce
  forward   all-gather of the ids -> K1 on every holder over the global batch for its shards
            (foreign lookups of row-wise tables skipped) into a peer-mapped (B_global, T_g*D)
            buffer; replicated tables gathered locally -> one barrier ->
            K4 pulls the rows of its local samples from the holders' buffers with TMA bulk
            copies over NVLink, inside the interaction kernel (no NCCL all-to-all, no staging)
  backward  K4 bwd stores each dX row straight into the holder's gradient buffer -> barrier ->
            K2 + fused sparse Adam on the holders (keys/sort on a side stream since the forward);
            replicated tables: reduce-only K2 per rank + async dense all-reduce + identical row
            update on every replica; MLP gradients: one flat async all-reduce overlapping K2.
  gather='direct' instead pulls the random table rows themselves from the remote shards
  (no K1, no id exchange in the forward; 2.2x slower over NVLink at 8 GPUs).

ShardedDLRM — table-wise only, K1 on owners, then either K4 pulls the pooled rows from the
owners' output buffers (exchange='p2p') or an async NCCL all-to-all that overlaps the bottom MLP
(exchange='nccl'); K4 reads the receive buffer in place (rows addressed by base+stride, blocks
ordered by source rank); reverse exchange symmetric.

The placement / exchange-layout helpers are device-agnostic (gloo on CPU in tests); only
lookups and the interaction need CUDA.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

from . import _lib as L
from .core import DNN, Dense, DenseAdam, Layer, binary_crossentropy
from .embedding import EmbeddingTables, SparseOptimizer, embed_bwd, embed_fwd
from .dlrm import build_dense_layers
from .interaction import dot_out_cols


def plan_table_owners(rows: Sequence[int], dims: Sequence[int], world: int) -> List[int]:
    """Greedy table-wise placement: biggest tables first, each to the rank with the fewest
    tables (every table costs one lookup per sample), ties broken by resident bytes, then rank."""
    order = sorted(range(len(rows)), key=lambda t: (-rows[t] * dims[t], t))
    count, nbytes = [0] * world, [0] * world
    owner = [0] * len(rows)
    for t in order:
        g = min(range(world), key=lambda r: (count[r], nbytes[r], r))
        owner[t] = g
        count[g] += 1
        nbytes[g] += rows[t] * dims[t] * 4
    return owner


class ShardLayout:
    """Static description of who owns what and where a table's rows sit in the exchange
    buffers.  slots[g] = tables owned by rank g (ascending); in a receive buffer the block of
    source g is (B_local, T_g, D) at element offset B_local*D*sum(T_<g)."""

    def __init__(self, rows, dims, world, rank, owners=None):
        self.world, self.rank = world, rank
        self.owners = list(owners) if owners is not None else plan_table_owners(rows, dims, world)
        self.slots = [[t for t, o in enumerate(self.owners) if o == g] for g in range(world)]
        self.T = [len(s) for s in self.slots]
        self.slot_of = {}
        for g, s in enumerate(self.slots):
            for j, t in enumerate(s):
                self.slot_of[t] = (g, j)
        self.n_tables = len(rows)

    def fwd_splits(self, B_local: int, D: int):
        """(input_split, output_split) element counts of the forward all-to-all on this rank."""
        me = self.T[self.rank]
        return [B_local * me * D] * self.world, [B_local * t * D for t in self.T]

    def block_offsets(self, B_local: int, D: int):
        off, acc = [], 0
        for t in self.T:
            off.append(acc)
            acc += B_local * t * D
        return off, acc

    def row_location(self, table: int, B_local: int, D: int):
        """(element offset, element stride between samples) of `table`'s row for local sample 0
        inside a receive / reverse-send buffer."""
        g, j = self.slot_of[table]
        off, _ = self.block_offsets(B_local, D)
        return off[g] + j * D, self.T[g] * D


def all_ranks_ok(ok: bool, device=None, group=None) -> bool:
    """Collective agreement on a per-rank success flag (MIN over ranks): every rank must take the
    same branch when the alternative is a different collective."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return bool(ok)
    if device is None or dist.get_backend(group) == "gloo":
        device = torch.device("cpu")
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    return bool(int(flag.item()))


def exchange_ids(ids_local: torch.Tensor, world: int, group=None) -> torch.Tensor:
    out = ids_local.new_empty((ids_local.shape[0] * world,) + tuple(ids_local.shape[1:]))
    dist.all_gather_into_tensor(out, ids_local.contiguous(), group=group)
    return out


def exchange_rows_fwd(local_out: torch.Tensor, layout: ShardLayout, B_local: int, D: int,
                      async_op=False, group=None):
    """owner-major -> sample-owner: local_out (B_global, T_me*D) -> flat receive buffer of
    source-ordered blocks.  Returns (recv, work)."""
    ins, outs = layout.fwd_splits(B_local, D)
    recv = local_out.new_empty(sum(outs))
    work = dist.all_to_all_single(recv, local_out.reshape(-1), outs, ins, group=group,
                                  async_op=async_op)
    return recv, work


def exchange_rows_bwd(grad_send: torch.Tensor, layout: ShardLayout, B_local: int, D: int,
                      async_op=False, group=None):
    """reverse of exchange_rows_fwd: flat source-ordered blocks -> (B_global, T_me*D) on owners."""
    ins, outs = layout.fwd_splits(B_local, D)          # reversed roles
    recv = grad_send.new_empty(sum(ins))
    work = dist.all_to_all_single(recv, grad_send, ins, outs, group=group, async_op=async_op)
    return recv.view(B_local * layout.world, layout.T[layout.rank] * D), work


def _row_tables(layout: ShardLayout, recv: torch.Tensor, dense: torch.Tensor, B_local: int, D: int):
    """HOST arrays (base pointers, strides) of the F1 rows of a sample in ORIGINAL table order:
    row 0 = dense (bottom-MLP output), row 1+t = table t's row inside `recv`."""
    base = [dense.data_ptr()]
    stride = [dense.stride(0)]
    p0 = recv.data_ptr()
    for t in range(layout.n_tables):
        off, st = layout.row_location(t, B_local, D)
        base.append(p0 + off * 4)
        stride.append(st)
    return (C.c_void_p * len(base))(*base), (C.c_int64 * len(stride))(*stride)


class _ShardedInteractFn(torch.autograd.Function):
    """K4 over [dense | received rows]; backward launches the reverse all-to-all asynchronously
    and parks the handle on the model (K2 runs in ShardedDLRM.finish_backward)."""

    @staticmethod
    def forward(ctx, model: "ShardedDLRM", dense, recv, pad_to):
        dense = dense.contiguous()
        lay = model.layout
        B, D = dense.shape
        F1 = lay.n_tables + 1
        cols = dot_out_cols(F1, D, pad_to)
        out = torch.empty((B, cols), dtype=torch.float32, device=dense.device)
        rb, rs = _row_tables(lay, recv, dense, B, D)
        L.check(L.lib().rtf_dot_rows_fwd(rb, rs, F1, D, B, out.data_ptr(), cols, cols, None, 0,
                                         L.current_stream_ptr()), "rtf_dot_rows_fwd")
        ctx.model = model
        ctx.save_for_backward(dense, recv)
        return out

    @staticmethod
    def backward(ctx, gout):
        dense, recv = ctx.saved_tensors
        model = ctx.model
        lay = model.layout
        gout = gout.contiguous()
        B, D = dense.shape
        F1 = lay.n_tables + 1
        gdense = torch.empty_like(dense)
        gsend = torch.empty_like(recv)
        rb, rs = _row_tables(lay, recv, dense, B, D)
        gb, gs = _row_tables(lay, gsend, gdense, B, D)
        L.check(L.lib().rtf_dot_rows_bwd(rb, rs, F1, D, B, gout.data_ptr(), gout.stride(0), gb, gs,
                                         L.current_stream_ptr()), "rtf_dot_rows_bwd")
        grecv, work = exchange_rows_bwd(gsend, lay, B, D, async_op=True)
        model._pending = (grecv, work, gsend)
        return None, gdense, None, None


def _peer_row_tables(layout: ShardLayout, peer_ptrs, dense: torch.Tensor, B_local: int, D: int):
    """Row (base, stride) tables when the rows are read from / written to the OWNERS' buffers
    over NVLink: rank g's buffer is (B_global, T_g, D) sample-major, so table t (owner g, slot j)
    of my local sample b sits at peer_ptrs[g] + ((me*B_local + b)*T_g + j)*D."""
    me = layout.rank
    base = [dense.data_ptr()]
    stride = [dense.stride(0)]
    for t in range(layout.n_tables):
        g, j = layout.slot_of[t]
        Tg = layout.T[g]
        base.append(peer_ptrs[g] + (me * B_local * Tg + j) * D * 4)
        stride.append(Tg * D)
    return (C.c_void_p * len(base))(*base), (C.c_int64 * len(stride))(*stride)


class _PeerInteractFn(torch.autograd.Function):
    """K4 fused with the exchange over NVLink peer memory (no NCCL all-to-all): the forward's
    TMA bulk copies pull each embedding row straight from its owner's HBM, the backward re-pulls
    them and stores the dX rows straight into the owners' gradient buffers.  The transfer is
    overlapped with the Gram arithmetic row by row inside the one kernel."""

    @staticmethod
    def forward(ctx, model: "ShardedDLRM", dense, pad_to):
        dense = dense.contiguous()
        lay = model.layout
        B, D = dense.shape
        F1 = lay.n_tables + 1
        cols = dot_out_cols(F1, D, pad_to)
        out = torch.empty((B, cols), dtype=torch.float32, device=dense.device)
        rb, rs = _peer_row_tables(lay, model._out_hdl.buffer_ptrs, dense, B, D)
        xsave = torch.empty((B, (F1 - 1) * D), dtype=torch.float32, device=dense.device)
        L.check(L.lib().rtf_dot_rows_fwd(rb, rs, F1, D, B, out.data_ptr(), cols, cols,
                                         xsave.data_ptr(), xsave.stride(0),
                                         L.current_stream_ptr()), "rtf_dot_rows_fwd")
        ctx.model = model
        ctx.save_for_backward(dense, xsave)
        return out

    @staticmethod
    def backward(ctx, gout):
        dense, xsave = ctx.saved_tensors
        model = ctx.model
        lay = model.layout
        gout = gout.contiguous()
        B, D = dense.shape
        F1 = lay.n_tables + 1
        gdense = torch.empty_like(dense)
        # rows come back from the local copy the forward kept (no second trip over NVLink) ...
        base = [dense.data_ptr()] + [xsave.data_ptr() + t * D * 4 for t in range(F1 - 1)]
        stride = [dense.stride(0)] + [xsave.stride(0)] * (F1 - 1)
        rb, rs = (C.c_void_p * F1)(*base), (C.c_int64 * F1)(*stride)
        # ... and the dX rows are stored straight into their owners' gradient buffers
        gb, gs = _peer_row_tables(lay, model._grad_hdl.buffer_ptrs, gdense, B, D)
        L.check(L.lib().rtf_dot_rows_bwd(rb, rs, F1, D, B, gout.data_ptr(), gout.stride(0), gb, gs,
                                         L.current_stream_ptr()), "rtf_dot_rows_bwd")
        model._pending = "p2p"
        return None, gdense, None


class ShardedDLRM(Layer):
    """DLRM (same constructor as dlrm.DLRM) with table-wise sharded embeddings."""

    def __init__(self, feature_columns, bot_dnn_hidden_units=(64, 32, 16),
                 top_dnn_hidden_units=(128, 64), activation="relu", dnn_dropout=0.0, embed_reg=1e-4,
                 sparse_optimizer: Optional[SparseOptimizer] = None, pad_to: int = 1,
                 input_bn: bool = True, seed: Optional[int] = None, owners=None,
                 exchange: str = "p2p"):
        super().__init__()
        if exchange not in ("p2p", "nccl"):
            raise ValueError(exchange)
        self.exchange = exchange
        self._sym_B = None
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.dense_feature_columns, self.sparse_feature_columns = feature_columns
        rows = [f["feat_num"] for f in self.sparse_feature_columns]
        dims = [f["embed_dim"] for f in self.sparse_feature_columns]
        if len(set(dims)) != 1 or bot_dnn_hidden_units[-1] != dims[0]:
            raise ValueError("dot interaction needs equal embed_dim == bot_dnn_hidden_units[-1]")
        self.D, self.pad_to, self.embed_reg = dims[0], pad_to, embed_reg
        self.layout = ShardLayout(rows, dims, self.world, self.rank, owners)
        mine = self.layout.slots[self.rank]
        # per-table seeds so that a table's init does not depend on the sharding
        self.embed_layers = EmbeddingTables([rows[t] for t in mine], [dims[t] for t in mine],
                                            "random_uniform", optimizer=sparse_optimizer,
                                            seed=None if seed is None else seed + 1 + self.rank)
        if seed is not None:
            torch.manual_seed(seed)                 # identical MLP init on every rank
        self.bot_dnn = DNN(bot_dnn_hidden_units, activation, dnn_dropout, input_bn=input_bn)
        self.top_dnn = DNN(top_dnn_hidden_units, activation, dnn_dropout, input_bn=input_bn)
        self.final_dense = Dense(1, activation=None)
        self._pending = None
        self._saved = None
        self._out_inflight = False      # peers may still be pulling the previous forward's rows
        self.register_buffer("_mine_idx", torch.as_tensor(mine, dtype=torch.int64,
                                                          device=self.embed_layers.err.device))

    def call(self, inputs, **kwargs):
        dense_inputs, sparse_inputs = inputs
        lay, D = self.layout, self.D
        B_local = sparse_inputs.shape[0]
        ids_global = exchange_ids(sparse_inputs, self.world)                       # (B_global, F)
        local_ids = ids_global.index_select(1, self._mine_idx).contiguous()        # (B_global, T_me)
        self._saved = local_ids
        if self.exchange == "p2p":
            ok = True
            try:
                self._ensure_symmetric(B_local)
            except Exception as e:  # no peer access on this box: use the NCCL all-to-all instead
                import warnings
                warnings.warn(f"symmetric memory unavailable ({e}); row exchange falls back to NCCL")
                ok = False
            if not all_ranks_ok(ok, sparse_inputs.device):   # the choice of exchange is collective
                self.exchange = "nccl"
        if self.exchange == "p2p":
            return self._call_p2p(dense_inputs, local_ids, B_local)
        with torch.no_grad():
            local_out = embed_fwd(list(self.embed_layers.weights), local_ids, "BF", None,
                                  err=self.embed_layers.err)
        recv, work = exchange_rows_fwd(local_out, lay, B_local, D, async_op=True)
        dense_fea = self.bot_dnn(dense_inputs)            # overlaps the all-to-all
        work.wait()
        x = _ShardedInteractFn.apply(self, dense_fea, recv, self.pad_to)
        x = _TopGradsReady.apply(x, self)
        return torch.sigmoid(self.final_dense(self.top_dnn(x)))

    # ---- exchange over NVLink peer memory (torch symmetric memory for the address exchange)
    def _ensure_symmetric(self, B_local: int):
        if self._sym_B == B_local:
            return
        import torch.distributed._symmetric_memory as symm
        if self._sym_B is not None:     # peers may still address the old buffers: quiesce first
            torch.cuda.synchronize()
            dist.barrier()
            self._out_inflight = False
        n = B_local * self.world * max(self.layout.T) * self.D       # same size on every rank
        dev = self.embed_layers.err.device
        self._out_buf = symm.empty(n, dtype=torch.float32, device=dev)
        self._grad_buf = symm.empty(n, dtype=torch.float32, device=dev)
        self._out_hdl = symm.rendezvous(self._out_buf, dist.group.WORLD)
        self._grad_hdl = symm.rendezvous(self._grad_buf, dist.group.WORLD)
        self._sym_B = B_local

    def _call_p2p(self, dense_inputs, local_ids, B_local):
        self._ensure_symmetric(B_local)
        D, Tme = self.D, self.layout.T[self.rank]
        Bg = B_local * self.world
        out_view = self._out_buf[: Bg * Tme * D].view(Bg, Tme * D)
        if self._out_inflight:      # no backward (grad barrier) since the last forward: peers may
            self._out_hdl.barrier(channel=1)          # still be reading the previous batch's rows
        with torch.no_grad():
            embed_fwd(list(self.embed_layers.weights), local_ids, "BF", None,
                      err=self.embed_layers.err, out=out_view)
        dense_fea = self.bot_dnn(dense_inputs)
        self._out_hdl.barrier(channel=0)      # every owner's rows are in place (stream-ordered)
        x = _PeerInteractFn.apply(self, dense_fea, self.pad_to)
        self._out_inflight = True
        x = _TopGradsReady.apply(x, self)
        return torch.sigmoid(self.final_dense(self.top_dnn(x)))

    def finish_backward(self):
        """Wait for the reverse exchange and run K2 (+ fused sparse optimizer) on the owner."""
        if self._pending is None:
            return
        if self._pending == "p2p":
            self._grad_hdl.barrier(channel=0)  # all peers' dX rows have landed in my buffer
            self._out_inflight = False         # ... so every peer is past its K4 forward as well
            D, Tme = self.D, self.layout.T[self.rank]
            Bg = self._sym_B * self.world
            grecv = self._grad_buf[: Bg * Tme * D].view(Bg, Tme * D)
            self.embed_layers.apply_sparse_grad(self._saved,
                                                list(range(len(self.embed_layers.weights))), grecv)
            self._pending = None
            return
        grecv, work, _keep = self._pending
        work.wait()
        self.embed_layers.apply_sparse_grad(self._saved, list(range(len(self.embed_layers.weights))),
                                            grecv)
        self._pending = None

    def dense_parameters(self):
        emb = {id(p) for p in self.embed_layers.parameters()}
        return [p for p in self.parameters() if id(p) not in emb]


class _TopGradsReady(torch.autograd.Function):
    """Identity on the interaction output.  Its backward runs when every gradient of the top MLP
    exists (autograd accumulates a layer's weight gradients before it walks further down), so the
    trainer's hook starts their all-reduce there — under K4's backward, the reverse exchange and
    the bottom MLP's backward instead of after them."""

    @staticmethod
    def forward(ctx, x, owner):
        ctx.owner = owner
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        cb = getattr(ctx.owner, "_top_grads_hook", None)
        if cb is not None:
            cb()
        return g, None


class ShardedDLRMTrainer:
    """Per-rank step: local loss / world (so that summed gradients equal the global-batch mean,
    as MirroredStrategy scales them, App. A18), reverse exchange + K2 on owners, one flat
    all-reduce of the MLP gradients, dense Adam (Keras form, core.DenseAdam).

    The dense replicas are made identical BEFORE the first training forward: the MLP layers are
    built on a dummy batch and rank 0's parameters and BatchNorm buffers are broadcast, so no
    rank ever back-propagates activations computed with weights that were later overwritten."""

    def __init__(self, model: ShardedDLRM, lr: float = 1e-3):
        self.model, self.lr = model, lr
        if model.embed_layers.optimizer is None:
            model.embed_layers.set_optimizer(SparseOptimizer("adam", lr=lr, l2=model.embed_reg))
        self.dense_opt = None

    def _setup(self, dense):
        m = self.model
        build_dense_layers(m, dense.shape[1], dense.device, m.layout.n_tables)
        params = m.dense_parameters()
        emb = {id(b) for b in m.embed_layers.buffers()}
        bufs = [b for n, b in m.named_buffers()
                if id(b) not in emb and b.is_floating_point() and not n.startswith("_")]
        for t in params + bufs:                            # same start on every rank
            dist.broadcast(t.data, 0)
        self.dense_opt = DenseAdam(params, lr=self.lr)
        # flat-buffer offset where the top MLP's (and the output layer's) gradients start: they are
        # all-reduced as soon as they exist (_TopGradsReady), the bottom MLP's after the backward
        top = {id(p) for p in list(m.top_dnn.parameters()) + list(m.final_dense.parameters())}
        first = [i for i, p in enumerate(self.dense_opt.params) if id(p) in top]
        tail = all(id(p) in top for p in self.dense_opt.params[first[0]:]) if first else False
        self._top_off = self.dense_opt.offsets[first[0]] if tail and first[0] > 0 else None
        if hasattr(m, "async_update"):
            m.async_update = True       # step() always ends with finish_backward()

    def step(self, dense, sparse, labels, next_sparse=None) -> torch.Tensor:
        """next_sparse: the NEXT batch's sparse ids, if the caller already has them (a prefetching
        feeder does): the embedding exchange of batch n+1 is then pipelined behind step n
        (PeerShardedDLRM.prefetch) and leaves the critical path."""
        m = self.model
        if self.dense_opt is None:
            self._setup(dense)
        m.embed_layers.begin_step()
        pred = m([dense, sparse])
        loss = binary_crossentropy(labels, pred)
        self.dense_opt.zero_grad()
        # all-reduce of the persistent flat MLP-gradient buffer (no concat / copy-back) in two
        # pieces: the top MLP's slice from inside the backward, the moment it is complete, the
        # rest after it; both asynchronous, overlapping K4's backward, the bottom MLP's backward,
        # the reverse exchange barrier and K2 on the embedding shards (finish_backward)
        fg, works = self.dense_opt.flat_grad, []
        if self._top_off is not None:
            m._top_grads_hook = lambda: works.append(dist.all_reduce(fg[self._top_off:], async_op=True))
        try:
            (loss / m.world).backward()
        finally:
            m._top_grads_hook = None
        work = dist.all_reduce(fg[:self._top_off] if works else fg, async_op=True)
        pipelined = next_sparse is not None and hasattr(m, "prefetch") and m.async_update
        if pipelined:
            m.prefetch(next_sparse)
        elif not getattr(m, "async_update", False):
            m.finish_backward()
        for w in works:
            w.wait()
        work.wait()
        self.dense_opt.step()
        if not pipelined:
            m.finish_backward()         # async: the row update ran behind the MLP backward / Adam
        return loss.detach()


# =============================================================================================
# Table-wise AND row-wise sharding with the exchange fused into K4 (exchange="peer")
# =============================================================================================
class PeerLayout:
    """Placement for the peer-memory path: tables with >= `row_wise_min_rows` rows are split
    row-wise over all ranks (row r lives on rank r % G at local row r // G — every rank then
    serves ~1/G of that table's lookups whatever the id distribution), the rest are placed
    table-wise by `plan_table_owners`.  fields[g] = tables with a shard on rank g (ascending);
    a rank's gradient buffer is (B_global, len(fields[g]) * D), column block j for fields[g][j].
    Tables with <= `replicate_max_rows` rows are REPLICATED on every rank (data-parallel, like the
    MLP weights): their lookups never cross NVLink, their gradients are summed by one dense
    all-reduce; they come last in fields[g] (rep_fields), after the sharded ones.
    Pure host logic (tested on CPU)."""

    def __init__(self, rows: Sequence[int], dims: Sequence[int], world: int,
                 row_wise_min_rows: int = 5_000_000, row_wise: Optional[Sequence[bool]] = None,
                 owners: Optional[Sequence[int]] = None, replicate_max_rows: int = 0):
        n = len(rows)
        if n > 64:
            raise ValueError("at most 64 sparse fields (row-wise mask is 64 bits)")
        self.world, self.rows, self.dims, self.n_tables = world, list(rows), list(dims), n
        self.replicated = [world > 1 and rows[t] <= replicate_max_rows for t in range(n)]
        if row_wise is None:
            row_wise = [world > 1 and rows[t] >= max(row_wise_min_rows, world) for t in range(n)]
        self.row_wise = [bool(x) and not self.replicated[t] for t, x in enumerate(row_wise)]
        for t in range(n):
            if self.row_wise[t] and rows[t] < world:
                raise ValueError(f"table {t} has {rows[t]} rows: too few to split over {world} ranks")
        tw = [t for t in range(n) if not self.row_wise[t] and not self.replicated[t]]
        if owners is None:
            own_tw = plan_table_owners([rows[t] for t in tw], [dims[t] for t in tw], world)
            owners = [-1] * n
            for t, g in zip(tw, own_tw):
                owners[t] = g
        self.owners = [(-1 if (self.row_wise[t] or self.replicated[t]) else int(owners[t]))
                       for t in range(n)]
        self.rw_mask = sum(1 << t for t in range(n) if self.row_wise[t])
        self.rep_fields = [t for t in range(n) if self.replicated[t]]
        self.shard_fields = [[t for t in range(n) if self.row_wise[t] or
                              (self.owners[t] == g and not self.replicated[t])] for g in range(world)]
        self.fields = [sf + self.rep_fields for sf in self.shard_fields]
        self.slot = [{t: j for j, t in enumerate(f)} for f in self.fields]

    def local_rows(self, g: int, t: int) -> int:
        if not self.row_wise[t]:
            return self.rows[t]
        return (self.rows[t] - g + self.world - 1) // self.world      # rows g, g+G, g+2G, ...

    def shard_offsets(self, g: int):
        """element offsets of rank g's shards inside its table buffer, and the total."""
        off, acc = {}, 0
        for t in self.fields[g]:
            off[t] = acc
            acc += self.local_rows(g, t) * self.dims[t]
        return off, acc

    def buffer_elems(self) -> int:
        return max(self.shard_offsets(g)[1] for g in range(self.world))   # same size everywhere

    def max_fields(self) -> int:
        return max(len(f) for f in self.fields)

    def holder(self, t: int, row: int):
        """(rank, local row) that stores global row `row` of table t."""
        if self.row_wise[t]:
            return row % self.world, row // self.world
        return self.owners[t], row

    def peer_pointer_tables(self, tab_ptrs: Sequence[int], grad_ptrs: Sequence[int], D: int,
                            rank: int = 0):
        """HOST lists [n_tables][G] for the kernels: table-shard base addresses, gradient-column
        base addresses and gradient sample strides (elements).  Table-wise fields use entry 0;
        a replicated field points at `rank`'s own copy / buffer."""
        G = self.world
        tab = [[0] * G for _ in range(self.n_tables)]
        gptr = [[0] * G for _ in range(self.n_tables)]
        gstr = [[0] * G for _ in range(self.n_tables)]
        offs = [self.shard_offsets(g)[0] for g in range(G)]
        for t in range(self.n_tables):
            ranks = range(G) if self.row_wise[t] else [rank if self.replicated[t] else self.owners[t]]
            for e, g in enumerate(ranks):
                tab[t][e] = tab_ptrs[g] + offs[g][t] * 4
                gptr[t][e] = grad_ptrs[g] + self.slot[g][t] * D * 4
                gstr[t][e] = len(self.fields[g]) * D
        return tab, gptr, gstr


def local_shard_ids(ids_global: torch.Tensor, layout: PeerLayout, rank: int, cache: Optional[dict] = None
                    ) -> torch.Tensor:
    """(B_global, n_tables) global ids -> (B_global, len(shard_fields[rank])) ids into this rank's
    shards; lookups of a row-wise table that another rank holds become -1 (K2 skips them).
    Replicated tables are not part of it: each rank handles them for its own samples.
    No host synchronisation: which fields are row-wise is host knowledge; the small index / mask
    tensors are built once and kept in `cache` (a per-step `.any()` on the device would stall the
    host every step and with it the CPU run-ahead that keeps the GPU fed)."""
    f = layout.shard_fields[rank]
    key = (rank, str(ids_global.device))
    if cache is not None and key in cache:
        idx, rw = cache[key]
    else:
        idx = torch.as_tensor(f, dtype=torch.int64, device=ids_global.device)
        rw = torch.as_tensor([layout.row_wise[t] for t in f], dtype=torch.bool,
                             device=ids_global.device)
        if cache is not None:
            cache[key] = (idx, rw)
    loc = ids_global.index_select(1, idx)
    if any(layout.row_wise[t] for t in f):
        G = layout.world
        mine = (loc % G) == rank
        loc = torch.where(rw.view(1, -1), torch.where(mine & (loc >= 0),
                                                     torch.div(loc, G, rounding_mode="floor"),
                                                     torch.full_like(loc, -1)), loc)
    return loc.contiguous()


def scatter_unique_rows(keys: torch.Tensor, sums: torch.Tensor, row_bits: int, first_table: int,
                        rep_off: torch.Tensor, R: int, D: int) -> torch.Tensor:
    """K2's unique-row output of the replicated tables -> dense block.  keys (cap,) int64 =
    table << row_bits | row (unused slots 0xFFFFFFFF), sums (cap, >= D); tables first_table ..
    first_table + Tr - 1 are the replicated ones, rep_off (Tr + 1,) their row offsets in the
    block (last entry = total R, also passed as a host int: no device sync here).  Returns G (R + 1, D + 1): summed gradient per row, column D =
    1 where the row was touched on this rank; row R collects the unused slots."""
    Tr = rep_off.numel() - 1
    tab = (keys >> row_bits) - first_table
    tab = torch.where(tab < 0, torch.full_like(tab, Tr), tab)
    ok = tab < Tr
    idx = torch.where(ok, rep_off[tab.clamp(max=Tr)] + (keys & ((1 << row_bits) - 1)),
                      torch.full_like(keys, R))
    G = torch.zeros((R + 1, D + 1), dtype=torch.float32, device=keys.device)
    G[:, :D].index_copy_(0, idx, sums[:, :D])                 # unique rows (duplicates only at R)
    G[:, D].index_fill_(0, idx, 1.0)
    return G


def apply_touched_rows(opt: SparseOptimizer, lr_t: float, W: torch.Tensor, m: Optional[torch.Tensor],
                       v: Optional[torch.Tensor], G: torch.Tensor) -> None:
    """K2's row update (csrc/embed_bwd.cu finish_row) on a dense block, in place, for the rows
    whose touched count G[:, D] is > 0: g = sum + 2*l2*w, then SGD / Adagrad / Adam (Keras form,
    lr_t already carries the bias corrections).  Untouched rows and their state do not move."""
    R, D = W.shape
    touched = (G[:R, D] > 0).unsqueeze(1)
    g = G[:R, :D]
    if opt.l2 > 0:
        g = g + (2.0 * opt.l2) * W
    if opt.kind == "sgd":
        W.copy_(torch.where(touched, W - lr_t * g, W))
    elif opt.kind == "adagrad":
        a = m + g * g
        W.copy_(torch.where(touched, W - (lr_t * g) / (a.sqrt() + opt.eps), W))
        m.copy_(torch.where(touched, a, m))
    elif opt.kind == "adam":
        m2 = opt.beta1 * m + (1.0 - opt.beta1) * g
        v2 = opt.beta2 * v + (1.0 - opt.beta2) * (g * g)
        W.copy_(torch.where(touched, W - (lr_t * m2) / (v2.sqrt() + opt.eps), W))
        m.copy_(torch.where(touched, m2, m))
        v.copy_(torch.where(touched, v2, v))
    else:
        raise ValueError(opt.kind)


class _PeerDotFn(torch.autograd.Function):
    """K1+K4 over tables sharded across the GPUs of the box: the forward's TMA bulk copies pull
    every embedding row from whichever GPU holds it (NVLink peer memory) and keep a local copy
    for the backward; the backward stores each dX row straight into the gradient buffer of the
    rank that owns that row.  No NCCL all-to-all, no staging buffer."""

    @staticmethod
    def forward(ctx, model: "PeerShardedDLRM", ids, dense, pad_to):
        L.require_cuda(ids, "peer embed_dot(ids)")
        dense = dense.contiguous()
        lay = model.layout
        B, F = ids.shape
        D = dense.shape[1]
        cols = dot_out_cols(F + 1, D, pad_to)
        out = torch.empty((B, cols), dtype=torch.float32, device=dense.device)
        owner = model.gather == "owner"
        if owner:
            # the rows of my samples were pulled into model._xsave by rtf_peer_pull_rows on the
            # exchange stream (behind the bottom MLP): the interaction reads local HBM only
            xsave = model._xsave
            base = [dense.data_ptr()] + [xsave.data_ptr() + t * D * 4 for t in range(F)]
            stride = [dense.stride(0)] + [xsave.stride(0)] * F
            rc = L.lib().rtf_dot_rows_fwd((C.c_void_p * (F + 1))(*base), (C.c_int64 * (F + 1))(*stride),
                                          F + 1, D, B, out.data_ptr(), cols, cols, None, 0,
                                          L.current_stream_ptr())
            L.check(rc, "rtf_dot_rows_fwd")
        else:
            xsave = torch.empty((B, F * D), dtype=torch.float32, device=dense.device)
            rc = L.lib().rtf_embed_dot_peer_fwd(
                model._d_peer_tab.data_ptr(), None, model.rank * B,
                lay.world, lay.rw_mask, model._rows_arr, F, D,
                ids.data_ptr(), int(ids.dtype == torch.int64), B, ids.stride(0), ids.stride(1),
                dense.data_ptr(), dense.stride(0), out.data_ptr(), cols, cols, xsave.data_ptr(),
                xsave.stride(0), model.embed_layers.err.data_ptr(), L.current_stream_ptr())
            L.check(rc, "rtf_embed_dot_peer_fwd")
        ctx.model, ctx.ids = model, ids
        ctx.save_for_backward(dense, xsave)
        return out

    @staticmethod
    def backward(ctx, gout):
        dense, xsave = ctx.saved_tensors
        model, ids = ctx.model, ctx.ids
        lay = model.layout
        gout = gout.contiguous()
        B, F = ids.shape
        D = dense.shape[1]
        gdense = torch.empty_like(dense)
        rc = L.lib().rtf_embed_dot_peer_bwd(
            xsave.data_ptr(), xsave.stride(0), model._rows_arr, F, D, ids.data_ptr(),
            int(ids.dtype == torch.int64), B, ids.stride(0), ids.stride(1), dense.data_ptr(),
            dense.stride(0), gout.data_ptr(), gout.stride(0), gdense.data_ptr(), gdense.stride(0),
            model._d_peer_gptr.data_ptr(), model._d_peer_gstr.data_ptr(), lay.world, lay.rw_mask,
            model.rank * B, L.current_stream_ptr())
        L.check(rc, "rtf_embed_dot_peer_bwd")
        model._pending = True
        if model.async_update:      # barrier + K2 on the exchange stream, behind the bottom MLP's
            model._launch_update()  # backward; the trainer's finish_backward() only waits
        return None, None, gdense, None


class PeerShardedDLRM(Layer):
    """DLRM (constructor of dlrm.DLRM) with embedding tables sharded table-wise and row-wise over
    the GPUs of one NVSwitch box; dense MLPs data-parallel.  Every rank keeps its shards in
    peer-mapped (symmetric) memory and the fused gather+interaction kernels address remote rows
    directly, so the "all-to-all of pooled embeddings" happens row by row inside K4, overlapped
    with its arithmetic.  Drives the same ShardedDLRMTrainer as ShardedDLRM."""

    exchange = "peer"

    def __init__(self, feature_columns, bot_dnn_hidden_units=(64, 32, 16),
                 top_dnn_hidden_units=(128, 64), activation="relu", dnn_dropout=0.0, embed_reg=1e-4,
                 sparse_optimizer: Optional[SparseOptimizer] = None, pad_to: int = 1,
                 input_bn: bool = True, seed: Optional[int] = None,
                 row_wise_min_rows: int = 5_000_000, row_wise=None, owners=None,
                 gather: str = "owner", replicate_max_rows: int = 0):
        """replicate_max_rows: tables with at most this many rows are replicated on every rank
        (data-parallel): each rank gathers them locally for its own samples, and their row
        gradients — segment-summed per rank by K2 — are combined by ONE dense all-reduce before the
        identical Adam update of the touched rows on every replica.  With the Criteo cardinalities
        and 16 384, 18 of the 26 lookups per sample stop crossing NVLink for 24 MB of all-reduce.
        gather='owner' (default): every holder runs K1 for the global batch over its shards
        (a lookup of a row-wise table on the rank that does not hold the row is skipped) into a
        peer-mapped (B_global, T_g*D) buffer, and K4 pulls each row BY SAMPLE from the holder's
        buffer — sequential addresses, measured 2.2x faster over NVLink at 8 GPUs than
        gather='direct', where K4 pulls the random table rows themselves from the remote shards
        (no K1, no buffer, no id exchange in the forward)."""
        super().__init__()
        if gather not in ("owner", "direct"):
            raise ValueError(gather)
        self.gather = gather
        import torch.distributed._symmetric_memory as symm
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.dense_feature_columns, self.sparse_feature_columns = feature_columns
        rows = [f["feat_num"] for f in self.sparse_feature_columns]
        dims = [f["embed_dim"] for f in self.sparse_feature_columns]
        if len(set(dims)) != 1 or bot_dnn_hidden_units[-1] != dims[0]:
            raise ValueError("dot interaction needs equal embed_dim == bot_dnn_hidden_units[-1]")
        self.D, self.pad_to, self.embed_reg = dims[0], pad_to, embed_reg
        self.layout = lay = PeerLayout(rows, dims, self.world, row_wise_min_rows, row_wise, owners,
                                       replicate_max_rows)
        dev = torch.device("cuda", torch.cuda.current_device())
        # this rank's shards: views into ONE symmetric allocation every peer can address
        self._tab_buf = symm.empty(lay.buffer_elems(), dtype=torch.float32, device=dev)
        self._tab_hdl = symm.rendezvous(self._tab_buf, dist.group.WORLD)
        offs, _ = lay.shard_offsets(self.rank)
        gen = None if seed is None else torch.Generator(device=dev).manual_seed(seed + 1 + self.rank)
        shards = []
        for t in lay.fields[self.rank]:
            n = lay.local_rows(self.rank, t)
            w = self._tab_buf[offs[t]: offs[t] + n * dims[t]].view(n, dims[t])
            w.uniform_(-0.05, 0.05, generator=gen)          # Keras 'random_uniform'
            if lay.replicated[t]:
                dist.broadcast(w, 0)                        # replicas start identical
            shards.append(w)
        self.embed_layers = EmbeddingTables.from_tensors(shards, optimizer=sparse_optimizer)
        if seed is not None:
            torch.manual_seed(seed)                 # identical MLP init on every rank
        self.bot_dnn = DNN(bot_dnn_hidden_units, activation, dnn_dropout, input_bn=input_bn)
        self.top_dnn = DNN(top_dnn_hidden_units, activation, dnn_dropout, input_bn=input_bn)
        self.final_dense = Dense(1, activation=None)
        self._rows_arr = L.host_array(C.c_int64, rows)
        self._grad_B = None
        self._pending = False
        self._prepared = None
        self._out_inflight = False      # peers may still be pulling the previous forward's rows
        self.async_update = False       # set by ShardedDLRMTrainer: K2 behind the MLP backward
        self._xstream = None            # exchange stream: K1 + barrier + row pull, barrier + K2
        self._loc_cache = {}
        self._prefetched = None         # exchange of a batch announced with prefetch()
        self._upd_ev = None
        # replicated block: the shards of rep_fields are contiguous at the end of the table buffer
        self._Ts, self._Tr = len(lay.shard_fields[self.rank]), len(lay.rep_fields)
        self._rep_rows = [rows[t] for t in lay.rep_fields]
        self._rep_state_ready = False
        if self._Tr:
            dev_ = self._tab_buf.device
            self.register_buffer("_rep_idx", torch.as_tensor(lay.rep_fields, dtype=torch.int64, device=dev_))
            off = [0]
            for r in self._rep_rows:
                off.append(off[-1] + r)
            self._rep_total = off[-1]
            # row offset of replicated table j in the block; extra entries send bad keys to the dummy row
            self.register_buffer("_rep_off", torch.as_tensor(off[:-1] + [off[-1]], dtype=torch.int64, device=dev_))
            o0 = offs[lay.rep_fields[0]]
            self._rep_W = self._tab_buf[o0: o0 + self._rep_total * self.D].view(self._rep_total, self.D)

    def _ensure_grad_buffer(self, B_local: int):
        if self._grad_B == B_local:
            return
        import torch.distributed._symmetric_memory as symm
        if self._grad_B is not None:    # peers may still address the old buffers: quiesce first
            torch.cuda.synchronize()
            dist.barrier()
            self._out_inflight = False
        lay, D = self.layout, self.D
        dev = self._tab_buf.device
        n = B_local * self.world * lay.max_fields() * D
        self._grad_buf = symm.empty(n, dtype=torch.float32, device=dev)
        self._grad_hdl = symm.rendezvous(self._grad_buf, dist.group.WORLD)
        tab, gptr, gstr = lay.peer_pointer_tables(self._tab_hdl.buffer_ptrs,
                                                  self._grad_hdl.buffer_ptrs, D, self.rank)
        self._d_peer_tab = torch.tensor(tab, dtype=torch.int64, device=dev)
        self._d_peer_gptr = torch.tensor(gptr, dtype=torch.int64, device=dev)
        self._d_peer_gstr = torch.tensor(gstr, dtype=torch.int64, device=dev)
        if self.gather == "owner":      # K1 output of every holder, same layout as the gradients
            self._out_buf = symm.empty(n, dtype=torch.float32, device=dev)
            self._out_hdl = symm.rendezvous(self._out_buf, dist.group.WORLD)
            _, optr, ostr = lay.peer_pointer_tables(self._tab_hdl.buffer_ptrs,
                                                    self._out_hdl.buffer_ptrs, D, self.rank)
            self._d_peer_optr = torch.tensor(optr, dtype=torch.int64, device=dev)
            self._d_peer_ostr = torch.tensor(ostr, dtype=torch.int64, device=dev)
            self._xsave = torch.empty((B_local, lay.n_tables * D), dtype=torch.float32, device=dev)
        self._grad_B = B_local

    def _exchange(self, sparse_inputs, train: bool):
        """Everything of a batch that depends on its ids only: the id all-gather, K2's keys / sort /
        segments (side stream) and — owner gather — K1 over my shards for the GLOBAL batch, the
        barrier and the NVLink pull of my samples' rows into the local staging buffer, all on the
        exchange stream.  Returns what the interaction / the backward need."""
        B_local = sparse_inputs.shape[0]
        self._ensure_grad_buffer(B_local)
        owner = self.gather == "owner"
        ex = {"key": (sparse_inputs.data_ptr(), tuple(sparse_inputs.shape), sparse_inputs._version),
              "prepared": None, "prepared_rep": None, "ids_rep": None, "pulled": None, "train": train}
        loc = None
        if train or owner:
            # the holders need the ids of the global batch for their shards (104 B/sample): K1 in
            # owner mode, and K2 — whose keys, sort and segments start now on a side stream
            ids_global = exchange_ids(sparse_inputs, self.world)
            loc = local_shard_ids(ids_global, self.layout, self.rank, self._loc_cache)
            if train and loc.shape[1]:
                ex["prepared"] = self.embed_layers.prepare_backward(loc, list(range(loc.shape[1])))
        ids_rep = None
        if self._Tr:
            ids_rep = sparse_inputs.index_select(1, self._rep_idx).contiguous()   # my samples only
            ex["ids_rep"] = ids_rep
            if train:   # keys / sort / segments of the replicated lookups: side stream as well
                ex["prepared_rep"] = self.embed_layers.prepare_backward(
                    ids_rep, [self._Ts + j for j in range(self._Tr)])
        if owner:
            # K1 over my shards for the GLOBAL batch -> barrier -> pull of my samples' rows, all on
            # the exchange stream: it overlaps the bottom MLP (tensor-core GEMMs) or, when the batch
            # was announced with prefetch(), the tail of the previous step
            cur = torch.cuda.current_stream()
            if self._xstream is None:
                self._xstream = torch.cuda.Stream()
            xs = self._xstream
            xs.wait_stream(cur)
            Tme = self._Ts + self._Tr
            Bg = B_local * self.world
            with torch.cuda.stream(xs):
                if self._out_inflight:   # no grad barrier since the last forward (eval, or a forward
                    self._out_hdl.barrier(channel=1)     # never followed by finish_backward): peers may
                    #                                      still be reading the previous batch's rows
                out_view = self._out_buf[: Bg * Tme * self.D].view(Bg, Tme * self.D)
                W = list(self.embed_layers.weights)
                with torch.no_grad():
                    if self._Ts:    # foreign lookups of row-wise tables (-1) are skipped silently
                        embed_fwd(W[: self._Ts], loc, "BF", None, err=None, out=out_view, skip_invalid=True)
                    if self._Tr:    # replicated tables: local gather for my own samples
                        mine = out_view[self.rank * B_local: (self.rank + 1) * B_local, self._Ts * self.D:]
                        embed_fwd(W[self._Ts:], ids_rep, "BF", None, err=self.embed_layers.err, out=mine)
                self._out_hdl.barrier(channel=0)     # every holder's rows are in place
                F = self.layout.n_tables
                rc = L.lib().rtf_peer_pull_rows(
                    self._d_peer_optr.data_ptr(), self._d_peer_ostr.data_ptr(), self.rank * B_local,
                    self.world, self.layout.rw_mask, self._rows_arr, F, self.D, sparse_inputs.data_ptr(),
                    int(sparse_inputs.dtype == torch.int64), B_local, sparse_inputs.stride(0),
                    sparse_inputs.stride(1), self._xsave.data_ptr(), self._xsave.stride(0),
                    self.embed_layers.err.data_ptr(), xs.cuda_stream)
                L.check(rc, "rtf_peer_pull_rows")
                ex["pulled"] = torch.cuda.Event()
                ex["pulled"].record(xs)
            for t in (loc, ids_rep, sparse_inputs):
                if t is not None:
                    t.record_stream(xs)
            self._out_inflight = True
        return ex

    def prefetch(self, sparse_next) -> None:
        """Announce the NEXT batch's sparse ids (call it after backward()): its exchange is
        enqueued right behind this step's row update on the exchange stream, so the id exchange,
        holder-side K1, the barrier and the NVLink pull all hide behind the tail of this step and
        the head of the next one.  Same arithmetic, same order of table reads and writes as the
        un-pipelined step (K2 of step n precedes K1 of step n+1 on the one exchange stream)."""
        if self.gather != "owner" or sparse_next is None:
            return
        if self._pending and self._upd_ev is None:      # no async update was launched: do it now
            self._finish_backward_impl()
        train = self.embed_layers.optimizer is not None
        self._prefetched = self._exchange(sparse_next, train)

    def call(self, inputs, **kwargs):
        dense_inputs, sparse_inputs = inputs
        train = torch.is_grad_enabled() and self.embed_layers.optimizer is not None
        owner = self.gather == "owner"
        ex, pf = None, self._prefetched
        self._prefetched = None
        if pf is not None and pf["key"] == (sparse_inputs.data_ptr(), tuple(sparse_inputs.shape),
                                            sparse_inputs._version) and (pf["train"] or not train):
            ex = pf                     # announced by prefetch(): the rows are (being) pulled already
        else:
            self.finish_backward()      # a pending row update must land before K1 reads the tables
            ex = self._exchange(sparse_inputs, train)
        self._prepared, self._prepared_rep = ex["prepared"], ex["prepared_rep"]
        self._saved_rep_ids = ex["ids_rep"]
        dense_fea = self.bot_dnn(dense_inputs)
        if owner:
            torch.cuda.current_stream().wait_event(ex["pulled"])
        else:
            # every rank's row updates of the previous step are complete before anyone pulls
            self._tab_hdl.barrier(channel=0)
        x = _PeerDotFn.apply(self, sparse_inputs, dense_fea, self.pad_to)
        x = _TopGradsReady.apply(x, self)
        return torch.sigmoid(self.final_dense(self.top_dnn(x)))

    def _launch_update(self):
        """barrier + K2 (+ replicated rows) on the exchange stream, ordered after the K4 backward
        that was just enqueued on the current stream."""
        cur = torch.cuda.current_stream()
        if self._xstream is None:
            self._xstream = torch.cuda.Stream()
        xs = self._xstream
        xs.wait_stream(cur)
        with torch.cuda.stream(xs):
            self._finish_backward_impl()
            self._upd_ev = torch.cuda.Event()
            self._upd_ev.record(xs)

    def finish_backward(self):
        """Complete the embedding update of the last backward on the current stream: either wait
        for the exchange stream (async_update) or run barrier + K2 here."""
        if self._upd_ev is not None:
            torch.cuda.current_stream().wait_event(self._upd_ev)
            self._upd_ev = None
        if self._pending:
            self._finish_backward_impl()

    def _finish_backward_impl(self):
        """Replicated tables: per-rank segment sums (K2, reduce only) -> dense all-reduce (async).
        Sharded tables: all peers' dX rows have landed in this rank's gradient buffer -> K2
        (+ sparse optimizer) on the local shards.  Then the replicated rows' update."""
        if not self._pending:
            return
        D, Ts, Tr = self.D, self._Ts, self._Tr
        Bl = self._grad_B
        Bg = Bl * self.world
        grad = self._grad_buf[: Bg * (Ts + Tr) * D].view(Bg, (Ts + Tr) * D)
        rep = self._reduce_replicated(grad[self.rank * Bl: (self.rank + 1) * Bl, Ts * D:]) if Tr else None
        self._grad_hdl.barrier(channel=0)
        self._out_inflight = False      # every peer is past its K4 forward (it has pushed its dX)
        if Ts:
            self.embed_layers.apply_prepared(self._prepared, grad)
        if rep is not None:
            self._apply_replicated(*rep)
        self._pending, self._prepared = False, None

    # ---- replicated (data-parallel) tables
    def _reduce_replicated(self, g_rep):
        """K2 without optimizer over my samples' lookups of the replicated tables -> dense
        (rows, D) gradient block + touched mask, summed over ranks by an asynchronous all-reduce."""
        Tr, Ts, D, R = self._Tr, self._Ts, self.D, self._rep_total
        cap = min(self._saved_rep_ids.numel(), R) + 1
        uk = torch.full((cap,), -1, dtype=torch.int32, device=g_rep.device)
        sums = torch.zeros((cap, D), dtype=torch.float32, device=g_rep.device)
        self.embed_layers.apply_prepared(self._prepared_rep, g_rep, reduce_only=(uk, sums))
        keys = uk.to(torch.int64) & 0xFFFFFFFF
        rows_all = [int(w.shape[0]) for w in self.embed_layers.weights]
        row_bits = max(1, (max(rows_all) - 1).bit_length())       # K2's key layout: table << row_bits | id
        G = scatter_unique_rows(keys, sums, row_bits, Ts, self._rep_off, R, D)
        work = dist.all_reduce(G, async_op=True)
        return G, work

    def _apply_replicated(self, G, work):
        """Identical update of the rows touched anywhere, on every replica (K2's row formulas)."""
        tl = self.embed_layers
        if not self._rep_state_ready:        # one (R, D) block per optimizer state, viewed per table
            n = tl.optimizer.n_states
            R, D = self._rep_total, self.D
            self._rep_m = torch.zeros((R, D), device=G.device) if n >= 1 else None
            self._rep_v = torch.zeros((R, D), device=G.device) if n >= 2 else None
            o = 0
            for j, r in enumerate(self._rep_rows):
                if n >= 1:
                    tl.state1[self._Ts + j] = self._rep_m[o:o + r]
                if n >= 2:
                    tl.state2[self._Ts + j] = self._rep_v[o:o + r]
                o += r
            self._rep_state_ready = True
        work.wait()
        opt = tl.optimizer
        st = opt.struct_for_step(max(opt.step, 1))
        R, D = self._rep_total, self.D
        # K2's row update on the rows touched anywhere, as ONE kernel (rtf_rows_apply_dense; the
        # torch restatement apply_touched_rows is what the CPU tests check against the oracle)
        g = G[:R, :D].contiguous()
        touched = G[:R, D].contiguous()
        rc = L.lib().rtf_rows_apply_dense(self._rep_W.data_ptr(),
                                          None if self._rep_m is None else self._rep_m.data_ptr(),
                                          None if self._rep_v is None else self._rep_v.data_ptr(),
                                          g.data_ptr(), touched.data_ptr(), R, D, C.byref(st),
                                          L.current_stream_ptr())
        L.check(rc, "rtf_rows_apply_dense")

    def dense_parameters(self):
        emb = {id(p) for p in self.embed_layers.parameters()}
        return [p for p in self.parameters() if id(p) not in emb]


# =============================================================================================
# Self-check: sharded == single-GPU replica (used by bench.py at WORLD_SIZE > 1 and by tests)
# =============================================================================================
def parity_self_check(exchange: str = "peer", gather: str = "owner", replicate_max_rows: int = 70,
                      steps: int = 3, F: int = 26, D: int = 32, B_local: int = 48,
                      seed: int = 5) -> dict:
    """Every rank trains the SAME small DLRM twice — sharded over the world (this rank's slice of
    the global batch) and as a single-GPU replica on the whole global batch — and compares
    predictions, loss, every table shard, its Adam moment and every MLP weight after `steps`
    steps.  26 tables of 37 .. 312 rows so that, with row_wise_min_rows=200 and
    replicate_max_rows=70, row-wise, table-wise AND replicated placements all occur (peer
    exchange).  Eval forwards are interleaved with training ones (no gradient barrier between
    two forwards: the peer-mapped row buffer must not be overwritten under a reader).  Returns
    {"world", "modes", "max_rel_err", "max_rel_err_state", "ok", ...}; never raises on a
    mismatch — the caller decides.  MirroredStrategy semantics (src/ctr/fm/train.py:43-50): the
    sharded run must be indistinguishable from one replica seeing the global batch.  The training
    steps hand the next batch's ids to the trainer, so the pipelined exchange is what is checked."""
    from .dlrm import DLRM, DLRMTrainer
    rank, world = dist.get_rank(), dist.get_world_size()
    rows = [37 + 11 * t for t in range(F)]
    fc = [[{"feat": f"I{i}"} for i in range(13)],
          [{"feat": f"C{t}", "feat_num": rows[t], "embed_dim": D} for t in range(F)]]
    kw = dict(bot_dnn_hidden_units=(64, D), top_dnn_hidden_units=(128, 64), input_bn=False)
    single = DLRM(fc, seed=seed, **kw)
    if exchange == "peer":
        sharded = PeerShardedDLRM(fc, seed=seed, row_wise_min_rows=200, gather=gather,
                                  replicate_max_rows=replicate_max_rows, **kw)
        lay = sharded.layout
        mine = lay.fields[rank]
        kinds = sorted({"row-wise" if lay.row_wise[t] else "replicated" if lay.replicated[t]
                        else "table-wise" for t in range(F)})

        def shard_of(w, t):
            return w[rank::world] if lay.row_wise[t] else w
    else:
        sharded = ShardedDLRM(fc, seed=seed, exchange=exchange, **kw)
        mine = sharded.layout.slots[rank]
        kinds = ["table-wise"]

        def shard_of(w, t):
            return w
    g = torch.Generator(device="cuda").manual_seed(99)
    B = B_local * world
    dense = torch.rand(B, 13, device="cuda", generator=g)
    sparse = torch.stack([torch.randint(0, r, (B,), device="cuda", generator=g) for r in rows],
                         1).to(torch.int32)
    y = (torch.rand(B, 1, device="cuda", generator=g) < 0.3).float()
    sl = slice(rank * B_local, (rank + 1) * B_local)

    err = {"fwd": 0.0, "state": 0.0}

    def rel(a, b, key):
        a, b = a.detach().double(), b.detach().double()
        scale = float(b.abs().max()) or 1.0
        err[key] = max(err[key], float((a - b).abs().max()) / scale)

    t1 = DLRMTrainer(single, lr=1e-2)
    t2 = ShardedDLRMTrainer(sharded, lr=1e-2)
    build_dense_layers(single, 13, dense.device)
    t2._setup(dense[sl])
    with torch.no_grad():       # same start: single-GPU weights -> shards / replicas
        for j, t in enumerate(mine):
            sharded.embed_layers.weights[j].copy_(shard_of(single.embed_layers.weights[t], t))
        for ps, pd in zip(single.dense_parameters(), sharded.dense_parameters()):
            pd.copy_(ps)
        rel(sharded([dense[sl], sparse[sl]]), single([dense, sparse])[sl], "fwd")
        rel(sharded([dense[sl], sparse[sl]]), single([dense, sparse])[sl], "fwd")   # 2 forwards, no bwd
    my_sparse = sparse[sl]
    for it in range(steps):
        l1 = t1.step(dense, sparse, y)
        # the next batch is announced: its exchange is pipelined behind this step (prefetch hit on
        # the next training step; after an eval forward consumed it, the inline path runs instead)
        l2 = t2.step(dense[sl], my_sparse, y[sl], next_sparse=my_sparse)
        lsum = l2.clone()
        dist.all_reduce(lsum)
        rel(lsum / world, l1, "fwd")
        if it >= 1:
            with torch.no_grad():   # eval forward between two training steps
                rel(sharded([dense[sl], my_sparse]), single([dense, sparse])[sl], "fwd")
    for j, t in enumerate(mine):
        rel(sharded.embed_layers.weights[j], shard_of(single.embed_layers.weights[t], t), "state")
        rel(sharded.embed_layers.state1[j], shard_of(single.embed_layers.state1[t], t), "state")
        rel(sharded.embed_layers.state2[j], shard_of(single.embed_layers.state2[t], t), "state")
    for ps, pd in zip(single.dense_parameters(), sharded.dense_parameters()):
        rel(pd, ps, "state")
    bad_ids = False
    try:
        sharded.embed_layers.check_ids()
    except IndexError:
        bad_ids = True
    e = torch.tensor([err["fwd"], err["state"], float(bad_ids)], dtype=torch.float64, device="cuda")
    dist.all_reduce(e, op=dist.ReduceOp.MAX)
    torch.cuda.synchronize()
    dist.barrier()      # nobody frees its peer-mapped buffers while a peer may still address them
    fwd, state, bad = float(e[0]), float(e[1]), bool(e[2])
    ok = (fwd <= 1e-5) and (state <= 1e-4) and not bad and fwd == fwd and state == state
    mode = exchange + (f"/{gather}" if exchange == "peer" else "")
    return {"world": world, "modes": [f"{mode}: {'+'.join(kinds)}"], "steps": steps,
            "max_rel_err": fwd, "max_rel_err_state": state, "tol": 1e-5, "tol_state": 1e-4,
            "ok": bool(ok)}

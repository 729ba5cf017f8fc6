"""Oracle / CPU baseline (test infrastructure): the reference op sequence of the non-DLRM
BASELINE configs on host cores, torch CPU tensors standing in for TensorFlow's Eigen kernels
(TensorFlow is not installable in this image).  Each module restates the reference file it
cites op for op — including what makes it slow (the one-hot FM, the materialised (B,L,4d) DIN
concat, the (B,H,L,L) attention logits) — and is trained with Keras BCE / the model's own loss
and Adam.  Embedding tables get sparse row updates (generous to the CPU: the reference's
l2-regularised tables get a dense Adam sweep).

Used only by bench.py's cpu_baseline / --impl reference legs and by tests.
"""
from __future__ import annotations

import math
import time

import torch
import torch.nn.functional as F

from .dlrm_ref import bce

PAD = -4294967296.0


def _glorot(lin, g):
    lim = math.sqrt(6.0 / (lin.in_features + lin.out_features))
    with torch.no_grad():
        lin.weight.uniform_(-lim, lim, generator=g)
        if lin.bias is not None:
            lin.bias.zero_()
    return lin


class FMRef(torch.nn.Module):
    """ctr FM model, ONE-HOT form exactly as src/ctr/fm/model.py:34-53: stack = concat(dense,
    one_hot(sparse_i)) (B, M); first = w0 + stack @ w; second = 0.5 sum((stack V^T)^2 -
    stack^2 (V^T)^2); sigmoid.  Dense (B, M) matrices and three GEMMs over mostly zeros."""

    def __init__(self, rows, n_dense=13, k=8, seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.rows = list(rows)
        M = n_dense + sum(rows)
        self.w0 = torch.nn.Parameter(torch.zeros(1))
        self.w = torch.nn.Parameter(torch.empty(M, 1).normal_(0, 0.05, generator=g))
        self.V = torch.nn.Parameter(torch.empty(k, M).normal_(0, 0.05, generator=g))

    def forward(self, dense, sparse):
        hots = [F.one_hot(sparse[:, i].long(), n).to(torch.float32) for i, n in enumerate(self.rows)]
        stack = torch.cat([dense] + hots, dim=-1)                                   # :37-42
        first = self.w0 + stack @ self.w                                            # :44
        second = 0.5 * (torch.pow(stack @ self.V.t(), 2) -
                        torch.pow(stack, 2) @ torch.pow(self.V.t(), 2)).sum(1, keepdim=True)  # :46-48
        return torch.sigmoid(first + second)


class DINRef(torch.nn.Module):
    """DIN on the classic input (hist (B,L,nb) ids, target (B,nb) ids): shared item tables
    (src/ctr/din/model.py:71-72), AttentionLayer (src/ctr/layers/modules.py:137-175) with the
    tiled q and the (B,L,4d) concat materialised, BN + FFN + Dense(1) + sigmoid."""

    def __init__(self, feature_nums, embed_dim=8, ffn=(80, 40), seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.tables = torch.nn.ParameterList(
            [torch.nn.Parameter(torch.empty(n, embed_dim).uniform_(-0.05, 0.05, generator=g))
             for n in feature_nums])
        d = embed_dim * len(feature_nums)
        self.att = _glorot(torch.nn.Linear(4 * d, 1), g)
        self.bn = torch.nn.BatchNorm1d(2 * d, eps=1e-3, momentum=0.01)
        dims = [2 * d] + list(ffn)
        self.ffn = torch.nn.ModuleList([_glorot(torch.nn.Linear(a, b), g) for a, b in zip(dims[:-1], dims[1:])])
        self.out = _glorot(torch.nn.Linear(dims[-1], 1), g)

    def forward(self, hist, target):
        B, L, nb = hist.shape
        k = torch.cat([F.embedding(hist[..., i].long(), self.tables[i], sparse=True) for i in range(nb)], -1)
        q = torch.cat([F.embedding(target[:, i].long(), self.tables[i], sparse=True) for i in range(nb)], -1)
        mask = (hist[..., 0] != 0).float()
        qt = q.repeat(1, L).reshape(B, L, -1)                                       # :150-151
        info = torch.cat([qt, k, qt - k, qt * k], dim=-1)                           # :154
        s = torch.sigmoid(self.att(info)).reshape(B, L)                             # :157-159
        s = torch.where(mask == 0, torch.full_like(s, PAD), s)                      # :161-163
        a = torch.softmax(s, -1)                                                    # :169
        user = torch.bmm(a.unsqueeze(1), k).squeeze(1)                              # :170-173
        x = self.bn(torch.cat([user, q], -1))
        for lin in self.ffn:
            x = F.relu(lin(x))
        return torch.sigmoid(self.out(x))


class AutoIntRef(torch.nn.Module):
    """(B, 39, d) field embeddings -> n_layers x ctr MultiHeadAttention (src/ctr/layers/
    modules.py:177-325: relu on Q, K, V, scores * sqrt(hs), softmax, PV, relu(out + relu(X W0)))
    -> Dense(1) -> sigmoid; one gather per sparse field + concat as src/ctr/autoint/model.py:46-47."""

    def __init__(self, rows, n_dense=13, d=16, heads=2, hs=16, n_layers=3, seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.H, self.hs = heads, hs
        self.tables = torch.nn.ParameterList(
            [torch.nn.Parameter(torch.empty(n, d).uniform_(-0.05, 0.05, generator=g)) for n in rows])
        self.dense_embed = torch.nn.Parameter(torch.empty(n_dense, d).uniform_(-0.05, 0.05, generator=g))
        self.layers = torch.nn.ModuleList()
        din = d
        for _ in range(n_layers):
            self.layers.append(torch.nn.ModuleList(
                [_glorot(torch.nn.Linear(din, heads * hs, bias=False), g) for _ in range(4)]))
            din = heads * hs
        self.out = _glorot(torch.nn.Linear((n_dense + len(rows)) * din, 1), g)

    def forward(self, dense, sparse):
        B = dense.shape[0]
        emb = torch.stack([F.embedding(sparse[:, i].long(), self.tables[i], sparse=True)
                           for i in range(sparse.shape[1])], 1)
        x = torch.cat([dense.unsqueeze(-1) * self.dense_embed.unsqueeze(0), emb], 1)
        H, hs = self.H, self.hs
        for wq, wk, wv, w0 in self.layers:
            q, k, v = F.relu(wq(x)), F.relu(wk(x)), F.relu(wv(x))                   # :255-270
            sp = lambda t: t.reshape(B, -1, H, hs).transpose(1, 2)                  # noqa: E731
            s = torch.matmul(sp(q), sp(k).transpose(-1, -2)) / (hs ** -0.5)         # :235-237
            o = torch.matmul(torch.softmax(s, -1), sp(v)).transpose(1, 2).reshape(B, -1, H * hs)
            x = F.relu(o + F.relu(w0(x)))                                           # :316-323
        return torch.sigmoid(self.out(x.reshape(B, -1)))


class SASRecRef(torch.nn.Module):
    """src/match/sasrec/model.py:60-97 with the match TransformerEncoder
    (src/match/layers/modules.py:76-185): three tables, query-row mask, LN(x + att), FFN as two
    1-wide convolutions, LN; last position, dot scores, log loss."""

    def __init__(self, item_num, d=64, blocks=2, ffn_hidden=128, seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.tables = torch.nn.ParameterList(
            [torch.nn.Parameter(torch.empty(item_num, d).uniform_(-0.05, 0.05, generator=g)) for _ in range(3)])
        self.blocks = torch.nn.ModuleList()
        for _ in range(blocks):
            self.blocks.append(torch.nn.ModuleDict(dict(
                wq=_glorot(torch.nn.Linear(d, d), g), wk=_glorot(torch.nn.Linear(d, d), g),
                wv=_glorot(torch.nn.Linear(d, d), g), ln1=torch.nn.LayerNorm(d, eps=1e-6),
                c1=_glorot(torch.nn.Linear(d, ffn_hidden), g), c2=_glorot(torch.nn.Linear(ffn_hidden, d), g),
                ln2=torch.nn.LayerNorm(d, eps=1e-6))))
        self.d = d

    def forward(self, seq, pos, neg):
        mask = (seq != 0).float().unsqueeze(-1)                                     # :72
        x = F.embedding(seq.long(), self.tables[0], sparse=True) * mask
        pe = F.embedding(pos.long(), self.tables[1], sparse=True)
        ne = F.embedding(neg.long(), self.tables[2], sparse=True)
        for b in self.blocks:
            q, k, v = b["wq"](x), b["wk"](x), b["wv"](x)
            s = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(self.d)            # modules.py:85-88
            s = torch.where(mask == 0, torch.full_like(s, PAD), s)                  # :90-91 (query rows)
            att = torch.matmul(torch.softmax(s, -1), v)
            o1 = b["ln1"](x + att)
            o2 = b["ln2"](o1 + b["c2"](F.relu(b["c1"](o1))))
            x = o2 * mask                                                           # model.py:86
        info = x[:, -1:, :]                                                         # :88
        ps, ns = (info * pe).sum(-1), (info * ne).sum(-1)
        loss = (-torch.log(torch.sigmoid(ps)) - torch.log(1 - torch.sigmoid(ns))).mean() / 2
        return loss


class YoutubeDNNRef(torch.nn.Module):
    """Two-tower user side + tf.nn.sampled_softmax_loss over a (N, D) item table (the
    conventional form of src/match/layers/modules.py:54-60; App. A13/A14), log-uniform samples
    injected per step."""

    def __init__(self, user_feature_nums, item_num, embed_dim=64, hidden=(64, 32), seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.user_tables = torch.nn.ParameterList(
            [torch.nn.Parameter(torch.empty(n, embed_dim).uniform_(-0.05, 0.05, generator=g))
             for n in user_feature_nums])
        dims = [embed_dim * len(user_feature_nums)] + list(hidden)
        self.dnn = torch.nn.ModuleList([_glorot(torch.nn.Linear(a, b), g) for a, b in zip(dims[:-1], dims[1:])])
        self.items = torch.nn.Parameter(torch.empty(item_num, hidden[-1]).uniform_(-0.05, 0.05, generator=g))
        self.N = item_num

    def forward(self, user_ids, item_ids, sampled, true_exp, samp_exp):
        x = torch.cat([F.embedding(user_ids[:, i].long(), self.user_tables[i], sparse=True)
                       for i in range(user_ids.shape[1])], -1)
        for lin in self.dnn:
            x = F.relu(lin(x))
        lab = item_ids.reshape(-1).long()
        tw = F.embedding(lab, self.items, sparse=True)
        sw = F.embedding(sampled.long(), self.items, sparse=True)
        true_logits = (x * tw).sum(1) - torch.log(true_exp)
        samp_logits = x @ sw.t()
        hit = lab.unsqueeze(1) == sampled.long().unsqueeze(0)
        samp_logits = samp_logits + torch.where(hit, torch.full_like(samp_logits, -torch.finfo(torch.float32).max),
                                                torch.zeros_like(samp_logits)) - torch.log(samp_exp).unsqueeze(0)
        logits = torch.cat([true_logits.unsqueeze(1), samp_logits], 1)
        return F.cross_entropy(logits, torch.zeros(len(lab), dtype=torch.long), reduction="mean")


class CpuStep:
    """SparseAdam on the embedding tables, Adam (Keras eps) on the rest."""

    def __init__(self, model, table_prefixes=("tables.", "user_tables.", "items"), lr=1e-3):
        self.model = model
        sparse = [p for n, p in model.named_parameters() if n.startswith(table_prefixes)]
        dense = [p for n, p in model.named_parameters() if not n.startswith(table_prefixes)]
        self.opts = []
        if sparse:
            self.opts.append(torch.optim.SparseAdam(sparse, lr=lr, eps=1e-7))
        if dense:
            self.opts.append(torch.optim.Adam(dense, lr=lr, eps=1e-7))

    def step(self, loss_of):
        for o in self.opts:
            o.zero_grad(set_to_none=True)
        loss = loss_of(self.model)
        loss.backward()
        for o in self.opts:
            o.step()
        return float(loss.detach())


def time_cpu(model, loss_of_batch, batches, warmup=1, threads=None):
    """-> (samples/s, threads, seconds/step); loss_of_batch(model, batch) -> scalar loss."""
    if threads:
        torch.set_num_threads(threads)
    st = CpuStep(model)
    for b in batches[:warmup]:
        st.step(lambda m: loss_of_batch(m, b))
    t0 = time.perf_counter()
    n = 0
    for b in batches[warmup:]:
        st.step(lambda m: loss_of_batch(m, b))
        n += b[0].shape[0]
    dt = time.perf_counter() - t0
    return n / dt, torch.get_num_threads(), dt / max(1, len(batches) - warmup)


__all__ = ["FMRef", "DINRef", "AutoIntRef", "SASRecRef", "YoutubeDNNRef", "CpuStep", "time_cpu", "bce"]

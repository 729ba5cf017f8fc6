"""Oracle / CPU baseline (test infrastructure): the reference DLRM op sequence on host cores.

Restates src/ctr/dlrm/model.py:16-54 op for op with torch CPU tensors (multi-threaded ATen
kernels standing in for TensorFlow's Eigen CPU kernels — TensorFlow itself is not installable
in this image): one gather per sparse field + concat (:45-46), DNN = BatchNormalization +
Dense stack (src/ctr/layers/modules.py:129-135), the interaction (:48 'cat', or the paper's
pairwise dot the file cites at :7), Dense(1) + sigmoid (:51-53), Keras binary_crossentropy and
Adam (src/ctr/fm/train.py:49-50).  Embedding rows are updated sparsely (touched rows only),
which is generous to the CPU: the reference's l2-regularised tables get a dense Adam sweep.

Used only by tests (cross-check of the product path) and by bench.py's cpu_baseline /
--impl reference legs.
"""
from __future__ import annotations

import math
import time

import torch
import torch.nn.functional as F


class DLRMRef(torch.nn.Module):
    def __init__(self, rows, dim, n_dense=13, bot=(64, 32, 16), top=(128, 64), interaction="dot",
                 input_bn=True, seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.interaction, self.dim = interaction, dim
        self.tables = torch.nn.ParameterList(
            [torch.nn.Parameter(torch.empty(n, dim).uniform_(-0.05, 0.05, generator=g)) for n in rows])
        F1 = len(rows) + 1
        top_in = dim + F1 * (F1 - 1) // 2 if interaction == "dot" else len(rows) * dim + bot[-1]

        def mlp(n_in, units):
            layers, bn = [], (torch.nn.BatchNorm1d(n_in, eps=1e-3, momentum=0.01) if input_bn else None)
            for u in units:
                lin = torch.nn.Linear(n_in, u)
                lim = math.sqrt(6.0 / (n_in + u))
                with torch.no_grad():
                    lin.weight.uniform_(-lim, lim, generator=g)
                    lin.bias.zero_()
                layers.append(lin)
                n_in = u
            return bn, torch.nn.ModuleList(layers)

        self.bot_bn, self.bot = mlp(n_dense, bot)
        self.top_bn, self.top = mlp(top_in, top)
        self.final = torch.nn.Linear(top[-1], 1)

    @staticmethod
    def _dnn(bn, layers, x):
        if bn is not None:
            x = bn(x)
        for lin in layers:
            x = F.relu(lin(x))
        return x

    def forward(self, dense, sparse):
        dense_fea = self._dnn(self.bot_bn, self.bot, dense)
        embs = [F.embedding(sparse[:, i].long(), self.tables[i], sparse=True)
                for i in range(sparse.shape[1])]                       # F separate gathers
        if self.interaction == "cat":
            x = torch.cat(embs + [dense_fea], dim=-1)                  # model.py:48
        else:
            X = torch.stack([dense_fea] + embs, dim=1)                 # (B, F1, D)
            Z = torch.bmm(X, X.transpose(1, 2))
            ii, jj = torch.tril_indices(X.shape[1], X.shape[1], -1)
            x = torch.cat([dense_fea, Z[:, ii, jj]], dim=1)
        return torch.sigmoid(self.final(self._dnn(self.top_bn, self.top, x)))


def bce(y, p):
    eps = 1e-7
    p = torch.clamp(p, eps, 1 - eps)
    return -(y * torch.log(p + eps) + (1 - y) * torch.log(1 - p + eps)).mean()


class CpuTrainer:
    def __init__(self, model: DLRMRef, lr=1e-3):
        self.model = model
        self.sparse_opt = torch.optim.SparseAdam(list(model.tables), lr=lr, eps=1e-7)
        dense = [p for n, p in model.named_parameters() if not n.startswith("tables.")]
        self.dense_opt = torch.optim.Adam(dense, lr=lr, eps=1e-7)

    def step(self, dense, sparse, labels):
        self.sparse_opt.zero_grad(set_to_none=True)
        self.dense_opt.zero_grad(set_to_none=True)
        loss = bce(labels, self.model(dense, sparse))
        loss.backward()
        self.sparse_opt.step()
        self.dense_opt.step()
        return float(loss)


def time_cpu_train(rows, dim, bot, top, batches, warmup=1, interaction="dot", threads=None):
    """-> (samples/s, threads used, seconds per step).  `batches`: list of (dense, sparse, y)."""
    if threads:
        torch.set_num_threads(threads)
    model = DLRMRef(rows, dim, bot=bot, top=top, interaction=interaction)
    tr = CpuTrainer(model)
    for d, s, y in batches[:warmup]:
        tr.step(d, s, y)
    t0 = time.perf_counter()
    n = 0
    for d, s, y in batches[warmup:]:
        tr.step(d, s, y)
        n += d.shape[0]
    dt = time.perf_counter() - t0
    steps = max(1, len(batches) - warmup)
    return n / dt, torch.get_num_threads(), dt / steps

"""The non-DLRM BASELINE.json configs (configs[0], [2], [3], [4]) as bench.py workloads.

Each workload = one training step (forward, loss, backward, fused sparse Adam on the tables,
dense Adam on the rest) of the model on synthetic data of SURVEY.md §8d's shape, plus
  * the dominant hot-path kernel of that model, timed alone with CUDA events, against the
    roofline that bounds it (§8d's algorithmic bytes / FLOPs),
  * the op-for-op CPU port of the reference model (oracle/models_ref.py) on a bounded sample.
bench.py owns the timing protocol and the JSON line; this file only describes the workloads.
"""
from __future__ import annotations

import math
import os

import numpy as np

CRITEO_ROWS = [min(n, 10_000_000) for n in (
    1460, 583, 10131227, 2202608, 305, 24, 12517, 633, 3, 93145, 5683, 8351593, 3194, 27, 14992,
    5461306, 10, 5652, 2173, 4, 7046547, 18, 15, 286181, 105, 142572)]
FMA_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12      # fp32 FFMA peak of a B200 at max SM clock


def _t(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a))


class Workload:
    name = ""
    metric = ""
    batch = 0
    cpu_batch = 0
    cuda_graph = False        # set by bench.py --cuda-graph: replay the step from a CUDA graph
    graphable = True          # False: the step draws a host-side value every step

    def config(self):
        raise NotImplementedError

    def build(self, pkg):
        """-> trainer-like object with .step(*dev_batch) -> loss tensor"""
        raise NotImplementedError

    def host_batch(self, rng, B, cpu=False):
        """-> tuple of numpy arrays (the model inputs + labels)"""
        raise NotImplementedError

    def roofline(self, pkg, stepper, dev_batches, peaks):
        return None

    def cpu_model(self):
        raise NotImplementedError

    def cpu_loss(self, model, batch):
        raise NotImplementedError


class _Stepper:
    def __init__(self, trainer, to_inputs):
        self.trainer, self.to_inputs = trainer, to_inputs
        self.model = trainer.model

    def step(self, *batch):
        inputs, labels = self.to_inputs(batch)
        return self.trainer.step(inputs, labels)


# ---------------------------------------------------------------------------------- configs[0] FM
class FM(Workload):
    """ctr FM model (src/ctr/fm/model.py:34-53), 13 dense + 26 sparse Criteo fields, k = 8,
    batch 1024 — gather form (K3b) + K2; the CPU port is the reference's one-hot form."""
    name, metric, batch, cpu_batch = "fm_criteo_synthetic", "FM train samples/sec", 1024, 1024
    K = 8
    CPU_ROW_CAP = 1000        # SURVEY §8d parity size: M = 26 013 one-hot columns

    def config(self):
        return {"workload": self.name, "model": "ctr FM (gather form of the one-hot model)", "k": self.K,
                "dense": 13, "sparse": 26, "rows_total": sum(CRITEO_ROWS), "batch_per_gpu": self.batch,
                "ids": "uniform", "optimizer": "adam (sparse rows in K2, dense rtf_dense_adam)"}

    def build(self, pkg):
        import torch
        fc = pkg.criteo_feature_columns(self.K, rows=CRITEO_ROWS)
        m = pkg.FMModel(fc, k=self.K, seed=1)
        tr = pkg.models.Trainer(m, lambda out, y: pkg.layers.binary_crossentropy(y, out), embed_l2=1e-4,
                                cuda_graph=self.cuda_graph)
        return _Stepper(tr, lambda b: ([b[0], b[1]], b[2]))

    def host_batch(self, rng, B, cpu=False):
        rows = [min(r, self.CPU_ROW_CAP) for r in CRITEO_ROWS] if cpu else CRITEO_ROWS
        from recommend_tf2_b200.data import synthetic_criteo_batch
        return synthetic_criteo_batch(rng, B, rows, "uniform")

    def roofline(self, pkg, st, dev, peaks):
        import torch
        m, kp = st.model, st.model.kp
        with torch.no_grad():
            ms = _timed(lambda i: m([dev[i][0], dev[i][1]]), len(dev))
        B = dev[0][0].shape[0]
        by = B * (26 * (4 + kp * 4) + 13 * 4 + 4)
        return _hbm_roof("fm_gather_kernel (K3b forward: 39 row gathers + FM cross per sample)", ms, by, peaks,
                         note="32-byte rows (k=8 padded to 12 floats = 48 B): HBM moves 64-B sectors, so "
                              "algorithmic bytes / time cannot approach the copy peak (SURVEY §7); B=1024 is "
                              "launch-latency sized")

    def cpu_model(self):
        from oracle.models_ref import FMRef
        return FMRef([min(r, self.CPU_ROW_CAP) for r in CRITEO_ROWS], k=self.K)

    def cpu_loss(self, model, b):
        from oracle.models_ref import bce
        return bce(b[2], model(b[0], b[1]))

    cpu_note = ("one-hot form as the reference writes it, tables capped at 1000 rows (M = 26 013 "
                "one-hot columns; the full cardinalities would need a 137 GB one-hot matrix)")


# ---------------------------------------------------------------------------------- configs[2] DIN
class DIN(Workload):
    """DIN local activation unit over a behaviour sequence of 100 (item, category) pairs."""
    name, metric, batch, cpu_batch = "din_behaviour_seq_synthetic", "DIN train samples/sec", 4096, 1024
    L, D, ITEMS, CATES = 100, 8, 1_000_000, 1000

    def config(self):
        return {"workload": self.name, "model": "DIN (AttentionLayer local activation unit)", "seq_len": self.L,
                "embed_dim": self.D, "d": 2 * self.D, "item_rows": self.ITEMS, "cate_rows": self.CATES,
                "batch_per_gpu": self.batch, "lengths": "U{1..100}, 0 = padding id",
                "optimizer": "adam (sparse rows in K2, dense rtf_dense_adam)"}

    def build(self, pkg):
        m = pkg.models.DIN([self.ITEMS, self.CATES], embed_dim=self.D, maxlen=self.L, seed=1)
        tr = pkg.models.Trainer(m, lambda out, y: pkg.layers.binary_crossentropy(y, out), embed_l2=1e-4,
                                cuda_graph=self.cuda_graph)
        return _Stepper(tr, lambda b: ([b[0], b[1]], b[2]))

    def host_batch(self, rng, B, cpu=False):
        lens = rng.integers(1, self.L + 1, B)
        hist = np.stack([rng.integers(1, self.ITEMS, (B, self.L)), rng.integers(1, self.CATES, (B, self.L))], -1)
        hist[np.arange(self.L)[None, :] >= lens[:, None]] = 0
        target = np.stack([rng.integers(1, self.ITEMS, B), rng.integers(1, self.CATES, B)], -1)
        y = (rng.random((B, 1)) < 0.25).astype(np.float32)
        return hist.astype(np.int32), target.astype(np.int32), y

    def roofline(self, pkg, st, dev, peaks):
        import torch
        m = st.model
        B, L, d = dev[0][0].shape[0], self.L, 2 * self.D
        q = torch.randn(B, d, device="cuda")
        k = torch.randn(B, L, d, device="cuda")
        mask = (dev[0][0][..., 0] != 0).float()
        with torch.no_grad():
            ms = _timed(lambda i: m.attention_layer([q, k, k, mask]), 6)
        by = B * (L * d * 4 + L * 4 + 2 * d * 4)
        return _hbm_roof("din_kernel (K5 forward: scores + masked softmax + weighted sum, one pass over k)", ms,
                         by, peaks)

    def cpu_model(self):
        from oracle.models_ref import DINRef
        return DINRef([self.ITEMS, self.CATES], self.D)

    def cpu_loss(self, model, b):
        from oracle.models_ref import bce
        return bce(b[2], model(b[0], b[1]))

    cpu_note = "same tables and sequence length; tiled q and the (B,L,4d) concat materialised as the reference does"


# ------------------------------------------------------------------------------ configs[3] AutoInt
class AutoInt(Workload):
    """AutoInt: 39 fields x 16, 3 interacting layers, 2 heads x 16, residual."""
    name, metric, batch, cpu_batch = "autoint_criteo_synthetic", "AutoInt train samples/sec", 4096, 1024
    D, H, HS, NL = 16, 2, 16, 3

    def config(self):
        return {"workload": self.name, "model": "AutoInt (3 x ctr MultiHeadAttention, use_res)", "fields": 39,
                "embed_dim": self.D, "heads": self.H, "head_size": self.HS, "layers": self.NL,
                "rows_total": sum(CRITEO_ROWS), "batch_per_gpu": self.batch, "ids": "uniform",
                "optimizer": "adam (sparse rows in K2, dense rtf_dense_adam)"}

    def build(self, pkg):
        fc = pkg.criteo_feature_columns(self.D, rows=CRITEO_ROWS)
        m = pkg.models.AutoInt(fc, self.HS, self.H, self.NL, use_res=True, seed=1)
        tr = pkg.models.Trainer(m, lambda out, y: pkg.layers.binary_crossentropy(y, out), embed_l2=1e-4,
                                cuda_graph=self.cuda_graph)
        return _Stepper(tr, lambda b: ([b[0], b[1]], b[2]))

    def host_batch(self, rng, B, cpu=False):
        from recommend_tf2_b200.data import synthetic_criteo_batch
        return synthetic_criteo_batch(rng, B, CRITEO_ROWS, "uniform")

    def roofline(self, pkg, st, dev, peaks):
        import torch
        B, F = dev[0][0].shape[0], 39
        layer = st.model.att_layers[1]             # a 32 -> 32 layer
        x = torch.randn(B, F, self.H * self.HS, device="cuda")
        with torch.no_grad():
            ms = _timed(lambda i: layer(x), 6)
        dm = self.H * self.HS
        fl = B * (4 * F * dm * dm * 2 + 2 * (self.H * F * F * self.HS * 2))
        return _fma_roof("AutoInt interacting layer forward (K6: QKV/W0 projections + softmax(QK^T sqrt hs)V + res)",
                         ms, fl)

    def cpu_model(self):
        from oracle.models_ref import AutoIntRef
        return AutoIntRef(CRITEO_ROWS, d=self.D, heads=self.H, hs=self.HS, n_layers=self.NL)

    def cpu_loss(self, model, b):
        from oracle.models_ref import bce
        return bce(b[2], model(b[0], b[1]))

    cpu_note = "full Criteo cardinalities at dim 16 (2.2 GB of tables), sparse-row Adam"


# ------------------------------------------------------------------------------- configs[4] SASRec
class SASRec(Workload):
    """SASRec: 2 blocks, sequence length 200, dim 64, 1 head, 100 negatives, 1 M-item vocab."""
    name, metric, batch, cpu_batch = "sasrec_seq200_synthetic", "SASRec train samples/sec", 1024, 256
    L, D, NEG, ITEMS = 200, 64, 100, 1_000_000

    def config(self):
        return {"workload": self.name, "model": "SASRec (2 x match TransformerEncoder)", "seq_len": self.L,
                "embed_dim": self.D, "blocks": 2, "heads": 1, "neg_len": self.NEG, "item_rows": self.ITEMS,
                "batch_per_gpu": self.batch, "padding": "left (pad_sequences 'pre'), lengths U{1..200}",
                "optimizer": "adam (sparse rows in K2, dense rtf_dense_adam)"}

    def build(self, pkg):
        m = pkg.models.SASRec(self.ITEMS, self.D, blocks=2, num_heads=1, seq_len=self.L, neg_len=self.NEG, seed=1)
        tr = pkg.models.Trainer(m, lambda out, y: out[1], cuda_graph=self.cuda_graph)
        return _Stepper(tr, lambda b: ([b[0], b[1], b[2]], None))

    def host_batch(self, rng, B, cpu=False):
        lens = rng.integers(1, self.L + 1, B)
        seq = rng.integers(1, self.ITEMS, (B, self.L))
        seq[np.arange(self.L)[None, :] < (self.L - lens)[:, None]] = 0        # pre-padding
        pos = rng.integers(1, self.ITEMS, (B, 1))
        neg = rng.integers(1, self.ITEMS, (B, self.NEG))
        return seq.astype(np.int32), pos.astype(np.int32), neg.astype(np.int32)

    def roofline(self, pkg, st, dev, peaks):
        import torch
        B, L, d = dev[0][0].shape[0], self.L, self.D
        qkv = torch.randn(B, L, d, device="cuda")
        rm = (dev[0][0] != 0).float()
        with torch.no_grad():
            ms = _timed(lambda i: pkg.attention(qkv, qkv, qkv, 1, 1.0 / math.sqrt(d), row_mask=rm), 6)
        return _fma_roof("attn_fwd8_kernel (K7 forward: QK^T / sqrt d, query-row mask, softmax, PV; L=200, d=64)",
                         ms, B * 2 * (L * L * d * 2))

    def cpu_model(self):
        from oracle.models_ref import SASRecRef
        return SASRecRef(self.ITEMS, self.D)

    def cpu_loss(self, model, b):
        return model(b[0], b[1], b[2])

    cpu_note = "same vocab, sequence length and blocks; (B,1,L,L) logits materialised as the reference does"


# --------------------------------------------------------------------------- configs[4] YoutubeDNN
class YoutubeDNN(Workload):
    """YoutubeDNN user tower + sampled softmax over a 1 M-item table, S = 1024 log-uniform samples."""
    name, metric, batch, cpu_batch = "youtubednn_sampled_softmax_synthetic", "YoutubeDNN train samples/sec", 4096, 1024
    ITEMS, D, S = 1_000_000, 64, 1024
    USER_FEATS = (1_000_000, 1000, 100, 10)

    def config(self):
        return {"workload": self.name, "model": "YoutubeDNN (conventional sampled softmax over the item table)",
                "item_rows": self.ITEMS, "embed_dim": self.D, "num_sampled": self.S,
                "user_features": list(self.USER_FEATS), "user_dnn": [64, 64], "batch_per_gpu": self.batch,
                "sampler": "device log-uniform, unique, shared by the batch (App. A14)",
                "optimizer": "adam (sparse rows in K2 incl. the class-weight table, dense rtf_dense_adam)"}

    def build(self, pkg):
        m = pkg.models.YoutubeDNN(self.USER_FEATS, self.ITEMS, self.D, (64, self.D), num_sampled=self.S,
                                  conventional=True, sparse_optimizer=pkg.SparseOptimizer("adam", lr=1e-3), seed=1)
        tr = pkg.models.Trainer(m, lambda out, y: out.mean(), cuda_graph=self.cuda_graph)         # loss_util.sampledsoftmaxloss
        return _Stepper(tr, lambda b: ([b[0], b[1]], None))

    def host_batch(self, rng, B, cpu=False):
        users = np.stack([rng.integers(0, n, B) for n in self.USER_FEATS], 1)
        items = (rng.zipf(1.2, B) - 1) % self.ITEMS
        out = [users.astype(np.int32), items.astype(np.int32).reshape(B, 1)]
        if cpu:       # injected log-uniform samples for the CPU port (unique, shared by the batch)
            lr = math.log(self.ITEMS + 1.0)
            seen, tries = [], 0
            sset = set()
            while len(seen) < self.S:
                tries += 1
                c = min(max(int(math.exp(rng.random() * lr)) - 1, 0), self.ITEMS - 1)
                if c not in sset:
                    sset.add(c)
                    seen.append(c)
            s = np.asarray(seen, np.int64)
            p = lambda c: (np.log(c + 2.0) - np.log(c + 1.0)) / lr                      # noqa: E731
            ec = lambda c: -np.expm1(tries * np.log1p(-p(c)))                            # noqa: E731
            out += [s, ec(items.astype(np.float64)).astype(np.float32), ec(s.astype(np.float64)).astype(np.float32)]
        return tuple(out)

    def roofline(self, pkg, st, dev, peaks):
        import torch
        from recommend_tf2_b200.layers import match as M
        B, N, D, S = dev[0][0].shape[0], self.ITEMS, self.D, self.S
        W = st.model.item_table.weights[0]
        x = torch.randn(B, D, device="cuda")
        labels = dev[0][1].reshape(-1).long()
        smp, tries = M.log_uniform_candidate_sampler(S, N, seed=1)
        te, se = M.log_uniform_expected(labels, N, tries), M.log_uniform_expected(smp, N, tries)
        with torch.no_grad():
            ms = _timed(lambda i: M.sampled_softmax_loss(W, None, labels, x, S, N, sampled_values=(smp, te, se)), 6)
        r = _fma_roof("K8 forward, GEMM form (ssm_gather + tcgen05 logits GEMM + ssm_logits_fwd epilogue: "
                      "true logit, hit removal, log-sum-exp)", ms, B * (S + 1) * D * 2)
        r["note"] = ("3 launches; shown against the fp32 FFMA peak for continuity with round 1's streaming "
                     "kernel (0.35 ms = 0.02) — the logits now run on the tensor cores, the step is "
                     "launch-latency sized at B = 4096")
        return r

    def cpu_model(self):
        from oracle.models_ref import YoutubeDNNRef
        return YoutubeDNNRef(self.USER_FEATS, self.ITEMS, self.D, (64, self.D))

    def cpu_loss(self, model, b):
        return model(b[0], b[1], b[2], b[3], b[4])

    cpu_note = "same tables; log-uniform samples drawn on the host and injected (shared by the batch)"


WORKLOADS = {"fm": FM, "din": DIN, "autoint": AutoInt, "sasrec": SASRec, "youtubednn": YoutubeDNN}


# -------------------------------------------------------------------------------------- helpers
def _timed(fn, n, warm=3):
    import statistics

    import torch
    for i in range(warm):
        fn(i % n)
    torch.cuda.synchronize()
    evs = []
    for i in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(300_000)
        a.record()
        fn(i)
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in evs)


def _hbm_roof(kernel, ms, nbytes, peaks, note=None):
    gbs = nbytes / (ms * 1e-3) / 1e9
    r = {"kernel": kernel, "bound": "hbm", "achieved": round(gbs, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
         "frac": round(gbs / peaks["hbm_gbs"], 4), "traffic": None,
         "algorithmic_bytes_per_launch": int(nbytes), "ms_per_launch": round(ms, 4)}
    if note:
        r["note"] = note
    return r


def _fma_roof(kernel, ms, flops):
    tf = flops / (ms * 1e-3) / 1e12
    return {"kernel": kernel, "bound": "fp32_fma", "achieved": round(tf, 2), "peak": round(FMA_TFLOPS, 1),
            "unit": "TFLOP/s", "frac": round(tf / FMA_TFLOPS, 4), "traffic": None,
            "peak_source": "148 SMs x 128 FMA lanes x 2 x 1.965 GHz (fp32 CUDA-core peak; the 1e-5 parity bar "
                           "rules out single-pass bf16/TF32 tensor-core math for these contractions)",
            "algorithmic_flops_per_launch": int(flops), "ms_per_launch": round(ms, 4)}


def run_cpu(wl: Workload, steps: int, warmup: int):
    """The workload's CPU port on a bounded sample -> cpu_baseline dict (+ ms_per_step)."""
    import torch
    from oracle import models_ref
    rng = np.random.default_rng(1)
    batches = [tuple(_t(a) for a in wl.host_batch(rng, wl.cpu_batch, cpu=True)) for _ in range(warmup + steps)]
    model = wl.cpu_model()
    sps, threads, sec = models_ref.time_cpu(model, wl.cpu_loss, batches, warmup=warmup, threads=os.cpu_count())
    return {"value": sps, "unit": "samples/s", "cores": threads, "kind": "port",
            "sample": f"{steps} steps x {wl.cpu_batch} samples; {wl.cpu_note}; torch-CPU restatement of the "
                      f"reference op sequence (oracle/models_ref.py), not TensorFlow",
            "ms_per_step": sec * 1e3}

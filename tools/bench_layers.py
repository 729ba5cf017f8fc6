#!/usr/bin/env python
"""Per-kernel timing of the non-DLRM hot-path kernels at the BASELINE.json config sizes
(configs[0], [2], [3], [4]) with their roofline: achieved GB/s (HBM-bound kernels) or TFLOP/s
(FMA-bound attention) against MEASURED_PEAKS.json / the fp32 FMA peak.  CUDA events on the
launch stream, stream kept busy so host launch latency is excluded, >= 3 warm-ups.

    python tools/bench_layers.py > profiles/r1_layers.jsonl
"""
import json
import math
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import bench
import recommend_tf2_b200 as pkg

PEAKS, _ = bench.load_peaks()
HBM = PEAKS["hbm_gbs"]
FMA_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12      # fp32 FFMA peak of a B200 at max clock


def timed(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(300_000)
        a.record()
        fn()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in evs)


def emit(name, ms, nbytes=None, flops=None, **extra):
    rec = {"kernel": name, "ms": round(ms, 4)}
    if nbytes:
        gbs = nbytes / ms / 1e6
        rec.update(algorithmic_bytes=int(nbytes), gbs=round(gbs, 1), frac_hbm=round(gbs / HBM, 4))
    if flops:
        tf = flops / ms / 1e9
        rec.update(flops=int(flops), tflops=round(tf, 2), frac_fp32_fma=round(tf / FMA_TFLOPS, 4))
    rec.update(extra)
    print(json.dumps(rec), flush=True)


def fwd_bwd(make_out, params):
    """returns (fwd_fn, bwd_fn) closures for timing; bwd re-runs fwd untimed once per call."""
    out = make_out()
    g = torch.randn_like(out)

    def bwd():
        for p in params:
            p.grad = None
        out.backward(g, retain_graph=True)
    return bwd


def main():
    torch.manual_seed(0)
    dev = "cuda"
    rows = bench.CRITEO_ROWS
    # ---- configs[0]: FM model (13 dense + 26 sparse, k = 8), Criteo cardinalities
    for B in (1024, 65536):
        fc = [[{"feat": f"I{i}"} for i in range(13)],
              [{"feat": f"C{i}", "feat_num": r, "embed_dim": 8} for i, r in enumerate(rows)]]
        m = pkg.FMModel(fc, k=8, seed=0, sparse_optimizer=pkg.SparseOptimizer("adam"))
        m.tables.begin_step()
        dense = torch.rand(B, 13, device=dev)
        sparse = torch.stack([torch.randint(0, r, (B,), device=dev) for r in rows], 1).to(torch.int32)
        with torch.no_grad():
            ms = timed(lambda: m([dense, sparse]))
        kp = m.kp
        emit(f"fm_gather_fwd(K3b) B={B}", ms, B * (26 * (4 + kp * 4) + 13 * 4 + kp * 4 + 4), config="FM k=8")
        out = m([dense, sparse])
        g = torch.randn_like(out)
        ms = timed(lambda: out.backward(g, retain_graph=True))
        emit(f"fm_gather_bwd+K2+colsum B={B}", ms, B * (26 * (4 + 2 * kp * 4) + 13 * kp * 4 + kp * 4 + 8),
             config="FM k=8, fused sparse Adam on the touched rows (K2)")
        del m
    # ---- FM layer as DeepFM calls it
    B = 65536
    layer = pkg.layers.FM(13 + 26 * 8)
    first = torch.randn(B, 221, device=dev, requires_grad=True)
    second = torch.randn(B, 208, device=dev, requires_grad=True)
    with torch.no_grad():
        ms = timed(lambda: layer([first, second]))
    emit("fm_layer_fwd(K3a) B=65536", ms, B * (221 + 208 + 1) * 4)
    out = layer([first, second])
    g = torch.randn_like(out)
    ms = timed(lambda: out.backward(g, retain_graph=True))
    emit("fm_layer_bwd(K3a) B=65536", ms, B * (2 * (221 + 208) + 1 + 221) * 4)
    # ---- K1 pooled lookup (bag of 10 ids, sum) D = 64
    B, L, D = 65536, 10, 64
    tabs = [torch.randn(1_000_000, D, device=dev) for _ in range(4)]
    ids = torch.randint(0, 1_000_000, (B, 4, L), device=dev, dtype=torch.int32)
    ms = timed(lambda: pkg.embed_fwd(tabs, ids, "BFL", "sum"))
    emit("embed_fwd_vec pooled sum L=10 D=64", ms, B * 4 * (L * (4 + D * 4) + D * 4))
    del tabs
    # ---- configs[2]: DIN local activation unit, L = 100
    for d in (16, 128):
        B, L = 4096, 100
        layer = pkg.layers.AttentionLayer(1, activation="sigmoid")
        q = torch.randn(B, d, device=dev, requires_grad=True)
        k = torch.randn(B, L, d, device=dev, requires_grad=True)
        mask = (torch.rand(B, L, device=dev) < 0.7).float()
        with torch.no_grad():
            ms = timed(lambda: layer([q, k, k, mask]))
        emit(f"din_kernel fwd(K5) L=100 d={d}", ms, B * (L * d * 4 + L * 4 + 2 * d * 4))
        out = layer([q, k, k, mask])
        g = torch.randn_like(out)
        ms = timed(lambda: out.backward(g, retain_graph=True))
        emit(f"din_kernel bwd(K5)+colsum L=100 d={d}", ms, B * (3 * L * d * 4 + L * 4 + 3 * d * 4 + (4 * d + 1) * 4))
    # ---- configs[3]: AutoInt interacting layer (39 fields, d 16, 2 heads x 16), core + whole layer
    B, F, dm, H, hs = 4096, 39, 16, 2, 16
    x = torch.randn(B, F, dm, device=dev, requires_grad=True)
    layer = pkg.layers.ctr.MultiHeadAttention(hs, H, use_res=True)
    with torch.no_grad():
        ms = timed(lambda: layer(x))
    fl_layer = B * (4 * F * dm * H * hs * 2 + 2 * (H * F * F * hs * 2))
    emit("autoint layer fwd (Dense proj + attn core K6)", ms, flops=fl_layer, config="39x16, 2 heads x16, res")
    qkv = torch.randn(B, F, H * hs, device=dev, requires_grad=True)
    with torch.no_grad():
        ms = timed(lambda: pkg.attention(qkv, qkv, qkv, H, math.sqrt(hs)))
    emit("attn_fwd_kernel core F=39 hs=16 H=2", ms, flops=B * 2 * (H * F * F * hs * 2), nbytes=B * F * H * hs * 4 * 4)
    out = pkg.attention(qkv, qkv, qkv, H, math.sqrt(hs))
    g = torch.randn_like(out)
    ms = timed(lambda: out.backward(g, retain_graph=True))
    emit("attn_bwd (dq + dkv kernels) F=39 hs=16 H=2", ms, flops=B * 7 * (H * F * F * hs * 2))
    # ---- configs[4]: SASRec attention L = 200, d = 64, 1 head
    B, L, d = 1024, 200, 64
    qkv = torch.randn(B, L, d, device=dev, requires_grad=True)
    rm = (torch.rand(B, L, device=dev) < 0.8).float()
    with torch.no_grad():
        ms = timed(lambda: pkg.attention(qkv, qkv, qkv, 1, 0.125, row_mask=rm))
    emit("attn_fwd_kernel(K7) L=200 hs=64", ms, flops=B * 2 * (L * L * d * 2), nbytes=B * L * d * 4 * 4)
    out = pkg.attention(qkv, qkv, qkv, 1, 0.125, row_mask=rm)
    g = torch.randn_like(out)
    ms = timed(lambda: out.backward(g, retain_graph=True))
    emit("attn_bwd(K7: dq + dkv kernels) L=200 hs=64", ms, flops=B * 7 * (L * L * d * 2))
    # ---- YoutubeDNN sampled softmax: 1 M-item table, D = 64
    B, N, D = 4096, 1_000_000, 64
    W = torch.randn(N, D, device=dev) * 0.05
    xx = torch.randn(B, D, device=dev, requires_grad=True)
    labels = torch.randint(0, N, (B, 1), device=dev)
    for S in (5, 1024):
        smp, tries = pkg.layers.match.log_uniform_candidate_sampler(S, N, seed=1)
        te = pkg.layers.match.log_uniform_expected(labels, N, tries)
        se = pkg.layers.match.log_uniform_expected(smp, N, tries)
        with torch.no_grad():
            ms = timed(lambda: pkg.layers.sampled_softmax_loss(W, None, labels, xx, S, N, sampled_values=(smp, te, se)))
        emit(f"sampled_softmax fwd(K8) S={S}", ms, nbytes=B * (2 * D * 4 + 8 + 4) + S * D * 4,
             flops=B * (S + 1) * D * 2)
        ms = timed(lambda: pkg.layers.match.log_uniform_candidate_sampler(S, N, seed=2))
        emit(f"log_uniform_sample_kernel S={S}", ms)


if __name__ == "__main__":
    main()

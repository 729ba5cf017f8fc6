"""K5/K6/K7/K8 parity through the C-ABI against the numpy oracle (fp64 restatement of the
reference layer code, quirks kept).  Tolerance: 1e-5 relative on fp32 outputs (north star)."""
import math

import numpy as np
import pytest
import torch

from oracle import attention as OA

pytestmark = pytest.mark.gpu


def _close(got, want, rtol=1e-5, atol=2e-6):
    np.testing.assert_allclose(np.asarray(got, np.float64), want, rtol=rtol, atol=atol)


def _t(a, grad=False):
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()
    return t.requires_grad_(True) if grad else t


# ------------------------------------------------------------------ core attention (K7)
def _torch_attn(q, k, v, H, scale, row_mask=None, key_mask=None, causal=False):
    B, Lq, HS = q.shape
    Lk, hs = k.shape[1], HS // H
    sp = lambda t: t.reshape(B, t.shape[1], H, hs).transpose(1, 2)
    s = (sp(q) @ sp(k).transpose(-1, -2)) * scale
    pad = torch.full_like(s, OA.PAD)
    if row_mask is not None:
        s = torch.where(row_mask.reshape(B, 1, Lq, 1) == 0, pad, s)
    if key_mask is not None:
        s = torch.where(key_mask.reshape(B, 1, 1, Lk) == 0, pad, s)
    if causal:
        tri = torch.ones(Lq, Lk, device=q.device).tril().bool()
        s = torch.where(tri, s, pad)
    return (torch.softmax(s, -1) @ sp(v)).transpose(1, 2).reshape(B, Lq, HS)


@pytest.mark.parametrize("B,H,Lq,Lk,hs", [(5, 1, 200, 200, 64), (9, 2, 39, 39, 16), (3, 4, 17, 33, 8),
                                          (2, 1, 10, 10, 64), (4, 1, 256, 256, 32), (3, 2, 1, 7, 128),
                                          (6, 3, 100, 100, 20)])
@pytest.mark.parametrize("masks", ["none", "row", "key", "causal", "row+key"])
def test_attention_core_fwd_bwd_vs_torch(rtf, B, H, Lq, Lk, hs, masks):
    rng = np.random.default_rng(0)
    q, k, v = (_t(rng.normal(0, 1, (B, L, H * hs)), True) for L in (Lq, Lk, Lk))
    rm = km = None
    if "row" in masks:
        rm = _t((rng.random((B, Lq)) < 0.7).astype(np.float32))
        rm[0] = 0                                   # a fully padded sequence
    if "key" in masks:
        km = _t((rng.random((B, Lk)) < 0.7).astype(np.float32))
        km[-1] = 0
    causal = masks == "causal" and Lq == Lk
    scale = 1.0 / math.sqrt(hs)
    out = rtf.attention(q, k, v, H, scale, rm, km, causal)
    g = torch.randn_like(out)
    out.backward(g)
    q2, k2, v2 = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
    want = _torch_attn(q2, k2, v2, H, scale, rm, km, causal)
    want.backward(g)
    torch.testing.assert_close(out, want, rtol=1e-5, atol=2e-6)
    for a, b_ in ((q, q2), (k, k2), (v, v2)):
        torch.testing.assert_close(a.grad, b_.grad, rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize("B,H,L,hs,form", [(4, 1, 200, 64, "match"), (8, 2, 39, 16, "ctr"),
                                           (3, 1, 200, 64, "match-nomask")])
def test_attention_core_vs_fp64_oracle(rtf, B, H, L, hs, form):
    """The core kernel at the BASELINE shapes against the fp64 ORACLE (not torch fp32): forward
    vs oracle.match_sdpa (src/match/layers/modules.py:76-96, query-row mask) / the ctr form's
    QK^T * sqrt(hs) (src/ctr/layers/modules.py:235-239); gradients vs fp64 autograd over the same
    formula.  Tolerance 1e-5 relative (north star), scale-aware atol."""
    rng = np.random.default_rng(3)
    q, k, v = (rng.normal(0, 1, (B, L, H * hs)) for _ in range(3))
    mask = None
    if form == "match":
        lens = rng.integers(1, L + 1, B)
        mask = (np.arange(L)[None, :] >= (L - lens)[:, None]).astype(np.float64)   # pre-padding
        mask[0] = 0
    sp = lambda t: np.transpose(t.reshape(B, L, H, hs), (0, 2, 1, 3))              # noqa: E731
    if form.startswith("match"):
        scale = 1.0 / math.sqrt(hs)
        m4 = np.ones((B, H, L, 1)) if mask is None else np.tile(mask[:, None, :, None], (1, H, 1, 1))
        want = OA.match_sdpa(sp(q), sp(k), sp(v), m4)
    else:
        scale = math.sqrt(hs)                       # the reference DIVIDES by hs ** -0.5
        want = OA.softmax((sp(q) @ np.swapaxes(sp(k), -1, -2)) / (hs ** -0.5)) @ sp(v)
    want = np.transpose(want, (0, 2, 1, 3)).reshape(B, L, H * hs)
    tq, tk, tv = (_t(a, True) for a in (q, k, v))
    rm = None if mask is None else _t(mask)
    out = rtf.attention(tq, tk, tv, H, scale, rm, None, False)
    g = rng.normal(0, 1, want.shape)
    out.backward(_t(g))
    _close(out.detach().cpu().numpy(), want, rtol=1e-5, atol=1e-5 * np.abs(want).max())
    # fp64 gradients of the same formula
    dq, dk, dv = (torch.from_numpy(a).double().requires_grad_(True) for a in (q, k, v))
    tsp = lambda t: t.reshape(B, L, H, hs).transpose(1, 2)                         # noqa: E731
    s = (tsp(dq) @ tsp(dk).transpose(-1, -2)) * scale
    if mask is not None:
        s = torch.where(torch.from_numpy(mask).reshape(B, 1, L, 1) == 0, torch.full_like(s, OA.PAD), s)
    ref = (torch.softmax(s, -1) @ tsp(dv)).transpose(1, 2).reshape(B, L, H * hs)
    ref.backward(torch.from_numpy(g))
    for got, w in ((tq.grad, dq.grad), (tk.grad, dk.grad), (tv.grad, dv.grad)):
        w = w.numpy()
        _close(got.cpu().numpy(), w, rtol=1e-5, atol=1e-5 * np.abs(w).max())


def test_fully_masked_row_is_exactly_uniform(rtf):
    """A6: every logit = pad -> softmax is exactly 1/L (SURVEY §8c known answer)."""
    B, L, hs = 2, 50, 16
    q, k = torch.randn(B, L, hs, device="cuda"), torch.randn(B, L, hs, device="cuda")
    v = torch.randn(B, L, hs, device="cuda")
    rm = torch.zeros(B, L, device="cuda")
    out = rtf.attention(q, k, v, 1, 0.25, row_mask=rm)
    torch.testing.assert_close(out, v.mean(1, keepdim=True).expand(-1, L, -1), rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------ match MHA / encoder (a9)
@pytest.mark.parametrize("B,L,d,H", [(4, 200, 64, 1), (6, 10, 64, 1), (3, 50, 32, 4)])
def test_match_mha_and_transformer_encoder(rtf, B, L, d, H):
    rng = np.random.default_rng(1)
    x = rng.normal(0, 1, (B, L, d)).astype(np.float32)
    lens = rng.integers(1, L + 1, B)
    mask = (np.arange(L)[None, :] >= (L - lens)[:, None]).astype(np.float32)[:, :, None]  # left pad
    enc = rtf.layers.TransformerEncoder(d, num_heads=H, ffn_hidden_unit=2 * d)
    xt = _t(x, True)
    out = enc([xt, _t(mask)])
    m = enc.mha
    p = {"wq": m.wq.kernel, "bq": m.wq.bias, "wk": m.wk.kernel, "bk": m.wk.bias, "wv": m.wv.kernel,
         "bv": m.wv.bias, "ln1_g": enc.layernorm1.gamma, "ln1_b": enc.layernorm1.beta,
         "ln2_g": enc.layernorm2.gamma, "ln2_b": enc.layernorm2.beta, "w1": enc.ffn.conv1.kernel,
         "b1": enc.ffn.conv1.bias, "w2": enc.ffn.conv2.kernel, "b2": enc.ffn.conv2.bias}
    p = {k_: v_.detach().cpu().numpy() for k_, v_ in p.items()}
    want_att = OA.match_mha(x, x, x, mask, p["wq"], p["bq"], p["wk"], p["bk"], p["wv"], p["bv"], H)
    got_att = m(xt, xt, xt, _t(mask))
    _close(got_att.detach().cpu().numpy(), want_att, atol=5e-6)
    want = OA.transformer_encoder(x, mask, p, H)
    _close(out.detach().cpu().numpy(), want, rtol=2e-5, atol=2e-5)
    out.sum().backward()
    assert torch.isfinite(xt.grad).all()


# ------------------------------------------------------------------ ctr MHA / AutoInt layer (a7)
@pytest.mark.parametrize("scale", ["reference", "paper"])
@pytest.mark.parametrize("use_res", [False, True])
@pytest.mark.parametrize("B,F,dm,H,hs", [(16, 39, 16, 2, 16), (8, 10, 192, 1, 64), (5, 26, 8, 1, 8)])
def test_ctr_multihead_attention_autoint_layer(rtf, B, F, dm, H, hs, use_res, scale):
    rng = np.random.default_rng(2)
    x = rng.normal(0, 0.3, (B, F, dm)).astype(np.float32)
    layer = rtf.layers.ctr.MultiHeadAttention(hs, H, use_res=use_res, scale=scale)
    xt = _t(x, True)
    out = layer(xt)
    W = lambda dn: dn.kernel.detach().cpu().numpy()
    want = OA.ctr_mha(x, x, x, W(layer.q_dense), W(layer.k_dense), W(layer.v_dense), H, hs, "relu",
                      W(layer.res_dense) if use_res else None, scale)
    assert out.shape == (B, F, H * hs)
    _close(out.detach().cpu().numpy(), want, rtol=1e-5, atol=5e-6)
    # list forms of the source (:294-309)
    torch.testing.assert_close(layer([xt]), out)
    torch.testing.assert_close(layer([xt, xt, xt]), out)
    out.sum().backward()
    assert torch.isfinite(xt.grad).all()


def _ctr_mha_fp64(x, Ws, H, hs, act, scale, use_res):
    """fp64 torch restatement of src/ctr/layers/modules.py:255-270,211-240,281-283,316-323."""
    f = {"relu": torch.relu, "sigmoid": torch.sigmoid, "tanh": torch.tanh, None: lambda z: z}[act]
    B, F, _ = x.shape
    q, k, v = (f(x @ W) for W in Ws[:3])
    sp = lambda t: t.reshape(B, F, H, hs).transpose(1, 2)                          # noqa: E731
    div = hs ** -0.5 if scale == "reference" else hs ** 0.5
    o = (torch.softmax(sp(q) @ sp(k).transpose(-1, -2) / div, -1) @ sp(v)).transpose(1, 2).reshape(B, F, H * hs)
    return torch.relu(o + f(x @ Ws[3])) if use_res else o


@pytest.mark.parametrize("act", ["relu", "sigmoid", "tanh", None])
@pytest.mark.parametrize("use_res", [False, True])
@pytest.mark.parametrize("B,F,dm,H,hs", [(37, 39, 16, 2, 16), (300, 39, 32, 2, 16), (9, 64, 16, 1, 16),
                                         (5, 39, 16, 1, 8), (7, 13, 64, 2, 32), (1, 1, 8, 2, 8)])
def test_autoint_fused_layer_fwd_bwd_vs_fp64(rtf, B, F, dm, H, hs, use_res, act):
    """K6 (one launch per direction): output, dX and the four weight gradients against fp64
    autograd over the reference formula, 1e-5 relative; the fused path must really be taken."""
    assert rtf.lib().rtf_autoint_layer_supported(F, dm, H, hs) == 1
    rng = np.random.default_rng(11)
    x = rng.normal(0, 0.5, (B, F, dm))
    layer = rtf.layers.ctr.MultiHeadAttention(hs, H, activation=act, use_res=use_res)
    xt = _t(x, True)
    out = layer(xt)
    assert type(out.grad_fn).__name__ == "_AutoIntLayerFnBackward"
    g = rng.normal(0, 1, (B, F, H * hs))
    out.backward(_t(g))
    dens = [layer.q_dense, layer.k_dense, layer.v_dense] + ([layer.res_dense] if use_res else [])
    Ws = [d.kernel.detach().cpu().double().requires_grad_(True) for d in dens]
    x64 = torch.from_numpy(x).requires_grad_(True)
    ref = _ctr_mha_fp64(x64, Ws + ([None] if not use_res else []), H, hs, act, "reference", use_res)
    ref.backward(torch.from_numpy(g))

    def close(got, want):
        want = want.detach().numpy()
        _close(got.detach().cpu().numpy(), want, rtol=1e-5, atol=1e-5 * max(np.abs(want).max(), 1e-30))
    close(out, ref)
    w = x64.grad.numpy()            # 128-term fp32 dots of O(1) values: 2e-5 of the largest entry
    _close(xt.grad.cpu().numpy(), w, rtol=1e-5, atol=2e-5 * np.abs(w).max())
    for d, W in zip(dens, Ws):      # a fp32 reduction over B*F terms: tolerance scaled accordingly
        w = W.grad.numpy()
        _close(d.kernel.grad.cpu().numpy(), w, rtol=1e-5, atol=5e-5 * np.abs(w).max())


def test_autoint_fused_layer_matches_unfused_and_is_deterministic(rtf):
    rng = np.random.default_rng(5)
    x = _t(rng.normal(0, 0.5, (513, 39, 16)))
    fused = rtf.layers.ctr.MultiHeadAttention(16, 2, use_res=True)
    plain = rtf.layers.ctr.MultiHeadAttention(16, 2, use_res=True)
    plain.fused = False
    fused(x), plain(x)
    for a, b_ in zip(fused.parameters(), plain.parameters()):
        b_.data.copy_(a.data)
    g = torch.randn(513, 39, 32, device="cuda")
    grads = []
    for layer in (fused, plain, fused):
        xi = x.clone().requires_grad_(True)
        for p in layer.parameters():
            p.grad = None
        out = layer(xi)
        out.backward(g)
        grads.append([out.detach(), xi.grad] + [p.grad.clone() for p in layer.parameters()])
    for a, b_ in zip(grads[0], grads[1]):      # two fp32 summation orders of B*F = 20 007 terms
        torch.testing.assert_close(a, b_, rtol=1e-3, atol=1e-4 * float(b_.abs().max()))
    for a, b_ in zip(grads[0], grads[2]):                      # fixed reduction order: bit-equal
        assert torch.equal(a, b_)


# ------------------------------------------------------------------ DIN local activation unit (a6)
@pytest.mark.parametrize("act", ["sigmoid", None, "relu", "tanh"])
@pytest.mark.parametrize("B,L,d", [(33, 100, 16), (9, 100, 128), (5, 10, 192), (64, 7, 8), (3, 37, 48)])
def test_din_attention_layer_fwd_bwd(rtf, B, L, d, act):
    rng = np.random.default_rng(3)
    q = rng.normal(0, 0.5, (B, d)).astype(np.float32)
    k = rng.normal(0, 0.5, (B, L, d)).astype(np.float32)
    lens = rng.integers(1, L + 1, B)
    mask = (np.arange(L)[None, :] < lens[:, None]).astype(np.float32)
    layer = rtf.layers.AttentionLayer(1, activation=act)
    qt, kt = _t(q, True), _t(k, True)
    out = layer([qt, kt, kt, _t(mask)])
    W = layer.att_dense_kernel.detach().cpu().numpy()
    bias = layer.att_dense_bias.detach().cpu().numpy()
    want = OA.din_attention_layer(q, k, k, mask, W, bias, act)
    _close(out.detach().cpu().numpy(), want, rtol=1e-5, atol=2e-6)
    # gradients vs torch autograd over the reference's own formulation (tile/concat/dense)
    g = torch.randn_like(out)
    out.backward(g)
    q2, k2 = _t(q, True), _t(k, True)
    W2, b2 = layer.att_dense_kernel.detach().clone().requires_grad_(True), \
        layer.att_dense_bias.detach().clone().requires_grad_(True)
    qq = q2.repeat(1, L).reshape(-1, L, d)
    info = torch.cat([qq, k2, qq - k2, qq * k2], -1)
    s = info @ W2 + b2
    s = {"sigmoid": torch.sigmoid, None: lambda z: z, "relu": torch.relu, "tanh": torch.tanh}[act](s)
    s = s.reshape(-1, L)
    s = torch.where(_t(mask) == 0, torch.full_like(s, OA.PAD), s)
    o2 = (torch.softmax(s, -1).unsqueeze(1) @ k2).squeeze(1)
    o2.backward(g)
    torch.testing.assert_close(qt.grad, q2.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(kt.grad, k2.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(layer.att_dense_kernel.grad, W2.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(layer.att_dense_bias.grad, b2.grad, rtol=1e-4, atol=1e-5)


def test_din_attention_quirks(rtf):
    """mask not a tensor -> uniform weights (:164-165); separate v tensor; all-masked row;
    the default activation string 'prelu' is rejected as Keras rejects it."""
    rng = np.random.default_rng(4)
    B, L, d = 6, 20, 16
    q, k, v = (rng.normal(0, 1, s).astype(np.float32) for s in ((B, d), (B, L, d), (B, L, d)))
    layer = rtf.layers.AttentionLayer(1, activation="sigmoid")
    out = layer([_t(q), _t(k), _t(v), None])
    _close(out.detach().cpu().numpy(), v.astype(np.float64).mean(1), atol=1e-6)
    mask = np.ones((B, L), np.float32)
    mask[0] = 0
    out = layer([_t(q), _t(k), _t(v), _t(mask)])
    W, b = layer.att_dense_kernel.detach().cpu().numpy(), layer.att_dense_bias.detach().cpu().numpy()
    _close(out.detach().cpu().numpy(), OA.din_attention_layer(q, k, v, mask, W, b, "sigmoid"), atol=2e-6)
    with pytest.raises(ValueError):
        rtf.layers.AttentionLayer(1)            # activation='prelu'
    with pytest.raises(ValueError):
        rtf.layers.AttentionLayer(8, activation="sigmoid")


# ------------------------------------------------------------------ sampled softmax (a11)
def test_sampled_softmax_gemm_form_is_taken_and_matches_streaming_form(rtf):
    """S >= 32: logits through the tensor-core GEMM + epilogues; same loss and gradients as the
    streaming kernel (rtf_sampled_softmax_*)."""
    from recommend_tf2_b200.layers import match as M
    torch.manual_seed(3)
    B, N, S, D = 256, 5000, 64, 16
    W = (torch.randn(N, D, device="cuda") * 0.1)
    x = torch.randn(B, D, device="cuda")
    labels = torch.randint(0, N, (B,), device="cuda")
    smp, tries = M.log_uniform_candidate_sampler(S, N, seed=3)
    labels[:8] = smp[:8]
    te, se = M.log_uniform_expected(labels, N, tries), M.log_uniform_expected(smp, N, tries)
    res = []
    for fn in (M._SampledSoftmaxGemmFn, M._SampledSoftmaxFn):
        Wi, xi = W.clone().requires_grad_(True), x.clone().requires_grad_(True)
        loss = fn.apply(Wi, None, labels, xi, smp, te, se, True, None, None)
        loss.sum().backward()
        res.append((loss.detach(), xi.grad, Wi.grad))
    assert M._gemm_form_ok(B, S, D)
    for a, b_ in zip(*res):
        torch.testing.assert_close(a, b_, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("B,N,S,D", [(64, 1000, 5, 32), (128, 100000, 1024, 64), (17, 50, 20, 10),
                                     (256, 5000, 64, 16)])
def test_sampled_softmax_injected_samples(rtf, B, N, S, D):
    rng = np.random.default_rng(5)
    W = rng.normal(0, 0.1, (N, D)).astype(np.float32)
    bias = rng.normal(0, 0.1, N).astype(np.float32)
    x = rng.normal(0, 1, (B, D)).astype(np.float32)
    sampled = rng.choice(N, S, replace=False)
    labels = rng.integers(0, N, B)
    labels[: min(B, S) // 2] = sampled[: min(B, S) // 2]          # accidental hits
    te = OA.log_uniform_expected(labels, N, 2 * S).astype(np.float32)
    se = OA.log_uniform_expected(sampled, N, 2 * S).astype(np.float32)
    want = OA.sampled_softmax_loss(W, bias, labels, x, sampled, te, se)
    Wt, xt = _t(W, True), _t(x, True)
    sv = (torch.from_numpy(sampled).cuda(), _t(te), _t(se))
    loss = rtf.layers.sampled_softmax_loss(Wt, _t(bias), torch.from_numpy(labels).cuda().view(-1, 1), xt,
                                           S, N, sampled_values=sv)
    _close(loss.detach().cpu().numpy(), want, rtol=1e-5, atol=1e-5)
    loss.sum().backward()
    # gradient vs torch autograd on the same formula
    W2, x2 = _t(W, True), _t(x, True)
    lab, smp = torch.from_numpy(labels).cuda(), torch.from_numpy(sampled).cuda()
    tl = (x2 * W2[lab]).sum(1) + _t(bias)[lab] - torch.log(_t(te))
    sl = x2 @ W2[smp].t() + _t(bias)[smp]
    sl = sl + torch.where(lab[:, None] == smp[None, :], -torch.finfo(torch.float32).max, 0.0)
    sl = sl - torch.log(_t(se))[None, :]
    lg = torch.cat([tl[:, None], sl], 1)
    (torch.logsumexp(lg, 1) - lg[:, 0]).sum().backward()
    torch.testing.assert_close(xt.grad, x2.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(Wt.grad, W2.grad, rtol=1e-4, atol=1e-5)


def test_log_uniform_sampler_properties(rtf):
    S, N = 1000, 1_000_000
    smp, tries = rtf.layers.match.log_uniform_candidate_sampler(S, N, seed=7)
    smp2, _ = rtf.layers.match.log_uniform_candidate_sampler(S, N, seed=7)
    assert torch.equal(smp, smp2)                              # deterministic given the seed
    a = smp.cpu().numpy()
    assert len(np.unique(a)) == S and a.min() >= 0 and a.max() < N and int(tries) >= S
    assert np.median(a) < N / 20                               # Zipfian: mass on small ids
    exp = rtf.layers.match.log_uniform_expected(smp, N, tries).cpu().numpy()
    _close(exp, OA.log_uniform_expected(a, N, int(tries)), rtol=1e-5, atol=1e-7)
    # statistical check of the distribution: P(c < 10) = log(11)/log(N+1)
    big, _ = rtf.layers.match.log_uniform_candidate_sampler(5, 10, seed=1)
    assert len(np.unique(big.cpu().numpy())) == 5


def test_sampled_softmax_layer_reference_form_and_sasrec_youtubednn(rtf):
    """SampledSoftmaxLayer with the source's odd wiring (weights = in-batch item tower output,
    num_classes = tower width) + the two matching models end to end."""
    torch.manual_seed(0)
    B, n = 64, 32
    item = torch.randn(B, 1, n, device="cuda", requires_grad=True)
    user = torch.randn(B, 1, n, device="cuda", requires_grad=True)
    labels = torch.randint(0, 2, (B, 1), device="cuda")
    layer = rtf.layers.SampledSoftmaxLayer(num_sampled=1)
    loss = layer([item, user, labels])
    assert loss.shape == (B, 1) and torch.isfinite(loss).all()
    rtf.layers.sampledsoftmaxloss(None, loss).backward()
    assert item.grad.abs().sum() > 0 and user.grad.abs().sum() > 0

    from recommend_tf2_b200.models import SASRec, YoutubeDNN
    # SASRec shape fixture of the reference (src/match/sasrec/model.py:122-127)
    m = SASRec(item_num=100, embed_dim=64, blocks=2, seq_len=10, neg_len=100, seed=0)
    seq = torch.randint(1, 100, (8, 10), device="cuda", dtype=torch.int32)
    seq[:, :4] = 0                                                   # left padding
    pos = torch.randint(1, 100, (8, 1), device="cuda", dtype=torch.int32)
    neg = torch.randint(1, 100, (8, 100), device="cuda", dtype=torch.int32)
    logits, l = m([seq, pos, neg])
    assert logits.shape == (8, 101) and torch.isfinite(l)
    l.backward()
    # oracle check of the scoring tail on the model's own activations
    y = YoutubeDNN([100, 50], item_num=1000, embed_dim=8, user_dnn_hidden_units=(64, 32), num_sampled=5)
    users = torch.randint(0, 50, (16, 2), device="cuda", dtype=torch.int32)
    items = torch.randint(0, 1000, (16,), device="cuda")
    out = y([users, items])
    assert out.shape == (16, 1)
    out.mean().backward()
    assert y.item_table.weights[0].grad is not None


def test_sasrec_tail_matches_oracle(rtf):
    rng = np.random.default_rng(6)
    att = rng.normal(0, 1, (5, 10, 16))
    pos, neg = rng.normal(0, 1, (5, 1, 16)), rng.normal(0, 1, (5, 7, 16))
    want_logits, want_loss = OA.sasrec_scores_loss(att, pos, neg)
    a, p, n = (_t(t) for t in (att, pos, neg))
    si = a[:, -1:, :]
    ps, ns = (si * p).sum(-1), (si * n).sum(-1)
    loss = (-torch.log(torch.sigmoid(ps)) - torch.log(1 - torch.sigmoid(ns))).mean() / 2
    _close(torch.cat([ps, ns], -1).cpu().numpy(), want_logits, atol=1e-5)
    _close(float(loss), want_loss, rtol=1e-5)


@pytest.mark.parametrize("B,NEG,D,N", [(5, 7, 16, 50), (33, 100, 64, 1000), (2, 1, 8, 3), (16, 100, 128, 40)])
def test_sasrec_score_kernel_vs_oracle_fwd_bwd(rtf, B, NEG, D, N):
    """a10 epilogue kernel (gathers + dots + log loss): logits / loss vs the oracle
    (src/match/sasrec/model.py:88-96), d seq_info and the per-table row gradients vs fp64
    autograd over the same formula (duplicate ids included: K2 sums them)."""
    from recommend_tf2_b200.models import _SasrecScoreFn
    rng = np.random.default_rng(8)
    tabs = [rng.normal(0, 0.5, (N, D)).astype(np.float32) for _ in range(3)]
    info = rng.normal(0, 0.5, (B, D)).astype(np.float32)
    pos = rng.integers(0, N, (B, 1))
    neg = rng.integers(0, N, (B, NEG))
    ts = rtf.EmbeddingTables.from_tensors([_t(t) for t in tabs])
    for w in ts.weights:
        w.requires_grad_(True)
    it = _t(info, True)
    logits, loss = _SasrecScoreFn.apply(ts, it, _t(pos).to(torch.int32), _t(neg).to(torch.int32), 1, 2,
                                        *ts.weights)
    att = info[:, None, :].astype(np.float64)
    want_logits, want_loss = OA.sasrec_scores_loss(att, tabs[1][pos].astype(np.float64),
                                                   tabs[2][neg].astype(np.float64))
    _close(logits.detach().cpu().numpy(), want_logits, rtol=1e-5, atol=1e-5 * np.abs(want_logits).max())
    _close(float(loss), want_loss, rtol=1e-5)
    gl = rng.normal(0, 1, (B, 1 + NEG))
    (loss * 3.0 + (logits * _t(gl)).sum()).backward()
    i64 = torch.from_numpy(info).double().requires_grad_(True)
    T64 = [torch.from_numpy(t).double().requires_grad_(True) for t in tabs]
    pe, ne = T64[1][torch.from_numpy(pos)], T64[2][torch.from_numpy(neg)]
    ps, ns = (i64[:, None, :] * pe).sum(-1), (i64[:, None, :] * ne).sum(-1)
    l64 = (-torch.log(torch.sigmoid(ps)) - torch.log(1 - torch.sigmoid(ns))).mean() / 2
    (l64 * 3.0 + (torch.cat([ps, ns], -1) * torch.from_numpy(gl)).sum()).backward()
    w = i64.grad.numpy()
    _close(it.grad.cpu().numpy(), w, rtol=1e-5, atol=1e-5 * np.abs(w).max())
    for t in (1, 2):
        w = T64[t].grad.numpy()
        _close(ts.weights[t].grad.to_dense().cpu().numpy(), w, rtol=1e-5, atol=1e-5 * np.abs(w).max())
    assert ts.weights[0].grad is None


def test_sasrec_fused_scores_match_unfused_training(rtf):
    from recommend_tf2_b200.models import SASRec
    torch.manual_seed(123)          # fixed inputs: three Adam steps at lr 1e-2 amplify any difference
    seq = torch.randint(1, 100, (8, 10), device="cuda", dtype=torch.int32)
    seq[:, :4] = 0
    pos = torch.randint(1, 100, (8, 1), device="cuda", dtype=torch.int32)
    neg = torch.randint(1, 100, (8, 100), device="cuda", dtype=torch.int32)
    outs = []
    for fused in (True, False):
        m = SASRec(item_num=100, embed_dim=64, blocks=2, seq_len=10, neg_len=100, seed=0)
        m.fused_scores = fused
        torch.manual_seed(0)
        tr = rtf.models.Trainer(m, lambda out, y: out[1], lr=1e-2)
        losses = [float(tr.step([seq, pos, neg])) for _ in range(3)]
        outs.append((losses, [w.detach().clone() for w in m.tables.weights]))
    # two fp32 formulations trained for 3 Adam steps at lr 1e-2: rounding differences grow a little
    np.testing.assert_allclose(outs[0][0], outs[1][0], rtol=1e-4)
    for a, b_ in zip(outs[0][1], outs[1][1]):
        torch.testing.assert_close(a, b_, rtol=1e-3, atol=1e-5)


def test_dice_pooling_and_models_smoke(rtf):
    x = torch.randn(64, 40, device="cuda")
    dice = rtf.layers.Dice()
    y = dice(x)
    bn = (x - x.mean(0)) / torch.sqrt(x.var(0, unbiased=False) + 1e-3)
    p = torch.sigmoid(bn)
    torch.testing.assert_close(y, dice.alpha * (1 - p) * x + p * x, rtol=1e-4, atol=1e-5)
    assert -math.sqrt(3) <= float(dice.alpha) <= math.sqrt(3)
    a, b_ = torch.randn(4, 3, device="cuda"), torch.randn(4, 3, device="cuda")
    pl = rtf.layers.PoolingLayer("mean")
    assert pl(a) is a and torch.equal(pl([a]), a)
    torch.testing.assert_close(pl([a, b_]), (a + b_) / 2)
    torch.testing.assert_close(rtf.layers.PoolingLayer("max")([a, b_]), torch.maximum(a, b_))
    with pytest.raises(ValueError):
        rtf.layers.PoolingLayer("min")

    from recommend_tf2_b200.models import DIN, AutoInt, DeepFM
    fc = [[{"feat": f"I{i}"} for i in range(13)],
          [{"feat": f"C{i}", "feat_num": 100, "embed_dim": 8} for i in range(26)]]
    dense = torch.rand(32, 13, device="cuda")
    sparse = torch.randint(0, 100, (32, 26), device="cuda", dtype=torch.int32)
    for model in (DeepFM(fc), AutoInt([fc[0], [dict(c, embed_dim=16) for c in fc[1]]])):
        out = model([dense, sparse])
        assert out.shape == (32, 1)
        out.sum().backward()
    din = DIN([1000, 50], embed_dim=8, maxlen=100)
    hist = torch.stack([torch.randint(1, 1000, (16, 100), device="cuda"),
                        torch.randint(1, 50, (16, 100), device="cuda")], -1).to(torch.int32)
    hist[:, 60:] = 0
    target = torch.stack([torch.randint(1, 1000, (16,), device="cuda"),
                          torch.randint(1, 50, (16,), device="cuda")], -1).to(torch.int32)
    out = din([hist, target])
    assert out.shape == (16, 1)
    out.sum().backward()

"""K3 parity: FM layer (reference quirks + paper variant) and the FM model in gather form
against the one-hot restatement of the reference.  fp32 outputs/gradients within 1e-5 rel."""
import numpy as np
import pytest
import torch

from oracle import interaction as OI

pytestmark = pytest.mark.gpu


def _close(got, want, rtol=1e-5, atol=1e-6):
    np.testing.assert_allclose(np.asarray(got, np.float64), want, rtol=rtol, atol=atol)


@pytest.mark.parametrize("B,P1,M", [(1024, 221, 208), (7, 5, 3), (300, 40, 64)])
def test_fm_layer_reference_2d_deepfm_shapes(rtf, B, P1, M):
    """DeepFM passes first (B, 13+26*8) and a 2-D second (B, 26*8): src/ctr/deep_fm/model.py:56-59."""
    from recommend_tf2_b200.fm import FM
    rng = np.random.default_rng(0)
    first = rng.normal(0, 0.3, (B, P1)).astype(np.float32)
    second = rng.normal(0, 0.3, (B, M)).astype(np.float32)
    layer = FM(P1)
    tf_, ts_ = torch.from_numpy(first).cuda().requires_grad_(True), torch.from_numpy(second).cuda().requires_grad_(True)
    out = layer([tf_, ts_])
    w = layer.w.detach().cpu().numpy()
    want = OI.fm_layer(first, second, w)
    assert out.shape == (B, 1)
    _close(out.detach().cpu().numpy(), want, atol=1e-5 * max(1.0, np.abs(want).max()))
    # gradients against torch autograd over the same formula
    g = torch.randn(B, 1, device="cuda")
    out.backward(g)
    f2, s2, w2 = (t.detach().clone().requires_grad_(True) for t in (tf_, ts_, layer.w))
    fo = (f2 @ w2).sum()
    so = 0.5 * (s2.sum(1, keepdim=True) ** 2 - (s2 ** 2).sum(1, keepdim=True)).sum(1)
    ((fo + so).reshape(-1, 1) * g).sum().backward()
    torch.testing.assert_close(tf_.grad, f2.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(ts_.grad, s2.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(layer.w.grad, w2.grad, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("D", [8, 16, 1, 40, 128])
def test_fm_layer_reference_3d_shape_quirk(rtf, D):
    """A 3-D (B,F,D) second_inputs yields (B*D, 1) in the reference (modules.py:70-71)."""
    from recommend_tf2_b200.fm import FM
    rng = np.random.default_rng(1)
    B, F, P1 = 33, 26, 50
    first = rng.normal(0, 0.3, (B, P1)).astype(np.float32)
    second = rng.normal(0, 0.3, (B, F, D)).astype(np.float32)
    layer = FM(P1)
    out = layer([torch.from_numpy(first).cuda(), torch.from_numpy(second).cuda()])
    want = OI.fm_layer(first, second, layer.w.detach().cpu().numpy())
    assert out.shape == want.shape == (B * D, 1)
    _close(out.detach().cpu().numpy(), want, atol=1e-5 * max(1.0, np.abs(want).max()))


@pytest.mark.parametrize("D", [8, 24, 64])
def test_fm_layer_paper_mode_and_pairwise_identity(rtf, D):
    from recommend_tf2_b200.fm import FM
    rng = np.random.default_rng(2)
    B, F, P1 = 65, 26, 39
    first = rng.normal(0, 0.3, (B, P1)).astype(np.float32)
    second = rng.normal(0, 0.3, (B, F, D)).astype(np.float32)
    layer = FM(P1, mode="paper")
    ts_ = torch.from_numpy(second).cuda().requires_grad_(True)
    out = layer([torch.from_numpy(first).cuda(), ts_])
    w = layer.w.detach().cpu().numpy()
    want = OI.fm_layer_paper(first, second, w)
    _close(out.detach().cpu().numpy(), want, atol=1e-5)
    # known-answer identity: 0.5((sum x)^2 - sum x^2) == sum_{i<j} <x_i, x_j>
    x = second.astype(np.float64)
    pair = sum((x[:, i] * x[:, j]).sum(-1) for i in range(F) for j in range(i))
    _close(out.detach().cpu().numpy()[:, 0] - (first.astype(np.float64) @ w.astype(np.float64))[:, 0], pair, atol=1e-5)
    out.sum().backward()
    gx = x.sum(1, keepdims=True) - x
    _close(ts_.grad.cpu().numpy(), gx, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("k", [8, 10, 3])
def test_fm_model_gather_equals_onehot_reference(rtf, k):
    """ctr.fm.model.FM: gather form == the reference's dense one-hot matmuls."""
    from recommend_tf2_b200.fm import FMModel
    rng = np.random.default_rng(3)
    B, nd = 257, 13
    feat_nums = [int(n) for n in rng.integers(2, 60, 26)]
    fc = [[{"feat": f"I{i}"} for i in range(nd)],
          [{"feat": f"C{i}", "feat_num": n, "embed_dim": k} for i, n in enumerate(feat_nums)]]
    model = FMModel(fc, k=k, seed=0)
    M = nd + sum(feat_nums)
    w0 = rng.normal(0, 0.1, 1).astype(np.float32)
    w = rng.normal(0, 0.05, (M, 1)).astype(np.float32)
    V = rng.normal(0, 0.05, (k, M)).astype(np.float32)
    model.load_reference_weights(w0, w, V)
    r0, rw, rV = model.reference_weights()
    assert np.array_equal(rw.cpu().numpy(), w) and np.array_equal(rV.cpu().numpy(), V)
    dense = rng.random((B, nd), dtype=np.float32)
    sparse = np.stack([rng.integers(0, n, B) for n in feat_nums], 1).astype(np.int32)
    out = model([torch.from_numpy(dense).cuda(), torch.from_numpy(sparse).cuda()])
    want = OI.fm_model_onehot(dense, sparse, feat_nums, w0, w, V)
    assert out.shape == (B, 1)
    _close(out.detach().cpu().numpy(), want, rtol=1e-5, atol=1e-6)

    # gradients: torch autograd through the one-hot formulation on the same weights
    g = torch.randn(B, 1, device="cuda")
    out.backward(g)
    tw0, tw, tV = (torch.from_numpy(a).cuda().requires_grad_(True) for a in (w0, w, V))
    hots = [torch.nn.functional.one_hot(torch.from_numpy(sparse[:, i]).long().cuda(), n).float()
            for i, n in enumerate(feat_nums)]
    stack = torch.cat([torch.from_numpy(dense).cuda()] + hots, -1)
    first = tw0 + stack @ tw
    second = 0.5 * ((stack @ tV.t()) ** 2 - (stack ** 2) @ (tV.t() ** 2)).sum(1, keepdim=True)
    (torch.sigmoid(first + second) * g).sum().backward()
    rows_grad = torch.cat([model.dense_table.grad] + [t.grad.to_dense() for t in model.tables.weights], 0)
    torch.testing.assert_close(rows_grad[:, k], tw.grad[:, 0], rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(rows_grad[:, :k], tV.grad.t(), rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(model.w0.grad, tw0.grad, rtol=1e-4, atol=1e-6)
    assert not rows_grad[:, k + 1:].any()


def test_fm_model_criteo_shaped_config0(rtf):
    """BASELINE configs[0]: 13 dense + 26 sparse fields, k = 8, batch 1024, fwd + bwd with the
    fused sparse optimizer."""
    from recommend_tf2_b200.fm import FMModel
    torch.manual_seed(0)
    B, k = 1024, 8
    fc = [[{"feat": f"I{i}"} for i in range(13)],
          [{"feat": f"C{i}", "feat_num": 1000, "embed_dim": k} for i in range(26)]]
    model = FMModel(fc, k=k, seed=0, sparse_optimizer=rtf.SparseOptimizer("adam", lr=1e-2))
    dense = torch.rand(B, 13, device="cuda")
    sparse = torch.randint(0, 1000, (B, 26), device="cuda", dtype=torch.int32)
    y = ((sparse[:, 0] % 2) == 0).float().unsqueeze(1)
    opt = torch.optim.Adam([model.w0, model.dense_table], lr=1e-2, eps=1e-7)
    losses = []
    for _ in range(80):
        model.tables.begin_step()
        opt.zero_grad()
        loss = rtf.layers.binary_crossentropy(y, model([dense, sparse]))
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < 0.6 * losses[0], losses[::10]


def test_colsum_is_deterministic_and_exact_order(rtf):
    from recommend_tf2_b200.fm import colsum
    x = torch.randn(1000, 37, device="cuda")
    s = torch.randn(1000, device="cuda")
    a, b = colsum(x, s), colsum(x, s)
    assert torch.equal(a, b)
    torch.testing.assert_close(a, (x * s[:, None]).sum(0), rtol=1e-4, atol=1e-4)
    # oracle of the exact order: chunks of 256 rows sequential, then chunks in order
    xs = (x * s[:, None]).cpu().numpy()
    parts = []
    for c0 in range(0, 1000, 256):
        acc = np.zeros(37, np.float32)
        for r in xs[c0:c0 + 256]:
            acc = acc + r
        parts.append(acc)
    tot = np.zeros(37, np.float32)
    for p in parts:
        tot = tot + p
    assert np.array_equal(a.cpu().numpy(), tot)


def test_dense_backward_epilogue_matches_autograd(rtf):
    """Dense (bias+ReLU epilogue, fused ReLU-mask + bias-grad kernel, split-K weight grad) ==
    plain torch autograd of relu(x W + b)."""
    torch.manual_seed(0)
    for B, K, N in [(16384, 64, 128), (300, 20, 12), (8192, 480, 256)]:
        layer = rtf.layers.Dense(N, activation="relu")
        x = torch.randn(B, K, device="cuda", requires_grad=True)
        y = layer(x)
        g = torch.randn_like(y)
        y.backward(g)
        x2 = x.detach().clone().requires_grad_(True)
        W, b = layer.kernel.detach().clone().requires_grad_(True), layer.bias.detach().clone().requires_grad_(True)
        y2 = torch.relu(x2 @ W + b)
        y2.backward(g)
        torch.testing.assert_close(y, y2, rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(x.grad, x2.grad, rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(layer.kernel.grad, W.grad, rtol=1e-4, atol=1e-3)
        torch.testing.assert_close(layer.bias.grad, b.grad, rtol=1e-4, atol=1e-3)
        layer.kernel.grad = layer.bias.grad = None

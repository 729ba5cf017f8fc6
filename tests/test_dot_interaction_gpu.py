"""K4 parity: DLRM pairwise dot interaction (stacked and fused-with-gather), fwd + bwd,
through the C-ABI against the fp64 oracle.  Tolerance: fp32 outputs/gradients within 1e-5
relative (north star), measured against the natural scale |x_i||x_j| of each dot product."""
import numpy as np
import pytest
import torch

from oracle import embedding as OE
from oracle import interaction as OI

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def _assert_close(got, want, scale):
    err = np.abs(got.astype(np.float64) - want)
    tol = RTOL * np.maximum(np.abs(want), scale)
    assert (err <= tol).all(), f"max err {err.max():.3e} vs tol {tol.flat[err.argmax()]:.3e}"


@pytest.mark.parametrize("B,F1,D", [(64, 27, 128), (33, 27, 64), (17, 4, 16), (5, 2, 8),
                                    (40, 40, 32), (9, 64, 12), (3, 27, 256), (200, 14, 128)])
def test_dot_stacked_fwd_bwd(rtf, B, F1, D):
    rng = np.random.default_rng(0)
    x = rng.normal(0, 1, (B, F1, D)).astype(np.float32)
    g = rng.normal(0, 1, (B, D + F1 * (F1 - 1) // 2)).astype(np.float32)
    want = OI.dot_interact(x)
    want_gx = OI.dot_interact_bwd(x, g)
    xt = torch.from_numpy(x).cuda().requires_grad_(True)
    out = rtf.dot_interact(xt)
    assert out.shape == want.shape
    nrm = np.linalg.norm(x.astype(np.float64), axis=-1)
    scale = np.concatenate([np.abs(x[:, 0, :]).astype(np.float64),
                            np.stack([nrm[:, i] * nrm[:, j] for i, j in zip(*np.tril_indices(F1, -1))], 1)
                            .reshape(B, -1)], 1)
    _assert_close(out.detach().cpu().numpy(), want, scale)
    out.backward(torch.from_numpy(g).cuda())
    gscale = (np.abs(g).max() * np.abs(x).sum(1, keepdims=True)).astype(np.float64) + np.abs(g).max()
    _assert_close(xt.grad.cpu().numpy(), want_gx, np.broadcast_to(gscale, want_gx.shape))
    # row 0 passes through bit-exactly
    assert np.array_equal(out.detach().cpu().numpy()[:, :D], x[:, 0, :])


def test_dot_padding_columns_are_zero(rtf):
    x = torch.randn(8, 27, 128, device="cuda")
    out = rtf.dot_interact(x, pad_to=8)
    assert out.shape == (8, 480)
    assert torch.equal(out[:, :479], rtf.dot_interact(x))
    assert not out[:, 479:].any()


def test_dot_matches_torch_fp32_reference(rtf):
    """plain PyTorch fp32 reference of the same op (floating-point kernel)."""
    x = torch.randn(128, 27, 128, device="cuda") * 0.05
    z = torch.bmm(x, x.transpose(1, 2))
    ii, jj = torch.tril_indices(27, 27, -1)
    want = torch.cat([x[:, 0], z[:, ii, jj]], 1)
    torch.testing.assert_close(rtf.dot_interact(x), want, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("idt", [torch.int32, torch.int64])
@pytest.mark.parametrize("B,F,D", [(130, 26, 128), (31, 3, 16), (64, 26, 64)])
def test_fused_gather_dot_equals_gather_then_dot(rtf, B, F, D, idt):
    rng = np.random.default_rng(1)
    rows = [int(r) for r in rng.integers(3, 5000, F)]
    tabs = [rng.uniform(-0.05, 0.05, (n, D)).astype(np.float32) for n in rows]
    ids = np.stack([(rng.zipf(1.05, B) - 1) % n for n in rows], 1)
    dense = rng.normal(0, 0.1, (B, D)).astype(np.float32)
    emb = OE.embed_lookup_concat(tabs, ids[:, :, None])[:, 0].reshape(B, F, D)
    x = np.concatenate([dense[:, None, :], emb], 1)
    want = OI.dot_interact(x)
    g = rng.normal(0, 1, want.shape).astype(np.float32)
    want_gx = OI.dot_interact_bwd(x, g)

    ts = rtf.EmbeddingTables(rows, [D] * F)
    with torch.no_grad():
        for w, t in zip(ts.weights, tabs):
            w.copy_(torch.from_numpy(t))
    dt = torch.from_numpy(dense).cuda().requires_grad_(True)
    dids = torch.from_numpy(ids).to(idt).cuda()
    out = rtf.embed_dot(ts, dids, dt)
    # fused == unfused product path, bit for bit (same kernel body, same summation order)
    unf = rtf.dot_interact(torch.from_numpy(x).cuda())
    assert torch.equal(out, unf)
    nrm = np.linalg.norm(x.astype(np.float64), axis=-1)
    scale = np.concatenate([np.abs(dense).astype(np.float64),
                            np.stack([nrm[:, i] * nrm[:, j] for i, j in zip(*np.tril_indices(F + 1, -1))], 1)], 1)
    _assert_close(out.detach().cpu().numpy(), want, scale)

    out.backward(torch.from_numpy(g).cuda())
    gs = (np.abs(g).max() * np.abs(x).sum(1, keepdims=True)).astype(np.float64) + np.abs(g).max()
    _assert_close(dt.grad.cpu().numpy(), want_gx[:, 0], np.broadcast_to(gs[:, 0], (B, D)))
    # table gradients: scatter the oracle's dX rows and compare with the sparse grads
    dense_g = OE.dense_reference_grad([(n, D) for n in rows], ids[:, :, None], list(range(F)),
                                      want_gx[:, 1:].reshape(B, 1, F * D))
    for t, w in enumerate(ts.weights):
        got = w.grad.to_dense().cpu().numpy()
        tol = RTOL * np.maximum(np.abs(dense_g[t]), gs.max() * 4)
        assert (np.abs(got - dense_g[t]) <= tol).all()
    ts.check_ids()


def test_fused_bad_id_reads_zero_row(rtf):
    ts = rtf.EmbeddingTables([10, 10], [8, 8], seed=0)
    ids = torch.tensor([[1, 2], [11, 3]], dtype=torch.int32, device="cuda")
    dense = torch.ones(2, 8, device="cuda")
    out = rtf.embed_dot(ts, ids, dense)
    w0, w1 = ts.weights[0].detach(), ts.weights[1].detach()
    x = torch.stack([dense, torch.stack([w0[1], torch.zeros(8, device="cuda")]),
                     torch.stack([w1[2], w1[3]])], 1)
    torch.testing.assert_close(out, rtf.dot_interact(x))
    with pytest.raises(IndexError):
        ts.check_ids()


def test_dlrm_model_matches_torch_reference(rtf):
    """Whole DLRM forward/backward through the drop-in classes == the same model written with
    plain torch ops on the same weights."""
    torch.manual_seed(0)
    F, D, B = 26, 16, 256
    rows = [50 + 7 * i for i in range(F)]
    fc = [[{"feat": f"I{i}"} for i in range(13)],
          [{"feat": f"C{i}", "feat_num": rows[i], "embed_dim": D} for i in range(F)]]
    m = rtf.DLRM(fc, bot_dnn_hidden_units=(64, 32, D), top_dnn_hidden_units=(128, 64), seed=1).cuda()
    dense = torch.rand(B, 13, device="cuda")
    sparse = torch.stack([torch.randint(0, r, (B,), device="cuda") for r in rows], 1).to(torch.int32)
    y = (torch.rand(B, 1, device="cuda") < 0.25).float()
    pred = m([dense, sparse])
    loss = rtf.layers.binary_crossentropy(y, pred)
    loss.backward()

    # reference: same parameters, torch indexing + bmm
    W = [w.detach().clone().requires_grad_(True) for w in m.embed_layers.weights]
    dfea = m.bot_dnn(dense)
    emb = torch.stack([W[i][sparse[:, i].long()] for i in range(F)], 1)
    x = torch.cat([dfea.unsqueeze(1), emb], 1)
    z = torch.bmm(x, x.transpose(1, 2))
    ii, jj = torch.tril_indices(F + 1, F + 1, -1)
    inter = torch.cat([dfea, z[:, ii, jj]], 1)
    pred2 = torch.sigmoid(m.final_dense(m.top_dnn(inter)))
    torch.testing.assert_close(pred, pred2, rtol=1e-5, atol=1e-6)
    loss2 = rtf.layers.binary_crossentropy(y, pred2)
    g_ref = torch.autograd.grad(loss2, W)
    for w, g in zip(m.embed_layers.weights, g_ref):
        torch.testing.assert_close(w.grad.to_dense(), g, rtol=1e-4, atol=1e-7)


def test_dlrm_trainer_loss_decreases(rtf):
    torch.manual_seed(0)
    F, D, B = 8, 16, 512
    rows = [100] * F
    fc = [[{"feat": f"I{i}"} for i in range(13)],
          [{"feat": f"C{i}", "feat_num": rows[i], "embed_dim": D} for i in range(F)]]
    m = rtf.DLRM(fc, bot_dnn_hidden_units=(32, D), top_dnn_hidden_units=(64, 32), seed=1).cuda()
    tr = rtf.DLRMTrainer(m, lr=1e-2)
    dense = torch.rand(B, 13, device="cuda")
    sparse = torch.stack([torch.randint(0, r, (B,), device="cuda") for r in rows], 1).to(torch.int32)
    y = ((sparse[:, 0] % 2) == 0).float().unsqueeze(1)      # learnable from one sparse field
    losses = [float(tr.step(dense, sparse, y)) for _ in range(60)]
    assert losses[-1] < 0.5 * losses[0], losses[::10]
    m.embed_layers.check_ids()

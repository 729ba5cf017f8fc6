"""core.StepGraph: a training step replayed from a CUDA graph must leave the model in exactly the
state the eager step leaves it in (same kernels, same order, same Adam step sizes — the device
scalar rtf_opt.lr_dev holds the fp32 value the eager path passes by value)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _fm_pair(rtf, graph):
    torch.manual_seed(11)       # dense layers draw from torch's global generator when they build
    rows = [50, 7, 1000, 33, 5000, 12]
    fc = [[{"feat": f"I{i}"} for i in range(4)],
          [{"feat": f"C{i}", "feat_num": r, "embed_dim": 8} for i, r in enumerate(rows)]]
    m = rtf.FMModel(fc, k=8, seed=3)
    tr = rtf.models.Trainer(m, lambda out, y: rtf.layers.binary_crossentropy(y, out), lr=1e-2,
                            embed_l2=1e-4, cuda_graph=graph)
    return m, tr, rows


def _batches(rows, n, B, n_dense, seed=0):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n):
        dense = torch.rand(B, n_dense, generator=g)
        sparse = torch.stack([torch.randint(0, r, (B,), generator=g) for r in rows], 1).to(torch.int32)
        y = torch.randint(0, 2, (B, 1), generator=g).float()
        out.append((dense.cuda(), sparse.cuda(), y.cuda()))
    return out


def test_fm_trainer_graph_matches_eager_bitwise(rtf):
    bs = _batches([50, 7, 1000, 33, 5000, 12], 9, 256, 4)
    m0, t0, rows = _fm_pair(rtf, False)
    l0 = [float(t0.step([d, s], y)) for d, s, y in bs]
    m1, t1, _ = _fm_pair(rtf, True)
    l1 = [float(t1.step([d, s], y)) for d, s, y in bs]
    assert t1.graph.graph is not None and t1.graph.n_replays == 9 - t1.graph.warmup
    assert l0 == l1
    w0, w1 = rtf.checkpoint.weights_dict(m0), rtf.checkpoint.weights_dict(m1)
    assert w0.keys() == w1.keys()
    for k in w0:
        np.testing.assert_array_equal(w0[k], w1[k], err_msg=k)
    for ts0, ts1 in zip(t0.tables, t1.tables):
        for a, b in zip(ts0.state1 + ts0.state2, ts1.state1 + ts1.state2):
            assert torch.equal(a, b)
        ts1.check_ids()


def test_graph_falls_back_to_eager_on_a_ragged_batch(rtf):
    rows = [50, 7, 1000, 33, 5000, 12]
    bs = _batches(rows, 6, 128, 4) + _batches(rows, 1, 40, 4, seed=5) + _batches(rows, 2, 128, 4, seed=6)
    m0, t0, rows = _fm_pair(rtf, False)
    l0 = [float(t0.step([d, s], y)) for d, s, y in bs]
    m1, t1, _ = _fm_pair(rtf, True)
    l1 = [float(t1.step([d, s], y)) for d, s, y in bs]
    assert l0 == l1
    assert t1.graph.n_replays == 6 - 3 + 2
    w0, w1 = rtf.checkpoint.weights_dict(m0), rtf.checkpoint.weights_dict(m1)
    for k in w0:
        np.testing.assert_array_equal(w0[k], w1[k], err_msg=k)


def test_dlrm_trainer_graph_matches_eager_bitwise(rtf):
    F, D, B = 8, 16, 512
    rows = [100, 5000, 17, 900, 100000, 3, 64, 2048]
    fc = [[{"feat": f"I{i}"} for i in range(13)],
          [{"feat": f"C{i}", "feat_num": rows[i], "embed_dim": D} for i in range(F)]]
    ms, trs, losses = [], [], []
    bs = _batches(rows, 8, B, 13)
    for graph in (False, True):
        torch.manual_seed(5)
        m = rtf.DLRM(fc, bot_dnn_hidden_units=(32, D), top_dnn_hidden_units=(64, 32), seed=1).cuda()
        tr = rtf.DLRMTrainer(m, lr=1e-2, cuda_graph=graph)
        losses.append([float(tr.step(d, s, y)) for d, s, y in bs])
        ms.append(m)
        trs.append(tr)
    assert losses[0] == losses[1]
    assert trs[1].graph.n_replays == 5
    for w0, w1 in zip(ms[0].embed_layers.weights, ms[1].embed_layers.weights):
        assert torch.equal(w0, w1)
    assert torch.equal(trs[0].dense_opt.flat, trs[1].dense_opt.flat)
    ms[1].embed_layers.check_ids()


def test_youtubednn_graph_draws_fresh_candidates_and_matches_eager(rtf):
    """The sampler's per-step seed lives in device memory under the graph: every replay draws the
    candidates the eager step with the same counter draws, so the two runs agree bit for bit."""
    g = torch.Generator().manual_seed(2)
    bs = [(torch.randint(0, 50, (128, 2), generator=g).to(torch.int32).cuda(),
           torch.randint(0, 5000, (128,), generator=g).cuda()) for _ in range(8)]
    runs = []
    for graph in (False, True):
        torch.manual_seed(9)
        m = rtf.models.YoutubeDNN([100, 50], item_num=5000, embed_dim=8, user_dnn_hidden_units=(64, 32),
                                  num_sampled=64, seed=4)
        tr = rtf.models.Trainer(m, lambda out, y: out.mean(), lr=1e-2, cuda_graph=graph)
        losses = [float(tr.step([u, i])) for u, i in bs]
        runs.append((m, tr, losses))
    assert runs[1][1].graph.n_replays == 5
    assert runs[0][2] == runs[1][2]
    assert len(set(runs[1][2][3:])) > 1
    assert torch.equal(runs[0][0].item_table.weights[0], runs[1][0].item_table.weights[0])
    assert torch.equal(runs[0][1].dense_opt.flat, runs[1][1].dense_opt.flat)


def test_log_uniform_sampler_device_seed_equals_host_seed(rtf):
    from recommend_tf2_b200.layers.match import log_uniform_candidate_sampler
    for seed in (1, 77, 2**40 + 3):
        a, ta = log_uniform_candidate_sampler(256, 100000, seed)
        b, tb = log_uniform_candidate_sampler(256, 100000, torch.tensor([seed], dtype=torch.int64, device="cuda"))
        assert torch.equal(a, b) and int(ta) == int(tb)

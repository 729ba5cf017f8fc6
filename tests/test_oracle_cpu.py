"""CPU tests of the oracle itself (known-answer identities) — no GPU."""
import numpy as np

from oracle import embedding as O


def test_gather_concat_matches_plain_indexing():
    rng = np.random.default_rng(0)
    tabs = [rng.normal(size=(n, d)).astype(np.float32) for n, d in [(10, 4), (7, 8)]]
    ids = np.stack([rng.integers(0, 10, (5, 3)), rng.integers(0, 7, (5, 3))], 1)
    out = O.embed_lookup_concat(tabs, ids)
    assert out.shape == (5, 3, 12)
    for b in range(5):
        for l in range(3):
            assert np.array_equal(out[b, l, :4], tabs[0][ids[b, 0, l]])
            assert np.array_equal(out[b, l, 4:], tabs[1][ids[b, 1, l]])
    s = O.embed_lookup_concat(tabs, ids, "sum")
    np.testing.assert_allclose(s, out.sum(1), rtol=1e-6)
    m = O.embed_lookup_concat(tabs, ids, "mean")
    np.testing.assert_allclose(m, out.mean(1), rtol=1e-6)


def test_grad_unique_matches_dense_scatter_add():
    rng = np.random.default_rng(1)
    rows, dims, ft = [5, 40], [8, 4], [0, 1, 0]
    ids = np.stack([rng.integers(0, 5, (300, 2)), rng.integers(0, 40, (300, 2)),
                    rng.integers(0, 5, (300, 2))], 1)
    grad = rng.normal(size=(300, 2, 20)).astype(np.float32)
    keys, tot, rb = O.embed_grad_unique(ids, ft, rows, dims, grad)
    assert np.all(np.diff(keys) > 0)
    dense = O.dense_reference_grad([(5, 8), (40, 4)], ids, ft, grad)
    tab, row = keys >> rb, keys & ((1 << rb) - 1)
    for t in range(2):
        np.testing.assert_allclose(tot[tab == t, : dims[t]], dense[t][row[tab == t]], rtol=1e-4, atol=1e-4)
    # table 0 has 5 rows and 1200 lookups -> segments longer than SEG_CHUNK exercise the chunk path
    assert (300 * 2 * 2) // 5 > O.SEG_CHUNK


def test_adam_matches_closed_form_first_step():
    W = [np.ones((4, 4), np.float32)]
    m = [np.zeros((4, 4), np.float32)]
    v = [np.zeros((4, 4), np.float32)]
    ids = np.array([[[2]]])
    grad = np.full((1, 1, 4), 0.5, np.float32)
    lr_t = O.adam_lr_t(1e-3, 0.9, 0.999, 1)
    O.embed_bwd_apply("adam", W, m, v, ids, [0], grad, lr=lr_t)
    # first Adam step moves by ~lr in the direction of the gradient sign
    np.testing.assert_allclose(W[0][2], 1 - 1e-3, rtol=1e-5)
    assert np.array_equal(W[0][[0, 1, 3]], np.ones((3, 4), np.float32))


def test_out_of_range_ids_get_sentinel_key():
    keys, rb = O.make_keys(np.array([[[1], [9]]]), [0, 1], [4, 4])
    assert keys[0] == 1 and keys[1] == (2 << rb)


def test_binary_crossentropy_host_path_matches_oracle_and_loop_restatement():
    """Keras binary_crossentropy (App. A11, src/ctr/fm/train.py:49): the package's CPU-tensor path
    (framework ops; CUDA tensors take rtf_bce_fwd, tests/test_loss_gpu.py), the oracle's
    restatement and a plain-Python loop over the formula agree."""
    import math
    import torch
    import recommend_tf2_b200 as pkg
    from oracle.dlrm_ref import bce
    g = torch.Generator().manual_seed(11)
    p = torch.rand(257, 1, generator=g, dtype=torch.float64)
    p[0], p[1], p[2] = 0.0, 1.0, 1e-9
    y = (torch.rand(257, generator=g) < 0.4).double()
    got = pkg.layers.binary_crossentropy(y, p)
    want = bce(y.reshape(p.shape), p)
    eps = 1e-7
    acc = 0.0
    for yi, pi in zip(y.tolist(), p.reshape(-1).tolist()):
        pc = min(max(pi, eps), 1 - eps)
        acc += yi * math.log(pc + eps) + (1 - yi) * math.log(1 - pc + eps)
    assert abs(float(got) - float(want)) < 1e-12
    assert abs(float(got) + acc / 257) < 1e-12

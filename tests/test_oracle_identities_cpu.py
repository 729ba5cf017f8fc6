"""Known-answer identities that pin the oracle independently of any kernel:
algebraic facts the reference's formulas must satisfy, checked with hypothesis over shapes, id
patterns and masks.  CPU only."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import attention as OA
from oracle import embedding as OE
from oracle import interaction as OI

SET = settings(max_examples=40, deadline=None)


@SET
@given(st.integers(1, 6), st.integers(2, 9), st.integers(1, 5), st.integers(0, 2 ** 31 - 1))
def test_fm_second_order_is_the_pairwise_sum(B, F, D, seed):
    """0.5((sum x)^2 - sum x^2) = sum_{i<j} x_i x_j — src/ctr/layers/modules.py:67-69."""
    rng = np.random.default_rng(seed)
    x = rng.normal(size=(B, F, D))
    w = rng.normal(size=(3, 1))
    first = rng.normal(size=(B, 3))
    got = OI.fm_layer_paper(first, x, w)
    pair = np.zeros((B, 1))
    for i in range(F):
        for j in range(i + 1, F):
            pair[:, 0] += (x[:, i] * x[:, j]).sum(-1)
    np.testing.assert_allclose(got, first @ w + pair, rtol=1e-10, atol=1e-10)
    # the layer as written (2-D second input): scalar first order over the WHOLE batch (:65)
    flat = x.reshape(B, F * D)
    lit = OI.fm_layer(first, flat, w)
    s = flat.sum(1)
    np.testing.assert_allclose(lit[:, 0], (first @ w).sum() + 0.5 * (s * s - (flat * flat).sum(1)),
                               rtol=1e-10, atol=1e-10)


@SET
@given(st.integers(1, 5), st.integers(1, 4), st.integers(1, 6), st.integers(0, 2 ** 31 - 1))
def test_onehot_fm_equals_gather_form(B, n_sparse, k, seed):
    """src/ctr/fm/model.py:37-51 on a one-hot matrix == gathering the rows 13+off_i+id_i."""
    rng = np.random.default_rng(seed)
    nd = 3
    feat_nums = [int(n) for n in rng.integers(2, 7, n_sparse)]
    M = nd + sum(feat_nums)
    dense = rng.random((B, nd))
    sparse = np.stack([rng.integers(0, n, B) for n in feat_nums], 1)
    w0, w, V = rng.normal(size=1), rng.normal(size=(M, 1)), rng.normal(size=(k, M))
    got = OI.fm_model_onehot(dense, sparse, feat_nums, w0, w, V)
    off = np.concatenate([[0], np.cumsum(feat_nums)[:-1]]) + nd
    out = np.zeros((B, 1))
    for b in range(B):
        idx = list(range(nd)) + [int(off[i] + sparse[b, i]) for i in range(n_sparse)]
        val = np.concatenate([dense[b], np.ones(n_sparse)])
        first = w0[0] + (val * w[idx, 0]).sum()
        vx = V[:, idx] * val                                        # (k, n)
        second = 0.5 * ((vx.sum(1)) ** 2 - (vx ** 2).sum(1)).sum()
        out[b, 0] = 1.0 / (1.0 + np.exp(-(first + second)))
    np.testing.assert_allclose(got, out, rtol=1e-10, atol=1e-12)


@SET
@given(st.integers(1, 4), st.integers(1, 12), st.integers(1, 6), st.integers(0, 2 ** 31 - 1))
def test_din_info_rearrangement_and_masking(B, L, d, seed):
    """w·[q, k, q-k, q∘k] = (w1+w3)·q + (w2-w3+w4∘q)·k (what K5 computes), masked positions get
    zero weight, an all-masked row is exactly uniform, mask=None pads everything (:161-165)."""
    rng = np.random.default_rng(seed)
    q, k, v = rng.normal(size=(B, d)), rng.normal(size=(B, L, d)), rng.normal(size=(B, L, d))
    W, bias = rng.normal(size=(4 * d, 1)), rng.normal(size=1)
    mask = (rng.random((B, L)) < 0.6).astype(np.float64)
    mask[0] = 0.0
    got = OA.din_attention_layer(q, k, v, mask, W, bias, "sigmoid")
    w1, w2, w3, w4 = (W[i * d:(i + 1) * d, 0] for i in range(4))
    c = q @ (w1 + w3) + bias[0]
    u = (w2 - w3)[None] + w4[None] * q
    s = 1.0 / (1.0 + np.exp(-(c[:, None] + np.einsum("bd,bld->bl", u, k))))
    s = np.where(mask == 0, np.float64(np.float32(OA.PAD)), s)
    a = OA.softmax(s)
    np.testing.assert_allclose(got, np.einsum("bl,bld->bd", a, v), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(a.sum(-1), 1.0, rtol=1e-12)
    some = mask.sum(1) > 0                                  # rows with at least one real position:
    assert np.all(a[some][mask[some] == 0] == 0.0)          # their masked positions weigh exactly 0
    np.testing.assert_allclose(got[0], v[0].mean(0), rtol=1e-12)             # all masked -> uniform
    none = OA.din_attention_layer(q, k, v, None, W, bias, "sigmoid")
    np.testing.assert_allclose(none, v.mean(1), rtol=1e-12)


@SET
@given(st.integers(1, 4), st.integers(2, 9), st.integers(1, 6), st.integers(0, 2 ** 31 - 1))
def test_dot_interaction_is_the_strict_lower_triangle_of_the_gram(B, F1, D, seed):
    rng = np.random.default_rng(seed)
    x = rng.normal(size=(B, F1, D))
    out = OI.dot_interact(x)
    assert out.shape == (B, D + F1 * (F1 - 1) // 2)
    np.testing.assert_array_equal(out[:, :D], x[:, 0])
    p = D
    for i in range(F1):
        for j in range(i):
            np.testing.assert_allclose(out[:, p], (x[:, i] * x[:, j]).sum(-1), rtol=1e-12, atol=1e-12)
            p += 1
    # the backward is the adjoint of the forward's Jacobian: <J dx, g> = <dx, J^T g>
    dx, g = rng.normal(size=x.shape), rng.normal(size=out.shape)
    eps = 1e-6
    jdx = (OI.dot_interact(x + eps * dx) - OI.dot_interact(x - eps * dx)) / (2 * eps)
    np.testing.assert_allclose((jdx * g).sum(), (dx * OI.dot_interact_bwd(x, g)).sum(), rtol=1e-5, atol=1e-6)


@SET
@given(st.integers(1, 40), st.integers(1, 3), st.integers(1, 6), st.integers(0, 2 ** 31 - 1))
def test_segment_sums_equal_scatter_add_for_any_duplicate_pattern(B, L, n_rows, seed):
    """K2's contract: sorted unique (table,row) keys + per-row sums == np.add.at on dense tables,
    whatever the duplicate structure (few rows -> long segments, incl. the chunked path)."""
    rng = np.random.default_rng(seed)
    rows, dims, ft = [n_rows, 3 * n_rows], [4, 8], [0, 1, 1]
    ids = np.stack([rng.integers(0, rows[t], (B, L)) for t in ft], 1)          # (B, F, L)
    grad = rng.normal(size=(B, L, 4 + 8 + 8)).astype(np.float32)
    keys, tot, rb = OE.embed_grad_unique(ids, ft, rows, dims, grad)
    assert np.all(np.diff(keys.astype(np.int64)) > 0)                           # sorted, unique
    dense = [np.zeros((rows[t], dims[t])) for t in range(2)]
    off = [0, 4, 12]
    for f, t in enumerate(ft):
        for l in range(L):
            np.add.at(dense[t], ids[:, f, l], grad[:, l, off[f]:off[f] + dims[t]].astype(np.float64))
    tab, row = keys >> rb, keys & ((1 << rb) - 1)
    touched = {(int(t), int(r)) for t, r in zip(tab, row)}
    want = {(t, int(r)) for f, t in enumerate(ft) for r in np.unique(ids[:, f])}
    assert touched == want
    for kk, (t, r) in enumerate(zip(tab, row)):
        np.testing.assert_allclose(tot[kk, :dims[t]], dense[t][r], rtol=2e-5, atol=2e-5)


@SET
@given(st.integers(1, 3), st.integers(1, 7), st.integers(1, 3), st.integers(0, 2 ** 31 - 1))
def test_match_attention_masks_query_rows_not_keys(B, L, H, seed):
    """src/match/layers/modules.py:90-91: the (B,L,1) mask blanks whole QUERY rows (uniform
    attention over all keys, padded ones included); unmasked rows are plain softmax(QK^T/sqrt(dk))V."""
    rng = np.random.default_rng(seed)
    dk = 4
    q, k, v = (rng.normal(size=(B, H, L, dk)) for _ in range(3))
    mask = (rng.random((B, L, 1)) < 0.6).astype(np.float64)
    out = OA.match_sdpa(q, k, v, mask[:, None])
    logits = np.einsum("bhld,bhmd->bhlm", q, k) / np.sqrt(dk)
    plain = np.einsum("bhlm,bhmd->bhld", OA.softmax(logits), v)
    for b in range(B):
        for l in range(L):
            if mask[b, l, 0] == 0:
                np.testing.assert_allclose(out[b, :, l], v[b].mean(1), rtol=1e-9, atol=1e-12)
            else:
                np.testing.assert_allclose(out[b, :, l], plain[b, :, l], rtol=1e-9, atol=1e-12)


def test_log_uniform_sampler_probabilities_sum_to_one_and_expected_counts():
    N = 1000
    p = OA.log_uniform_prob(np.arange(N), N)
    np.testing.assert_allclose(p.sum(), 1.0, rtol=1e-12)
    assert np.all(np.diff(p) < 0)                                   # Zipfian: decreasing in the id
    e = OA.log_uniform_expected(np.arange(N), N, num_tries=50)
    np.testing.assert_allclose(e, -np.expm1(50 * np.log1p(-p)), rtol=1e-12)
    assert np.all((e > 0) & (e < 1))

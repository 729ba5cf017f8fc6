"""K1/K2 parity (GPU, through the C-ABI) against the numpy oracle.

Bar (north star): ids / gathered rows / index work bit-exact; fp32 sums bit-exact because the
summation order is fixed on both sides; optimizer updates bit-exact (separately rounded ops).
"""
import numpy as np
import pytest
import torch

from oracle import embedding as O

pytestmark = pytest.mark.gpu


def _zipf_ids(rng, B, F, L, rows, a=1.05):
    ids = np.empty((B, F, L), np.int64)
    for f in range(F):
        ids[:, f, :] = (rng.zipf(a, size=(B, L)) - 1) % rows[f]
    return ids


def _tables(rng, rows, dims):
    return [rng.uniform(-0.05, 0.05, size=(n, d)).astype(np.float32) for n, d in zip(rows, dims)]


CASES = [
    # name, B, rows, dims, L
    ("dlrm_like", 257, [1460, 583, 100003, 3, 24, 12517], [128] * 6, 1),
    ("fm_d8", 1024, [1000] * 26, [8] * 26, 1),
    ("autoint_d16", 130, [50, 7, 999], [16] * 3, 1),
    ("mixed_dims", 65, [10, 20, 30, 40], [4, 8, 64, 256], 1),
    ("d512", 33, [100, 50], [512, 384], 1),
    ("odd_dims_scalar", 37, [11, 13, 17], [3, 5, 10], 1),
    ("one_sample", 1, [5], [128], 1),
]


@pytest.mark.parametrize("name,B,rows,dims,L", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("idt", [torch.int32, torch.int64])
def test_gather_concat_bit_exact(rtf, name, B, rows, dims, L, idt):
    rng = np.random.default_rng(0)
    tabs = _tables(rng, rows, dims)
    ids = _zipf_ids(rng, B, len(rows), L, rows)
    want = O.embed_lookup_concat(tabs, ids)[:, 0, :]
    dtabs = [torch.from_numpy(t).cuda() for t in tabs]
    dids = torch.from_numpy(ids[:, :, 0]).to(idt).cuda()
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    got = rtf.embed_fwd(dtabs, dids, "BF", None, err=err)
    assert got.shape == want.shape
    assert np.array_equal(got.cpu().numpy().view(np.uint32), want.view(np.uint32))
    assert int(err.item()) == 0


@pytest.mark.parametrize("pool", [None, "sum", "mean"])
@pytest.mark.parametrize("layout", ["BFL", "BLF"])
@pytest.mark.parametrize("dims", [[8, 8, 8], [64, 64], [128], [6, 10]])
def test_sequence_lookup_and_pooling(rtf, pool, layout, dims):
    rng = np.random.default_rng(1)
    F, B, L = len(dims), 41, 13
    rows = [37 + 11 * f for f in range(F)]
    tabs = _tables(rng, rows, dims)
    ids = _zipf_ids(rng, B, F, L, rows)
    want = O.embed_lookup_concat(tabs, ids, pool)
    dtabs = [torch.from_numpy(t).cuda() for t in tabs]
    t_ids = torch.from_numpy(ids).to(torch.int32).cuda()
    if layout == "BLF":
        t_ids = t_ids.permute(0, 2, 1).contiguous()
    got = rtf.embed_fwd(dtabs, t_ids, layout, pool)
    assert got.shape == want.shape
    assert np.array_equal(got.cpu().numpy().view(np.uint32), want.view(np.uint32))


def test_shared_table_and_strided_ids(rtf):
    """DIN shares the item tables between the target item and the behaviour sequence
    (src/ctr/din/model.py:71-72) and indexes columns of one wide id matrix."""
    rng = np.random.default_rng(2)
    item = rng.normal(0, 0.05, (1000, 8)).astype(np.float32)
    cate = rng.normal(0, 0.05, (50, 8)).astype(np.float32)
    B, L = 64, 10
    wide = np.stack([rng.integers(0, 1000, (B, L)), rng.integers(0, 50, (B, L))], -1)  # (B,L,2)
    want = O.embed_lookup_concat([item, cate], np.transpose(wide, (0, 2, 1)))
    flat = torch.from_numpy(wide.reshape(B, 2 * L)).to(torch.int32).cuda()   # (B, 2*maxlen)
    view = flat.view(B, L, 2)
    got = rtf.embed_fwd([torch.from_numpy(item).cuda(), torch.from_numpy(cate).cuda()], view, "BLF")
    assert np.array_equal(got.cpu().numpy(), want)


def test_out_of_range_id_sets_flag_and_reads_zero(rtf):
    tab = torch.arange(40, dtype=torch.float32, device="cuda").view(10, 4) + 1
    ids = torch.tensor([[3], [10], [-1], [9]], dtype=torch.int32, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    got = rtf.embed_fwd([tab], ids, "BF", None, err=err).cpu().numpy()
    assert int(err.item()) == 1
    assert np.array_equal(got[0], tab[3].cpu().numpy()) and np.array_equal(got[3], tab[9].cpu().numpy())
    assert not got[1].any() and not got[2].any()
    with pytest.raises(IndexError):
        O.embed_lookup_concat([tab.cpu().numpy()], ids.cpu().numpy()[:, :, None])
    ts = rtf.EmbeddingTables([10], [4])
    ts.lookup(ids)
    with pytest.raises(IndexError):
        ts.check_ids()


def test_skip_invalid_leaves_rows_untouched_and_unflagged(rtf):
    """RTF_POOL_SKIP_INVALID (owner-gather exchange): a lookup with id -1 / >= rows is skipped —
    its output row keeps what was there, no error flag — the others are gathered bit-exactly."""
    g = torch.Generator(device="cuda").manual_seed(4)
    rows, D, B = [50, 7, 1000], 128, 777
    tabs = [torch.randn(n, D, device="cuda", generator=g) for n in rows]
    ids = torch.stack([torch.randint(0, n, (B,), device="cuda", generator=g) for n in rows], 1).to(torch.int32)
    bad = torch.rand(B, 3, device="cuda", generator=g) < 0.4
    ids_bad = torch.where(bad, torch.full_like(ids, -1), ids)
    ids_bad[5, 2] = 1000                       # too large counts as invalid as well
    bad[5, 2] = True
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    out = torch.full((B, 3 * D), 7.25, device="cuda")
    rtf.embed_fwd(tabs, ids_bad, "BF", None, err=err, out=out, skip_invalid=True)
    want = torch.cat([t[ids[:, f].long().clamp(0, rows[f] - 1)] for f, t in enumerate(tabs)], 1)
    keep = bad.repeat_interleave(D, dim=1)
    assert torch.equal(out[~keep], want[~keep])
    assert bool((out[keep] == 7.25).all())
    assert int(err.item()) == 0


@pytest.mark.parametrize("dims,frac_bad", [((128, 64, 256, 8), 0.875), ((512, 128), 0.5), ((16,), 0.99)])
def test_skip_invalid_mixed_dims_mostly_foreign(rtf, dims, frac_bad):
    """The holder-side gather of a row-wise sharded table: 7 of 8 lookups are foreign (-1)."""
    g = torch.Generator(device="cuda").manual_seed(9)
    B = 4099
    rows = [1000 + 7 * i for i in range(len(dims))]
    tabs = [torch.randn(n, d, device="cuda", generator=g) for n, d in zip(rows, dims)]
    ids = torch.stack([torch.randint(0, n, (B,), device="cuda", generator=g) for n in rows], 1)
    bad = torch.rand(B, len(dims), device="cuda", generator=g) < frac_bad
    ids_bad = torch.where(bad, torch.full_like(ids, -1), ids).to(torch.int32)
    out = torch.full((B, sum(dims)), -3.5, device="cuda")
    rtf.embed_fwd(tabs, ids_bad, "BF", None, err=None, out=out, skip_invalid=True)
    off = 0
    for f, (t, d) in enumerate(zip(tabs, dims)):
        got = out[:, off:off + d]
        ok = ~bad[:, f]
        assert torch.equal(got[ok], t[ids[ok, f]])
        assert bool((got[~ok] == -3.5).all())
        off += d


def test_empty_batch(rtf):
    tab = torch.zeros(10, 8, device="cuda")
    ids = torch.zeros((0, 1), dtype=torch.int32, device="cuda")
    assert rtf.embed_fwd([tab], ids).shape == (0, 8)


def test_cpu_tensor_raises(rtf):
    with pytest.raises(rtf.RtfError):
        rtf.embed_fwd([torch.zeros(4, 4)], torch.zeros((2, 1), dtype=torch.int32))


# ------------------------------------------------------------------ backward
BWD_CASES = [
    # name, B, rows, dims, field_table, L, pool
    ("dlrm_like", 300, [1460, 3, 100003, 24], [128] * 4, [0, 1, 2, 3], 1, None),
    ("hot_rows_long_segments", 1500, [3, 2, 50], [128, 128, 128], [0, 1, 2], 1, None),
    ("fm_d8", 1024, [1000] * 26, [8] * 26, list(range(26)), 1, None),
    ("shared_table_seq", 50, [200, 20], [16, 16], [0, 1, 0, 1], 7, None),
    ("pool_sum", 40, [30, 300], [64, 64], [0, 1], 9, "sum"),
    ("pool_mean", 40, [30, 300], [32, 32], [0, 1], 9, "mean"),
    ("mixed_dims", 70, [10, 20, 30], [4, 64, 256], [0, 1, 2], 1, None),
    ("odd_dims_scalar", 90, [11, 13], [3, 10], [0, 1], 2, None),
    # > RTF_SEG_CHUNK * RTF_SEG_GROUP = 4096 lookups of one row (a padding id in a behaviour
    # sequence): chunk partials are combined per group of 64 chunks, then the groups
    ("giant_segments_d8_seq", 600, [2, 40], [8, 8], [0, 1], 20, None),
    ("giant_segments_d128", 9000, [1, 3], [128, 128], [0, 1], 1, None),
]


def _bwd_inputs(case, seed=3):
    name, B, rows, dims, ft, L, pool = case
    rng = np.random.default_rng(seed)
    F = len(ft)
    frows = [rows[t] for t in ft]
    ids = _zipf_ids(rng, B, F, L, frows)
    sumD = sum(dims[t] for t in ft)
    gshape = (B, L, sumD) if pool is None else (B, sumD)
    grad = rng.normal(0, 1, gshape).astype(np.float32)
    return rng, ids, grad


@pytest.mark.parametrize("case", BWD_CASES, ids=[c[0] for c in BWD_CASES])
def test_grad_segments_bit_exact(rtf, case):
    name, B, rows, dims, ft, L, pool = case
    rng, ids, grad = _bwd_inputs(case)
    keys_w, tot_w, rb_w = O.embed_grad_unique(ids, ft, rows, dims, grad, pool)
    weights = [torch.zeros(n, d, device="cuda") for n, d in zip(rows, dims)]
    dids = torch.from_numpy(ids).to(torch.int32).cuda()
    keys, tot, rb = rtf.embed_bwd(weights, ft, dids, torch.from_numpy(grad).cuda(), "BFL", pool,
                                  want_unique=True)
    assert rb == rb_w
    assert np.array_equal(keys.cpu().numpy(), keys_w)                      # sorted unique keys
    assert np.array_equal(tot.cpu().numpy().view(np.uint32), tot_w.view(np.uint32))
    # and against an order-free fp64 dense scatter-add
    dense = O.dense_reference_grad([(n, d) for n, d in zip(rows, dims)], ids, ft, grad, pool)
    tab, row = keys_w >> rb_w, keys_w & ((1 << rb_w) - 1)
    for t in range(len(rows)):
        sel = tab == t
        np.testing.assert_allclose(tot_w[sel, : dims[t]], dense[t][row[sel]], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("kind", ["sgd", "adagrad", "adam"])
@pytest.mark.parametrize("case", [BWD_CASES[0], BWD_CASES[1], BWD_CASES[3], BWD_CASES[5], BWD_CASES[7],
                                  BWD_CASES[8]], ids=lambda c: c[0])
def test_sparse_optimizer_in_place_bit_exact(rtf, kind, case):
    name, B, rows, dims, ft, L, pool = case
    rng, ids, grad = _bwd_inputs(case, seed=4)
    W = _tables(rng, rows, dims)
    S1 = [np.abs(rng.normal(0, 0.01, w.shape)).astype(np.float32) for w in W]
    S2 = [np.abs(rng.normal(0, 0.01, w.shape)).astype(np.float32) for w in W]
    dW = [torch.from_numpy(w.copy()).cuda() for w in W]
    dS1 = [torch.from_numpy(w.copy()).cuda() for w in S1]
    dS2 = [torch.from_numpy(w.copy()).cuda() for w in S2]
    opt = rtf.SparseOptimizer(kind, lr=0.01, l2=1e-4)
    st = opt.struct_for_step(3)
    O.embed_bwd_apply(kind, W, S1, S2, ids, ft, grad, pool, lr=np.float32(st.lr), l2=1e-4)
    rtf.embed_bwd(dW, ft, torch.from_numpy(ids).to(torch.int64).cuda(), torch.from_numpy(grad).cuda(),
                  "BFL", pool, opt=st, state1=dS1, state2=dS2)
    for t in range(len(W)):
        assert np.array_equal(dW[t].cpu().numpy().view(np.uint32), W[t].view(np.uint32)), f"W[{t}]"
        if kind in ("adagrad", "adam"):
            assert np.array_equal(dS1[t].cpu().numpy().view(np.uint32), S1[t].view(np.uint32))
        if kind == "adam":
            assert np.array_equal(dS2[t].cpu().numpy().view(np.uint32), S2[t].view(np.uint32))


def test_backward_is_run_to_run_deterministic(rtf):
    case = ("det", 4096, [5, 1000, 100000], [128] * 3, [0, 1, 2], 1, None)
    rng, ids, grad = _bwd_inputs(case)
    dids = torch.from_numpy(ids).to(torch.int32).cuda()
    g = torch.from_numpy(grad).cuda()
    outs = []
    for _ in range(3):
        W = [torch.zeros(n, 128, device="cuda") for n in case[2]]
        k, t, _ = rtf.embed_bwd(W, case[4], dids, g, "BFL", None, want_unique=True)
        outs.append((k.clone(), t.clone()))
    for k, t in outs[1:]:
        assert torch.equal(k, outs[0][0]) and torch.equal(t, outs[0][1])


def test_large_sort_matches_torch_sort(rtf):
    """Full-size index work: 2^20+ lookups, sorted unique keys must equal torch.unique."""
    g = torch.Generator(device="cpu").manual_seed(5)
    B, F = 70001, 16
    rows = [10_000_000] * 4 + [100_000] * 4 + [1000] * 4 + [7] * 4
    ids = torch.stack([torch.randint(0, r, (B,), generator=g) for r in rows], 1).to(torch.int32).cuda()
    weights = [torch.empty((r, 4), device="cuda") for r in rows]   # 4 floats/row: 640 MB total max
    grad = torch.ones(B, F * 4, device="cuda")
    keys, tot, rb = rtf.embed_bwd(weights, list(range(F)), ids, grad, "BF", None, want_unique=True)
    want = torch.unique((torch.arange(F, device="cuda").view(1, F).long() << rb) | ids.long())
    assert torch.equal(keys, want)
    # counts: every summed gradient equals the multiplicity of the key (grad rows are all ones)
    _, counts = torch.unique((torch.arange(F, device="cuda").view(1, F).long() << rb) | ids.long(),
                             return_counts=True)
    assert torch.equal(tot[:, 0], counts.float())


def test_autograd_lookup_fused_and_sparse(rtf):
    rng = np.random.default_rng(6)
    rows, dims = [50, 60], [16, 16]
    ids = torch.from_numpy(_zipf_ids(rng, 32, 2, 1, rows)[:, :, 0]).to(torch.int32).cuda()
    ts = rtf.EmbeddingTables(rows, dims, seed=0)
    out = ts.lookup(ids)
    (out * out).sum().backward()
    ref = [w.detach().clone().requires_grad_(True) for w in ts.weights]
    o2 = torch.cat([ref[0][ids[:, 0].long()], ref[1][ids[:, 1].long()]], 1)
    (o2 * o2).sum().backward()
    for w, r in zip(ts.weights, ref):
        torch.testing.assert_close(w.grad.to_dense(), r.grad, rtol=1e-5, atol=1e-6)
    # fused SGD: W <- W - lr * g
    ts2 = rtf.EmbeddingTables(rows, dims, seed=0, optimizer=rtf.SparseOptimizer("sgd", lr=0.5))
    ts2.begin_step()
    out = ts2.lookup(ids)
    (out * out).sum().backward()
    for w2, r in zip(ts2.weights, ref):
        torch.testing.assert_close(w2.detach(), r.detach() - 0.5 * r.grad, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("kind", ["adam", "adagrad", "sgd"])
def test_rows_apply_dense_matches_row_update_restatement(rtf, kind):
    """rtf_rows_apply_dense (replicated small tables, multi-GPU) == the torch restatement of K2's
    row update that the CPU tests check against the oracle, to rounding; untouched rows keep
    their values and state."""
    import ctypes as C
    from recommend_tf2_b200 import _lib as L
    from recommend_tf2_b200.sharded import apply_touched_rows
    torch.manual_seed(0)
    R, D = 1000, 128
    opt = rtf.SparseOptimizer(kind, lr=1e-2, l2=1e-4)
    W = torch.randn(R, D, device="cuda") * 0.05
    m = torch.rand(R, D, device="cuda") * 0.01
    v = torch.rand(R, D, device="cuda") * 0.001
    G = torch.zeros(R + 1, D + 1, device="cuda")
    G[:R, :D] = torch.randn(R, D, device="cuda")
    G[:R, D] = (torch.rand(R, device="cuda") < 0.4).float() * 3.0
    st = opt.struct_for_step(5)
    want = [t.clone() for t in (W, m, v)]
    apply_touched_rows(opt, st.lr, want[0], want[1] if opt.n_states >= 1 else None,
                       want[2] if opt.n_states >= 2 else None, G)
    g, touched = G[:R, :D].contiguous(), G[:R, D].contiguous()
    rc = L.lib().rtf_rows_apply_dense(W.data_ptr(), m.data_ptr() if opt.n_states >= 1 else None,
                                      v.data_ptr() if opt.n_states >= 2 else None, g.data_ptr(),
                                      touched.data_ptr(), R, D, C.byref(st), L.current_stream_ptr())
    assert rc == 0
    untouched = G[:R, D] == 0
    # (the torch restatement forms 1 - beta in double, the kernel in fp32 like K2 and the oracle:
    #  equal to rounding; untouched rows must not move at all)
    # fp32(1) - fp32(0.999) differs from fp32(1 - 0.999) by 1.3e-5 relative: that is the tolerance
    torch.testing.assert_close(W, want[0], rtol=5e-5, atol=1e-7)
    assert torch.equal(W[untouched], want[0][untouched])
    if opt.n_states >= 1:
        torch.testing.assert_close(m, want[1], rtol=5e-5, atol=1e-8)
        assert torch.equal(m[untouched], want[1][untouched])
    if opt.n_states >= 2:
        torch.testing.assert_close(v, want[2], rtol=5e-5, atol=1e-9)


def test_dense_adam_matches_keras_formula(rtf):
    """core.DenseAdam (rtf_dense_adam over the flat buffer) == Keras Adam (App. A12) in fp64."""
    from recommend_tf2_b200.core import DenseAdam
    torch.manual_seed(1)
    ps = [torch.nn.Parameter(torch.randn(n, device="cuda")) for n in (33, 1024, 7)]
    ref = [p.detach().double().cpu() for p in ps]
    mm = [torch.zeros_like(r) for r in ref]
    vv = [torch.zeros_like(r) for r in ref]
    opt = DenseAdam(ps, lr=1e-2)
    for t in range(1, 6):
        opt.zero_grad()
        gs = [torch.randn_like(p) for p in ps]
        for p, g in zip(ps, gs):
            p.grad.add_(g)
        opt.step()
        lr_t = 1e-2 * (1 - 0.999 ** t) ** 0.5 / (1 - 0.9 ** t)
        for r, m_, v_, g in zip(ref, mm, vv, gs):
            g = g.double().cpu()
            m_.mul_(0.9).add_(0.1 * g)
            v_.mul_(0.999).add_(0.001 * g * g)
            r.sub_(lr_t * m_ / (v_.sqrt() + 1e-7))
    for p, r in zip(ps, ref):
        torch.testing.assert_close(p.detach().double().cpu(), r, rtol=1e-5, atol=1e-6)


def test_prepared_work_list_can_be_applied_twice_incl_giant_segments(rtf):
    """K2 split (prepare on the ids, apply on the gradient): the work list of one batch is applied
    twice (the multi-GPU path reduces, then updates) — the arrival counters of long segments, incl.
    the two-level ones, must be back at zero after every apply.  Sums equal the one-shot K2's."""
    case = ("giant", 5000, [1, 7, 3000], [32, 32, 32], [0, 1, 2], 2, None)
    rng, ids, grad = _bwd_inputs(case)
    name, B, rows, dims, ft, L, pool = case
    dids = torch.from_numpy(ids).to(torch.int32).cuda()          # (B, F, L)
    g = torch.from_numpy(grad).cuda()
    ts = rtf.EmbeddingTables(rows, dims, seed=1, optimizer=rtf.SparseOptimizer("sgd", lr=0.0))
    weights = [torch.zeros(n, d, device="cuda") for n, d in zip(rows, dims)]
    keys_w, tot_w, rb = rtf.embed_bwd(weights, ft, dids, g, "BFL", None, want_unique=True)
    h = ts.prepare_backward(dids, ft, "BFL")
    n = B * L * len(ft)
    outs = []
    for _ in range(2):
        uk = torch.full((n,), -1, dtype=torch.int32, device="cuda")
        ug = torch.zeros((n, max(dims)), device="cuda")
        ts.apply_prepared(h, g, None, reduce_only=(uk, ug))
        torch.cuda.synchronize()
        k = keys_w.numel()
        outs.append((uk[:k].clone(), ug[:k].clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert torch.equal(outs[0][0].to(torch.int64) & 0xFFFFFFFF, keys_w)
    assert torch.equal(outs[0][1], tot_w)

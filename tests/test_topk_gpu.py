"""f3: exact top-k inner-product retrieval == ranking the fp64 scores (ties: lower index first)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_topk(u, it, k, chunk=131072):
    """np.argsort of the fp64 scores, chunked; stable sort on (-score) keeps lower indices first."""
    B = u.shape[0]
    best_s = np.full((B, 0), -np.inf)
    best_i = np.zeros((B, 0), np.int64)
    u64 = u.astype(np.float64)
    for s0 in range(0, it.shape[0], chunk):
        sc = u64 @ it[s0:s0 + chunk].astype(np.float64).T
        idx = np.arange(s0, s0 + sc.shape[1])[None, :].repeat(B, 0)
        cs, ci = np.concatenate([best_s, sc], 1), np.concatenate([best_i, idx], 1)
        order = np.argsort(-cs, axis=1, kind="stable")[:, :k]
        best_s, best_i = np.take_along_axis(cs, order, 1), np.take_along_axis(ci, order, 1)
    return best_s, best_i


@pytest.mark.parametrize("B,N,D,k", [(64, 1_000_000, 64, 10), (300, 50_001, 32, 10), (7, 40, 8, 5),
                                     (1030, 4099, 64, 16), (5, 31, 16, 10)])
def test_topk_ip_equals_fp64_argsort(rtf, B, N, D, k):
    rng = np.random.default_rng(4)
    u = rng.normal(0, 1, (B, D)).astype(np.float32)
    it = rng.normal(0, 1, (N, D)).astype(np.float32)
    sc, idx = rtf.topk_ip(torch.from_numpy(u).cuda(), torch.from_numpy(it).cuda(), k)
    ws, wi = _ref_topk(u, it, k)
    kk = min(k, N)
    assert np.array_equal(idx.cpu().numpy()[:, :kk], wi[:, :kk])
    np.testing.assert_allclose(sc.cpu().numpy()[:, :kk], ws[:, :kk], rtol=1e-6, atol=1e-6)


def test_topk_ip_ties_prefer_lower_index_and_index_api(rtf):
    it = torch.zeros(1000, 16, device="cuda")
    it[:, 0] = 1.0                                   # every item scores the same
    it[500:510, 0] = 2.0
    u = torch.zeros(3, 16, device="cuda")
    u[:, 0] = torch.tensor([1.0, 2.0, 0.5], device="cuda")
    index = rtf.IndexFlatIP(16)
    index.add(it[:600])
    index.add(it[600:])
    assert index.ntotal == 1000
    D_, I_ = index.search(u, 12)
    want = list(range(500, 510)) + [0, 1]
    assert I_.cpu().tolist() == [want] * 3
    torch.testing.assert_close(D_[0], torch.tensor([2.0] * 10 + [1.0] * 2, device="cuda"))

"""Input contract: feature-column dictionaries and synthetic batches have the
reference's format (src/ctr/utils/data_process.py:13-30,80-91; src/match/utils/feature_util.py)."""
import numpy as np

from recommend_tf2_b200 import data


def test_feature_dicts_match_reference_format():
    assert data.sparseFeature("C1", 100, embed_dim=8) == {"feat": "C1", "feat_num": 100, "embed_dim": 8}
    assert data.sparseFeature("C1", 100) == {"feat": "C1", "feat_num": 100, "embed_dim": 4}
    assert data.denseFeature("I1") == {"feat": "I1"}
    assert data.sparseFeature("u", 10, feat_len=1, embed_dim=4) == {
        "feat": "u", "feat_num": 10, "feat_len": 1, "embed_dim": 4}
    assert data.varLenSparseFeat("h", 10, 50) == {"feat": "h", "feat_num": 10, "maxlen": 50, "embed_dim": 4}


def test_criteo_feature_columns_and_batches():
    fc = data.criteo_feature_columns(embed_dim=128, row_cap=10_000_000)
    dense, sparse = fc
    assert [f["feat"] for f in dense] == [f"I{i}" for i in range(1, 14)]
    assert [f["feat"] for f in sparse] == [f"C{i}" for i in range(1, 27)]
    rows = [f["feat_num"] for f in sparse]
    assert max(rows) == 10_000_000 and sum(rows) == 33_631_350
    for kind in ("uniform", "zipf"):
        d, s, y = data.synthetic_criteo_batch(np.random.default_rng(3), 512, rows, kind)
        assert d.dtype == np.float32 and d.shape == (512, 13) and 0 <= d.min() and d.max() < 1
        assert s.dtype == np.int32 and s.shape == (512, 26)
        assert (s >= 0).all() and (s < np.asarray(rows)).all()
        assert y.shape == (512, 1) and set(np.unique(y)) <= {0.0, 1.0}
    a = data.synthetic_criteo_batch(np.random.default_rng(3), 64, rows)
    b = data.synthetic_criteo_batch(np.random.default_rng(3), 64, rows)
    assert all(np.array_equal(x, z) for x, z in zip(a, b))          # seeded => reproducible

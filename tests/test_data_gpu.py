"""DeviceFeeder: batches arrive on the device in order, bit-equal to the host copies, with the
copy one step ahead on its own stream."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from recommend_tf2_b200 import data  # noqa: E402


@pytest.mark.parametrize("depth", [1, 2, 4])
def test_feeder_order_and_content(rtf, depth):
    rng = np.random.default_rng(0)
    rows = [100, 7, 5000]
    host = []
    for _ in range(7):
        d, s, y = data.synthetic_criteo_batch(rng, 1024, rows)
        host.append((torch.from_numpy(d), torch.from_numpy(s).pin_memory(), torch.from_numpy(y)))
    feeder = data.DeviceFeeder(host, depth=depth)
    seen = 0
    for i, (d, s, y) in enumerate(feeder):
        assert d.is_cuda and s.is_cuda and y.is_cuda
        # consume on the current stream right away (no explicit sync): the feeder ordered it
        assert torch.equal((d + 0).cpu(), host[i][0]) and torch.equal(s.cpu(), host[i][1])
        assert torch.equal(y.cpu(), host[i][2])
        keep = d * 2.0                              # work enqueued on the batch, read after the
        seen += 1                                   # slot may already be refilling
        if i:
            assert torch.equal(prev.cpu(), host[i - 1][0] * 2.0)
        prev = keep
    assert seen == 7
    assert feeder.h2d_bytes == sum(t.numel() * t.element_size() for b in host for t in b)


def test_feeder_ragged_last_batch(rtf):
    host = [(torch.arange(12.).view(4, 3),), (torch.arange(12., 24.).view(4, 3),), (torch.ones(1, 3),)]
    got = [b[0].cpu().clone() for b in data.DeviceFeeder(host, depth=2)]
    assert [g.shape for g in got] == [(4, 3), (4, 3), (1, 3)]
    assert all(torch.equal(g, h[0]) for g, h in zip(got, host))


def test_feeder_empty(rtf):
    assert list(data.DeviceFeeder([])) == []

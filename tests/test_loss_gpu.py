"""rtf_bce_fwd (Keras binary_crossentropy on probabilities, the loss the reference's CTR scripts
compile with: src/ctr/fm/train.py:49) against the oracle's restatement
(oracle/dlrm_ref.py::bce) in fp64, incl. the gradient through the clip and run-to-run determinism."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _oracle(y, p, dtype=torch.float64):
    from oracle.dlrm_ref import bce
    pr = p.to(dtype).cpu().requires_grad_(True)
    loss = bce(y.to(dtype).cpu().reshape(pr.shape), pr)
    loss.backward()
    return loss.detach().double(), pr.grad.double()


@pytest.mark.parametrize("n", [1, 7, 1024, 1025, 4096, 65536, 100003])
def test_bce_matches_fp64_oracle(rtf, n):
    g = torch.Generator(device="cuda").manual_seed(n)
    p = torch.rand(n, 1, device="cuda", generator=g)
    # predictions saturating the lower clip bound (value and zero gradient outside the interval)
    if n >= 7:
        p[0], p[2], p[3] = 0.0, 1e-7, 1e-9
    y = (torch.rand(n, device="cuda", generator=g) < 0.3).float()
    p.requires_grad_(True)
    loss = rtf.layers.binary_crossentropy(y, p)
    assert loss.shape == ()
    (3.0 * loss).backward()
    want, want_g = _oracle(y, p.detach())
    assert abs(float(loss) - float(want)) <= 1e-6 * abs(float(want))
    got_g = p.grad.double().cpu() / 3.0
    # fp32 evaluation of y/(p+eps) near the clip bounds: 1e-5 relative to the largest entry
    assert float((got_g - want_g).abs().max()) <= 1e-5 * float(want_g.abs().max())
    # outside the clip interval the gradient is exactly zero, on it the clip passes it
    if n >= 7:
        assert float(p.grad[3]) == 0.0 and float(p.grad[2]) != 0.0 and float(p.grad[0]) == 0.0
    p.grad = None
    loss2 = rtf.layers.binary_crossentropy(y, p)
    loss2.backward()
    assert torch.equal(loss, loss2)


def test_bce_upper_clip_bound_in_fp32(rtf):
    """1 - 1e-7 is not representable in fp32: the upper clip bound TF (and this kernel) use is
    fp32(1 - 1e-7) = 1 - 2^-23, so p = 1 gives log(2^-23 + 1e-7), not the fp64 value log(2e-7).
    Checked against the oracle evaluated in fp32 (torch CPU kernels: an independent implementation)."""
    n = 4099
    g = torch.Generator(device="cuda").manual_seed(5)
    p = torch.rand(n, device="cuda", generator=g)
    p[0], p[1], p[2] = 1.0, float(torch.tensor(1.0 - 1e-7, dtype=torch.float32)), 0.0
    y = (torch.rand(n, device="cuda", generator=g) < 0.5).float()
    y[0], y[1] = 0.0, 0.0
    p.requires_grad_(True)
    loss = rtf.layers.binary_crossentropy(y, p)
    loss.backward()
    want, want_g = _oracle(y, p.detach(), torch.float32)
    assert abs(float(loss) - float(want)) <= 2e-6 * abs(float(want))
    assert float((p.grad.double().cpu() - want_g).abs().max()) <= 1e-5 * float(want_g.abs().max())
    assert float(p.grad[0]) == 0.0 and float(p.grad[1]) != 0.0


def test_bce_integer_labels_and_no_grad(rtf):
    p = torch.rand(513, device="cuda")
    y = torch.randint(0, 2, (513, 1), device="cuda")
    with torch.no_grad():
        loss = rtf.layers.binary_crossentropy(y, p)
    want, _ = _oracle(y.reshape(-1).float(), p)
    assert abs(float(loss) - float(want)) <= 1e-6 * abs(float(want))

"""rtf_layernorm_fwd / rtf_layernorm_bwd (the LayerNormalization of the reference's
TransformerEncoder, src/match/layers/modules.py:173-185; Keras semantics App. A8: last axis, biased
variance) against the fp64 formulas, incl. gamma / beta gradients and run-to-run determinism."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(x, gamma, beta, eps, dy):
    x, gamma, beta, dy = (t.double() for t in (x, gamma, beta, dy))
    x = x.detach().requires_grad_(True)
    gamma = gamma.detach().requires_grad_(True)
    beta = beta.detach().requires_grad_(True)
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    y = (x - mu) / torch.sqrt(var + eps) * gamma + beta
    y.backward(dy)
    return y.detach(), x.grad, gamma.grad, beta.grad


def _close(got, want, rtol=1e-5):
    got, want = got.double().cpu(), want.cpu()
    err = (got - want).abs().max().item()
    assert err <= rtol * max(want.abs().max().item(), 1e-30), (err, want.abs().max().item())


@pytest.mark.parametrize("shape,mu", [((1024, 200, 64), 0.0), ((37, 5, 50), 3.0), ((4, 256), 0.0),
                                      ((100000, 16), 10.0), ((3, 7, 130), -2.0), ((1, 1), 0.0)])
def test_layernorm_matches_fp64(rtf, shape, mu):
    torch.manual_seed(sum(shape))
    C = shape[-1]
    x = (torch.randn(*shape, device="cuda") * 1.5 + mu).requires_grad_(True)
    ln = rtf.layers.LayerNormalization(epsilon=1e-6)
    ln(x.detach())
    with torch.no_grad():
        ln.gamma.copy_(torch.rand(C, device="cuda") + 0.5)
        ln.beta.copy_(torch.randn(C, device="cuda"))
    y = ln(x)
    dy = torch.randn(*shape, device="cuda")
    y.backward(dy)
    wy, wdx, wdg, wdb = _ref(x.detach().cpu(), ln.gamma.detach().cpu(), ln.beta.detach().cpu(), 1e-6, dy.cpu())
    tol = 1e-5 if C > 1 else 1e-3        # C = 1: y = beta exactly, gradients of x vanish
    _close(y, wy, tol)
    if C > 1:
        _close(x.grad, wdx, 3e-5)
    rows = x.numel() // C
    # column sums over `rows` terms: scale-aware bound (sum of |terms|)
    xh = (wy - ln.beta.detach().double().cpu()) / ln.gamma.detach().double().cpu()
    sg = (dy.double().cpu() * xh).abs().reshape(rows, C).sum(0).max().item()
    sb = dy.double().cpu().abs().reshape(rows, C).sum(0).max().item()
    assert (ln.gamma.grad.double().cpu() - wdg).abs().max().item() <= 1e-5 * max(sg, 1e-30)
    assert (ln.beta.grad.double().cpu() - wdb).abs().max().item() <= 1e-5 * max(sb, 1e-30)


def test_layernorm_deterministic_and_no_input_grad(rtf):
    torch.manual_seed(1)
    x = torch.randn(5000, 64, device="cuda")
    ln = rtf.layers.LayerNormalization(epsilon=1e-6)
    outs = []
    for _ in range(2):
        ln.zero_grad()
        y = ln(x)
        y.square().sum().backward()
        outs.append((y.detach().clone(), ln.gamma.grad.clone(), ln.beta.grad.clone()))
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)


def test_layernorm_wide_rows_use_the_library_path(rtf):
    x = torch.randn(8, 300, device="cuda", requires_grad=True)
    ln = rtf.layers.LayerNormalization(epsilon=1e-6)
    y = ln(x)
    y.sum().backward()
    torch.testing.assert_close(y, torch.nn.functional.layer_norm(x, (300,), ln.gamma, ln.beta, 1e-6))

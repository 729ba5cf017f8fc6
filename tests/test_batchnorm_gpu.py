"""rtf_bn_fwd / rtf_bn_bwd (BatchNormalization at the head of ctr.layers.modules.DNN,
src/ctr/layers/modules.py:129-135) against an fp64 restatement of the Keras formulas (App. A9):
batch mean, BIASED batch variance, eps 1e-3, moving = moving*momentum + batch*(1-momentum)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

RTOL = 1e-5      # BASELINE north_star: fp32 outputs / gradients within a relative 1e-5


def _ref(x, gamma, beta, eps, dy=None):
    x = x.double()
    mean = x.mean(0)
    var = x.var(0, unbiased=False)
    invstd = 1.0 / torch.sqrt(var + eps)
    g = torch.ones_like(mean) if gamma is None else gamma.double()
    b = torch.zeros_like(mean) if beta is None else beta.double()
    y = (x - mean) * invstd * g + b
    out = {"y": y, "mean": mean, "var": var}
    if dy is not None:
        dy = dy.double()
        n = x.shape[0]
        xh = (x - mean) * invstd
        out["dbeta"] = dy.sum(0)
        out["dgamma"] = (dy * xh).sum(0)
        out["dx"] = (dy - out["dbeta"] / n - xh * out["dgamma"] / n) * g * invstd
    return out


def _close(got, want, rtol=RTOL, scale=None):
    got, want = got.double().cpu(), want.cpu()
    s = want.abs().max().item() if scale is None else scale
    err = (got - want).abs().max().item()
    assert err <= rtol * max(s, 1e-30), (err, s)


@pytest.mark.parametrize("B,C,mu,sigma", [(65536, 13, 0.5, 0.3), (4096, 480, 0.0, 1.0),
                                          (1000, 479, 3.0, 2.0), (33, 7, 0.0, 1.0), (5, 1024, 1.0, 1.0),
                                          (20000, 128, 100.0, 0.1), (1, 8, 0.0, 1.0)])
def test_bn_layer_matches_fp64_formulas(rtf, B, C, mu, sigma):
    torch.manual_seed(B + C)
    x = (torch.randn(B, C, device="cuda") * sigma + mu).requires_grad_(True)
    bn = rtf.layers.BatchNormalization()
    bn.train()
    y = bn(x)
    with torch.no_grad():
        bn.gamma.copy_(torch.rand(C, device="cuda") + 0.5)
        bn.beta.copy_(torch.randn(C, device="cuda"))
        bn.moving_mean.zero_()
        bn.moving_variance.fill_(1.0)
    y = bn(x)
    dy = torch.randn(B, C, device="cuda")
    y.backward(dy)
    want = _ref(x.detach(), bn.gamma.detach(), bn.beta.detach(), 1e-3, dy)
    _close(y, want["y"])
    _close(bn.moving_mean, 0.01 * want["mean"], scale=0.01 * max(want["mean"].abs().max().item(), sigma))
    _close(bn.moving_variance, 0.99 + 0.01 * want["var"])
    _close(bn.beta.grad, want["dbeta"], scale=dy.abs().sum(0).max().item())
    _close(bn.gamma.grad, want["dgamma"], scale=(dy.abs().double().cpu() * ((x.detach().double().cpu() - want["mean"].cpu()).abs() / torch.sqrt(want["var"].cpu() + 1e-3))).sum(0).max().item())
    if B > 1:
        _close(x.grad, want["dx"], rtol=2e-5)


def test_bn_without_scale_center_and_without_input_grad(rtf):
    torch.manual_seed(3)
    x = torch.rand(2048, 36, device="cuda")             # no grad: the raw dense features
    bn = rtf.layers.BatchNormalization(center=False, scale=False)
    bn.train()
    y = bn(x)
    want = _ref(x, None, None, 1e-3)
    _close(y, want["y"])
    assert not y.requires_grad
    bn2 = rtf.layers.BatchNormalization()
    bn2.train()
    y2 = bn2(x)
    y2.sum().backward()
    _close(bn2.beta.grad, torch.full((36,), 2048.0, dtype=torch.float64))
    assert bn2.gamma.grad.abs().max().item() < 1e-2      # sum of x_hat over the batch = 0


def test_bn_strided_rows_and_determinism(rtf):
    torch.manual_seed(4)
    big = torch.randn(3000, 96, device="cuda")
    x = big[:, 16:80].requires_grad_(True)               # ldx = 96, C = 64, 16-byte aligned
    bn = rtf.layers.BatchNormalization()
    bn.train()
    y = bn(x)
    dy = torch.randn(3000, 64, device="cuda")
    (gx,) = torch.autograd.grad(y, x, dy)
    want = _ref(x.detach(), bn.gamma.detach(), bn.beta.detach(), 1e-3, dy)
    _close(y, want["y"])
    _close(gx, want["dx"], rtol=2e-5)
    y_again = bn(x)
    assert torch.equal(y, y_again)                       # fixed summation order
    bn.eval()
    ye = bn(x)
    mm, mv = bn.moving_mean.double(), bn.moving_variance.double()
    _close(ye, ((x.detach().double() - mm) / torch.sqrt(mv + 1e-3)).cpu())


def test_bn_capi_rejects_bad_arguments(rtf):
    import ctypes as C
    L = rtf._lib
    x = torch.zeros(8, 8, device="cuda")
    st = torch.zeros(2, 8, device="cuda")
    ws = torch.zeros(16, dtype=torch.uint8, device="cuda")       # too small
    rc = L.lib().rtf_bn_fwd(x.data_ptr(), 8, 8, 8, None, None, 1e-3, 0.99, x.data_ptr(), 8,
                            st[0].data_ptr(), st[1].data_ptr(), None, None, ws.data_ptr(), ws.numel(), None)
    assert rc == -4      # RTF_E_WORKSPACE
    rc = L.lib().rtf_bn_fwd(None, 8, 8, 8, None, None, 1e-3, 0.99, None, 8, None, None, None, None, None, 0, None)
    assert rc == -1            # RTF_E_ARG

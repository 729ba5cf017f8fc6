"""Dense-MLP GEMMs on the tcgen05 tensor cores (csrc/dense_gemm.cuh) against an
fp64 reference: the 3-band bf16 split must stay at fp32-GEMM accuracy (parity bar 1e-5)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from recommend_tf2_b200 import core  # noqa: E402

TOL = 2e-6     # max |err| / max |ref|; an IEEE fp32 GEMM sits at ~1e-7 for these K


def _rel(out, ref):
    return float((out.double() - ref).abs().max() / ref.abs().max())


@pytest.mark.parametrize("M,K,N", [(2048, 512, 256), (1000, 480, 1024), (4096, 128, 64), (260, 36, 132)])
@pytest.mark.parametrize("relu", [False, True])
def test_nn_bias_relu(rtf, M, K, N, relu):
    g = torch.Generator(device="cuda").manual_seed(M + N)
    x = torch.randn(M, K, device="cuda", generator=g)
    w = torch.randn(K, N, device="cuda", generator=g) * 0.1
    b = torch.randn(N, device="cuda", generator=g)
    ref = x.double() @ w.double() + b.double()
    if relu:
        ref = ref.clamp_min(0)
    out = core.dense_gemm("nn", x, w, b, relu)
    assert out.shape == (M, N)
    assert _rel(out, ref) < TOL
    if relu:
        assert float(out.min()) >= 0.0
    out0 = core.dense_gemm("nn", x, w, None, False)
    assert _rel(out0, x.double() @ w.double()) < TOL


@pytest.mark.parametrize("M,K,N", [(2048, 256, 512), (1000, 1024, 480), (4096, 64, 128)])
def test_nt_dgrad(rtf, M, K, N):
    g = torch.Generator(device="cuda").manual_seed(K)
    dy = torch.randn(M, K, device="cuda", generator=g)
    w = torch.randn(N, K, device="cuda", generator=g) * 0.1
    out = core.dense_gemm("nt", dy, w)
    assert _rel(out, dy.double() @ w.double().t()) < TOL


@pytest.mark.parametrize("B,Kin,N,splits", [(4096, 512, 256, 1), (8192, 480, 1024, 8), (4096, 128, 64, 4)])
def test_tn_wgrad_with_splits(rtf, B, Kin, N, splits):
    g = torch.Generator(device="cuda").manual_seed(B + N)
    x = torch.randn(B, Kin, device="cuda", generator=g)
    dy = torch.randn(B, N, device="cuda", generator=g)
    out = core.dense_gemm("tn", x, dy, splits=splits)
    assert out.shape == (Kin, N)
    assert _rel(out, x.double().t() @ dy.double()) < TOL
    again = core.dense_gemm("tn", x, dy, splits=splits)
    assert torch.equal(out, again)          # fixed split + ordered sum: run-to-run identical


def test_wide_dynamic_range_operands(rtf):
    """Entries spanning 2^-20 .. 2^20: the bf16 split keeps 24 significant bits of each."""
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(1024, 256, device="cuda", generator=g) * torch.exp2(
        torch.randint(-20, 21, (1024, 256), device="cuda", generator=g).float())
    w = torch.randn(256, 128, device="cuda", generator=g)
    out = core.dense_gemm("nn", x, w)
    ref = x.double() @ w.double()
    bound = (x.double().abs() @ w.double().abs())         # componentwise error scale
    assert float(((out.double() - ref).abs() / bound).max()) < 1e-6


def test_dense_layer_matches_library_path(rtf):
    """layers.Dense forward + backward: tensor-core path == framework fp32 GEMM path."""
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(2048, 480, device="cuda", generator=g, requires_grad=True)
    res = {}
    for kind in ("library", "bf16x6"):
        core.set_dense_gemm(kind)
        torch.manual_seed(0)
        layer = core.Dense(256, activation="relu")
        y = layer(x)
        (y * torch.arange(256, device="cuda")).sum().backward()
        res[kind] = (y.detach(), x.grad.clone(), layer.kernel.grad.clone(), layer.bias.grad.clone())
        x.grad = None
    core.set_dense_gemm("bf16x6")
    for a, b in zip(res["library"], res["bf16x6"]):
        torch.testing.assert_close(b, a, rtol=1e-5, atol=1e-5 * float(a.abs().max()))


@pytest.mark.parametrize("B,N", [(65536, 256), (5000, 128), (130, 4), (1, 1024)])
@pytest.mark.parametrize("masked", [True, False])
def test_relu_mask_and_bias_grad_one_pass(rtf, B, N, masked):
    """rtf_relu_bwd_colsum: g = gy * (y > 0) bit-exact, column sums == fp64 sums to fp32 accuracy
    and identical from run to run (fixed two-stage order)."""
    g = torch.Generator(device="cuda").manual_seed(B + N)
    gy = torch.randn(B, N, device="cuda", generator=g)
    y = torch.randn(B, N, device="cuda", generator=g).clamp_min(0) if masked else None
    got_g, got_db = core._relu_bwd_bias_grad(gy, y)
    want_g = gy * (y > 0) if masked else gy
    assert torch.equal(got_g, want_g)
    want_db = want_g.double().sum(0)
    scale = float(want_g.double().abs().sum(0).max()) + 1e-30
    assert float((got_db.double() - want_db).abs().max()) / scale < 1e-6
    _, again = core._relu_bwd_bias_grad(gy, y)
    assert torch.equal(got_db, again)


def test_dense_layer_unaligned_input_width_is_padded(rtf):
    """13 dense features (K % 4 != 0): the layer zero-pads K for the tensor-core path and returns
    gradients of the original shapes, equal to the framework path's."""
    g = torch.Generator(device="cuda").manual_seed(13)
    x = torch.rand(4096, 13, device="cuda", generator=g, requires_grad=True)
    res = {}
    for kind in ("library", "bf16x6"):
        core.set_dense_gemm(kind)
        torch.manual_seed(1)
        layer = core.Dense(512, activation="relu")
        y = layer(x)
        (y * y).sum().backward()
        res[kind] = (y.detach(), x.grad.clone(), layer.kernel.grad.clone(), layer.bias.grad.clone())
        assert x.grad.shape == (4096, 13) and layer.kernel.grad.shape == (13, 512)
        x.grad = None
    core.set_dense_gemm("bf16x6")
    for a, b in zip(res["library"], res["bf16x6"]):
        torch.testing.assert_close(b, a, rtol=1e-5, atol=1e-5 * float(a.abs().max()))

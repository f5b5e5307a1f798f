"""The TMA-fed tcgen05 TF32 GEMM kernels of the training step (csrc/gemm_tf32.cu) against fp64 matmuls.
TF32 products (10-bit mantissa operands, fp32 accumulation): tolerance 2e-3 of the result's scale,
the same contract as the library's TF32 GEMMs these kernels replace (torch allow_tf32)."""
import pytest
import torch

from demf_b200 import _lib
from demf_b200.mm import point_ops as P

pytestmark = pytest.mark.gpu
TOL = 2e-3


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "these tests need a CUDA device"
    _lib.load()
    return torch.device("cuda:0")


def _rel(got, want):
    return (got.double() - want).abs().max().item() / max(want.abs().max().item(), 1e-12)


# (R, K, N): the training shapes of configs/demf/demf_votenet.py:48-62,142-162 at small row counts, ragged rows,
# K / N that are not multiples of 32, N > 256 (two column blocks), K = 8 (SA1 layer 0), K = 512 (FP layers)
SHAPES = [(4096, 8, 64), (4096, 64, 64), (4096, 64, 128), (2048, 132, 128), (2048, 128, 256), (1024, 260, 128),
          (1000, 260, 256), (777, 256, 256), (512, 512, 256), (300, 128, 12), (130, 256, 260), (128 * 150 + 5, 64, 128)]


@pytest.mark.parametrize("R,K,N", SHAPES)
def test_rows_gemm_forward(dev, R, K, N):
    g = torch.Generator(device=dev).manual_seed(R + K + N)
    x = torch.randn(R, K, generator=g, device=dev)
    w = torch.randn(N, K, generator=g, device=dev) / K ** 0.5
    b = torch.randn(N, generator=g, device=dev)
    y = P.gemm_rows_fwd(x, w)
    assert _rel(y, x.double() @ w.double().t()) < TOL
    y = P.gemm_rows_fwd(x, w, bias=b, relu=True)
    assert _rel(y, torch.relu(x.double() @ w.double().t() + b.double())) < TOL
    assert P.gemm_error() == 0


@pytest.mark.parametrize("R,K,N", SHAPES)
def test_rows_gemm_dgrad(dev, R, K, N):
    g = torch.Generator(device=dev).manual_seed(R + K + N + 1)
    dy = torch.randn(R, N, generator=g, device=dev)
    w = torch.randn(N, K, generator=g, device=dev) / N ** 0.5
    dx = P.gemm_rows_dgrad(dy, w)
    assert dx.shape == (R, K)
    assert _rel(dx, dy.double() @ w.double()) < TOL
    assert P.gemm_error() == 0


@pytest.mark.parametrize("R,K,N", SHAPES + [(100000, 64, 128), (33, 128, 128), (1024, 1024, 256)])
def test_wgrad(dev, R, K, N):
    g = torch.Generator(device=dev).manual_seed(R + K + N + 2)
    dy = torch.randn(R, N, generator=g, device=dev)
    x = torch.randn(R, K, generator=g, device=dev)
    base = torch.randn(N, K, generator=g, device=dev)
    dw = base.clone()
    P.gemm_wgrad_(dw, dy, x)
    want = base.double() + dy.double().t() @ x.double()
    assert _rel(dw, want) < TOL
    assert P.gemm_error() == 0


def test_strided_views(dev):
    """Operands and results as column slices of wider row-major buffers (row stride > width)."""
    g = torch.Generator(device=dev).manual_seed(5)
    big = torch.randn(1500, 320, generator=g, device=dev)
    x = big[:, 32:32 + 132]
    w = torch.randn(128, 132, generator=g, device=dev)
    out = torch.zeros(1500, 256, device=dev)
    P.gemm_rows_fwd(x, w, out=out[:, 64:192])
    assert _rel(out[:, 64:192], x.double() @ w.double().t()) < TOL
    assert float(out[:, :64].abs().max()) == 0.0 and float(out[:, 192:].abs().max()) == 0.0


@pytest.mark.parametrize("R,K,N", [(4096, 64, 64), (128 * 150 + 5, 64, 128), (1000, 260, 256), (300, 128, 12)])
def test_forward_bn_statistics_epilogue(dev, R, K, N):
    """Per-channel sum / sum of squares from the GEMM epilogue -> the BatchNorm statistics torch computes on y."""
    g = torch.Generator(device=dev).manual_seed(R + N)
    x = torch.randn(R, K, generator=g, device=dev) + 0.3
    w = torch.randn(N, K, generator=g, device=dev) / K ** 0.5
    state = P.bn_rows_state(N if P.bn_rows_supported(N) else 256, dev)
    rm, rv = torch.zeros(N, device=dev), torch.ones(N, device=dev)
    for _ in range(2):      # the accumulators are left zero: a second call gives the same statistics
        y = P.gemm_rows_fwd(x, w, bn_state=state)
        mean, invstd = P.bn_finalize(state, R, N, 1e-5, 0.1, rm, rv)
        yd = y.double()
        assert torch.allclose(mean.double(), yd.mean(0), atol=1e-5, rtol=1e-5)
        assert torch.allclose(invstd.double(), 1.0 / torch.sqrt(yd.var(0, unbiased=False) + 1e-5), rtol=1e-4)
    assert float(state.abs().max()) == 0.0
    want_rm = 0.1 * yd.mean(0) * (1 + 0.9)
    assert torch.allclose(rm.double(), want_rm, atol=1e-5, rtol=1e-4)
    assert P.gemm_error() == 0


@pytest.mark.parametrize("R,N,K", [(4096, 64, 64), (2000, 128, 64), (128 * 9 + 77, 256, 128), (640, 128, 256)])
def test_dgrad_through_batchnorm_relu(dev, R, N, K):
    """gemm_rows_dgrad_bn + bn_bwd_from_masked == autograd through z = relu(batch_norm(y_prev)); out = z @ w^T
    (fp64 reference): masked data gradient, grad_gamma, grad_beta and dL/dy_prev."""
    g = torch.Generator(device=dev).manual_seed(R + N + K)
    y_prev = torch.randn(R, K, generator=g, device=dev) * 1.5 + 0.2
    w = torch.randn(N, K, generator=g, device=dev) / K ** 0.5
    dy = torch.randn(R, N, generator=g, device=dev)
    gamma = torch.rand(K, generator=g, device=dev) + 0.5
    beta = torch.randn(K, generator=g, device=dev) * 0.3
    mean = y_prev.mean(0)
    invstd = 1.0 / torch.sqrt(y_prev.var(0, unbiased=False) + 1e-5)
    state = P.bn_rows_state(K, dev)
    gm = P.gemm_rows_dgrad_bn(dy, w, y_prev, mean, invstd, gamma, beta, state)
    gyp, ggamma, gbeta = P.bn_bwd_from_masked(gm, y_prev, gamma, mean, invstd, state)
    assert float(state.abs().max()) == 0.0          # accumulators handed back zeroed
    yd = y_prev.double().requires_grad_(True)
    gd, bd = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    z = torch.relu(torch.nn.functional.batch_norm(yd, None, None, gd, bd, True, 0.0, 1e-5))
    (z @ w.double().t()).backward(dy.double())
    mask = (z > 0).double()
    assert _rel(gm, (dy.double() @ w.double()) * mask) < TOL
    assert _rel(gyp, yd.grad) < 2 * TOL
    assert _rel(ggamma, gd.grad) < 2 * TOL and _rel(gbeta, bd.grad) < 2 * TOL
    assert P.gemm_error() == 0


@pytest.mark.parametrize("R,K,N", [(1024, 256, 259), (2048, 128, 30), (1024, 1024, 256), (1024, 256, 1024)])
def test_linear_rows_autograd(dev, R, K, N):
    """bricks.linear_rows (padded output widths, K up to 2048) forward and all three gradients vs fp64 autograd."""
    from demf_b200.mm import bricks
    g = torch.Generator(device=dev).manual_seed(R + K + N)
    x = torch.randn(R, K, generator=g, device=dev, requires_grad=True)
    w = (torch.randn(N, K, generator=g, device=dev) / K ** 0.5).requires_grad_(True)
    b = torch.randn(N, generator=g, device=dev, requires_grad=True)
    gy = torch.randn(R, N, generator=g, device=dev)
    tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n0 = _lib.launch_count()
        y = bricks.linear_rows(x, w, b)
        y.backward(gy)
        assert _lib.launch_count() - n0 >= 3          # our kernels, not the library
    finally:
        torch.backends.cuda.matmul.allow_tf32 = tf32
    xd, wd, bd = (t.detach().double().requires_grad_(True) for t in (x, w, b))
    yd = torch.nn.functional.linear(xd, wd, bd)
    yd.backward(gy.double())
    assert y.shape == (R, N)
    assert _rel(y.detach(), yd.detach()) < TOL and _rel(x.grad, xd.grad) < TOL
    assert _rel(w.grad, wd.grad) < TOL and _rel(b.grad, bd.grad) < TOL
    assert P.gemm_error() == 0


@pytest.mark.gpu
@pytest.mark.parametrize("R,N", [(1, 3), (1024, 259), (21760, 256), (37, 30), (100000, 8)])
def test_col_sum_add_matches_torch(R, N):
    """Bias gradient in one launch: out += x.sum(0) (csrc/bn_rows.cu col_sum_add_kernel), also on a strided view."""
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(R + N)
    base = torch.randn(R, N + 4, generator=g, device=dev)
    x = base[:, :N]                              # row stride N + 4
    out = torch.randn(N, generator=g, device=dev)
    want = out.double() + x.double().sum(0)
    P.col_sum_add_(out, x)
    torch.cuda.synchronize()
    scale = x.abs().double().sum(0).max().item() + 1.0
    assert (out.double() - want).abs().max().item() <= 1e-6 * scale

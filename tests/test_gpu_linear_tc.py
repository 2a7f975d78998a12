"""The tcgen05 fully connected kernel (sdn_linear_forward, csrc/sdn_linear_tc.cu) against a plain PyTorch fp32
reference of the same op.  The kernel rounds its operands to bf16 and accumulates in fp32, so against the SAME rounded
operands only the summation order differs (<= 1e-5 rel-L2, stated here); against unrounded fp32 the bf16 operand
rounding bounds the error at ~2^-8 relative per product (<= 1e-2 rel-L2 for N(0,1) data)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(128, 128, 64), (128, 128, 16), (1, 1552, 256), (32, 256, 260), (200, 300, 512), (7, 5, 3), (513, 1552, 256),
          (129, 129, 65), (2048, 256, 512)]


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available()
    import sdn_b200
    sdn_b200.lib()
    return torch.device('cuda:0')


def _rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize('M,N,K', SHAPES)
@pytest.mark.parametrize('relu', [False, True])
def test_linear_tc_forward(dev, M, N, K, relu):
    import sdn_b200
    g = torch.Generator(device='cpu').manual_seed(M * 7 + N * 3 + K)
    x = torch.randn(M, K, generator=g).to(dev)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev)
    b = torch.randn(N, generator=g).to(dev)
    y = sdn_b200.linear_tc(x, w, b, relu)
    exact = x.bfloat16().float() @ w.bfloat16().float().t() + b
    full = x @ w.t() + b
    if relu:
        exact, full = exact.clamp_min(0), full.clamp_min(0)
    assert y.shape == (M, N) and y.dtype == torch.float32
    assert _rel(y, exact) <= 1e-5
    assert _rel(y, full) <= 1e-2
    # no bias, leading batch dimensions
    y2 = sdn_b200.linear_tc(x.reshape(1, M, K), w, None, False)
    assert y2.shape == (1, M, N) and _rel(y2[0], x.bfloat16().float() @ w.bfloat16().float().t()) <= 1e-5


def test_linear_tc_is_deterministic_and_stream_ordered(dev):
    import sdn_b200
    torch.manual_seed(1)
    x, w, b = torch.randn(300, 260, device=dev), torch.randn(1552, 260, device=dev), torch.randn(1552, device=dev)
    y0 = sdn_b200.linear_tc(x, w, b, True)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        y1 = sdn_b200.linear_tc(x, w, b, True)
    s.synchronize()
    assert torch.equal(y0, y1)


@pytest.mark.parametrize('relu', [False, True])
def test_linear_tc_backward_matches_fp32_linear(dev, relu):
    """The backward is the fp32 matmul pair on the unrounded operands; the ReLU mask is the forward output's."""
    import sdn_b200
    torch.manual_seed(2)
    x = torch.randn(64, 260, device=dev, requires_grad=True)
    lin = torch.nn.Linear(260, 256).to(dev)
    gy = torch.randn(64, 256, device=dev)
    y = sdn_b200.linear_tc(x, lin.weight, lin.bias, relu)
    gx, gw, gb = torch.autograd.grad(y, [x, lin.weight, lin.bias], gy)
    gm = gy * (y > 0).float() if relu else gy
    rx, rw, rb = torch.autograd.grad(lin(x), [x, lin.weight, lin.bias], gm)
    for a, r in ((gx, rx), (gw, rw), (gb, rb)):
        assert _rel(a, r) <= 1e-5


def test_encoder_head_on_tensor_cores_matches_fp32_head(dev):
    """derender3d/models/derenderer.py:35-43 with the four Linear layers on tcgen05 vs on cuBLAS fp32."""
    from derender3d.models.derenderer import Derenderer
    torch.manual_seed(3)
    enc = Derenderer().to(dev).eval()
    images = torch.randn(5, 3, 224, 224, device=dev)
    mroi, droi = torch.rand(5, 2, device=dev), torch.rand(5, 2, device=dev)
    with torch.no_grad():
        enc.tensor_core_head = True
        a = enc(images, mroi, droi)
        enc.tensor_core_head = False
        b = enc(images, mroi, droi)
    for k in b:
        assert a[k].shape == b[k].shape
        assert _rel(a[k].float(), b[k].float()) <= 2e-2, k


def test_linear_tc_errors(dev):
    import sdn_b200
    with pytest.raises(sdn_b200.SdnError):
        sdn_b200.linear_tc(torch.randn(4, 8), torch.randn(3, 8))                         # CPU tensors: no fallback
    with pytest.raises(sdn_b200.SdnError):
        sdn_b200.linear_tc(torch.randn(4, 8, device=dev), torch.randn(3, 9, device=dev))  # K mismatch
    assert sdn_b200.linear_tc(torch.randn(0, 8, device=dev), torch.randn(3, 8, device=dev)).shape == (0, 3)

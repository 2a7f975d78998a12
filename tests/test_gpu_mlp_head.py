"""The encoder head as one launch (sdn_mlp_head_forward, csrc/sdn_linear_tc.cu: fp32 FMA, a warp per output feature, grid-wide
barriers between the four layers) against the plain torch.nn.Linear head of derender3d/models/derenderer.py:26-32,35-43 in
fp32: only the summation order differs (<= 1e-5 rel-L2)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available()
    import sdn_b200
    sdn_b200.lib()
    return torch.device('cuda:0')


def _rel(a, b):
    return float((a.detach().double() - b.detach().double()).norm() / (b.detach().double().norm() + 1e-30))


def _layers(dev, K0=512, N0=1000, E=4, N1=256, N2=256, N3=43, seed=0):
    torch.manual_seed(seed)
    return [torch.nn.Linear(K0, N0).to(dev), torch.nn.Linear(N0 + E, N1).to(dev), torch.nn.Linear(N1, N2).to(dev),
            torch.nn.Linear(N2, N3).to(dev)]


def _torch_head(x, extra, fcs):
    h = torch.cat([F.relu(fcs[0](x)), extra], dim=1)
    return fcs[3](F.relu(fcs[2](F.relu(fcs[1](h)))))


@pytest.mark.timeout(120)
@pytest.mark.parametrize('M', [1, 16, 37, 64])
def test_fused_head_forward_and_backward(dev, M):
    import sdn_b200
    fcs = _layers(dev)
    g = torch.Generator().manual_seed(M)
    x = torch.randn(M, 512, generator=g).to(dev).requires_grad_(True)
    extra = torch.randn(M, 4, generator=g).to(dev).requires_grad_(True)
    gy = torch.randn(M, 43, generator=g).to(dev)
    ref = _torch_head(x, extra, fcs)
    ref.backward(gy)
    want = [x.grad.clone(), extra.grad.clone()] + [p.grad.clone() for fc in fcs for p in (fc.weight, fc.bias)]
    x.grad = extra.grad = None
    for fc in fcs:
        fc.zero_grad(set_to_none=True)
    n0 = sdn_b200.launch_count()
    y = sdn_b200.mlp_head(x, extra, *fcs)
    assert sdn_b200.launch_count() - n0 == 1
    assert y.shape == ref.shape and _rel(y, ref) <= 1e-5, _rel(y, ref)
    y.backward(gy)
    got = [x.grad, extra.grad] + [p.grad for fc in fcs for p in (fc.weight, fc.bias)]
    for a, b in zip(got, want):
        assert a.shape == b.shape and _rel(a, b) <= 1e-4, _rel(a, b)


@pytest.mark.timeout(120)
def test_fused_head_in_the_encoder_and_as_cuda_graph(dev):
    """Derenderer(fused_head=True) gives the default module's outputs; the launch (with its barrier clear) replays from a
    CUDA graph."""
    import sdn_b200
    from derender3d.models.derenderer import Derenderer
    torch.manual_seed(3)
    enc = Derenderer().to(dev).eval()
    g = torch.Generator().manual_seed(4)
    images = torch.randn(5, 3, 224, 224, generator=g).to(dev)
    m, d = (torch.randn(5, 2, generator=g) * 0.1).to(dev), (0.2 + torch.rand(5, 2, generator=g)).to(dev)
    with torch.no_grad():
        ref = enc(images, m, d)
        enc.fused_head = True
        out = enc(images, m, d)
    for k in ref:
        assert _rel(out[k], ref[k]) <= 1e-4, (k, _rel(out[k], ref[k]))
    fcs = _layers(dev, seed=5)
    x, extra = torch.randn(16, 512, device=dev), torch.randn(16, 4, device=dev)
    with torch.no_grad():
        want = sdn_b200.mlp_head(x, extra, *fcs)
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            sdn_b200.mlp_head(x, extra, *fcs)
        torch.cuda.current_stream().wait_stream(s)
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            ys = sdn_b200.mlp_head(x, extra, *fcs)
        for _ in range(3):
            gr.replay()
        torch.cuda.synchronize()
    assert torch.equal(ys, want)

"""Backward of the tcgen05 convolution: weight gradient (sdn_conv2d_wgrad: GEMM over pixels with MN-major operands) and data
gradient (the forward kernel on flipped / transposed weights, parity phases for stride 2) against torch autograd in fp32 on
the same bf16-rounded operands."""
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
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


CASES = [   # N, H, W, Cin, Cout, K, stride, pad: the ResNet-18 layer shapes (smaller batch) + ragged ones
    (4, 56, 56, 64, 64, 3, 1, 1), (4, 56, 56, 64, 128, 3, 2, 1), (4, 56, 56, 64, 128, 1, 2, 0), (4, 28, 28, 128, 128, 3, 1, 1),
    (4, 28, 28, 128, 256, 3, 2, 1), (4, 14, 14, 256, 256, 3, 1, 1), (4, 14, 14, 256, 512, 3, 2, 1), (4, 14, 14, 256, 512, 1, 2, 0),
    (6, 7, 7, 512, 512, 3, 1, 1), (3, 10, 18, 64, 192, 3, 1, 1), (1, 112, 112, 192, 64, 1, 1, 0),
]


@pytest.mark.parametrize('N,H,W,Cin,Cout,K,stride,pad', CASES)
def test_weight_and_data_gradients(dev, N, H, W, Cin, Cout, K, stride, pad):
    import sdn_b200
    g_ = torch.Generator().manual_seed(N + H + Cin + Cout + K + stride)
    x = torch.randn(N, Cin, H, W, generator=g_).bfloat16().float().requires_grad_(True)
    w = (torch.randn(Cout, Cin, K, K, generator=g_) / (Cin * K * K) ** 0.5).bfloat16().float().requires_grad_(True)
    y = F.conv2d(x, w, stride=stride, padding=pad)
    go = torch.randn(y.shape, generator=g_).bfloat16().float()
    y.backward(go)
    xn = x.detach().to(dev).bfloat16().permute(0, 2, 3, 1).contiguous()
    gn = go.to(dev).bfloat16().permute(0, 2, 3, 1).contiguous()
    gw = sdn_b200.conv2d_wgrad_tc(gn, xn, K, K, stride, pad)                   # [Cout, K, K, Cin] fp32
    assert _rel(gw.permute(0, 3, 1, 2), w.grad) <= 2e-3, _rel(gw.permute(0, 3, 1, 2), w.grad)
    gx = sdn_b200.conv2d_dgrad_tc(gn, w.detach().to(dev), (H, W), stride, pad)  # [N, H, W, Cin] bf16
    assert _rel(gx.permute(0, 3, 1, 2).float(), x.grad) <= 4e-3, _rel(gx.permute(0, 3, 1, 2).float(), x.grad)


def test_data_gradient_adds_the_skip_gradient(dev):
    import sdn_b200
    g_ = torch.Generator().manual_seed(5)
    w = (torch.randn(64, 64, 3, 3, generator=g_) / 24.0).bfloat16().float()
    go = torch.randn(2, 64, 12, 20, generator=g_).bfloat16().float()
    skip = torch.randn(2, 64, 12, 20, generator=g_).bfloat16()
    want = F.conv_transpose2d(go, w, padding=1) + skip.float()
    got = sdn_b200.conv2d_dgrad_tc(go.to(dev).bfloat16().permute(0, 2, 3, 1).contiguous(), w.to(dev), (12, 20), 1, 1,
                                   residual=skip.to(dev).permute(0, 2, 3, 1).contiguous())
    assert _rel(got.permute(0, 3, 1, 2).float(), want) <= 4e-3


@pytest.mark.parametrize('shape,relu,skip', [((4, 28, 28, 128), True, True), ((8, 7, 7, 512), True, False),
                                             ((3, 10, 18, 64), False, False), ((2, 56, 56, 64), True, True)])
def test_batch_norm_training_forward_and_backward(dev, shape, relu, skip):
    """sdn_affine_act_nhwc + sdn_batch_norm_backward_nhwc against torch's BatchNorm2d (training) autograd in fp32 on the
    same bf16-rounded z / residual / upstream gradient.  The ReLU mask is taken from the stored bf16 output, so elements
    whose fp32 pre-activation is within bf16 rounding of 0 may differ: compared in rel-L2."""
    import sdn_b200
    N, H, W, C = shape
    g_ = torch.Generator().manual_seed(C + H)
    z = (torch.randn(N, C, H, W, generator=g_) * 1.5 + 0.3).bfloat16().float().to(dev).requires_grad_(True)
    res = torch.randn(N, C, H, W, generator=g_).bfloat16().float().to(dev).requires_grad_(True) if skip else None
    gamma = (1 + 0.2 * torch.randn(C, generator=g_)).to(dev).requires_grad_(True)
    beta = (0.1 * torch.randn(C, generator=g_)).to(dev).requires_grad_(True)
    go = torch.randn(N, C, H, W, generator=g_).bfloat16().float().to(dev)
    y = F.batch_norm(z, None, None, gamma, beta, True, 0.1, 1e-5)
    if skip:
        y = y + res
    if relu:
        y = torch.relu(y)
    y.backward(go)

    nh = lambda t: t.detach().permute(0, 2, 3, 1).contiguous().bfloat16()
    zz = z.detach()
    mean = zz.mean(dim=(0, 2, 3))
    rstd = torch.rsqrt(zz.var(dim=(0, 2, 3), unbiased=False) + 1e-5)
    scale = gamma.detach() * rstd
    yn = sdn_b200.affine_act_nhwc(nh(z), scale, beta.detach() - mean * scale, nh(res) if skip else None, relu)
    assert _rel(yn.float().permute(0, 3, 1, 2), y) <= 4e-3
    gz, gg, gb, gs = sdn_b200.batch_norm_backward_nhwc(nh(go), yn if relu else None, nh(z), mean, rstd, gamma, relu, skip)
    assert _rel(gz.float().permute(0, 3, 1, 2), z.grad) <= 1e-2, _rel(gz.float().permute(0, 3, 1, 2), z.grad)
    assert _rel(gg, gamma.grad) <= 5e-3 and _rel(gb, beta.grad) <= 5e-3, (_rel(gg, gamma.grad), _rel(gb, beta.grad))
    if skip:
        assert _rel(gs.float().permute(0, 3, 1, 2), res.grad) <= 5e-3
    else:
        assert gs is None


def test_batch_norm_finalize_updates_running_statistics_like_torch(dev):
    """sdn_batch_norm_finalize: batch mean / rstd / scale / shift and the in-place running statistics against
    torch.nn.BatchNorm2d in training mode fed the same (bf16-rounded) convolution result."""
    import sdn_b200
    g_ = torch.Generator().manual_seed(11)
    N, H, W, C = 6, 14, 14, 256
    z = (torch.randn(N, C, H, W, generator=g_) * 2 + 0.5).bfloat16().float().to(dev)
    bn = torch.nn.BatchNorm2d(C).to(dev).train()
    mine = torch.nn.BatchNorm2d(C).to(dev).train()
    with torch.no_grad():
        bn.weight.copy_(1 + 0.1 * torch.randn(C, generator=g_)); bn.bias.copy_(0.1 * torch.randn(C, generator=g_))
        bn.running_mean.copy_(torch.randn(C, generator=g_)); bn.running_var.copy_(1 + torch.rand(C, generator=g_))
    mine.load_state_dict(bn.state_dict())
    y = bn(z)
    stats = torch.stack([z.sum(dim=(2, 3)), (z * z).sum(dim=(2, 3))], dim=2).contiguous()      # [N, C, 2]
    mean, rstd, scale, shift = sdn_b200.batch_norm_finalize(stats, N * H * W, mine)
    assert _rel(mean, z.mean(dim=(0, 2, 3))) <= 1e-5
    assert _rel(rstd, torch.rsqrt(z.var(dim=(0, 2, 3), unbiased=False) + bn.eps)) <= 1e-5
    assert _rel(z * scale[None, :, None, None] + shift[None, :, None, None], y) <= 1e-5
    assert _rel(mine.running_mean, bn.running_mean) <= 1e-6 and _rel(mine.running_var, bn.running_var) <= 1e-5
    assert int(mine.num_batches_tracked) == int(bn.num_batches_tracked) == 1


@pytest.mark.parametrize('co,ci,k,stride,chlast', [(128, 64, 3, 1, True), (64, 192, 3, 1, False), (128, 64, 3, 2, True),
                                                   (512, 256, 1, 2, False), (64, 64, 1, 1, True)])
def test_transposed_weight_operands(dev, co, ci, k, stride, chlast):
    """sdn_pack_conv_weight_transposed against the torch expression of the same operand (bit-exact: one bf16 rounding)."""
    import sdn_b200
    w = torch.randn(co, ci, k, k, generator=torch.Generator().manual_seed(co + ci)).to(dev)
    if chlast:
        w = w.contiguous(memory_format=torch.channels_last)
    ops = sdn_b200.pack_conv_weight_transposed(w, stride)
    wT = w.transpose(0, 1)                                  # [Cin, Cout, KH, KW]
    if stride == 2 and k == 3:
        taps = {0: [1], 1: [2, 0]}
        want = [wT[:, :, taps[ph], :][:, :, :, taps[pw]].permute(0, 2, 3, 1).bfloat16()
                for ph, pw in ((0, 0), (0, 1), (1, 0), (1, 1))]
    else:
        want = [wT.flip(2, 3).permute(0, 2, 3, 1).bfloat16()]
    assert len(ops) == len(want)
    for a, b in zip(ops, want):
        assert a.shape == b.shape and a.is_contiguous() and torch.equal(a, b)


@pytest.mark.parametrize('shape,k,s,p', [((4, 112, 112, 64), 3, 2, 1), ((2, 17, 23, 64), 3, 2, 1), ((3, 12, 12, 128), 2, 2, 0),
                                         ((2, 9, 9, 64), 3, 1, 1)])
def test_max_pool_forward_and_backward_match_torch_exactly(dev, shape, k, s, p):
    """sdn_max_pool_nhwc / sdn_max_pool_backward_nhwc against torch's max_pool2d + autograd on the same bf16 tensor with MANY
    ties (post-ReLU values on a coarse grid): the gradient of a window must go to the same element torch picks."""
    import sdn_b200
    N, H, W, C = shape
    g_ = torch.Generator().manual_seed(H * W + C)
    x = (torch.relu(torch.randn(N, C, H, W, generator=g_)) * 4).round().div(4).bfloat16().to(dev)       # ties, zeros
    xr = x.clone().float().requires_grad_(True)
    y = F.max_pool2d(xr, k, s, p)
    go = torch.randn(y.shape, generator=g_).bfloat16().to(dev)
    y.backward(go.float())
    xn = x.permute(0, 2, 3, 1).contiguous()
    yn, am = sdn_b200.max_pool_nhwc(xn, k, s, p, return_argmax=True)
    assert torch.equal(sdn_b200.max_pool_nhwc(xn, k, s, p), yn)
    assert torch.equal(yn.permute(0, 3, 1, 2).float(), y.detach())
    gx = sdn_b200.max_pool_backward_nhwc(am, go.permute(0, 2, 3, 1).contiguous(), (H, W), k, s, p)
    want = xr.grad.bfloat16()          # sums of at most 4 bf16 values in fp32, rounded once
    assert torch.equal(gx.permute(0, 3, 1, 2), want), float((gx.permute(0, 3, 1, 2).float() - want.float()).abs().max())


def test_gradients_accumulated_in_place(dev):
    """accumulate_into: the weight gradient (vector reductions into an existing tensor, one and several pixel ranges) and
    BatchNorm's parameter gradients are ADDED to what the targets hold."""
    import sdn_b200
    g_ = torch.Generator().manual_seed(5)
    for (N, H, W, Cin, Cout, K, s, p) in [(4, 28, 28, 128, 128, 3, 1, 1), (6, 7, 7, 512, 512, 3, 1, 1), (4, 56, 56, 64, 128, 1, 2, 0)]:
        x = torch.randn(N, H, W, Cin, generator=g_).bfloat16().to(dev)
        Ho, Wo = (H + 2 * p - K) // s + 1, (W + 2 * p - K) // s + 1
        go = torch.randn(N, Ho, Wo, Cout, generator=g_).bfloat16().to(dev)
        base = torch.randn(Cout, K, K, Cin, generator=g_).to(dev)
        fresh = sdn_b200.conv2d_wgrad_tc(go, x, K, K, s, p)
        acc = base.clone()
        out = sdn_b200.conv2d_wgrad_tc(go, x, K, K, s, p, accumulate_into=acc)
        assert out.data_ptr() == acc.data_ptr()
        assert _rel(acc - base, fresh) <= 2e-4, _rel(acc - base, fresh)
    C, M = 128, 4 * 14 * 14
    z = torch.randn(4, 14, 14, C, generator=g_).bfloat16().to(dev)
    gy = torch.randn(4, 14, 14, C, generator=g_).bfloat16().to(dev)
    mean, rstd, gamma = z.float().mean(dim=(0, 1, 2)), torch.ones(C, device=dev), torch.ones(C, device=dev)
    gz0, gg0, gb0, _ = sdn_b200.batch_norm_backward_nhwc(gy, None, z, mean, rstd, gamma, False)
    ag, ab = torch.full((C,), 2.0, device=dev), torch.full((C,), -1.0, device=dev)
    gz1, _, _, _ = sdn_b200.batch_norm_backward_nhwc(gy, None, z, mean, rstd, gamma, False, accumulate_into=(ag, ab))
    assert _rel(gz1, gz0) <= 2e-3           # the per-channel sums are float atomics: bf16 roundings of gz may flip
    assert _rel(ag - 2.0, gg0) <= 1e-5 and _rel(ab + 1.0, gb0) <= 1e-5

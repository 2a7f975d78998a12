"""sdn_conv2d_forward (implicit GEMM: TMA tensor-map loads, tcgen05.mma, fp32 accumulation in tensor memory) against
torch's fp32 convolution on the SAME bf16-rounded operands.  The output is stored as bf16 (relative rounding 2^-9), so the
bar is rel-L2 <= 4e-3 on the result and <= 2e-2 against the unrounded fp32 convolution (operand rounding 2^-9 each)."""
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
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _case(dev, N, H, W, Cin, Cout, K, stride, pad, seed, relu=False, residual=False, affine=False):
    import sdn_b200
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, K, K, generator=g) / (Cin * K * K) ** 0.5
    xb, wb = x.bfloat16().float(), w.bfloat16().float()
    ref = F.conv2d(xb, wb, stride=stride, padding=pad)
    scale = shift = res = None
    if affine:
        scale, shift = 0.5 + torch.rand(Cout, generator=g), torch.randn(Cout, generator=g)
        ref = ref * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    if residual:
        res = torch.randn(ref.shape, generator=g).bfloat16()
        ref = ref + res.float()
    if relu:
        ref = ref.clamp_min(0)
    xn = x.to(dev).bfloat16().permute(0, 2, 3, 1).contiguous()
    wp = sdn_b200.pack_conv_weight(w.to(dev))
    if xn.shape[-1] % 64:
        xn = F.pad(xn, (0, wp.shape[-1] - xn.shape[-1]))
    out = sdn_b200.conv2d_tc(xn, wp, stride, pad, None if scale is None else scale.to(dev),
                             None if shift is None else shift.to(dev),
                             None if res is None else res.to(dev).permute(0, 2, 3, 1).contiguous(), relu)
    got = out[..., :Cout].permute(0, 3, 1, 2).float().cpu()
    assert got.shape == ref.shape
    return _rel(got, ref), _rel(got, F.conv2d(x, w, stride=stride, padding=pad)) if not (affine or residual or relu) else 0.0


RESNET18_B16 = [    # every distinct convolution shape of torchvision resnet18 at 224x224 after the stem (H = W)
    (56, 64, 64, 3, 1, 1), (56, 64, 128, 3, 2, 1), (56, 64, 128, 1, 2, 0), (28, 128, 128, 3, 1, 1),
    (28, 128, 256, 3, 2, 1), (28, 128, 256, 1, 2, 0), (14, 256, 256, 3, 1, 1), (14, 256, 512, 3, 2, 1),
    (14, 256, 512, 1, 2, 0), (7, 512, 512, 3, 1, 1),
]


@pytest.mark.parametrize('HW,Cin,Cout,K,stride,pad', RESNET18_B16)
def test_resnet18_layer_shapes_batch16(dev, HW, Cin, Cout, K, stride, pad):
    e_same, e_fp32 = _case(dev, 16, HW, HW, Cin, Cout, K, stride, pad, seed=HW + Cout + K)
    assert e_same <= 4e-3, e_same
    assert e_fp32 <= 2e-2, e_fp32


def test_stem_7x7_through_im2col(dev):
    """conv1 of ResNet-18 (3 -> 64, 7x7, stride 2, pad 3): im2col to 192 channels, then a 1x1 convolution."""
    import sdn_b200
    g = torch.Generator().manual_seed(5)
    x = torch.randn(4, 3, 224, 224, generator=g)
    w = torch.randn(64, 3, 7, 7, generator=g) / 12.0
    ref = F.conv2d(x.bfloat16().float(), w.bfloat16().float(), stride=2, padding=3)
    cols = sdn_b200.im2col_nhwc(x.to(dev).bfloat16().permute(0, 2, 3, 1).contiguous(), 7, 7, 2, 3, 192)
    assert cols.shape == (4, 112, 112, 192)
    wk = torch.zeros(64, 1, 1, 192, device=dev, dtype=torch.bfloat16)
    wk[:, 0, 0, :147] = w.to(dev).permute(0, 2, 3, 1).reshape(64, 147).bfloat16()        # (kh, kw, ci) order
    out = sdn_b200.conv2d_tc(cols, wk, 1, 0)
    assert _rel(out.permute(0, 3, 1, 2).float().cpu(), ref) <= 4e-3


@pytest.mark.parametrize('relu,residual,affine', [(True, False, True), (True, True, True), (False, True, False)])
def test_epilogue_scale_shift_residual_relu(dev, relu, residual, affine):
    e, _ = _case(dev, 3, 20, 28, 128, 192, 3, 1, 1, seed=9, relu=relu, residual=residual, affine=affine)
    assert e <= 4e-3, e


@pytest.mark.parametrize('N,H,W,Cin,Cout,K,stride,pad', [
    (1, 9, 13, 64, 64, 3, 1, 1),        # ragged: partial tiles in both directions
    (5, 7, 7, 64, 128, 3, 1, 1),        # several images per 128-pixel tile, odd batch
    (2, 33, 65, 64, 64, 3, 2, 1),       # stride 2 with odd sizes
    (1, 16, 48, 128, 64, 7, 1, 0),      # 7x7 valid convolution (the generator's stem after reflection padding)
    (2, 8, 8, 1024, 1024, 3, 1, 0),     # 1024 channels (the generator's residual blocks), valid
    (1, 12, 20, 40, 48, 3, 1, 1),       # channel counts that need padding to 64
])
def test_shapes_and_edges(dev, N, H, W, Cin, Cout, K, stride, pad):
    e, _ = _case(dev, N, H, W, Cin, Cout, K, stride, pad, seed=N + H + W)
    assert e <= 4e-3, e


def test_few_output_channels_strided_output_and_extent_override(dev):
    import sdn_b200
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, 64, 10, 14, generator=g)
    w = torch.randn(3, 64, 7, 7, generator=g) / 56.0
    ref = torch.tanh(F.conv2d(x.bfloat16().float(), w.bfloat16().float(), padding=3))
    wp = sdn_b200.pack_conv_weight(w.to(dev))
    assert wp.shape[0] == 16                                   # 3 outputs -> the 16-column tile, not 64
    out = sdn_b200.conv2d_tc(x.to(dev).bfloat16().permute(0, 2, 3, 1).contiguous(), wp, 1, 3, relu='tanh')
    assert _rel(out[..., :3].permute(0, 3, 1, 2).float().cpu(), ref) <= 4e-3
    # extent override + strided destination: a 2x2-tap kernel evaluated at H x W outputs (one-sided zero padding),
    # written into the odd / odd positions of a twice-as-large tensor
    w2 = torch.randn(64, 64, 2, 2, generator=g) / 16.0
    xr = x.bfloat16().float()
    ref2 = F.conv2d(F.pad(xr, (0, 1, 0, 1)), w2.bfloat16().float())
    big = torch.zeros(2, 20, 28, 64, device=dev, dtype=torch.bfloat16)
    sdn_b200.conv2d_tc(x.to(dev).bfloat16().permute(0, 2, 3, 1).contiguous(), sdn_b200.pack_conv_weight(w2.to(dev)), 1, 0,
                       out_hw=(10, 14), out=big[:, 1::2, 1::2])
    assert _rel(big[:, 1::2, 1::2].permute(0, 3, 1, 2).float().cpu(), ref2) <= 4e-3
    assert float(big[:, 0::2].abs().sum()) == 0 and float(big[:, :, 0::2].abs().sum()) == 0


def test_epilogue_statistics_feed_instance_norm(dev):
    """conv2d_tc(stats=...) leaves per (image, channel) sum / sum of squares of its fp32 result; the normalising pass alone
    then equals InstanceNorm2d of the convolution."""
    import sdn_b200
    g = torch.Generator().manual_seed(21)
    x = torch.randn(2, 64, 24, 40, generator=g)
    w = torch.randn(128, 64, 3, 3, generator=g) / 24.0
    b = torch.randn(128, generator=g)
    y = F.conv2d(x.bfloat16().float(), w.bfloat16().float(), b, padding=1)
    stats = torch.full((2, 128, 2), 7.0, device=dev)          # garbage: the call clears it
    out = sdn_b200.conv2d_tc(x.to(dev).bfloat16().permute(0, 2, 3, 1).contiguous(), sdn_b200.pack_conv_weight(w.to(dev)), 1, 1,
                             None, b.to(dev), stats=stats)
    want = torch.stack([y.sum(dim=(2, 3)), (y * y).sum(dim=(2, 3))], dim=-1)
    assert _rel(stats.cpu(), want) <= 1e-3
    normed = sdn_b200.instance_norm_nhwc(out, None, True, stats=stats)
    ref = F.instance_norm(y).clamp_min(0)
    assert _rel(normed.permute(0, 3, 1, 2).float().cpu(), ref) <= 6e-3
    # two launches adding into one buffer (the phases of a transposed convolution do this)
    sdn_b200.conv2d_tc(x.to(dev).bfloat16().permute(0, 2, 3, 1).contiguous(), sdn_b200.pack_conv_weight(w.to(dev)), 1, 1,
                       None, b.to(dev), stats=stats, stats_keep=True)
    assert _rel(stats.cpu(), 2 * want) <= 1e-3


def test_reflect_pad_kernel(dev):
    import sdn_b200
    x = torch.randn(2, 9, 11, 48, generator=torch.Generator().manual_seed(3)).to(dev).bfloat16()
    got = sdn_b200.reflect_pad_nhwc(x, 3, 64)
    want = F.pad(x.permute(0, 3, 1, 2).float(), (3, 3, 3, 3), mode='reflect').permute(0, 2, 3, 1)
    assert got.shape == (2, 15, 17, 64)
    assert torch.equal(got[..., :48].float(), want) and float(got[..., 48:].abs().sum()) == 0


def test_argument_errors(dev):
    import sdn_b200
    x = torch.zeros(1, 8, 8, 64, device=dev, dtype=torch.bfloat16)
    w = torch.zeros(64, 3, 3, 64, device=dev, dtype=torch.bfloat16)
    with pytest.raises(sdn_b200.SdnError):
        sdn_b200.conv2d_tc(x.float(), w)                           # not bf16
    with pytest.raises(sdn_b200.SdnError):
        sdn_b200.conv2d_tc(x[..., :32].contiguous(), w[..., :32].contiguous())     # channels not a multiple of 64
    with pytest.raises(sdn_b200.SdnError):
        sdn_b200.conv2d_tc(x.cpu(), w.cpu())                       # no CPU fallback
    assert sdn_b200.conv2d_tc(x, w, 1, 1).abs().sum() == 0


@pytest.mark.parametrize('N,H,W,Cout,K,act,stats', [(1, 40, 300, 64, 7, None, True), (2, 39, 200, 3, 7, 'tanh', False),
                                                    (1, 24, 134, 64, 5, 'relu', True), (1, 70, 518, 64, 7, None, False)])
def test_large_filters_row_ring_form(dev, N, H, W, Cout, K, act, stats, monkeypatch):
    """k_conv_rows (opt-in with SDN_CONV_ROWS=1; 7x7 / 5x5, stride 1, pre-padded 64-channel input): every input row of a 128-column strip is loaded once
    into a ring of shared-memory row slots and all K x K taps address it through row-shifted UMMA descriptors.  Strips
    with a ragged last column block, odd output heights (a last group of one row), several images, bias + tanh / ReLU,
    InstanceNorm statistics from the epilogue; against torch fp32 on the same bf16-rounded operands."""
    import sdn_b200
    monkeypatch.setenv('SDN_CONV_ROWS', '1')
    g = torch.Generator().manual_seed(H + W + Cout)
    x = torch.randn(N, 64, H, W, generator=g)
    w = torch.randn(Cout, 64, K, K, generator=g) / (64 * K * K) ** 0.5
    bias = torch.randn(Cout, generator=g) * 0.1
    ref = F.conv2d(x.bfloat16().float(), w.bfloat16().float(), bias)
    if act == 'tanh':
        refa = torch.tanh(ref)
    elif act == 'relu':
        refa = ref.clamp_min(0)
    else:
        refa = ref
    xn = x.to(dev).bfloat16().permute(0, 2, 3, 1).contiguous()
    wp = sdn_b200.pack_conv_weight(w.to(dev))
    b = F.pad(bias, (0, wp.shape[0] - Cout)).to(dev)
    st = torch.full((N, wp.shape[0], 2), float('nan'), device=dev) if stats else None
    sdn_b200.timing_enable(True); sdn_b200.timing_read()
    out = sdn_b200.conv2d_tc(xn, wp, 1, 0, None, b, None, {'tanh': 'tanh', 'relu': True, None: False}[act], stats=st)
    torch.cuda.synchronize()
    names = set(sdn_b200.timing_read())
    sdn_b200.timing_enable(False)
    assert 'k_conv_rows' in names, names
    got = out[..., :Cout].permute(0, 3, 1, 2).float().cpu()
    assert got.shape == refa.shape
    assert _rel(got, refa) <= 5e-3, _rel(got, refa)
    if stats:
        s = st[:, :Cout].cpu()
        assert _rel(s[..., 0], ref.sum(dim=(2, 3))) <= 2e-3 and _rel(s[..., 1], (ref * ref).sum(dim=(2, 3))) <= 2e-3

"""GlobalGenerator of the textural branch (reference: textural/models/networks.py:211-283, built by define_G :45-63 with
input_nc = 48, output_nc = 3, ngf = 64, n_downsampling = 4, n_blocks = 9, instance norm: base_options.py:40,57-59,74,80).

The module keeps the reference's `self.model` nn.Sequential (same layer order, so `model.<i>.weight` checkpoint keys load
unchanged).  On a CUDA device `forward` does not run that Sequential: it walks it and runs every convolution -- the 7x7
stems, the four stride-2 convolutions, the eighteen 3x3 convolutions of the 1024-channel ResnetBlocks and the four
transposed convolutions -- through the hand-written tcgen05 implicit-GEMM kernel (sdn_b200.conv2d_tc: bf16 operands,
fp32 accumulation in tensor memory, NHWC), instance normalisation through sdn_instance_norm_nhwc, reflection padding as
an NHWC gather.  Transposed convolutions (k 3, s 2, p 1, output_padding 1) are decomposed into their four output-parity
phases (1, 2, 2 and 4 taps): no zero-insertion, exactly the layer's own FLOPs.  Inference only (no autograd).
On the CPU the torch Sequential itself runs (it IS the reference definition)."""
import functools

import torch
import torch.nn as nn
import torch.nn.functional as F


def get_norm_layer(norm_type='instance'):
    """networks.py:23-30."""
    if norm_type == 'batch':
        return functools.partial(nn.BatchNorm2d, affine=True, track_running_stats=True)
    if norm_type == 'instance':
        return functools.partial(nn.InstanceNorm2d, affine=False, track_running_stats=True)
    raise NotImplementedError('normalization layer [%s] is not found' % norm_type)


class ResnetBlock(nn.Module):
    """networks.py:244-283 (reflect padding, no dropout)."""

    def __init__(self, dim, padding_type='reflect', norm_layer=nn.BatchNorm2d, activation=nn.ReLU(True), use_dropout=False):
        super(ResnetBlock, self).__init__()
        assert padding_type == 'reflect' and not use_dropout
        self.conv_block = nn.Sequential(nn.ReflectionPad2d(1), nn.Conv2d(dim, dim, kernel_size=3, padding=0),
                                        norm_layer(dim), activation,
                                        nn.ReflectionPad2d(1), nn.Conv2d(dim, dim, kernel_size=3, padding=0),
                                        norm_layer(dim))

    def forward(self, x):
        return x + self.conv_block(x)


def _reflect_index(n, p, device):
    i = torch.arange(-p, n + p, device=device)
    i = i.abs()
    return torch.where(i >= n, 2 * (n - 1) - i, i)


def reflect_pad_nhwc(x, p, cpad=None):
    """nn.ReflectionPad2d(p) on an NHWC tensor; optionally zero-pads the channels to `cpad`.  One native pass on CUDA
    bf16 tensors (sdn_reflect_pad_nhwc), two torch gathers otherwise."""
    if x.is_cuda and x.dtype == torch.bfloat16 and (x.shape[-1] if cpad is None else cpad) % 8 == 0:
        import sdn_b200
        return sdn_b200.reflect_pad_nhwc(x, p, cpad)
    N, H, W, C = x.shape
    x = x.index_select(1, _reflect_index(H, p, x.device)).index_select(2, _reflect_index(W, p, x.device))
    if cpad is not None and cpad > C:
        x = F.pad(x, (0, cpad - C))
    return x


class GlobalGenerator(nn.Module):
    def __init__(self, input_nc, output_nc, ngf=64, n_downsampling=3, n_blocks=9, norm_layer=nn.BatchNorm2d,
                 padding_type='reflect'):
        assert n_blocks >= 0
        super(GlobalGenerator, self).__init__()
        activation = nn.ReLU(True)
        model = [nn.ReflectionPad2d(3), nn.Conv2d(input_nc, ngf, kernel_size=7, padding=0), norm_layer(ngf), activation]
        for i in range(n_downsampling):
            mult = 2 ** i
            model += [nn.Conv2d(ngf * mult, ngf * mult * 2, kernel_size=3, stride=2, padding=1),
                      norm_layer(ngf * mult * 2), activation]
        mult = 2 ** n_downsampling
        for i in range(n_blocks):
            model += [ResnetBlock(ngf * mult, padding_type=padding_type, activation=activation, norm_layer=norm_layer)]
        for i in range(n_downsampling):
            mult = 2 ** (n_downsampling - i)
            model += [nn.ConvTranspose2d(ngf * mult, int(ngf * mult / 2), kernel_size=3, stride=2, padding=1,
                                         output_padding=1), norm_layer(int(ngf * mult / 2)), activation]
        model += [nn.ReflectionPad2d(3), nn.Conv2d(ngf, output_nc, kernel_size=7, padding=0), nn.Tanh()]
        self.model = nn.Sequential(*model)
        self._packed = {}

    # ---- weights in the kernel's layout (packed once per device; call `repack()` after loading new weights) ----
    def repack(self):
        self._packed = {}

    def _conv_w(self, conv):
        import sdn_b200
        key = id(conv)
        if key not in self._packed:
            w = sdn_b200.pack_conv_weight(conv.weight)
            self._packed[key] = (w, None if conv.bias is None else F.pad(conv.bias.detach().float(),
                                                                         (0, w.shape[0] - conv.out_channels)))
        return self._packed[key]

    def _deconv_w(self, m):
        """ConvTranspose2d(k3, s2, p1, op1) weight [Cin, Cout, 3, 3] -> four phase kernels.  Output (2a + ph, 2b + pw):
        phase 0 along an axis uses tap k = 1 at input offset 0; phase 1 uses k = 2 at offset 0 and k = 0 at offset +1."""
        import sdn_b200
        key = id(m)
        if key not in self._packed:
            w = m.weight.detach()                                   # [Cin, Cout, kh, kw]
            taps = {0: [1], 1: [2, 0]}
            phases = {}
            for ph in (0, 1):
                for pw in (0, 1):
                    wk = w[:, :, taps[ph], :][:, :, :, taps[pw]]    # [Cin, Cout, KH', KW']
                    phases[(ph, pw)] = sdn_b200.pack_conv_weight(wk.permute(1, 0, 2, 3).contiguous())
            bias = None if m.bias is None else F.pad(m.bias.detach().float(), (0, (-m.out_channels) % 64))
            self._packed[key] = (phases, bias)
        return self._packed[key]

    @staticmethod
    def _wants_stats(m):
        """The following norm layer takes per-instance statistics: let the convolution's epilogue produce them."""
        return isinstance(m, nn.InstanceNorm2d) and not ((not m.training) and m.track_running_stats and
                                                         m.running_mean is not None)

    def _norm(self, x, m, relu, residual=None, stats=None):
        """InstanceNorm2d(affine=False, track_running_stats=True): instance statistics in train mode, the running
        statistics in eval mode (torch semantics); BatchNorm2d in eval mode likewise folds to scale / shift."""
        import sdn_b200
        C = m.num_features
        use_running = (not m.training) and getattr(m, 'track_running_stats', False) and m.running_mean is not None
        if isinstance(m, nn.InstanceNorm2d) and not use_running:
            return sdn_b200.instance_norm_nhwc(x, residual, relu, m.eps, stats=stats)
        if not use_running:
            raise NotImplementedError('GlobalGenerator tcgen05 path: BatchNorm2d in training mode')
        rstd = torch.rsqrt(m.running_var.float() + m.eps)
        scale = rstd * (m.weight.float() if getattr(m, 'affine', False) and m.weight is not None else 1.0)
        shift = -m.running_mean.float() * scale + (m.bias.float() if getattr(m, 'affine', False) and m.bias is not None else 0.0)
        y = x[..., :C].float() * scale + shift
        if residual is not None:
            y = y + residual[..., :C].float()
        y = y.clamp_min(0) if relu else y
        return F.pad(y.to(torch.bfloat16), (0, x.shape[-1] - C))

    def _conv(self, x, conv, pad, act=None, stats=None):
        import sdn_b200
        w, b = self._conv_w(conv)
        return sdn_b200.conv2d_tc(x, w, conv.stride[0], pad, None, b, None, act, stats=stats)

    @staticmethod
    def _new_stats(x, channels):
        return torch.empty(x.shape[0], channels, 2, device=x.device, dtype=torch.float32)

    def _deconv(self, x, m, stats=None):
        """Four parity phases, each written straight into its interleaved positions of the [N, 2H, 2W, Co] output (strided
        store in the kernel's epilogue); phase taps that reach one pixel beyond the input read TMA's zeros."""
        import sdn_b200
        phases, b = self._deconv_w(m)
        N, H, W, C = x.shape
        Co = phases[(0, 0)].shape[0]
        out = torch.empty(N, 2 * H, 2 * W, Co, device=x.device, dtype=torch.bfloat16)
        for k, ((ph, pw), w) in enumerate(phases.items()):
            sdn_b200.conv2d_tc(x, w, 1, 0, None, b, None, None, out_hw=(H, W), out=out[:, ph::2, pw::2], stats=stats,
                               stats_keep=k > 0)
        return out

    def forward_tc(self, input):
        """input [N, input_nc, H, W] float (CUDA) -> [N, output_nc, H, W] float32."""
        layers = list(self.model)
        x = input.detach().permute(0, 2, 3, 1).to(torch.bfloat16)
        i, pending_pad, pending_stats, prepadded = 0, 0, None, False
        while i < len(layers):
            m = layers[i]
            if isinstance(m, nn.ReflectionPad2d):
                pending_pad = int(m.padding[0])
                i += 1
            elif isinstance(m, nn.Conv2d):
                if prepadded:                   # the producer already wrote the reflection-padded tensor
                    pad, prepadded = 0, False
                    if x.shape[-1] % 64:
                        x = F.pad(x, (0, (-x.shape[-1]) % 64))
                elif pending_pad:
                    x = reflect_pad_nhwc(x, pending_pad, (x.shape[-1] + 63) // 64 * 64)
                    pad, pending_pad = 0, 0
                else:
                    pad = int(m.padding[0])
                    if x.shape[-1] % 64:
                        x = F.pad(x, (0, (-x.shape[-1]) % 64))
                nxt = layers[i + 1] if i + 1 < len(layers) else None
                pending_stats = None
                if isinstance(nxt, nn.Tanh):
                    x = self._conv(x, m, pad, 'tanh')
                    i += 2
                else:
                    if self._wants_stats(nxt):
                        pending_stats = self._new_stats(x, self._conv_w(m)[0].shape[0])
                    x = self._conv(x, m, pad, stats=pending_stats)
                    i += 1
            elif isinstance(m, nn.ConvTranspose2d):
                nxt = layers[i + 1] if i + 1 < len(layers) else None
                pending_stats = self._new_stats(x, self._deconv_w(m)[0][(0, 0)].shape[0]) if self._wants_stats(nxt) else None
                x = self._deconv(x, m, pending_stats)
                i += 1
            elif isinstance(m, (nn.InstanceNorm2d, nn.BatchNorm2d)):
                relu = i + 1 < len(layers) and isinstance(layers[i + 1], nn.ReLU)
                after = layers[i + (2 if relu else 1)] if i + (2 if relu else 1) < len(layers) else None
                if pending_stats is not None and isinstance(after, nn.ReflectionPad2d):
                    # norm (+ReLU) and the reflection padding of the next convolution in one pass
                    import sdn_b200
                    x = sdn_b200.instance_norm_pad_nhwc(x, pending_stats, int(after.padding[0]), None, relu, m.eps)
                    pending_stats, prepadded = None, True
                    i += 3 if relu else 2
                    continue
                x = self._norm(x, m, relu, stats=pending_stats)
                pending_stats = None
                i += 2 if relu else 1
            elif isinstance(m, ResnetBlock):
                cb = list(m.conv_block)     # pad, conv, norm, relu, pad, conv, norm
                st1 = self._new_stats(x, x.shape[-1]) if self._wants_stats(cb[2]) else None
                y = self._conv(reflect_pad_nhwc(x, 1), cb[1], 0, stats=st1)
                if st1 is not None:             # norm + ReLU + the second convolution's reflection padding in one pass
                    import sdn_b200
                    y = sdn_b200.instance_norm_pad_nhwc(y, st1, 1, None, True, cb[2].eps)
                else:
                    y = reflect_pad_nhwc(self._norm(y, cb[2], True), 1)
                st2 = self._new_stats(x, x.shape[-1]) if self._wants_stats(cb[6]) else None
                y = self._conv(y, cb[5], 0, stats=st2)
                x = self._norm(y, cb[6], False, residual=x, stats=st2)
                i += 1
            else:
                raise NotImplementedError('GlobalGenerator tcgen05 path: layer %r' % (m,))
        C = layers[-2].out_channels if isinstance(layers[-1], nn.Tanh) else x.shape[-1]
        return x[..., :C].permute(0, 3, 1, 2).float().contiguous()      # NCHW fp32 like the reference module

    def forward(self, input):
        if input.is_cuda and not torch.is_grad_enabled():
            return self.forward_tc(input)
        if input.is_cuda and not any(p.requires_grad for p in self.parameters()):
            return self.forward_tc(input)
        return self.model(input)

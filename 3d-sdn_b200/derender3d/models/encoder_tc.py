"""ResNet-18 trunk of the image encoder on the hand-written tcgen05 kernels (csrc/sdn_conv_tc.cu): an inference form
(`ResNet18TC`, BatchNorm folded) and a training form (`resnet18_trunk_train`, batch statistics, forward AND backward).

Reference: geometric/derender3d/models/derenderer.py:25-27,35-43 (torchvision resnet18, `net.avgpool` replaced by
AdaptiveAvgPool2d(1), `net.fc` by Linear(512, 256)).  In eval mode BatchNorm is a per-channel affine map, folded here into
the convolution's epilogue (scale / shift), together with the residual add and the ReLU of every BasicBlock:

    stem   : im2col(7x7, s2, p3: 147 -> 192 columns) -> 1x1 conv [64 x 192] * bn1 -> ReLU -> max-pool 3x3 s2 (torch)
    block  : y = ReLU(bn1(conv3x3_s(x)));  x = ReLU(bn2(conv3x3(y)) + (downsample: bn(conv1x1_s(x)) | x))
    head   : global average pool (torch) -> [N, 512] fp32

Activations are NHWC bf16, accumulation fp32.

Training form (second half of this file): one torch.autograd.Function for the whole trunk.  Forward per unit: convolution with
the per-(image, channel) sums from its epilogue -> sdn_batch_norm_finalize (batch mean / rstd, running statistics) ->
sdn_affine_act_nhwc (normalise, + identity, ReLU); the stem's max-pool is sdn_max_pool_nhwc.  Backward per unit:
sdn_batch_norm_backward_nhwc -> sdn_conv2d_wgrad (weight gradient, accumulated in place when the parameter already holds a
gradient) -> data gradient on the forward kernel with transposed operands (sdn_pack_conv_weight_transposed).  The global
average pool and its backward are the only torch kernels on activations."""
import os

import torch
import torch.nn.functional as F

import sdn_b200


def _fold(bn):
    scale = bn.weight.detach().float() * torch.rsqrt(bn.running_var.detach().float() + bn.eps)
    return scale.contiguous(), (bn.bias.detach().float() - bn.running_mean.detach().float() * scale).contiguous()


class ResNet18TC(object):
    def __init__(self, net):
        """net: torchvision.models.ResNet (resnet18 layout: BasicBlocks) already on its CUDA device."""
        dev = net.conv1.weight.device
        if dev.type != 'cuda':
            raise sdn_b200.SdnError('ResNet18TC needs the module on a CUDA device (no CPU fallback)')
        self.device = dev
        w = net.conv1.weight.detach()                       # [64, 3, 7, 7] -> 1x1 kernel over (kh, kw, ci) columns
        self.stem_k = (int(w.shape[2]), int(net.conv1.stride[0]), int(net.conv1.padding[0]))
        cols = w.shape[1] * w.shape[2] * w.shape[3]
        self.stem_kpad = (cols + 63) // 64 * 64
        wk = torch.zeros(w.shape[0], 1, 1, self.stem_kpad, device=dev, dtype=torch.bfloat16)
        wk[:, 0, 0, :cols] = w.permute(0, 2, 3, 1).reshape(w.shape[0], cols).to(torch.bfloat16)
        self.stem_w = wk
        self.stem_bn = _fold(net.bn1)
        self.pool = (net.maxpool.kernel_size, net.maxpool.stride, net.maxpool.padding)
        self.blocks = []
        for layer in (net.layer1, net.layer2, net.layer3, net.layer4):
            for blk in layer:
                ds = None
                if blk.downsample is not None:
                    ds = (sdn_b200.pack_conv_weight(blk.downsample[0].weight), int(blk.downsample[0].stride[0]),
                          _fold(blk.downsample[1]))
                self.blocks.append((sdn_b200.pack_conv_weight(blk.conv1.weight), int(blk.conv1.stride[0]), _fold(blk.bn1),
                                    sdn_b200.pack_conv_weight(blk.conv2.weight), _fold(blk.bn2), ds))
        self.signature = ResNet18TC.weights_signature(net)

    @staticmethod
    def weights_signature(net):
        """Changes whenever a parameter or BatchNorm buffer of the trunk is written in place (optimizer step,
        load_state_dict): the folded / packed copy is then rebuilt."""
        return tuple(t._version for t in list(net.parameters()) + list(net.buffers()))

    @torch.no_grad()
    def __call__(self, images):
        """images [N, 3, H, W] float -> features [N, 512] float32."""
        x = images.detach().permute(0, 2, 3, 1).to(torch.bfloat16)
        k, s, p = self.stem_k
        x = sdn_b200.im2col_nhwc(x, k, k, s, p, self.stem_kpad)
        x = sdn_b200.conv2d_tc(x, self.stem_w, 1, 0, self.stem_bn[0], self.stem_bn[1], None, True)
        x = F.max_pool2d(x.permute(0, 3, 1, 2), *self.pool).permute(0, 2, 3, 1)      # stays channels-last in memory
        for w1, s1, bn1, w2, bn2, ds in self.blocks:
            identity = x
            if ds is not None:
                identity = sdn_b200.conv2d_tc(x, ds[0], ds[1], 0, ds[2][0], ds[2][1], None, False)
            y = sdn_b200.conv2d_tc(x, w1, s1, 1, bn1[0], bn1[1], None, True)
            x = sdn_b200.conv2d_tc(y, w2, 1, 1, bn2[0], bn2[1], identity, True)
        return x.float().mean(dim=(1, 2))


# ----------------------------------------------------------------------------------------------------------------------
# Training form: forward AND backward of the trunk on the hand-written kernels (BatchNorm in batch-statistics mode).
# ----------------------------------------------------------------------------------------------------------------------
def _bn_train(z, stats, bn, relu, residual=None):
    """BatchNorm2d (training) of a convolution result z [N,H,W,C] bf16 whose per-(image, channel) sums the convolution's
    epilogue left in `stats` [N,C,2]: batch mean / biased variance and the running-statistics update in one launch
    (sdn_batch_norm_finalize), then y = relu?(z * scale + shift (+ residual)).  -> (y, mean [C], rstd [C])."""
    mean, rstd, scale, shift = sdn_b200.batch_norm_finalize(stats, z.numel() // z.shape[-1], bn)
    return sdn_b200.affine_act_nhwc(z, scale, shift, residual, relu), mean, rstd


# Parameter gradients of the trunk are added straight into existing `.grad` tensors by the wgrad / BatchNorm-backward kernels
# (what Megatron calls gradient-accumulation fusion); SDN_FUSED_GRAD_ACC=0 returns them through autograd instead.
FUSED_GRAD_ACCUMULATION = os.environ.get('SDN_FUSED_GRAD_ACC', '1') != '0'


def _can_accumulate(p):
    g = p.grad
    return g is not None and g.dtype == torch.float32 and g.is_contiguous() and g.device == p.device


class _TrunkTrainFunction(torch.autograd.Function):
    """features = ResNet18(images) in training mode; every convolution (forward, data gradient, weight gradient) on the
    tcgen05 kernels of csrc/sdn_conv_tc.cu, BatchNorm forward / backward on sdn_batch_norm_finalize / sdn_affine_act_nhwc /
    sdn_batch_norm_backward_nhwc, the stem's max-pool on sdn_max_pool_nhwc / sdn_max_pool_backward_nhwc, the average pool
    through torch.  Activations and activation gradients are bf16 NHWC, parameter gradients fp32.  The intermediate
    activations live on `ctx` (no double backward)."""

    @staticmethod
    def forward(ctx, images, net, *params):
        x = images.detach().permute(0, 2, 3, 1).to(torch.bfloat16)
        N = x.shape[0]
        dev = x.device
        saved = {}
        convs = [m for m in net.modules() if isinstance(m, torch.nn.Conv2d)]
        chans = sum(c.out_channels for c in convs)
        stats_all = torch.zeros(N * chans * 2, device=dev, dtype=torch.float32)     # one clearing pass for all layers
        cursor = [0]

        def conv_bn(xin, conv, bn, relu, residual=None, stem=False):
            if stem:
                k, s, p = int(conv.kernel_size[0]), int(conv.stride[0]), int(conv.padding[0])
                cols_n = conv.in_channels * k * k
                kpad = (cols_n + 63) // 64 * 64
                xin = sdn_b200.im2col_nhwc(xin, k, k, s, p, kpad)
                w = torch.zeros(conv.out_channels, 1, 1, kpad, device=dev, dtype=torch.bfloat16)
                w[:, 0, 0, :cols_n] = conv.weight.detach().permute(0, 2, 3, 1).reshape(conv.out_channels, cols_n).to(torch.bfloat16)
                stride, pad = 1, 0
            else:
                w = sdn_b200.pack_conv_weight(conv.weight)
                stride, pad = int(conv.stride[0]), int(conv.padding[0])
            n_st = N * conv.out_channels * 2
            stats = stats_all[cursor[0]:cursor[0] + n_st].view(N, conv.out_channels, 2)
            cursor[0] += n_st
            z = sdn_b200.conv2d_tc(xin, w, stride, pad, stats=stats, stats_keep=True)
            y, mean, rstd = _bn_train(z, stats, bn, relu, residual)
            return xin, z, y, mean, rstd

        units = []          # (name, conv, bn, relu, input, z, y, mean, rstd)
        xin, z, y, mean, rstd = conv_bn(x, net.conv1, net.bn1, True, stem=True)
        units.append(('stem', net.conv1, net.bn1, True, xin, z, y, mean, rstd))
        mp = net.maxpool
        pool = (int(_pair(mp.kernel_size)[0]), int(_pair(mp.stride)[0]), int(_pair(mp.padding)[0]))
        x, pool_arg = sdn_b200.max_pool_nhwc(y, *pool, return_argmax=True)
        saved['pool'] = (pool, pool_arg, (y.shape[1], y.shape[2]))
        blocks = []
        for layer in (net.layer1, net.layer2, net.layer3, net.layer4):
            for blk in layer:
                xb = x
                ds = None
                identity = xb
                if blk.downsample is not None:
                    dxin, dz, dy, dmean, drstd = conv_bn(xb, blk.downsample[0], blk.downsample[1], False)
                    ds = (blk.downsample[0], blk.downsample[1], dz, dmean, drstd)
                    identity = dy
                _, z1, y1, m1, r1 = conv_bn(xb, blk.conv1, blk.bn1, True)
                _, z2, out, m2, r2 = conv_bn(y1, blk.conv2, blk.bn2, True, residual=identity)
                blocks.append((blk, xb, z1, y1, m1, r1, z2, out, m2, r2, ds))
                x = out
        feat = x.float().mean(dim=(1, 2))
        ctx.net, ctx.units, ctx.blocks, ctx.saved, ctx.last_shape = net, units, blocks, saved, x.shape
        ctx.chans = chans
        ctx.param_ids = [id(p) for p in params]
        return feat

    @staticmethod
    def backward(ctx, g_feat):
        net = ctx.net
        grads = {}
        N, H, W, C = ctx.last_shape
        sums_all = torch.zeros(ctx.chans * 2, device=g_feat.device, dtype=torch.float32)     # (grad_beta, grad_gamma) of all layers
        cursor = [0]
        fused_acc = FUSED_GRAD_ACCUMULATION
        g = (g_feat.detach().float() / (H * W)).to(torch.bfloat16)[:, None, None, :].expand(N, H, W, C).contiguous()

        def bn_conv_backward(gy, y, z, mean, rstd, conv, bn, relu, xin, want_skip=False, need_dx=True, skip_add=None, stem=False):
            nc = z.shape[-1]
            sums = sums_all[cursor[0]:cursor[0] + 2 * nc].view(nc, 2)
            cursor[0] += 2 * nc
            # parameters that already hold a gradient get this one ADDED in place by the kernels (no temporary, no add
            # kernel); the Function then returns None for them
            acc_bn = None
            if fused_acc and _can_accumulate(bn.weight) and _can_accumulate(bn.bias):
                acc_bn = (bn.weight.grad, bn.bias.grad)
            gz, g_gamma, g_beta, g_skip = sdn_b200.batch_norm_backward_nhwc(gy, y if relu else None, z, mean, rstd, bn.weight, relu,
                                                                              want_skip, sums=sums, accumulate_into=acc_bn)
            if acc_bn is None:
                grads[id(bn.weight)], grads[id(bn.bias)] = g_gamma, g_beta
            if stem:
                k = int(conv.kernel_size[0])
                gw = sdn_b200.conv2d_wgrad_tc(gz, xin, 1, 1, 1, 0)              # [64, 1, 1, Kpad] over the im2col columns
                cols_n = conv.in_channels * k * k
                grads[id(conv.weight)] = gw[:, 0, 0, :cols_n].reshape(conv.out_channels, k, k, conv.in_channels).permute(0, 3, 1, 2)
                return None, g_skip
            k, s, p = int(conv.kernel_size[0]), int(conv.stride[0]), int(conv.padding[0])
            wg = conv.weight.grad
            if fused_acc and wg is not None and wg.dtype == torch.float32 and wg.permute(0, 2, 3, 1).is_contiguous():
                sdn_b200.conv2d_wgrad_tc(gz, xin, k, k, s, p, accumulate_into=wg.permute(0, 2, 3, 1))
            else:
                grads[id(conv.weight)] = sdn_b200.conv2d_wgrad_tc(gz, xin, k, k, s, p).permute(0, 3, 1, 2)
            dx = None
            if need_dx:
                dx = sdn_b200.conv2d_dgrad_tc(gz, conv.weight, (xin.shape[1], xin.shape[2]), s, p, residual=skip_add)
            return dx, g_skip

        for blk, xb, z1, y1, m1, r1, z2, out, m2, r2, ds in reversed(ctx.blocks):
            # out = relu(bn2(conv2(y1)) + identity)
            g_y1, g_id = bn_conv_backward(g, out, z2, m2, r2, blk.conv2, blk.bn2, True, y1, want_skip=True)
            if ds is None:
                g, _ = bn_conv_backward(g_y1, y1, z1, m1, r1, blk.conv1, blk.bn1, True, xb, skip_add=g_id if int(blk.conv1.stride[0]) == 1 else None)
                if int(blk.conv1.stride[0]) != 1:
                    g = g + g_id
            else:
                dconv, dbn, dz, dmean, drstd = ds
                g_main, _ = bn_conv_backward(g_y1, y1, z1, m1, r1, blk.conv1, blk.bn1, True, xb)
                g_ds, _ = bn_conv_backward(g_id, None, dz, dmean, drstd, dconv, dbn, False, xb)
                g = g_main + g_ds
        # max-pool backward (torch), then the stem
        pool, pool_arg, pool_hw = ctx.saved['pool']
        g_pool = sdn_b200.max_pool_backward_nhwc(pool_arg, g, pool_hw, *pool)
        _, conv, bn, relu, xin, z, y, mean, rstd = ctx.units[0]
        bn_conv_backward(g_pool, y, z, mean, rstd, conv, bn, True, xin, stem=True)
        out = [None, None]
        for pid in ctx.param_ids:
            gp = grads.get(pid)
            out.append(None if gp is None else gp.float())
        return tuple(out)


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def resnet18_trunk_train(net, images):
    """Training-mode trunk: features [N, 512] fp32 with autograd through the hand-written kernels.  Gradients flow to the
    convolution weights and BatchNorm affine parameters of `net` (not to `images`: the encoder's input is data)."""
    params = [p for n, p in net.named_parameters() if not n.startswith('fc.')]
    return _TrunkTrainFunction.apply(images, net, *params)

"""ResNet-18 trunk of the image encoder on the hand-written tcgen05 convolution (csrc/sdn_conv_tc.cu) -- inference form.

Reference: geometric/derender3d/models/derenderer.py:25-27,35-43 (torchvision resnet18, `net.avgpool` replaced by
AdaptiveAvgPool2d(1), `net.fc` by Linear(512, 256)).  In eval mode BatchNorm is a per-channel affine map, folded here into
the convolution's epilogue (scale / shift), together with the residual add and the ReLU of every BasicBlock:

    stem   : im2col(7x7, s2, p3: 147 -> 192 columns) -> 1x1 conv [64 x 192] * bn1 -> ReLU -> max-pool 3x3 s2 (torch)
    block  : y = ReLU(bn1(conv3x3_s(x)));  x = ReLU(bn2(conv3x3(y)) + (downsample: bn(conv1x1_s(x)) | x))
    head   : global average pool (torch) -> [N, 512] fp32

Activations are NHWC bf16, accumulation fp32.  Training (batch statistics + autograd) stays on the torch / cuDNN path."""
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

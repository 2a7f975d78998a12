"""Image encoder of the geometric branch (reference: derender3d/models/derenderer.py:7-65): ResNet-18 trunk + three
fully connected layers + head split.  Attribute names (net, relu, fc1, fc2, _fc3) match the reference so that its
checkpoints load unchanged.  The convolutions are library GEMMs (cuDNN through torchvision); the four fully
connected layers run on the hand-written tcgen05 kernel (sdn_b200.linear_tc: bf16 operands, fp32 accumulate in TMEM)
when the module is on a GPU and `tensor_core_head` is set.  That is OPT-IN (SDN_TC_HEAD=1, or under torch.autocast):
the reference head is fp32 and a loaded reference checkpoint must give the same theta / class argmax, so the default
keeps fp32 nn.Linear.  DESIGN.md "Encoder"."""
import os

import torch
import torchvision
from torch.nn.modules import Module


class Derenderer(Module):
    in_size = 4
    hidden_size = 256

    def __init__(self, num_classes=8, grid_size=4):
        super(Derenderer, self).__init__()
        self.num_classes = num_classes
        self.grid_size = grid_size
        self.out_sizes = {
            '_theta_deltas': 2,
            '_translation2ds': 2,
            '_log_scales': 3,
            '_log_depths': 1,
            '_class_probs': num_classes,
            '_ffd_coeffs': num_classes * (grid_size ** 3) * 3,
        }
        # derenderer.py:25 asks for ImageNet weights; there is no network here, so they are loaded only if a local
        # file is named (RESNET18_WEIGHTS); a 3D-SDN checkpoint overwrites them anyway.
        self.net = torchvision.models.resnet18(weights=None)
        path = os.environ.get('RESNET18_WEIGHTS')
        if path and os.path.isfile(path):
            self.net.load_state_dict(torch.load(path, map_location='cpu'))
        self.net.avgpool = torch.nn.AdaptiveAvgPool2d(1)
        self.net.fc = torch.nn.Linear(512, Derenderer.hidden_size)
        self.relu = torch.nn.ReLU(inplace=True)
        self.fc1 = torch.nn.Linear(self.hidden_size + self.in_size, self.hidden_size)
        self.fc2 = torch.nn.Linear(self.hidden_size, self.hidden_size)
        self._fc3 = torch.nn.Linear(self.hidden_size, sum(self.out_sizes.values()))
        self.tensor_core_head = os.environ.get('SDN_TC_HEAD', '0') == '1'
        # the four FC layers as ONE fp32 launch (sdn_b200.mlp_head; opt-in: SDN_FUSED_HEAD=1 or `fused_head = True`)
        self.fused_head = os.environ.get('SDN_FUSED_HEAD', '0') == '1'
        # eval-mode trunk on the hand-written tcgen05 convolution (encoder_tc.ResNet18TC: bf16 operands, BatchNorm folded
        # into the epilogue).  Opt-in like the head (SDN_TC_TRUNK=1 or `tensor_core_trunk = True`): the reference is fp32.
        self.tensor_core_trunk = os.environ.get('SDN_TC_TRUNK', '0') == '1'
        self._trunk_tc = None

    def _trunk(self, images):
        """torchvision ResNet._forward_impl without its last Linear."""
        if self.tensor_core_trunk and images.is_cuda and not self.training:
            from derender3d.models.encoder_tc import ResNet18TC
            if self._trunk_tc is None or self._trunk_tc.signature != ResNet18TC.weights_signature(self.net) or \
                    self._trunk_tc.device != images.device:
                self._trunk_tc = ResNet18TC(self.net)
            return self._trunk_tc(images)
        if self.tensor_core_trunk and images.is_cuda and self.training:
            # training form: forward, data and weight gradients of every convolution on the tcgen05 kernels, BatchNorm
            # in batch-statistics mode on native kernels (encoder_tc.resnet18_trunk_train)
            from derender3d.models.encoder_tc import resnet18_trunk_train
            return resnet18_trunk_train(self.net, images)
        n = self.net
        x = n.maxpool(n.relu(n.bn1(n.conv1(images))))
        x = n.layer4(n.layer3(n.layer2(n.layer1(x))))
        return torch.flatten(n.avgpool(x), 1)

    def forward(self, images, mroi_norms, droi_norms):
        if images.is_cuda and self.fused_head:
            from sdn_b200 import mlp_head
            extra = torch.cat([mroi_norms.float(), droi_norms.float()], dim=1)
            x = mlp_head(self._trunk(images).float(), extra, self.net.fc, self.fc1, self.fc2, self._fc3)
        elif images.is_cuda and (self.tensor_core_head or torch.is_autocast_enabled() or
                               (self.tensor_core_trunk and not self.training)):
            from sdn_b200 import linear_tc
            x = linear_tc(self._trunk(images), self.net.fc.weight, self.net.fc.bias, relu=True)
            x = torch.cat([x, mroi_norms.float(), droi_norms.float()], dim=1)
            x = linear_tc(x, self.fc1.weight, self.fc1.bias, relu=True)
            x = linear_tc(x, self.fc2.weight, self.fc2.bias, relu=True)
            x = linear_tc(x, self._fc3.weight, self._fc3.bias, relu=False)
        else:
            if self.tensor_core_trunk and images.is_cuda:       # training: native trunk, fp32 head
                x = self.relu(self.net.fc(self._trunk(images)))
            else:
                x = self.relu(self.net(images))
            x = torch.cat([x.float(), mroi_norms, droi_norms], dim=1)
            x = self.relu(self.fc1(x))
            x = self.relu(self.fc2(x))
            x = self._fc3(x)
        (theta, t2d, log_scales, log_depths, class_logits, ffd) = torch.split(x, list(self.out_sizes.values()), dim=1)
        return {
            '_theta_deltas': theta / torch.norm(theta, p=2, dim=1, keepdim=True),
            '_translation2ds': t2d,
            '_log_scales': log_scales,
            '_log_depths': log_depths,
            '_class_probs': torch.nn.functional.softmax(class_logits, dim=1),
            '_ffd_coeffs': ffd.reshape(-1, self.num_classes, (self.grid_size ** 3) * 3),
        }

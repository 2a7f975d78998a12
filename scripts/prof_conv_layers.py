#!/usr/bin/env python
"""One launch of sdn_conv2d_forward per ResNet-18 layer shape at B = 16 (after a warm-up launch each) for `ncu -k k_conv_tc`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, '3d-sdn_b200')]
import torch
import sdn_b200
dev = torch.device('cuda:0')
SHAPES = [('stem 1x1 over im2col 192', 112, 192, 64, 1, 1, 0), ('layer1 3x3', 56, 64, 64, 3, 1, 1), ('layer2.0 3x3 s2', 56, 64, 128, 3, 2, 1),
          ('layer2.0 1x1 s2', 56, 64, 128, 1, 2, 0), ('layer2 3x3', 28, 128, 128, 3, 1, 1), ('layer3.0 3x3 s2', 28, 128, 256, 3, 2, 1),
          ('layer3.0 1x1 s2', 28, 128, 256, 1, 2, 0), ('layer3 3x3', 14, 256, 256, 3, 1, 1), ('layer4.0 3x3 s2', 14, 256, 512, 3, 2, 1),
          ('layer4.0 1x1 s2', 14, 256, 512, 1, 2, 0), ('layer4 3x3', 7, 512, 512, 3, 1, 1)]
for name, hw, ci, co, k, s, p in SHAPES:
    x = torch.randn(16, hw, hw, ci, device=dev).bfloat16()
    w = torch.randn(co, k, k, ci, device=dev).bfloat16()
    for _ in range(2):
        sdn_b200.conv2d_tc(x, w, s, p)
    torch.cuda.synchronize()
    print(name, flush=True)

#!/usr/bin/env python
"""BASELINE config 2b: silhouette only (the reference's training path), B = 32, 375x1242, fwd+bwd."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, '3d-sdn_b200')): sys.path.insert(0, p)
import torch, sdn_b200
from sdn_b200 import workloads as wl
from derender3d.models.renderer import Renderer, RenderType
dev = torch.device('cuda:0'); B, H, W = 32, 375, 1242
verts, faces = wl.vkitti_scene_batch(B, seed=0)
v = verts.to(dev).requires_grad_(True); f = faces.to(dev).int()
g = torch.randn(B, 1, H, W, generator=torch.Generator().manual_seed(1)).to(dev)
r = Renderer(image_size=(H, W), viewing_angle=wl.vkitti_viewing_angle())
sdn_b200.DEFAULT_FAST_MATH = True
def step():
    out = r(v, f, render_type=RenderType.Silhouette)
    torch.autograd.backward([out], [g]); gr = v.grad; v.grad = None; return gr
ref = step()
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
sdn_b200.timing_enable(True); sdn_b200.timing_read()
for _ in range(5): step()
torch.cuda.synchronize()
kern = {k: round(t[0] / 5, 4) for k, t in sorted(sdn_b200.timing_read().items(), key=lambda kv: -kv[1][0])[:6]}
print(json.dumps({'workload': 'config2b silhouette only B=32 375x1242', 'ms_per_step': round(ms, 4), 'images_per_s': round(B / ms * 1e3, 1), 'kernels_ms': kern, 'grad_norm': float(ref.norm())}))

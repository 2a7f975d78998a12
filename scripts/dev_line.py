"""Dev check of the line-sweep kernels: new (default) vs the IEEE item kernel (fast_math=0), per output set."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, '3d-sdn_b200')]
import torch
import sdn_b200
from sdn_b200 import workloads as wl
from derender3d.models.renderer import Renderer

dev = torch.device('cuda:0')
B = int(os.environ.get('B', 4))
verts, faces = wl.vkitti_scene_batch(B, seed=0)
angle = wl.vkitti_viewing_angle()
H, W = wl.VKITTI_H, wl.VKITTI_W
r = Renderer(image_size=(H, W), viewing_angle=angle)
f = faces.to(dev).int()
gen = torch.Generator().manual_seed(1)
g = {'silhouette': torch.randn(B, 1, H, W, generator=gen).to(dev), 'depth': torch.randn(B, 1, H, W, generator=gen).to(dev),
     'normal': torch.randn(B, 3, H, W, generator=gen).to(dev)}
for want in (dict(silhouette=True, depth=False, normal=False), dict(silhouette=False, depth=False, normal=True),
             dict(silhouette=True, depth=True, normal=True)):
    res = {}
    for fm in (0, 1):
        v = verts.to(dev).requires_grad_(True)
        out = r.forward_multi(v, f, fast_math=fm, **want)
        keys = [k for k in want if want[k]]
        torch.autograd.backward([out[k] for k in keys], [g[k] for k in keys])
        torch.cuda.synchronize()
        res[fm] = v.grad.clone()
    a, b = res[1].double(), res[0].double()
    print(want, 'nan fast', int(torch.isnan(res[1]).sum()), 'nan exact', int(torch.isnan(res[0]).sum()),
          'rel', float((a - b).norm() / b.norm()), flush=True)

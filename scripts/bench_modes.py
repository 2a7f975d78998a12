"""Per-kernel device time (ms per step) of config 2 restricted to one output set: s (config 2b), n, sdn."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, '3d-sdn_b200')]
import torch
import sdn_b200
from sdn_b200 import workloads as wl
from derender3d.models.renderer import Renderer

dev = torch.device('cuda:0')
B = 32
verts, faces = wl.vkitti_scene_batch(B, seed=0)
H, W = wl.VKITTI_H, wl.VKITTI_W
r = Renderer(image_size=(H, W), viewing_angle=wl.vkitti_viewing_angle())
f = faces.to(dev).int()
gen = torch.Generator().manual_seed(1)
g = {'silhouette': torch.randn(B, 1, H, W, generator=gen).to(dev), 'depth': torch.randn(B, 1, H, W, generator=gen).to(dev),
     'normal': torch.randn(B, 3, H, W, generator=gen).to(dev)}
v = verts.to(dev).requires_grad_(True)
out = {}
for name, want in (('s', dict(silhouette=True, depth=False, normal=False)), ('n', dict(silhouette=False, depth=False, normal=True)),
                   ('sdn', dict(silhouette=True, depth=True, normal=True))):
    keys = [k for k in want if want[k]]
    def step():
        o = r.forward_multi(v, f, **want)
        torch.autograd.backward([o[k] for k in keys], [g[k] for k in keys])
        v.grad = None
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        step()
    e1.record(); torch.cuda.synchronize()
    sdn_b200.timing_enable(True); sdn_b200.timing_read()
    for _ in range(10):
        step()
    torch.cuda.synchronize()
    kern = {k: round(t[0] / 10, 4) for k, t in sdn_b200.timing_read().items() if t[0] / 10 > 0.03}
    sdn_b200.timing_enable(False)
    out[name] = {'ms_per_step_eager': round(e0.elapsed_time(e1) / 20, 4), 'images_per_s': round(B / (e0.elapsed_time(e1) / 20) * 1e3, 1), 'kernels_ms': kern}
print(json.dumps(out))

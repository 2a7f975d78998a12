import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, '3d-sdn_b200')): sys.path.insert(0, p)
import torch, sdn_b200
from sdn_b200 import workloads as wl
from derender3d.models.renderer import Renderer, RenderType
B = int(sys.argv[1]); S = 384
dev = torch.device('cuda:0')
t0 = time.time()
mesh = wl.superellipsoid_mesh(192, 190, 0.6, 0.6)
verts, faces = wl.vkitti_scene_batch(B, seed=0, H=S, W=S, focal=725.0 * S / 256, z_range=(9.0, 12.0), mesh=mesh)
print('mesh', time.time() - t0, flush=True)
r = Renderer(image_size=S, viewing_angle=wl.vkitti_viewing_angle(S, 725.0 * S / 256))
v = verts.to(dev).requires_grad_(True); f = faces.to(dev).int()
gs = torch.randn(B, 1, S, S, device=dev)
sdn_b200.DEFAULT_FAST_MATH = True
sdn_b200.timing_enable(True)
for it in range(3):
    t0 = time.time()
    out = r(v, f, render_type=RenderType.Silhouette)
    torch.cuda.synchronize(); t1 = time.time()
    print('fwd', it, t1 - t0, 'ws bytes', sdn_b200.lib().sdn_workspace_bytes(B, 2 * len(faces), 2 * S, 2 * S), flush=True)
    torch.autograd.backward([out], [gs])
    torch.cuda.synchronize(); print('bwd', it, time.time() - t1, flush=True)
    print({k: round(t[0], 2) for k, t in sdn_b200.timing_read().items()}, flush=True)

import os, sys, time, faulthandler
faulthandler.dump_traceback_later(70, exit=True)
import threading, ctypes
def watchdog():
    time.sleep(35)
    buf = ctypes.create_string_buffer(1 << 14)
    n = sdn_b200.lib().sdn_debug_pending(buf, len(buf))
    print('PENDING:\n' + buf.raw[:n].decode()[:400], flush=True)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, '3d-sdn_b200')): sys.path.insert(0, p)
import torch, sdn_b200
from sdn_b200 import workloads as wl
from derender3d.models.renderer import Renderer, RenderType
B = int(sys.argv[1]); S = 384; sync_each = int(sys.argv[2])
dev = torch.device('cuda:0')
mesh = wl.superellipsoid_mesh(192, 190, 0.6, 0.6)
verts, faces = wl.vkitti_scene_batch(B, seed=0, H=S, W=S, focal=725.0 * S / 256, z_range=(9.0, 12.0), mesh=mesh)
verts[:, :, 0] -= verts[:, :, 0].mean(1, keepdim=True); verts[:, :, 1] -= verts[:, :, 1].mean(1, keepdim=True)
r = Renderer(image_size=S, viewing_angle=wl.vkitti_viewing_angle(S, 725.0 * S / 256))
v = verts.to(dev).requires_grad_(True); f = faces.to(dev).int()
gs = torch.randn(B, 1, S, S, device=dev)
sdn_b200.DEFAULT_FAST_MATH = True
sdn_b200.timing_enable(True)
sdn_b200.lib().sdn_debug_pending.argtypes = [ctypes.c_char_p, ctypes.c_int]
threading.Thread(target=watchdog, daemon=True).start()
for it in range(6):
    t0 = time.time()
    out = r(v, f, render_type=RenderType.Silhouette)
    torch.autograd.backward([out], [gs])
    v.grad = None
    if sync_each:
        torch.cuda.synchronize()
    print('step', it, round(time.time() - t0, 4), 'mem GB', round(torch.cuda.memory_allocated() / 1e9, 2), round(torch.cuda.memory_reserved() / 1e9, 2), flush=True)
torch.cuda.synchronize()
print('done', flush=True)

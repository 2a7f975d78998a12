import sys, torch
sys.path[:0] = ['.', '3d-sdn_b200']
import sdn_b200
from textural.models.networks import GlobalGenerator, get_norm_layer
dev = torch.device('cuda:0')
torch.manual_seed(0)
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm())
# 1. single conv layers, repeated
for name, N, H, W, ci, co, k, s, p in [('stem', 1, 518, 1542, 64, 64, 7, 1, 0), ('down1', 1, 512, 1536, 64, 128, 3, 2, 1),
                                       ('block', 1, 34, 98, 1024, 1024, 3, 1, 0), ('up4 ph', 1, 256, 768, 128, 64, 2, 1, 0), ('final', 1, 518, 1542, 64, 16, 7, 1, 0)]:
    x = torch.randn(N, H, W, ci, device=dev).bfloat16()
    w = (torch.randn(co, k, k, ci, device=dev) / (ci * k * k) ** 0.5).bfloat16()
    st = torch.empty(N, co, 2, device=dev)
    a = sdn_b200.conv2d_tc(x, w, s, p).float()
    b = sdn_b200.conv2d_tc(x, w, s, p).float()
    c = sdn_b200.conv2d_tc(x, w, s, p, stats=st).float() if co != 16 else a
    ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).float(), w.permute(0, 3, 1, 2).float(), stride=s, padding=p).permute(0, 2, 3, 1)
    print('%-8s repeat %.2e  with-stats %.2e  vs torch %.2e' % (name, rel(a, b), rel(c, a), rel(a, ref)), flush=True)
g = GlobalGenerator(48, 3, 64, 4, 9, get_norm_layer('instance')).to(dev).train()
for p_ in g.parameters(): p_.requires_grad_(False)
x = torch.randn(1, 48, 512, 1536, device=dev)
with torch.no_grad():
    y1 = g(x); y2 = g(x)
    print('generator eager vs eager', rel(y1, y2))
    ref = g.model(x)
    print('generator vs torch fp32 module', rel(y1, ref))

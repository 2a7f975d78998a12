import os, sys, torch
sys.path[:0] = ['.', '3d-sdn_b200']
import sdn_b200
dev = torch.device('cuda:0')
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for name, N, H, W, ci, co, k, s, p in [('stem', 1, 518, 1542, 64, 64, 7, 1, 0), ('down1', 1, 512, 1536, 64, 128, 3, 2, 1),
                                       ('block', 1, 34, 98, 1024, 1024, 3, 1, 0), ('up4 phase 2x2', 1, 256, 768, 128, 64, 2, 1, 0),
                                       ('up1 phase 2x2', 1, 32, 96, 1024, 512, 2, 1, 0)]:
    x = torch.randn(N, H, W, ci, device=dev).bfloat16()
    w = torch.randn(co, k, k, ci, device=dev).bfloat16()
    st = torch.empty(N, co, 2, device=dev)
    a = t(lambda: sdn_b200.conv2d_tc(x, w, s, p))
    b = t(lambda: sdn_b200.conv2d_tc(x, w, s, p, stats=st))
    print('%-16s plain %7.1f us   with stats %7.1f us' % (name, a, b), flush=True)

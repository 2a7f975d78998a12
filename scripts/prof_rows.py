import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, '3d-sdn_b200')]
import torch, sdn_b200
dev = torch.device('cuda:0')
Cout = int(os.environ.get('PROF_COUT', 64))
x = torch.randn(1, 518, 1542, 64, device=dev).bfloat16()
w = sdn_b200.pack_conv_weight(torch.randn(Cout, 64, 7, 7, device=dev) / 56.0)
for _ in range(3):
    y = sdn_b200.conv2d_tc(x, w, 1, 0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    y = sdn_b200.conv2d_tc(x, w, 1, 0)
e1.record(); torch.cuda.synchronize()
print('us per call', e0.elapsed_time(e1) / 5 * 1e3)

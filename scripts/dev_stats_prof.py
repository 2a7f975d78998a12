import sys, torch
sys.path[:0] = ['.', '3d-sdn_b200']
import sdn_b200
dev = torch.device('cuda:0')
x = torch.randn(1, 512, 1536, 64, device=dev).bfloat16()
w = torch.randn(128, 3, 3, 64, device=dev).bfloat16()
st = torch.empty(1, 128, 2, device=dev)
for _ in range(3):
    sdn_b200.conv2d_tc(x, w, 2, 1, stats=st)
torch.cuda.synchronize()

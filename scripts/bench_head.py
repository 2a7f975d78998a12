#!/usr/bin/env python
"""Encoder head at M = 16: one fused fp32 launch (sdn_mlp_head_forward) vs four k_linear_tc launches vs torch.nn.Linear
(cuBLAS fp32), CUDA-graph replay of 20 forwards, CUDA events."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, '3d-sdn_b200')]
import torch, torch.nn.functional as F
import sdn_b200
dev = torch.device('cuda:0')
torch.manual_seed(0)
fcs = [torch.nn.Linear(512, 1000).to(dev), torch.nn.Linear(1004, 256).to(dev), torch.nn.Linear(256, 256).to(dev), torch.nn.Linear(256, 43).to(dev)]
M = int(os.environ.get('M', 16))
x, extra = torch.randn(M, 512, device=dev), torch.randn(M, 4, device=dev)


def graph_us(fn):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn(); fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / 100 * 1e3, 2)


def torch_head():
    h = torch.cat([F.relu(fcs[0](x)), extra], dim=1)
    return fcs[3](F.relu(fcs[2](F.relu(fcs[1](h)))))


def tc_head():
    h = sdn_b200.linear_tc(x, fcs[0].weight, fcs[0].bias, relu=True)
    h = torch.cat([h, extra], dim=1)
    h = sdn_b200.linear_tc(h, fcs[1].weight, fcs[1].bias, relu=True)
    h = sdn_b200.linear_tc(h, fcs[2].weight, fcs[2].bias, relu=True)
    return sdn_b200.linear_tc(h, fcs[3].weight, fcs[3].bias, relu=False)


with torch.no_grad():
    out = {'M': M, 'fused_fp32_one_launch_us': graph_us(lambda: sdn_b200.mlp_head(x, extra, *fcs)),
           'four_k_linear_tc_launches_us': graph_us(tc_head), 'torch_cublas_fp32_us': graph_us(torch_head)}
print(json.dumps(out))

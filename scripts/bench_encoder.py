#!/usr/bin/env python
"""Encoder forward at B = 16, 224x224 (eval): ResNet-18 trunk on the tcgen05 convolution vs torchvision / cuDNN (fp32 and
bf16 channels-last), CUDA events.  One JSON line."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, '3d-sdn_b200')]
import torch
import sdn_b200
from derender3d.models.derenderer import Derenderer


def time_ms(fn, steps=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


dev = torch.device('cuda:0')
torch.manual_seed(0)
enc = Derenderer().to(dev).eval()
x = torch.randn(16, 3, 224, 224, device=dev)
out = {}
with torch.no_grad():
    out['cudnn_fp32_ms'] = round(time_ms(lambda: enc._trunk(x)), 4)
    encb = Derenderer().to(dev).eval()
    encb.net = encb.net.to(memory_format=torch.channels_last).bfloat16()
    xb = x.bfloat16().contiguous(memory_format=torch.channels_last)
    out['cudnn_bf16_channels_last_ms'] = round(time_ms(lambda: encb._trunk(xb)), 4)
    enc.tensor_core_trunk = True
    out['tcgen05_ms'] = round(time_ms(lambda: enc._trunk(x)), 4)
    g = torch.cuda.CUDAGraph()
    enc._trunk(x)
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        enc._trunk(x)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        y = enc._trunk(x)
    out['tcgen05_cuda_graph_ms'] = round(time_ms(g.replay), 4)
    sdn_b200.timing_enable(True); sdn_b200.timing_read()
    for _ in range(5):
        enc._trunk(x)
    torch.cuda.synchronize()
    out['kernels_ms'] = {k: round(t[0] / 5, 4) for k, t in sdn_b200.timing_read().items()}
    sdn_b200.timing_enable(False)
out['workload'] = 'ResNet-18 trunk forward, eval mode, B = 16, 224x224 (1.82 GMAC per image)'
out['tflops_tcgen05_graph'] = round(16 * 1.82e9 * 2 / out['tcgen05_cuda_graph_ms'] / 1e9, 1)
print(json.dumps(out))

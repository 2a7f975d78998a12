#!/usr/bin/env python
"""Encoder trunk forward + backward at B = 16, 224x224 (train mode, batch-statistics BatchNorm): the native path
(tcgen05 convolution forward / dgrad / wgrad + BatchNorm kernels, encoder_tc.resnet18_trunk_train) against torchvision on
cuDNN in fp32 and under bf16 autocast (channels-last), eager and as CUDA graphs.  CUDA events.  One JSON line."""
import copy, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, '3d-sdn_b200')]
import torch
import sdn_b200
from derender3d.models.derenderer import Derenderer
from derender3d.models.encoder_tc import resnet18_trunk_train


def time_ms(fn, steps=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def graphed(fn):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g


B = int(os.environ.get('B', 16))
dev = torch.device('cuda:0')
torch.manual_seed(0)
base = Derenderer().to(dev).train()
x = torch.randn(B, 3, 224, 224, device=dev)
t = torch.randn(B, 512, device=dev)
out = {}


def make(kind):
    enc = copy.deepcopy(base)
    for p in enc.parameters():
        p.grad = torch.zeros_like(p)
    if kind == 'fp32':
        def step():
            (enc._trunk(x) * t).sum().backward()
    elif kind == 'autocast':
        enc.net = enc.net.to(memory_format=torch.channels_last)
        xc = x.contiguous(memory_format=torch.channels_last)

        def step():
            with torch.autocast('cuda', dtype=torch.bfloat16):
                f = enc._trunk(xc)
            (f.float() * t).sum().backward()
    else:
        enc.net = enc.net.to(memory_format=torch.channels_last)     # [Cout,KH,KW,Cin] in memory: what the kernels read
        for p in enc.parameters():
            p.grad = torch.zeros_like(p)

        def step():
            (resnet18_trunk_train(enc.net, x) * t).sum().backward()
    return step


ONLY = os.environ.get('ONLY')
if ONLY:                                # profiling runs: a few eager steps of one variant, nothing else
    step = make(ONLY)
    for _ in range(int(os.environ.get('STEPS', 2))):
        step()
    torch.cuda.synchronize()
    sys.exit(0)
for kind in ('fp32', 'autocast', 'native'):
    step = make(kind)
    out[kind + '_eager_ms'] = round(time_ms(step), 4)
    try:
        g = graphed(step)
        out[kind + '_graph_ms'] = round(time_ms(g.replay), 4)
    except Exception as e:      # noqa: BLE001
        out[kind + '_graph_ms'] = 'failed: %s' % str(e)[:120]
        torch.cuda.synchronize()
step = make('native')
sdn_b200.timing_enable(True); sdn_b200.timing_read()
for _ in range(5):
    step()
torch.cuda.synchronize()
out['native_kernels_ms'] = {k: [round(v[0] / 5, 4), v[1] // 5] for k, v in sdn_b200.timing_read().items()}
sdn_b200.timing_enable(False)
out['workload'] = 'ResNet-18 trunk forward + backward, train mode, B = %d, 224x224 (3 x 1.82 GMAC per image)' % B
print(json.dumps(out))

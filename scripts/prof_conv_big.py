#!/usr/bin/env python
"""The generator's 3x3 convolution at 1024 channels ([1,34,98,1024] reflect-padded -> [1,32,96,1024]): tile variants of
k_conv_tc (SDN_CONV_TN / SDN_CONV_PAIR are read once per process, so each variant is a subprocess), with and without the
InstanceNorm statistics in the epilogue.  CUDA-graph replay of 20 calls, CUDA events."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == 'child':
    sys.path[:0] = [ROOT, os.path.join(ROOT, '3d-sdn_b200')]
    import torch, sdn_b200
    dev = torch.device('cuda:0')
    H, W, C = int(os.environ.get('PROF_H', 34)), int(os.environ.get('PROF_W', 98)), int(os.environ.get('PROF_C', 1024))
    x = torch.randn(1, H, W, C, device=dev).bfloat16()
    w = sdn_b200.pack_conv_weight(torch.randn(C, C, 3, 3, device=dev) / (C * 9) ** 0.5)
    stats = torch.zeros(1, C, 2, device=dev)
    out = {}
    for name, kw in (('plain', {}), ('stats', {'stats': stats})):
        fn = lambda: sdn_b200.conv2d_tc(x, w, 1, 0, **kw)
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
        us = e0.elapsed_time(e1) / 100 * 1e3
        out[name + '_us'] = round(us, 1)
        out[name + '_tflops'] = round(2.0 * (H - 2) * (W - 2) * C * C * 9 / us / 1e6, 1)
    print(json.dumps(out))
    sys.exit(0)
for label, env in (('TN=128', {'SDN_CONV_TN': '128', 'SDN_CONV_PAIR': '0'}), ('TN=256', {'SDN_CONV_TN': '256', 'SDN_CONV_PAIR': '0'}),
                   ('pair 256x256', {'SDN_CONV_PAIR': '1'})):
    e = dict(os.environ); e.update(env)
    r = subprocess.run([sys.executable, __file__, 'child'], env=e, capture_output=True, text=True, timeout=200)
    print(label, (r.stdout.strip().splitlines() or [r.stderr[-300:]])[-1], flush=True)

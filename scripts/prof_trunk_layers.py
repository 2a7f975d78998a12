#!/usr/bin/env python
"""Per-layer device time of the training trunk's kernels at B = 16 (ResNet-18 shapes): forward conv (+stats), dgrad, wgrad,
BatchNorm backward, affine.  Each op is captured REPS times in one CUDA graph and replayed (no per-launch event overhead)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, '3d-sdn_b200')]
import torch
import sdn_b200

dev = torch.device('cuda:0')
REPS = 20
B = int(os.environ.get('B', 16))
SHAPES = [  # name, H, W, Cin, Cout, K, stride, pad
    ('stem(im2col 1x1)', 112, 112, 192, 64, 1, 1, 0), ('l1 3x3 64', 56, 56, 64, 64, 3, 1, 1), ('l2.0 3x3 s2', 56, 56, 64, 128, 3, 2, 1),
    ('l2 ds 1x1 s2', 56, 56, 64, 128, 1, 2, 0), ('l2 3x3 128', 28, 28, 128, 128, 3, 1, 1), ('l3.0 3x3 s2', 28, 28, 128, 256, 3, 2, 1),
    ('l3 ds 1x1 s2', 28, 28, 128, 256, 1, 2, 0), ('l3 3x3 256', 14, 14, 256, 256, 3, 1, 1), ('l4.0 3x3 s2', 14, 14, 256, 512, 3, 2, 1),
    ('l4 ds 1x1 s2', 14, 14, 256, 512, 1, 2, 0), ('l4 3x3 512', 7, 7, 512, 512, 3, 1, 1)]


def graph_ms(fn):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn(); fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(REPS):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5 / REPS * 1e3      # us per call


rows = []
for name, H, W, Cin, Cout, K, s, p in SHAPES:
    Ho, Wo = (H + 2 * p - K) // s + 1, (W + 2 * p - K) // s + 1
    x = torch.randn(B, H, W, Cin, device=dev).bfloat16()
    w = (torch.randn(Cout, Cin, K, K, device=dev) / (Cin * K * K) ** 0.5).contiguous(memory_format=torch.channels_last)
    wp = sdn_b200.pack_conv_weight(w)
    go = torch.randn(B, Ho, Wo, Cout, device=dev).bfloat16()
    stats = torch.empty(B, Cout, 2, device=dev)
    z = sdn_b200.conv2d_tc(x, wp, s, p, stats=stats)
    mean, rstd, gamma = torch.zeros(Cout, device=dev), torch.ones(Cout, device=dev), torch.ones(Cout, device=dev)
    flop = 2.0 * B * Ho * Wo * Cout * Cin * K * K
    r = {'layer': name, 'gflop': round(flop / 1e9, 2)}
    r['fwd_us'] = round(graph_ms(lambda: sdn_b200.conv2d_tc(x, wp, s, p, stats=stats)), 1)
    r['fwd_nostats_us'] = round(graph_ms(lambda: sdn_b200.conv2d_tc(x, wp, s, p)), 1)
    if 'stem' not in name:
        r['dgrad_us'] = round(graph_ms(lambda: sdn_b200.conv2d_dgrad_tc(go, w, (H, W), s, p)), 1)
    r['wgrad_us'] = round(graph_ms(lambda: sdn_b200.conv2d_wgrad_tc(go, x, K, K, s, p)), 1)
    r['bn_bwd_us'] = round(graph_ms(lambda: sdn_b200.batch_norm_backward_nhwc(go, z, z, mean, rstd, gamma, True, False)), 1)
    r['affine_us'] = round(graph_ms(lambda: sdn_b200.affine_act_nhwc(z, gamma, mean, None, True)), 1)
    r['wgrad_tflops'] = round(flop / r['wgrad_us'] / 1e6, 1)
    r['fwd_tflops'] = round(flop / r['fwd_nostats_us'] / 1e6, 1)
    rows.append(r)
    print(json.dumps(r), flush=True)

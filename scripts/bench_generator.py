#!/usr/bin/env python
"""BASELINE config 5: GlobalGenerator(48 -> 3, ngf 64, 4 down-samplings, 9 blocks, instance norm) forward at
[1, 48, 512, 1536], bf16 tensor-core convolution path (tcgen05).  Prints one JSON line: ms per image, TFLOP/s against
the measured bf16 peak, per-kernel device time; --layers adds a per-layer-shape microbenchmark of the convolution kernel
(every ResNet-18 shape at B = 16 and the generator's shapes)."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, '3d-sdn_b200')]
import torch
import sdn_b200
from textural.models.networks import GlobalGenerator, get_norm_layer


def conv_flops(model, H, W):
    """MACs * 2 of every convolution of the Sequential at input H x W (transposed convolutions: their own taps)."""
    total = 0
    h, w = H, W
    for m in model.modules():
        if isinstance(m, torch.nn.Conv2d):
            if m.stride[0] == 2:
                h, w = h // 2, w // 2
            total += 2 * m.out_channels * m.in_channels * m.kernel_size[0] * m.kernel_size[1] * h * w
        elif isinstance(m, torch.nn.ConvTranspose2d):
            total += 2 * m.out_channels * m.in_channels * 9 * h * w        # 9 taps per INPUT pixel
            h, w = h * 2, w * 2
    return total


def time_ms(fn, steps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--H', type=int, default=512)
    ap.add_argument('--W', type=int, default=1536)
    ap.add_argument('--blocks', type=int, default=9)
    ap.add_argument('--layers', action='store_true')
    ap.add_argument('--cudnn', action='store_true', help='also time the torch / cuDNN bf16 channels-last module (comparator)')
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    g = GlobalGenerator(48, 3, 64, 4, args.blocks, get_norm_layer('instance')).to(dev).train()
    for p in g.parameters():
        p.requires_grad_(False)
    x = torch.randn(1, 48, args.H, args.W, device=dev)
    with torch.no_grad():
        ms_eager = time_ms(lambda: g(x), args.steps)
        # the forward as ONE CUDA graph (static shapes; the TMA tensor maps are encoded on the host at capture time)
        ms = ms_eager
        launch = 'eager'
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                g(x)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                y_static = g(x)
            ref = g(x)
            graph.replay()
            torch.cuda.synchronize()
            # Two runs are not bit-identical: the instance-norm statistics are summed with float atomics, and after 27
            # layers of bf16 activations a last-bit difference decorrelates the rounding noise (two eager runs differ by
            # ~2e-2 rel-L2, each is ~3e-2 from the fp32 module).  The check guards against a broken capture, not rounding.
            err = float((ref - y_static).norm() / ref.norm())
            assert err < 6e-2, 'graph replay differs from the eager forward: %g' % err
            ms = time_ms(graph.replay, args.steps)
            launch = 'cuda graph replay'
        except Exception as e:      # noqa: BLE001
            launch = 'eager (graph capture failed: %s)' % repr(e)[:100]
            torch.cuda.synchronize()
        sdn_b200.timing_enable(True); sdn_b200.timing_read()
        for _ in range(3):
            g(x)
        torch.cuda.synchronize()
        kern = {k: round(t[0] / 3, 4) for k, t in sorted(sdn_b200.timing_read().items(), key=lambda kv: -kv[1][0])}
        sdn_b200.timing_enable(False)
    peaks = {}
    pth = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(pth):
        peaks = json.load(open(pth))
    fl = conv_flops(g.model, args.H, args.W)
    out = {'workload': 'config5: GlobalGenerator(48->3, ngf 64, 4 down, %d blocks, instance norm) forward [1,48,%d,%d], bf16 '
                       'tcgen05 convolutions, fp32 accumulate' % (args.blocks, args.H, args.W),
           'ms_per_image': round(ms, 3), 'launch': launch, 'graph_vs_eager_rel_l2': round(err, 4) if launch.startswith('cuda') else None, 'ms_per_image_eager': round(ms_eager, 3), 'conv_tflop': round(fl / 1e12, 3), 'tflops': round(fl / ms / 1e9, 1),
           'bf16_peak_sustained_tflops': peaks.get('bf16_tflops_sustained'),
           'frac_of_sustained_peak': round(fl / ms / 1e9 / peaks['bf16_tflops_sustained'], 4) if peaks else None,
           'kernels_ms': kern}
    if args.cudnn:
        gc = GlobalGenerator(48, 3, 64, 4, args.blocks, get_norm_layer('instance')).to(dev).train()
        gc = gc.to(memory_format=torch.channels_last).bfloat16()
        xc = x.bfloat16().contiguous(memory_format=torch.channels_last)
        with torch.no_grad():
            out['cudnn_bf16_channels_last_ms'] = round(time_ms(lambda: gc.model(xc), args.steps), 3)
    if args.layers:
        shapes = [('resnet18 B16 ' + n, 16, hw, hw, ci, co, k, s, p) for n, hw, ci, co, k, s, p in [
            ('layer1 3x3', 56, 64, 64, 3, 1, 1), ('layer2.0 3x3 s2', 56, 64, 128, 3, 2, 1), ('layer2.0 1x1 s2', 56, 64, 128, 1, 2, 0),
            ('layer2 3x3', 28, 128, 128, 3, 1, 1), ('layer3.0 3x3 s2', 28, 128, 256, 3, 2, 1), ('layer3 3x3', 14, 256, 256, 3, 1, 1),
            ('layer4.0 3x3 s2', 14, 256, 512, 3, 2, 1), ('layer4 3x3', 7, 512, 512, 3, 1, 1)]]
        shapes += [('generator stem 7x7', 1, 518, 1542, 64, 64, 7, 1, 0), ('generator down1', 1, 512, 1536, 64, 128, 3, 2, 1),
                   ('generator down4', 1, 64, 192, 512, 1024, 3, 2, 1), ('generator block 3x3 @1024', 1, 34, 98, 1024, 1024, 3, 1, 0),
                   ('gemm-like 1x1 @1024 M=65536', 1, 256, 256, 1024, 1024, 1, 1, 0)]
        rows = []
        for name, N, H, W, ci, co, k, s, p in shapes:
            xx = torch.randn(N, H, W, ci, device=dev).bfloat16()
            ww = torch.randn(co, k, k, ci, device=dev).bfloat16()
            t = time_ms(lambda: sdn_b200.conv2d_tc(xx, ww, s, p), 20)
            Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
            f = 2.0 * N * Ho * Wo * co * ci * k * k
            rows.append({'layer': name, 'us': round(t * 1e3, 1), 'tflops': round(f / t / 1e9, 1)})
        out['conv_layers'] = rows
    print(json.dumps(out))


if __name__ == '__main__':
    main()

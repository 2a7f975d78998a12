"""dev check of the tcgen05 linear kernel against torch (bf16-rounded operands, fp32 accumulate)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '3d-sdn_b200'))
import sdn_b200 as S

torch.manual_seed(0)
bad = 0
for (M, N, K, relu) in [(128, 128, 64, 0), (128, 128, 16, 0), (128, 128, 128, 0), (32, 256, 260, 1), (1, 1552, 256, 0),
                        (200, 300, 512, 1), (7, 5, 3, 0), (513, 1552, 256, 0), (4096, 256, 512, 1)]:
    x = torch.randn(M, K, device='cuda'); w = torch.randn(N, K, device='cuda') / K ** 0.5; b = torch.randn(N, device='cuda')
    y = S.linear_tc(x, w, b, bool(relu))
    torch.cuda.synchronize()
    ref = x.bfloat16().float() @ w.bfloat16().float().t() + b
    if relu: ref = ref.clamp_min(0)
    err = ((y - ref).norm() / ref.norm()).item()
    mx = (y - ref).abs().max().item()
    print('M=%d N=%d K=%d relu=%d  rel-L2 %.3e  max-abs %.3e' % (M, N, K, relu, err, mx), flush=True)
    bad += err > 1e-5
# timing: device time dominates at these sizes (python + ctypes cost ~20 us per call)
for (M, N, K) in [(4096, 256, 512), (65536, 256, 512), (65536, 1552, 256)]:
    x = torch.randn(M, K, device='cuda'); w = torch.randn(N, K, device='cuda'); b = torch.randn(N, device='cuda')
    xb, wb, bb = x.bfloat16(), w.bfloat16(), b.bfloat16()
    fs = {'tcgen05 (fp32 in/out)': lambda: S.linear_tc(x, w, b, True),
          'cuBLAS fp32': lambda: torch.relu(torch.nn.functional.linear(x, w, b)),
          'cuBLAS bf16 (bf16 in/out)': lambda: torch.relu(torch.nn.functional.linear(xb, wb, bb))}
    for name, f in fs.items():
        for _ in range(5): f()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(30): f()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 30
        print('M=%d N=%d K=%d %-28s %.4f ms  %.1f TFLOP/s' % (M, N, K, name, ms, 2.0 * M * N * K / ms / 1e9))
print('BAD' if bad else 'OK')

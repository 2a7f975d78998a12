#!/usr/bin/env python
"""BASELINE config 3: one full geometric-branch training step -- encoder (ResNet-18 + FC, bf16 channels-last through
cuDNN) -> pose algebra -> class sampling -> batched FFD / camera transform -> render 384x384 (768x768 internal) ->
masked MSE + REINFORCE -> Adam -- on synthetic VKITTI-shaped inputs with ShapeNet-sized template meshes."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, '3d-sdn_b200')):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import sdn_b200  # noqa: E402
from sdn_b200 import workloads as wl  # noqa: E402
from derender3d import TargetType  # noqa: E402
from derender3d.models import Derenderer3d, ShapenetObj  # noqa: E402
from derender3d.train_step import geometric_losses, GraphedTrainStep  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=16)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--render-size', type=int, default=384)
    ap.add_argument('--faces', type=int, default=72576, help='triangles per template mesh (ShapeNet: 29k-72k)')
    ap.add_argument('--fp32', action='store_true')
    ap.add_argument('--synthetic', action='store_true', help='synthetic templates even when the real OBJ files are present')
    ap.add_argument('--encoder-only', action='store_true')
    ap.add_argument('--graph', action='store_true', help='capture the whole step in a CUDA graph (GraphedTrainStep)')
    ap.add_argument('--native-encoder', action='store_true', help='encoder trunk trained on the tcgen05 kernels')
    ap.add_argument('--profile', action='store_true', help='torch.profiler table of one step (kernel time vs wall)')
    args = ap.parse_args()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:       # BASELINE config 4: one process per GPU, one flat-bucket all-reduce per step
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    from derender3d.bench_setup import train_setup
    B, S = args.batch, args.render_size
    st = train_setup(dev, B, S, args.faces, seed=rank, real=not args.synthetic, capturable=bool(args.graph),
                     native_encoder=args.native_encoder)
    model, opt, batch, objs = st['model'], st['opt'], st['batch'], st['objs']
    roi = batch['roi_norms']
    sync = None
    if world > 1 or args.native_encoder:
        from sdn_b200.dist import FlatGradSync
        sync = FlatGradSync(model, bucket_views=True)
    amp = not args.fp32

    def step():
        with torch.autocast('cuda', dtype=torch.bfloat16, enabled=amp):
            if args.encoder_only:
                m = torch.stack([roi[:, 2] + roi[:, 0], roi[:, 3] + roi[:, 1]], dim=1) / 2.0
                dr = torch.stack([roi[:, 2] - roi[:, 0], roi[:, 3] - roi[:, 1]], dim=1)
                out = model.derenderer(batch['images'], m, dr)
                loss = sum(v.float().pow(2).mean() for v in out.values())
                cov = 0.0
            else:
                blob = model(batch['images'], batch['roi_norms'], batch['focals'])
                loss = sum(geometric_losses(blob, batch, TargetType.full).values())
                cov = blob['_masks']
        if sync is not None:
            sync.zero()
        else:
            opt.zero_grad(set_to_none=True)
        loss.backward()
        if sync is not None:
            sync()
        opt.step()
        return loss, cov

    if args.graph:
        assert not args.encoder_only
        gs = GraphedTrainStep(model, opt, batch, TargetType.full, autocast_dtype=torch.bfloat16 if amp else None,
                              grad_sync=sync)

        def step():     # noqa: F811
            return gs.step(), gs.blob['_masks']

    loss, cov = step()
    coverage = float((cov > 0).float().mean()) if not args.encoder_only else 0.0
    for _ in range(3):
        step()
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss, _ = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:       # max over ranks
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    barrier()
    sdn_b200.timing_enable(True)
    sdn_b200.timing_read()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    kern = {k: round(t[0] / 3, 3) for k, t in sorted(sdn_b200.timing_read().items(), key=lambda kv: -kv[1][0])[:6]}
    sdn_b200.timing_enable(False)
    if args.profile:
        from torch.profiler import profile, ProfilerActivity
        import time
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            step()
        host_ms = (time.perf_counter() - t0) / 5 * 1e3
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                step()
            torch.cuda.synchronize()
        ev = prof.key_averages()
        tot = sum(e.self_device_time_total for e in ev) / 3e3
        print('host dispatch ms/step (no sync): %.2f   device kernel ms/step: %.2f   wall ms/step: %.2f' % (host_ms, tot, ms))
        print(ev.table(sort_by='self_cuda_time_total', row_limit=30, max_name_column_width=60))
        print(ev.table(sort_by='self_cpu_time_total', row_limit=25, max_name_column_width=60))
    nparam = sum(p.numel() for p in model.parameters())
    if world > 1:
        w0 = model.derenderer._fc3.weight.detach().clone()
        dist.broadcast(w0, src=0)
        assert torch.equal(w0, model.derenderer._fc3.weight.detach()), 'replicas diverged'
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    B = B * world
    print(json.dumps({'workload': ('config4' if world > 1 else 'config3') + ': B=%d, encoder 224x224 %s, %d-triangle templates, render %dx%d, mode full, Adam'
                      % (B, 'fp32' if args.fp32 else 'bf16 autocast channels-last', len(objs[0].faces), S, S),
                      'encoder_only': args.encoder_only, 'cuda_graph': bool(args.graph), 'n_gpus': world, 'global_batch': B, 'ms_per_step': round(ms, 3),
                      'objects_per_s': round(B / ms * 1e3, 1), 'loss': float(loss), 'mask_coverage': round(coverage, 3),
                      'parameters': nparam, 'render_kernels_ms': kern}))


if __name__ == '__main__':
    main()

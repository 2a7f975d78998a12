"""Synthetic set-up of BASELINE configs 3 / 4: one full geometric-branch training step -- encoder ->
pose algebra -> class sampling -> fused FFD / camera transform -> render 384x384 (768x768 internal) -> masked MSE +
REINFORCE -> [one flat-bucket gradient all-reduce] -> Adam -- shared by bench.py and scripts/bench_train_step.py."""
import os

import numpy as np
import torch

from derender3d import TargetType
from derender3d.models import SHAPENET_OBJS, Derenderer3d, ShapenetObj


def reference_assets_dir():
    """geometric/assets of the reference (the eight real ShapeNet templates) when a copy is reachable: the git-ignored
    travelling copy baseline/_ref/geometric/assets (staged by __graft_entry__.build()) or $SHAPENET_ROOT_DIR."""
    here = os.path.dirname(os.path.abspath(__file__))
    repo = os.path.dirname(os.path.dirname(here))
    for cand in (os.environ.get('SHAPENET_ROOT_DIR'), os.path.join(repo, 'baseline', '_ref', 'geometric', 'assets')):
        if cand and os.path.isfile(os.path.join(cand, SHAPENET_OBJS[0][0], SHAPENET_OBJS[0][1], 'models',
                                                'model_normalized.obj')):
            return cand
    return None


def template_meshes(faces=72576, real=True):
    """-> (list of ShapenetObj, description).  The reference's eight OBJ files when present, else closed superellipsoids
    of `faces` triangles each (ShapeNet templates have 29k-72k)."""
    from sdn_b200 import workloads as wl
    root = reference_assets_dir() if real else None
    if root:
        objs = [ShapenetObj(c, o, root_dir=root) for c, o in SHAPENET_OBJS]
        return objs, 'the 8 ShapeNet templates of geometric/assets (%d-%d triangles)' % (
            min(len(o.faces) for o in objs), max(len(o.faces) for o in objs))
    n_lon = 192
    n_lat = max(4, faces // (2 * n_lon) + 1)
    objs = []
    for k in range(8):
        v, f = wl.superellipsoid_mesh(n_lon, n_lat, 0.5 + 0.04 * k, 0.6)
        objs.append(ShapenetObj(vertices=v * np.array([1.0, 0.35, 0.45], np.float32), faces=f))
    return objs, '8 synthetic %d-triangle superellipsoids' % len(objs[0].faces)


def train_setup(dev, B=16, S=384, faces=72576, seed=0, real=True, capturable=True, native_encoder=False):
    """-> dict(model, opt, batch, desc): Derenderer3d(mode full) in train mode, Adam(lr 1e-4, wd 1e-3), one synthetic
    VKITTI-shaped batch of B objects (images ~ N(0,1), ROIs, focals 725, random GT masks, ignores 0, targets full)."""
    torch.manual_seed(0)
    objs, desc = template_meshes(faces, real)
    model = Derenderer3d(TargetType.full, image_size=256, render_size=S, objs=objs).to(dev).train()
    model.derenderer.net = model.derenderer.net.to(memory_format=torch.channels_last)
    # native_encoder: the ResNet-18 trunk's forward, dgrad, wgrad and BatchNorm on this repo's kernels instead of cuDNN
    model.derenderer.tensor_core_trunk = bool(native_encoder)
    with torch.no_grad():       # a pose head that puts the objects in view (a real run starts from the pre-trained stage)
        model.derenderer._fc3.weight.mul_(0.01)
        model.derenderer._fc3.bias.zero_()
        model.derenderer._fc3.bias[0] = 1.0
        model.derenderer._fc3.bias[4:7] = torch.tensor([1.3, 0.4, 0.5])
        model.derenderer._fc3.bias[7] = 2.2
    # torch's fused multi-tensor Adam: one kernel over all 62 parameter tensors instead of ~a dozen foreach passes
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=1e-3, capturable=capturable, fused=dev.type == 'cuda')
    g = torch.Generator().manual_seed(seed)
    c = 0.3 * (torch.rand(B, 2, generator=g) - 0.5)
    d = 0.25 + 0.15 * torch.rand(B, 2, generator=g)
    roi = torch.cat([c - d / 2, c + d / 2], dim=1).to(dev)
    batch = {'images': torch.randn(B, 3, 224, 224, generator=g).to(dev).contiguous(memory_format=torch.channels_last),
             'roi_norms': roi, 'focals': torch.full((B, 1), 725.0, device=dev),
             'targets': torch.full((B,), TargetType.full, device=dev, dtype=torch.long),
             'thetas': torch.rand(B, 1, generator=g).to(dev), 'translation2ds': torch.zeros(B, 2, device=dev),
             'log_scales': torch.tensor([1.3, 0.4, 0.5], device=dev).expand(B, 3).contiguous(),
             'log_depths': torch.full((B, 1), 2.2, device=dev),
             'masks': (torch.rand(B, 1, 256, 256, generator=g) > 0.5).float().to(dev),
             'ignores': torch.zeros(B, 1, 256, 256, device=dev)}
    return {'model': model, 'opt': opt, 'batch': batch, 'templates': desc, 'objs': objs}


def measure_train(dev, world, rank, B=16, S=384, steps=10, graph=True, real=True, faces=72576, native_encoder=False):
    """Times `steps` training steps (CUDA events; max over ranks).  world > 1: one FlatGradSync all-reduce per step inside
    the timed region (config 4).  -> dict for the bench line (rank 0 fills it; the other ranks return None)."""
    import torch.distributed as dist
    from derender3d.train_step import GraphedTrainStep, geometric_losses
    from sdn_b200.dist import FlatGradSync
    st = train_setup(dev, B, S, faces, seed=0, real=real, native_encoder=native_encoder)   # same data on every rank: the scaling line measures the
    model, opt, batch = st['model'], st['opt'], st['batch']      # machine, not the heaviest scene
    # flat gradient bucket: the all-reduce buffer of config 4; at one GPU it makes the gradients persistent views, which the
    # native trunk's wgrad / BatchNorm kernels accumulate into in place (one memset per step instead of ~60 add kernels)
    sync = FlatGradSync(model, bucket_views=True) if (world > 1 or native_encoder) else None
    amp = torch.bfloat16
    launch = 'eager'
    gs = None

    def eager_step():
        with torch.autocast('cuda', dtype=amp):
            blob = model(batch['images'], batch['roi_norms'], batch['focals'])
            loss = sum(geometric_losses(blob, batch, TargetType.full).values())
        if sync is not None:
            sync.zero()
        else:
            opt.zero_grad(set_to_none=True)
        loss.backward()
        if sync is not None:
            sync()
        opt.step()
        return loss.detach()
    step = eager_step
    if world > 1:
        dist.all_reduce(torch.zeros(1, device=dev))             # NCCL communicator exists before any capture
    if graph:
        try:
            gs = GraphedTrainStep(model, opt, batch, TargetType.full, autocast_dtype=amp, grad_sync=sync)
            step, launch = gs.step, 'cuda graph replay' + (' (all-reduce captured in the graph)' if world > 1 else '')
        except Exception as e:      # noqa: BLE001 -- report and time the eager step
            launch = 'eager (graph capture failed: %s)' % repr(e)[:120]
            torch.cuda.synchronize()
    for _ in range(3):
        loss = step()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    ranks_ms = [ms]
    in_sync = True
    allreduce_ms = None
    if world > 1:
        t = torch.tensor([ms], device=dev)
        all_t = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(all_t, t)
        ranks_ms = [float(x.item()) for x in all_t]
        ms = max(ranks_ms)
        # the collective on its own: 11.84 M fp32 = 47.4 MB, NCCL ring / NVLS over NVSwitch
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(10):
            dist.all_reduce(sync.flat)
        a1.record()
        torch.cuda.synchronize()
        allreduce_ms = a0.elapsed_time(a1) / 10
        sync.flat.zero_()
        in_sync = gs.replicas_in_sync() if gs is not None else True
        if gs is None:
            w0 = model.derenderer._fc3.weight.detach().clone()
            dist.broadcast(w0, src=0)
            ok = torch.tensor([1.0 if torch.equal(w0, model.derenderer._fc3.weight.detach()) else 0.0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            in_sync = bool(ok.item() > 0)
        barrier()
    if rank != 0:
        return None
    nparam = sum(p.numel() for p in model.parameters())
    out = {'workload': ('config4' if world > 1 else 'config3') + ': B=%d/GPU, encoder 224x224 %s, %s, '
                       'render %dx%d 2xSSAA, mode full, Adam' % (
                           B, 'trunk fwd/dgrad/wgrad/BatchNorm on the tcgen05 kernels + tcgen05 head (bf16)' if native_encoder
                           else 'bf16 autocast channels-last (cuDNN)', st['templates'], S, S),
           'n_gpus': world, 'global_batch': B * world, 'ms_per_step': round(ms, 3),
           'objects_per_s': round(B * world / ms * 1e3, 1), 'launch': launch, 'loss': float(loss),
           'rank_ms': {'min': round(min(ranks_ms), 3), 'median': round(sorted(ranks_ms)[len(ranks_ms) // 2], 3),
                       'max': round(max(ranks_ms), 3)}}
    if world > 1:
        out['collective'] = {'what': 'one flat fp32 gradient bucket, ncclAllReduce(sum) + scale, inside the timed step',
                             'elements': nparam, 'mbytes': round(nparam * 4 / 1e6, 1),
                             'allreduce_alone_ms': round(allreduce_ms, 3),
                             'bus_gb_s': round(2 * (world - 1) / world * nparam * 4 / (allreduce_ms * 1e-3) / 1e9, 1),
                             'replicas_in_sync': bool(in_sync)}
    return out

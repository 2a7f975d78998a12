#!/usr/bin/env python
"""bench.py -- images/sec of render forward+backward at 375x1242 (BASELINE.json metric, config 2).

One "step" = forward (silhouette + depth + normal, 2x SSAA) and backward (N(0,1) upstream gradients) of ONE batch of
32 synthetic VKITTI-shaped scenes per GPU.  See DESIGN.md section "Measurement" for the algorithmic-byte model.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch 32]
  N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N

Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import math

if '--impl' in sys.argv and 'reference' in sys.argv:
    # torch.distributed.run exports OMP_NUM_THREADS=1 to its workers; the CPU arm is meant to use the host's cores
    # (set before libgomp / torch are loaded; `cores` in the JSON line reports what OpenMP then really used)
    os.environ['OMP_NUM_THREADS'] = str(max(1, (os.cpu_count() or 2) // 2))

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, '3d-sdn_b200')):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

H, W, SS = 375, 1242, 2
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures (profiles/)
NCU_TRAFFIC = {'k_line_sweep': 195101696,      # profiles/r1_prof_line_v4.txt
               'k_line_split': 186921984 + 14710016}      # profiles/r2_raster_kernels.md (dram read + write per launch)
NF_EFF = 2 * 1984
C_OUT = 5


def algorithmic_bytes(c_out=C_OUT, f_eff=NF_EFF):
    """per image."""
    P = H * W
    fwd = 36 * f_eff + 4 * c_out * P + 4 * SS * SS * P
    bwd = 4 * c_out * P + 4 * SS * SS * P + 36 * f_eff + 36 * f_eff
    return fwd, bwd


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(path):
        with open(path) as fh:
            return float(json.load(fh)['hbm_gbs']), 'measured'
    return 6650.0, 'fallback'


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super(ClockSampler, self).__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                self.samples.append([x.strip() for x in out.strip().split(',')])
            except Exception:
                pass
            time.sleep(0.02)

    def summary(self):
        sm = sorted(int(s[0]) for s in self.samples if s and s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if len(s) > 1 and s[1].isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({n for s in self.samples if len(s) >= 6 for n, v in zip(names, s[2:6]) if v == 'Active'})
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(sm)}


def make_inputs(batch, seed, dev):
    from sdn_b200 import workloads as wl
    verts, faces = wl.vkitti_scene_batch(batch, seed=seed)
    gen = torch.Generator().manual_seed(1000 + seed)
    grads = {'silhouette': torch.randn(batch, 1, H, W, generator=gen),
             'depth': torch.randn(batch, 1, H, W, generator=gen),
             'normal': torch.randn(batch, 3, H, W, generator=gen)}
    return verts, faces, grads


# ----------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    import sdn_b200
    from sdn_b200 import workloads as wl
    from derender3d.models.renderer import Renderer

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device; there is no CPU fallback'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    sdn_b200.lib()

    B = args.batch
    # every rank renders the SAME 32 scenes unless --rank-seeds: the weak-scaling line then measures the machine and not
    # the heaviest rank's scenes (the work per image is data dependent); results per rank are printed either way
    verts, faces, grads = make_inputs(B, rank if args.rank_seeds else 0, dev)
    renderer = Renderer(image_size=(H, W), viewing_angle=wl.vkitti_viewing_angle())
    fast = not args.exact
    v_dev = verts.to(dev).requires_grad_(True)
    f_dev = faces.to(dev).int()
    g_dev = {k: v.to(dev) for k, v in grads.items()}
    v_pin = verts.pin_memory()
    # end-to-end loop: two steps in flight (what a training / fitting loop does): step i's vertex gradients and loss
    # land in pinned slot i % 2 and are consumed on the host while step i + 1 runs on the device
    gv_host = [torch.empty_like(verts).pin_memory() for _ in range(2)]
    loss_host = [torch.empty(1).pin_memory() for _ in range(2)]
    done = [None, None]
    e2e_state = {'i': 0, 'consumed': 0, 'loss': 0.0}

    def step_resident():
        out = renderer.forward_multi(v_dev, f_dev, fast_math=fast)
        torch.autograd.backward([out['silhouette'], out['depth'], out['normal']],
                                [g_dev['silhouette'], g_dev['depth'], g_dev['normal']])
        v_dev.grad = None

    def e2e_consume(slot):
        if done[slot] is not None:
            done[slot].synchronize()
            e2e_state['loss'] = float(loss_host[slot][0]) + float(gv_host[slot][0, 0, 0]) * 0.0
            e2e_state['consumed'] += 1
            done[slot] = None

    def step_e2e():
        # the call a user makes: host vertices in; loss = <outputs, upstream> and its vertex gradients back on the host
        slot = e2e_state['i'] & 1
        e2e_state['i'] += 1
        e2e_consume(slot)                       # this slot's previous step must have been read before it is reused
        v = v_pin.to(dev, non_blocking=True).requires_grad_(True)
        out = renderer.forward_multi(v, f_dev, fast_math=fast)
        outs = [out['silhouette'], out['depth'], out['normal']]
        ups = [g_dev['silhouette'], g_dev['depth'], g_dev['normal']]
        with torch.no_grad():
            loss = sum(torch.dot(o.reshape(-1), g.reshape(-1)) for o, g in zip(outs, ups))
        torch.autograd.backward(outs, ups)      # d loss / d out = upstream
        gv_host[slot].copy_(v.grad, non_blocking=True)
        loss_host[slot].copy_(loss.view(1), non_blocking=True)
        done[slot] = torch.cuda.Event()
        done[slot].record()
        e2e_consume(slot ^ 1)                   # read the previous step's result while this one runs

    def e2e_drain():
        e2e_consume(0)
        e2e_consume(1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    rank_ms = []

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        rank_ms[:] = [ms]
        if world > 1:
            t = torch.tensor([ms], device=dev)
            all_t = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(all_t, t)
            rank_ms[:] = [float(x.item()) for x in all_t]
            ms = max(rank_ms)
        barrier()
        return ms

    for _ in range(max(args.warmup, 3)):
        step_resident()
    # host-side dispatch cost of one eager step (python + autograd + ctypes), reported next to the device time
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        step_resident()
    host_ms = (time.perf_counter() - t0) / 20 * 1e3
    torch.cuda.synchronize()
    launches0 = sdn_b200.launch_count()
    step_resident()
    launches_per_step = sdn_b200.launch_count() - launches0

    # the whole step (forward + backward of all three outputs) as ONE CUDA graph: shapes, launch grids and the
    # workspace are static, nothing in the library reads back to the host
    graph, timed_step = None, step_resident
    if args.graph:
        try:
            step_resident_keep = lambda: torch.autograd.backward(      # noqa: E731 -- eager result to check against
                list(renderer.forward_multi(v_dev, f_dev, fast_math=fast)[k] for k in ('silhouette', 'depth', 'normal')),
                [g_dev['silhouette'], g_dev['depth'], g_dev['normal']])
            v_dev.grad = None
            step_resident_keep()
            gref = v_dev.grad.clone()
            v_dev.grad = None
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                step_resident()
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = renderer.forward_multi(v_dev, f_dev, fast_math=fast)
                torch.autograd.backward([out['silhouette'], out['depth'], out['normal']],
                                        [g_dev['silhouette'], g_dev['depth'], g_dev['normal']])
            graph.replay()
            torch.cuda.synchronize()
            err = float((gref - v_dev.grad).norm() / gref.norm())
            assert err < 1e-5, 'graph replay differs from the eager step: rel-L2 %g' % err
            timed_step = graph.replay
        except Exception as e:      # noqa: BLE001 -- report and time the eager step instead
            print('[bench] CUDA graph capture failed (%s); timing the eager step' % repr(e)[:300], file=sys.stderr)
            graph = None
            torch.cuda.synchronize()
        for _ in range(3):
            timed_step()
    sampler = ClockSampler(local)      # rank 0 samples its own GPU; eight concurrent nvidia-smi pollers only add noise
    if rank == 0:
        sampler.start()
    ms = timed(timed_step, args.steps)
    step_rank_ms = sorted(x / args.steps for x in rank_ms)
    launches = launches_per_step * args.steps
    sampler.stop_flag = True
    v_dev.grad = None

    for _ in range(3):
        step_e2e()
    e2e_drain()

    # the same end-to-end step as one CUDA graph per pinned slot: H2D copy of the vertices, forward, loss, backward and
    # the D2H copies of loss + vertex gradients are all nodes of the graph (pinned host memory on both ends)
    e2e_fn, e2e_launch = step_e2e, 'eager'
    if args.graph:
        try:
            ref_loss, ref_gv = float(loss_host[0][0]), gv_host[0].clone()
            e2e_graphs, v_static = [], []
            for slot in range(2):
                vs = torch.empty_like(v_dev).requires_grad_(True)
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    with torch.no_grad():
                        vs.copy_(v_pin, non_blocking=True)
                    out = renderer.forward_multi(vs, f_dev, fast_math=fast)
                    outs = [out['silhouette'], out['depth'], out['normal']]
                    ups = [g_dev['silhouette'], g_dev['depth'], g_dev['normal']]
                    with torch.no_grad():
                        loss = sum(torch.dot(o.reshape(-1), g.reshape(-1)) for o, g in zip(outs, ups))
                    torch.autograd.backward(outs, ups)
                    gv_host[slot].copy_(vs.grad, non_blocking=True)
                    loss_host[slot].copy_(loss.view(1), non_blocking=True)
                e2e_graphs.append(gr)
                v_static.append(vs)
                gv_host[slot].zero_()
                loss_host[slot].zero_()
                gr.replay()
                torch.cuda.synchronize()
                err = float((gv_host[slot] - ref_gv).norm() / ref_gv.norm())
                lerr = abs(float(loss_host[slot][0]) - ref_loss) / max(abs(ref_loss), 1e-30)
                assert err < 1e-5 and lerr < 1e-4, 'e2e graph differs from the eager step: %g %g' % (err, lerr)

            # the two slots are independent steps with their own static buffers: each replays on its own stream, so step
            # i + 1's host-to-device copy and kernels overlap step i's device-to-host copies
            e2e_streams = [torch.cuda.Stream(), torch.cuda.Stream()]
            for st_ in e2e_streams:
                st_.wait_stream(torch.cuda.current_stream())

            def step_e2e_graph():
                slot = e2e_state['i'] & 1
                e2e_state['i'] += 1
                e2e_consume(slot)
                with torch.cuda.stream(e2e_streams[slot]):
                    e2e_graphs[slot].replay()
                    done[slot] = torch.cuda.Event()
                    done[slot].record()
                e2e_consume(slot ^ 1)

            for _ in range(3):
                step_e2e_graph()
            e2e_drain()
            e2e_fn, e2e_launch = step_e2e_graph, 'cuda graph replay (H2D, step, D2H are graph nodes; the two slots on two streams)'
        except Exception as e:      # noqa: BLE001
            print('[bench] e2e graph capture failed (%s); timing the eager e2e step' % repr(e)[:300], file=sys.stderr)
            torch.cuda.synchronize()
            done[0] = done[1] = None

    def e2e_run():          # timed as one region: K pipelined steps, every result read on the host before it ends
        for _ in range(args.steps):
            e2e_fn()
        e2e_drain()
    consumed0 = e2e_state['consumed']
    ms_e2e = timed(e2e_run, 1)
    assert e2e_state['consumed'] - consumed0 == args.steps and math.isfinite(e2e_state['loss'])

    # per-call device time of the four C-ABI calls (events on the launching stream = torch's current stream)
    seg = segment_times(renderer, v_dev, f_dev, g_dev, fast, steps=max(5, min(args.steps, 20)))
    # per-kernel device time: the library brackets each of its launches with CUDA events on that same stream
    ksteps = max(5, min(args.steps, 20))
    sdn_b200.timing_enable(True)
    sdn_b200.timing_read()
    for _ in range(ksteps):
        step_resident()
    torch.cuda.synchronize()
    kern = {k: v[0] / ksteps for k, v in sdn_b200.timing_read().items()}
    sdn_b200.timing_enable(False)

    images = B * world * args.steps
    value = images / (ms * 1e-3)
    line = {
        'metric': 'images/sec fwd+bwd render 375x1242', 'value': round(value, 1), 'unit': 'images/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': round(ms / args.steps, 4),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'config2: B=%d/GPU, 375x1242 (750x2484 2xSSAA), 1984-triangle superellipsoid '
                               '(3968 with fill_back), VKITTI camera f=725, silhouette+depth+normal, N(0,1) grads'
                               % B,
                   'batch_per_gpu': B, 'math': 'fast' if fast else 'exact',
                   'launch': 'cuda graph replay' if graph is not None else 'eager', 'host_dispatch_ms_per_step': round(host_ms, 3),
                   'l2': 'inputs+outputs per step (%.0f MB) exceed the 126 MB L2' % (
                       B * sum(algorithmic_bytes()) / 1e6),
                   'sharding': 'by image, no data-path collective',
                   'scenes': 'per-rank seeds' if args.rank_seeds else 'the same 32 scenes on every rank',
                   'rank_ms_per_step': {'min': round(step_rank_ms[0], 4), 'median': round(step_rank_ms[len(step_rank_ms) // 2], 4),
                                        'max': round(step_rank_ms[-1], 4)}},
        'e2e': {'value': round(images / (ms_e2e * 1e-3), 1), 'unit': 'images/s',
                'h2d_bytes_per_step': int(verts.numel() * 4), 'd2h_bytes_per_step': int(verts.numel() * 4 + 4),
                'launch': e2e_launch,
                'pipeline': 'two independent steps in flight, each slot replaying on its own stream: every step copies its '
                            'vertices from pinned host memory, renders forward + backward, computes loss = <outputs, '
                            'upstream> and copies loss + vertex gradients back; each result is read on the host inside the '
                            'timed region.  The copies AND the kernel tails of one step overlap the other step, which is why '
                            'this number can exceed `value` (steps strictly one after another on one stream)'},
        'gpu_launches': int(launches),
        'clocks': sampler.summary(),
    }
    extra = None
    if not args.no_extra:
        del graph
        torch.cuda.empty_cache()
        extra = other_configs(renderer, dev, world, rank, fast, args)
    if rank == 0:
        if extra:
            line.update(extra)
        fwd_b, bwd_b = algorithmic_bytes()
        peak, how = measured_peaks()
        dom = max(kern, key=lambda k: kern[k])
        fwd_kernels = ('k_face_setup', 'k_fill_background', 'k_tile_prefix', 'k_raster_tiles')
        # the dominant kernel is charged the algorithmic bytes of the pass it belongs to: it is the
        # kernel that has to stream them; the other kernels of the pass only add time
        algo = (fwd_b if dom in fwd_kernels else bwd_b) * B
        achieved = algo / (kern[dom] * 1e-3) / 1e9
        line['roofline'] = {'bound': 'hbm', 'kernel': dom, 'achieved': round(achieved, 1), 'peak': peak,
                            'peak_source': how, 'unit': 'GB/s', 'frac': round(achieved / peak, 4),
                            'traffic': NCU_TRAFFIC.get(dom), 'algorithmic_bytes_per_launch': int(algo),
                            'launch_ms': round(kern[dom], 4),
                            'kernels_ms': {k: round(v, 4) for k, v in sorted(kern.items(), key=lambda kv: -kv[1])},
                            'segments_ms': {k: round(v, 4) for k, v in seg.items()},
                            'whole_step_frac': round(value / world * sum(algorithmic_bytes()) / 1e9 / peak, 4)}
        if world == 1 and not args.no_cpu:
            line['cpu_baseline'] = cpu_baseline()
        if world == 1 and not args.no_extra:
            line['reference_cuda'] = reference_cuda_arm()
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _time_steps(fn, steps, warm=3):
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


def other_configs(renderer, dev, world, rank, fast, args):
    """The other BASELINE configs, measured briefly in the same process so that the driver's JSON carries them:
    N = 1: config 2 in exact (IEEE-division) mode, config 2b (silhouette only: the reference's training path), config 2c
    (the reference's real 58k-145k-face templates at its own 384x384) and config 3 (full training step);
    N > 1: config 4 = config 3 per rank + ONE flat-bucket gradient all-reduce inside the timed step."""
    from sdn_b200 import workloads as wl
    from derender3d.bench_setup import measure_train, template_meshes
    from derender3d.models.renderer import Renderer
    out = {}
    if world > 1:
        out['config4'] = measure_train(dev, world, rank, B=16, S=384, steps=10, graph=not args.no_train_graph,
                                       native_encoder=not args.cudnn_encoder)
        return out
    B = args.batch
    verts, faces, grads = make_inputs(B, 0, dev)
    v = verts.to(dev).requires_grad_(True)
    f = faces.to(dev).int()
    g = {k: t.to(dev) for k, t in grads.items()}

    def step(want, fm):
        def fn():
            o = renderer.forward_multi(v, f, fast_math=fm, **want)
            keys = [k for k in ('silhouette', 'depth', 'normal') if want[k]]
            torch.autograd.backward([o[k] for k in keys], [g[k] for k in keys])
            v.grad = None
        return fn
    all3 = dict(silhouette=True, depth=True, normal=True)
    ms = _time_steps(step(all3, False), 3, warm=1)
    out['config2_exact_math'] = {'images_per_s': round(B / ms * 1e3, 1), 'ms_per_step': round(ms, 3),
                                 'what': 'same step with fast_math=0: IEEE divisions in the op order of rasterize.py:646-653 '
                                         '(one warp per (face, edge, axis) item), eager'}
    ms = _time_steps(step(dict(silhouette=True, depth=False, normal=False), fast), 20)
    fb, bb = algorithmic_bytes(c_out=1)
    out['config2b'] = {'workload': 'config 2 with the silhouette output only (the reference training path, '
                                   'derender3d/models/__init__.py:203-208), eager',
                       'images_per_s': round(B / ms * 1e3, 1), 'ms_per_step': round(ms, 3),
                       'hbm_roofline_frac': round(B / ms * 1e3 * (fb + bb) / 1e9 / measured_peaks()[0], 4)}
    # config 2c: the real templates (or 72.6k-triangle stand-ins) at the reference's own 384x384, B = 32, three outputs
    objs, desc = template_meshes()
    S, Bc = 384, 32
    nv_max = max(o.vertices.shape[0] for o in objs)
    nf_max = max(o.faces.shape[0] for o in objs)
    gen = torch.Generator().manual_seed(3)
    vb = torch.zeros(Bc, nv_max, 3)
    fbt = torch.zeros(Bc, nf_max, 3, dtype=torch.int32)
    for i in range(Bc):
        o = objs[i % len(objs)]
        yaw = float(torch.rand(1, generator=gen)) * 6.283
        c, s_ = math.cos(yaw), math.sin(yaw)
        R = torch.tensor([[c, 0, s_], [0, 1, 0], [-s_, 0, c]])
        size = torch.tensor([4.0, 1.6, 1.8])
        vv = (o.vertices * size) @ R.T + torch.tensor([0.0, 0.0, -(5.5 + 3.0 * float(torch.rand(1, generator=gen)))])
        vb[i, :vv.shape[0]] = vv
        vb[i, vv.shape[0]:] = vv[0]
        fbt[i, :o.faces.shape[0]] = o.faces
    vc = vb.to(dev).requires_grad_(True)
    fc = fbt.to(dev)
    gc = {'silhouette': torch.randn(Bc, 1, S, S, generator=gen).to(dev), 'depth': torch.randn(Bc, 1, S, S, generator=gen).to(dev),
          'normal': torch.randn(Bc, 3, S, S, generator=gen).to(dev)}
    rc = Renderer(image_size=S, viewing_angle=wl.vkitti_viewing_angle(S, 500.0))

    def step_c():
        o = rc.forward_multi(vc, fc, fast_math=fast)
        torch.autograd.backward([o['silhouette'], o['depth'], o['normal']], [gc['silhouette'], gc['depth'], gc['normal']])
        vc.grad = None
    ms = _time_steps(step_c, 10)
    cov = float((rc.forward_multi(vc, fc)['silhouette'] > 0).float().mean())
    out['config2c'] = {'workload': 'B=32, %s, 384x384 (768x768 2xSSAA), silhouette+depth+normal fwd+bwd, eager' % desc,
                       'images_per_s': round(Bc / ms * 1e3, 1), 'ms_per_step': round(ms, 3), 'coverage': round(cov, 3)}
    del vc, fc, gc, rc
    torch.cuda.empty_cache()
    # the encoder trunk (forward, dgrad, wgrad, BatchNorm) on this repo's tcgen05 kernels; the cuDNN bf16-autocast encoder
    # of round 1 is kept beside it as the comparator
    out['config3'] = measure_train(dev, 1, 0, B=16, S=384, steps=10, graph=not args.no_train_graph, native_encoder=True)
    try:
        out['config3_cudnn_encoder'] = measure_train(dev, 1, 0, B=16, S=384, steps=10, graph=not args.no_train_graph,
                                                     native_encoder=False)
    except Exception as e:      # noqa: BLE001
        out['config3_cudnn_encoder'] = {'failed': repr(e)[:200]}
        torch.cuda.synchronize()
    return out


def reference_cuda_arm(timeout=240):
    """The reference's own CUDA kernels re-hosted on this GPU (baseline/ref_cuda.py: rasterize.py's kernel strings through
    NVRTC, driven per object like derender3d/models/__init__.py:161-224), in a subprocess under a timeout (its visibility
    kernel spins on per-pixel locks).  Needs the git-ignored copy baseline/_ref staged by __graft_entry__.build()."""
    script = os.path.join(ROOT, 'baseline', 'ref_cuda.py')
    try:
        r = subprocess.run([sys.executable, script, '--objects', '16', '--check'], capture_output=True, text=True,
                           timeout=timeout)
        last = [ln for ln in r.stdout.strip().splitlines() if ln.startswith('{')]
        if r.returncode == 0 and last:
            return json.loads(last[-1])
        return {'unavailable': 'baseline/ref_cuda.py exited %d: %s' % (r.returncode, (r.stderr or '')[-200:])}
    except subprocess.TimeoutExpired:
        return {'unavailable': 'baseline/ref_cuda.py did not finish in %d s' % timeout}
    except Exception as e:      # noqa: BLE001
        return {'unavailable': repr(e)[:200]}


def segment_times(renderer, v_dev, f_dev, g_dev, fast, steps):
    """CUDA-event time of each C-ABI call (faces fwd, rasterise fwd, rasterise bwd, faces bwd)."""
    import sdn_b200
    r = renderer._native()
    r.eye, r.camera_mode, r.camera_direction, r.up = renderer._camera_args(v_dev)
    cam = r._camera(v_dev)
    acc = {'faces_forward': 0.0, 'rasterize_forward': 0.0, 'rasterize_backward': 0.0, 'faces_backward': 0.0}

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e
    for _ in range(steps):
        v = v_dev.detach().requires_grad_(True)
        e0 = ev()
        f, n = sdn_b200.FacesFunction.apply(v, f_dev, cam, True, True)
        e1 = ev()
        fd, nd = f.detach().requires_grad_(True), n.detach().requires_grad_(True)
        cfg = sdn_b200.RasterConfig((H, W), True, 0.1, 100, 1e-3, (0, 0, 0), True, True, True,
                                    split_eps=(1e-4, 1e-3), fast_math=fast)
        e2 = ev()
        rgb, alpha, dep = sdn_b200.RasterizeFunction.apply(fd, nd, cfg)
        e3 = ev()
        torch.autograd.backward([alpha, dep, rgb], [g_dev['silhouette'][:, 0], g_dev['depth'][:, 0], g_dev['normal']])
        e4 = ev()
        torch.autograd.backward([f, n], [fd.grad, nd.grad])
        e5 = ev()
        torch.cuda.synchronize()
        acc['faces_forward'] += e0.elapsed_time(e1)
        acc['rasterize_forward'] += e2.elapsed_time(e3)
        acc['rasterize_backward'] += e3.elapsed_time(e4)
        acc['faces_backward'] += e4.elapsed_time(e5)
    return {k: v / steps for k, v in acc.items()}


# ----------------------------------------------------------------------------------------------------------------
# CPU arms (the only place bench.py executes oracle/)
# ----------------------------------------------------------------------------------------------------------------
def cpu_reference_step(sample_images, seed=0):
    """What 3D-SDN does per object on its render path (derender3d/models/__init__.py:203-224): three
    rasterisations (silhouette, depth, normal) and their backward, here on the host with the oracle port.
    The reference has no CPU rasteriser (rasterize.py:890-894) and renders squares only, so its own kernels
    (oracle/_ref) cannot run the 375x1242 workload: kind = "port"."""
    from oracle import oracle as O
    from sdn_b200 import workloads as wl
    # forward visibility with the per-face scanline kernel 3D-SDN actually runs (NEURAL_RENDERER_UNSAFE=1,
    # rasterize.py:102-236), not the O(S^2 F) safe kernel that defines parity: the fair CPU baseline
    O.SCANLINE = True
    verts, faces, grads = make_inputs(sample_images, seed, None)
    fb = faces[None].expand(sample_images, -1, -1).contiguous()
    angle = wl.vkitti_viewing_angle()
    t0 = time.time()
    v = verts.clone().requires_grad_(True)
    loss = 0
    for rt, key in ((O.RenderType.Silhouette, 'silhouette'), (O.RenderType.Depth, 'depth'),
                    (O.RenderType.Normal, 'normal')):
        loss = loss + (O.derender_render(v, fb, (H, W), angle, rt) * grads[key]).sum()
    loss.backward()
    return time.time() - t0, O.num_threads()


CPU_SAMPLE_IMAGES = 8


def cpu_baseline(sample_images=8 * CPU_SAMPLE_IMAGES):
    """~10-30 s of host work: 64 images of config 2 in chunks of 8."""
    cpu_reference_step(1)                                   # warm the OpenMP pool / page in the library
    secs, threads, done = 0.0, 1, 0
    while done < sample_images:
        c = min(CPU_SAMPLE_IMAGES, sample_images - done)
        dt, threads = cpu_reference_step(c, seed=done)
        secs += dt
        done += c
    return {'value': round(sample_images / secs, 4), 'unit': 'images/s', 'cores': threads, 'kind': 'port',
            'host_cores': os.cpu_count(),
            'sample': '%d images of config2 in chunks of %d, fwd+bwd, 3 rasterisations per object as the reference issues '
                      'them (scanline forward kernel, K6/K7/K8 backward), OpenMP, %.1f s'
                      % (sample_images, CPU_SAMPLE_IMAGES, secs)}


REFERENCE_ARM_IMAGES = 600      # ~2 minutes of host work at 4-5 images/s


def run_reference(args):
    """K timed steps of a bounded sample: a step is n_img images of config 2 (n_img = 600 // K clamped to 1..8); the
    K * n_img images are pushed through the oracle port in chunks of 8 -- the chunk size at which its image-parallel
    loops (scanline forward, K7, K8) use the host cores best -- so the throughput does not depend on K."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps, warm = max(1, args.steps), max(0, args.warmup)
    n_img = max(1, min(CPU_SAMPLE_IMAGES, REFERENCE_ARM_IMAGES // steps))
    warm_images = min(warm * n_img, 2 * CPU_SAMPLE_IMAGES)
    done = 0
    while done < warm_images:
        c = min(CPU_SAMPLE_IMAGES, warm_images - done)
        cpu_reference_step(c, seed=1000 + done)
        done += c
    total, t, threads, done = steps * n_img, 0.0, 1, 0
    while done < total:
        c = min(CPU_SAMPLE_IMAGES, total - done)
        dt, threads = cpu_reference_step(c, seed=done)
        t += dt
        done += c
    value = total / t
    line = {'impl': 'reference', 'metric': 'images/sec fwd+bwd render 375x1242', 'value': round(value, 4),
            'unit': 'images/s', 'n_gpus': int(os.environ.get('WORLD_SIZE', '1')), 'steps': steps, 'warmup': warm,
            'ms_per_step': round(t / steps * 1e3, 1), 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'config2 sample: %d image(s) per step, 375x1242 2xSSAA, 1984-triangle superellipsoid, '
                                   'silhouette+depth+normal fwd+bwd on the host cores' % n_img},
            'cpu_baseline': {'value': round(value, 4), 'unit': 'images/s', 'cores': threads, 'kind': 'port',
                             'host_cores': os.cpu_count(),
                             'sample': '%d step(s) x %d image(s) (+%d warm-up images), run in chunks of %d; the reference '
                                       'has no CPU path and no rectangular raster, so this is the oracle port (OpenMP) of '
                                       'its kernels: scanline forward (rasterize.py:102-236), K6/K7/K8 backward'
                                       % (steps, n_img, warm_images, CPU_SAMPLE_IMAGES)},
            'e2e': {'value': round(value, 4), 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--exact', action='store_true', help='IEEE-division distance math in the backward (default: fast)')
    ap.add_argument('--no-graph', dest='graph', action='store_false', help='time the eager step, not its CUDA graph')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--no-extra', action='store_true', help='skip the other configs (2b, 2c, 3 / 4) measured after the headline')
    ap.add_argument('--cudnn-encoder', action='store_true', help='config 4: cuDNN bf16-autocast encoder instead of the native one')
    ap.add_argument('--no-train-graph', action='store_true', help='config 3 / 4: time the eager step, not its CUDA graph')
    ap.add_argument('--rank-seeds', action='store_true', help='N > 1: different synthetic scenes on every rank')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()

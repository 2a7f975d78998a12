#!/usr/bin/env python
"""Reference-shaped workload (SURVEY.md fact 4 / BASELINE config 2c): meshes of ShapeNet size (tens of thousands of
triangles) rendered the way 3D-SDN renders them -- 384x384 (768x768 internal), one object per image filling the crop.
The ShapeNet OBJ files are not redistributable, so a 72.6k-triangle superellipsoid stands in for them."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, '3d-sdn_b200')):
    sys.path.insert(0, p)
import torch  # noqa: E402
import sdn_b200  # noqa: E402
from sdn_b200 import workloads as wl  # noqa: E402
from derender3d.models.renderer import Renderer, RenderType  # noqa: E402


def main():
    import faulthandler
    faulthandler.dump_traceback_later(60, exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--lon', type=int, default=192)
    ap.add_argument('--lat', type=int, default=190)
    ap.add_argument('--size', type=int, default=384)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--multi', action='store_true', help='silhouette+depth+normal instead of silhouette only')
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    mesh = wl.superellipsoid_mesh(args.lon, args.lat, 0.6, 0.6)
    B, S = args.batch, args.size
    focal = 725.0
    # objects filling ~60% of the crop, like the ROI crops of derender3d/datasets.py
    verts, faces = wl.vkitti_scene_batch(B, seed=0, H=S, W=S, focal=focal * S / 256, z_range=(9.0, 12.0), mesh=mesh)
    verts[:, :, 0] -= verts[:, :, 0].mean(1, keepdim=True)
    verts[:, :, 1] -= verts[:, :, 1].mean(1, keepdim=True)
    angle = wl.vkitti_viewing_angle(S, focal * S / 256)
    r = Renderer(image_size=S, viewing_angle=angle)
    v = verts.to(dev).requires_grad_(True)
    f = faces.to(dev).int()
    gen = torch.Generator().manual_seed(1)
    gs = torch.randn(B, 1, S, S, generator=gen).to(dev)
    gd = torch.randn(B, 1, S, S, generator=gen).to(dev)
    gn = torch.randn(B, 3, S, S, generator=gen).to(dev)
    sdn_b200.DEFAULT_FAST_MATH = True

    def step():
        if args.multi:
            out = r.forward_multi(v, f)
            torch.autograd.backward([out['silhouette'], out['depth'], out['normal']], [gs, gd, gn])
        else:
            out = r(v, f, render_type=RenderType.Silhouette)
            torch.autograd.backward([out], [gs])
        v.grad = None
        return out

    out = step()
    cov = float(((out['silhouette'] if args.multi else out) > 0).float().mean())
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    sdn_b200.timing_enable(True)
    sdn_b200.timing_read()
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    kern = {k: round(t[0] / 5, 3) for k, t in sorted(sdn_b200.timing_read().items(), key=lambda kv: -kv[1][0])}
    sdn_b200.timing_enable(False)
    print(json.dumps({'workload': 'B=%d, %dx%d, %d triangles (%d with fill_back), %s' % (
        B, S, S, len(faces), 2 * len(faces), 'silhouette+depth+normal' if args.multi else 'silhouette only'),
        'coverage': round(cov, 3), 'ms_per_step': round(ms, 3), 'images_per_s': round(B / ms * 1e3, 1),
        'kernels_ms': kern}))


if __name__ == '__main__':
    main()

#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE's own kernel bodies (oracle/_ref, built by
oracle/build_ref.py from /root/reference/geometric/neural_renderer/rasterize.py) on seeded inputs.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
The .npz files are committed; this script is the provenance record.  Each file holds the inputs (NDC faces,
textures, upstream gradients, raster size, near/far/eps, flags) and every buffer the reference kernels produced
(face_index_map, weight_map, depth_map, rgb_map, alpha_map, sampling maps, grad_faces, grad_textures)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, '3d-sdn_b200'))
from oracle import oracle as O  # noqa: E402
from sdn_b200 import workloads as wl  # noqa: E402


def random_faces(rng, B, F, spread=0.9, size=0.5, zlo=0.5, zhi=3.0):
    c = rng.uniform(-spread, spread, (B, F, 1, 2))
    xy = c + rng.uniform(-size, size, (B, F, 3, 2))
    z = rng.uniform(zlo, zhi, (B, F, 3, 1))
    return np.concatenate([xy, z], axis=3).astype(np.float32)


def cube_faces(eye=(1.2, 1.0, -2.3), angle=30.):
    import torch
    v, f = wl.cube_mesh()
    V = torch.tensor(v)[None]
    Fc = O.fill_back_faces(torch.tensor(f)[None])
    faces = O.vertices_to_faces(O.perspective(O.look_at(V, list(eye)), angle), Fc)
    return faces.numpy()


def case(name, faces, S, ts, near, far, eps, flags, seed):
    rng = np.random.RandomState(seed)
    B, F = faces.shape[:2]
    rgb, alpha, depth = flags
    tex = rng.uniform(0, 1, (B, F, ts, ts, ts, 3)).astype(np.float32) if rgb else None
    st = O.raster_forward(faces, tex, S, S, near, far, eps, (0.1, 0.2, 0.3), rgb, alpha, depth, use_ref=True)
    g_rgb = rng.randn(B, S, S, 3).astype(np.float32) if rgb else None
    g_alpha = rng.randn(B, S, S).astype(np.float32) if alpha else None
    g_depth = rng.randn(B, S, S).astype(np.float32) if depth else None
    gf, gt = O.raster_backward(st, g_rgb, g_alpha, g_depth)
    out = dict(faces=faces, S=S, ts=ts, near=near, far=far, eps=eps, flags=np.array(flags, np.int32),
               background=np.array((0.1, 0.2, 0.3), np.float32),
               face_index_map=st.face_index_map, weight_map=st.weight_map, depth_map=st.depth_map,
               grad_faces=gf)
    if rgb:
        out.update(textures=tex, rgb_map=st.rgb_map, sampling_index_map=st.sampling_index_map,
                   sampling_weight_map=st.sampling_weight_map, grad_rgb_map=g_rgb, grad_textures=gt)
    if alpha:
        out.update(alpha_map=st.alpha_map, grad_alpha_map=g_alpha)
    if depth:
        out.update(face_inv_map=st.face_inv_map, grad_depth_map=g_depth)
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, 'covered', float((st.face_index_map >= 0).mean()))


if __name__ == '__main__':
    assert O.ref_lib() is not None, 'build oracle/_ref first: python oracle/build_ref.py'
    rng = np.random.RandomState(7)
    case('cube64_all', cube_faces(), 64, 2, 0.1, 100, 1e-3, (1, 1, 1), 1)
    case('cube32_alpha', cube_faces((0.3, -1.5, -2.0)), 32, 2, 0.1, 100, 1e-4, (0, 1, 0), 2)
    # RGB cases stay at B == 1: the reference's K2 drops the batch offset of the face pointer (rasterize.py:390)
    # and then indexes textures out of bounds for B > 1.
    case('rand48_rgb_alpha', random_faces(rng, 1, 40), 48, 4, 0.1, 100, 1e-3, (1, 1, 0), 3)
    case('rand36_b2_alpha_depth', random_faces(rng, 2, 30), 36, 2, 0.1, 100, 1e-4, (0, 1, 1), 6)
    case('rand40_depth', random_faces(rng, 1, 60, size=0.3), 40, 2, 0.1, 100, 1e-4, (0, 0, 1), 4)
    case('rand56_clip', random_faces(rng, 1, 50, zlo=0.05, zhi=2.0), 56, 2, 0.3, 1.5, 1e-4, (0, 1, 1), 5)

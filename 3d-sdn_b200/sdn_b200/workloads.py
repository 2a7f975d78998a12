"""Synthetic VKITTI-shaped workloads (SURVEY.md section 8d / BASELINE.md section 3).  Pure torch/numpy on the CPU; callers move
the tensors to the device.  Used by bench.py and the tests; no dataset or network access."""
import math

import numpy as np
import torch

VKITTI_FOCAL = 725.0        # derender3d/datasets.py:207-213
VKITTI_H, VKITTI_W = 375, 1242


def cube_mesh():
    """8 vertices (+-0.5)^3, 12 counter-clockwise (outward) triangles -- BASELINE config 1."""
    v = np.array([[x, y, z] for x in (-.5, .5) for y in (-.5, .5) for z in (-.5, .5)], np.float32)
    quads = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
    f = []
    for a, b, c, d in quads:
        f += [(a, b, c), (a, c, d)]
    return v, np.array(f, np.int32)


def superellipsoid_mesh(n_lon=32, n_lat=32, e1=0.5, e2=0.5):
    """Closed lat-long superellipsoid in [-0.5,0.5]^3: 2 + (n_lat-1)*n_lon vertices, 2*n_lon*(n_lat-1) triangles
    (994 / 1,984 at the defaults) -- the car-like stand-in of BASELINE config 2."""
    def spow(c, e):
        return np.sign(c) * np.abs(c) ** e
    verts = [(0.0, 0.5, 0.0)]
    for i in range(1, n_lat):
        phi = math.pi / 2 - math.pi * i / n_lat
        for j in range(n_lon):
            th = 2 * math.pi * j / n_lon
            verts.append((0.5 * spow(math.cos(phi), e1) * spow(math.cos(th), e2),
                          0.5 * spow(math.sin(phi), e1),
                          0.5 * spow(math.cos(phi), e1) * spow(math.sin(th), e2)))
    verts.append((0.0, -0.5, 0.0))
    last = len(verts) - 1
    ring = lambda i, j: 1 + (i - 1) * n_lon + (j % n_lon)
    faces = []
    for j in range(n_lon):
        faces.append((0, ring(1, j + 1), ring(1, j)))
        faces.append((last, ring(n_lat - 1, j), ring(n_lat - 1, j + 1)))
    for i in range(1, n_lat - 1):
        for j in range(n_lon):
            a, b, c, d = ring(i, j), ring(i, j + 1), ring(i + 1, j + 1), ring(i + 1, j)
            faces.append((a, b, c))
            faces.append((a, c, d))
    return np.asarray(verts, np.float32), np.asarray(faces, np.int32)


def vkitti_viewing_angle(H=VKITTI_H, focal=VKITTI_FOCAL):
    """Vertical half-angle in degrees, as derender3d/models/__init__.py:202 derives it from the focal length."""
    return math.degrees(math.atan(H / (2.0 * focal)))


def vkitti_scene_batch(B, seed=0, H=VKITTI_H, W=VKITTI_W, focal=VKITTI_FOCAL, z_range=(6.0, 40.0), mesh=None):
    """B posed objects in camera space (+x right, +y up, looking down -z), one per image.
    Returns vertices [B,nv,3] float32, faces [nf,3] int32 (shared topology)."""
    g = torch.Generator().manual_seed(seed)
    v0, f = superellipsoid_mesh() if mesh is None else mesh
    v0 = torch.from_numpy(v0)
    u = lambda lo, hi: lo + (hi - lo) * torch.rand(B, generator=g)
    size = torch.stack([u(3.5, 4.5), u(1.4, 1.8), u(1.6, 2.0)], dim=1)     # l, h, w (metres)
    yaw = u(-math.pi, math.pi)
    z = u(*z_range)
    uu, vv = u(0.0, W - 1.0), u(0.0, H - 1.0)
    cx = (uu - (W - 1) / 2.0) / focal * z
    cy = -(vv - (H - 1) / 2.0) / focal * z
    v = v0[None] * size[:, None, :]
    c, s = torch.cos(yaw)[:, None], torch.sin(yaw)[:, None]
    x = c * v[:, :, 0] + s * v[:, :, 2]
    zz = -s * v[:, :, 0] + c * v[:, :, 2]
    verts = torch.stack([x + cx[:, None], v[:, :, 1] + cy[:, None], zz - z[:, None]], dim=2)
    return verts.float().contiguous(), torch.from_numpy(f)

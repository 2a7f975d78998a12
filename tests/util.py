import glob
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = sorted(p for p in glob.glob(os.path.join(HERE, 'golden', '*.npz')) if not os.path.basename(p).startswith('derender3d_'))


def rel_l2(a, b):
    a = torch.as_tensor(np.asarray(a.detach().cpu() if isinstance(a, torch.Tensor) else a)).double()
    b = torch.as_tensor(np.asarray(b.detach().cpu() if isinstance(b, torch.Tensor) else b)).double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def random_faces(rng, B, F, spread=0.9, size=0.5, zlo=0.5, zhi=3.0):
    """Random NDC triangles [B,F,3,3] (about half are back-facing)."""
    c = rng.uniform(-spread, spread, (B, F, 1, 2))
    xy = c + rng.uniform(-size, size, (B, F, 3, 2))
    z = rng.uniform(zlo, zhi, (B, F, 3, 1))
    return np.concatenate([xy, z], axis=3).astype(np.float32)


def load_golden(path):
    d = dict(np.load(path))
    d['name'] = os.path.basename(path)[:-4]
    return d

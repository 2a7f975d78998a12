"""Small torch helpers shared by look / look_at / lighting (host-side glue, a few [B,3] tensors)."""
import numpy as np
import torch

from .cross import cross


def normalize(x, eps=1e-5):
    """chainer.functions.normalize as the reference uses it (look.py:29-31): x / (||x||_2 + eps) over axis 1."""
    sq = x * x
    n2 = (sq[:, 0] + sq[:, 1]) + sq[:, 2]
    return x / (torch.sqrt(n2) + eps).unsqueeze(1)


def as_tensor(x, like):
    if isinstance(x, torch.Tensor):
        return x.to(device=like.device, dtype=torch.float32)
    return torch.as_tensor(np.asarray(x, dtype=np.float32), device=like.device)


def camera_rotation(z_dir, up):
    """Rows x_axis, y_axis, z_axis of the view rotation (look.py:28-34, look_at.py:29-35) -> [bs,3,3]."""
    z_axis = normalize(z_dir)
    x_axis = normalize(cross(up, z_axis))
    y_axis = normalize(cross(z_axis, x_axis))
    return torch.stack((x_axis, y_axis, z_axis), dim=1)


def rotate(vertices, eye, r):
    """(v - eye) @ r^T with the 3-term sums spelled out left to right (look.py:40-43)."""
    v = vertices - eye.unsqueeze(1)
    cols = []
    for i in range(3):
        s = v[:, :, 0] * r[:, i, 0:1] + v[:, :, 1] * r[:, i, 1:2]
        cols.append(s + v[:, :, 2] * r[:, i, 2:3])
    return torch.stack(cols, dim=2)

import torch

from ._ops import as_tensor, normalize
from .cross import cross


def lighting(faces, textures, intensity_ambient=0.5, intensity_directional=0.5, color_ambient=(1, 1, 1),
             color_directional=(1, 1, 1), direction=(0, 1, 0)):
    """Ambient + Lambert light per face multiplied into textures [bs,nf,ts,ts,ts,3] (reference: lighting.py:8-52)."""
    bs, nf = faces.shape[:2]
    color_ambient = as_tensor(color_ambient, faces)
    color_directional = as_tensor(color_directional, faces)
    direction = as_tensor(direction, faces)
    color_ambient, color_directional, direction = (
        t.unsqueeze(0).expand(bs, 3) if t.dim() == 1 else t for t in (color_ambient, color_directional, direction))
    light = torch.zeros(bs, nf, 3, device=faces.device)
    if intensity_ambient != 0:
        light = light + intensity_ambient * color_ambient.unsqueeze(1)
    if intensity_directional != 0:
        f = faces.reshape(bs * nf, 3, 3)
        normals = normalize(cross(f[:, 0] - f[:, 1], f[:, 2] - f[:, 1])).reshape(bs, nf, 3)
        d = direction.unsqueeze(1)
        cos = normals[:, :, 0] * d[:, :, 0] + normals[:, :, 1] * d[:, :, 1]
        cos = torch.relu(cos + normals[:, :, 2] * d[:, :, 2])
        light = light + intensity_directional * (color_directional.unsqueeze(1) * cos.unsqueeze(2))
    return textures * light[:, :, None, None, None, :]

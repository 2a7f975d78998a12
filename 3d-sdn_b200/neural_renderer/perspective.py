import math

import numpy as np
import torch


def perspective_width(angle):
    """tan(angle / 180. * 3.1416) evaluated in float32 like perspective.py:10-13 (note the literal 3.1416).
    The tangent is float32(tan(float64(x))) so that host code everywhere produces the same bits."""
    a = np.float32(angle)
    a = np.float32(a / np.float32(180.))
    a = np.float32(a * np.float32(3.1416))
    return np.float32(math.tan(float(a)))


def perspective(vertices, angle=30., aspect=None):
    """Perspective division (reference: perspective.py:5-19): x/z/tan, y/z/tan, z.
    `aspect` = W/H is this repo's rectangular-raster extension: `angle` is then the vertical half-angle and
    the horizontal width is tan*aspect.  aspect=None (or 1) is the reference."""
    assert vertices.dim() == 3
    if isinstance(angle, torch.Tensor):
        angle = float(angle)
    width_y = perspective_width(angle)
    width_x = width_y if aspect is None else np.float32(width_y * np.float32(aspect))
    z = vertices[:, :, 2]
    x = vertices[:, :, 0] / z / float(width_x)
    y = vertices[:, :, 1] / z / float(width_y)
    return torch.stack((x, y, z), dim=2)

"""Camera position on a sphere around the origin (reference: neural_renderer/get_points_from_angles.py:6-24)."""
import math

import torch


def _on_sphere(radius, elevation, azimuth, cos, sin):
    """(x, y, z) with y up, azimuth measured from -z: the reference's convention."""
    ring = radius * cos(elevation)
    return ring * sin(azimuth), radius * sin(elevation), -ring * cos(azimuth)


def get_points_from_angles(distance, elevation, azimuth, degrees=True):
    """Python numbers in -> a tuple of floats; array-likes in -> a float32 tensor [..., 3]."""
    if isinstance(distance, (float, int)):
        if degrees:
            elevation, azimuth = math.radians(elevation), math.radians(azimuth)
        return _on_sphere(distance, elevation, azimuth, math.cos, math.sin)
    radius, el, az = (torch.as_tensor(t, dtype=torch.float32) for t in (distance, elevation, azimuth))
    if degrees:
        el, az = torch.deg2rad(el), torch.deg2rad(az)
    return torch.stack(_on_sphere(radius, el, az, torch.cos, torch.sin), dim=-1)

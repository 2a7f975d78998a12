import math

import torch


def get_points_from_angles(distance, elevation, azimuth, degrees=True):
    """Camera position on a sphere (reference: get_points_from_angles.py:6-24)."""
    if isinstance(distance, (float, int)):
        if degrees:
            elevation, azimuth = math.radians(elevation), math.radians(azimuth)
        return (distance * math.cos(elevation) * math.sin(azimuth),
                distance * math.sin(elevation),
                -distance * math.cos(elevation) * math.cos(azimuth))
    distance, elevation, azimuth = (torch.as_tensor(t, dtype=torch.float32) for t in (distance, elevation, azimuth))
    if degrees:
        elevation, azimuth = torch.deg2rad(elevation), torch.deg2rad(azimuth)
    return torch.stack((distance * torch.cos(elevation) * torch.sin(azimuth),
                        distance * torch.sin(elevation),
                        -distance * torch.cos(elevation) * torch.cos(azimuth)), dim=-1)

import torch


def cross(a, b):
    """Row-wise 3-vector cross product of two [n,3] tensors (reference: cross.py:25-35; the backward rule of
    cross.py:53-57 falls out of autograd).  Written component-wise so no FMA contraction can enter."""
    assert a.dim() == 2 and b.dim() == 2 and a.shape[1] == 3 and b.shape[1] == 3
    a, b = torch.broadcast_tensors(a, b)        # [1,3] against [n,3] is accepted (the reference asserts equal n)
    a0, a1, a2 = a.unbind(1)
    b0, b1, b2 = b.unbind(1)
    return torch.stack((a1 * b2 - a2 * b1, a2 * b0 - a0 * b2, a0 * b1 - a1 * b0), dim=1)

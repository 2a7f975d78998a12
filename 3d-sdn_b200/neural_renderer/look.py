from ._ops import as_tensor, camera_rotation, rotate


def look(vertices, eye, direction=None, up=None):
    """"Look" transformation of vertices [bs,nv,3] (reference: look.py:7-45): camera at `eye` looking along
    `direction` with `up`; eye/direction/up may be [3] or [bs,3]."""
    assert vertices.dim() == 3
    bs = vertices.shape[0]
    direction = as_tensor([0, 0, 1] if direction is None else direction, vertices)
    up = as_tensor([0, 1, 0] if up is None else up, vertices)
    eye = as_tensor(eye, vertices)
    eye, direction, up = (t.unsqueeze(0) if t.dim() == 1 else t for t in (eye, direction, up))
    r = camera_rotation(direction, up)
    return rotate(vertices, eye.expand(bs, 3), r.expand(bs, 3, 3))

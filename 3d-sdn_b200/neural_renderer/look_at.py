from ._ops import as_tensor, camera_rotation, rotate


def look_at(vertices, eye, at=None, up=None):
    """"Look at" transformation of vertices [bs,nv,3] (reference: look_at.py:7-46): camera at `eye` aimed at `at`."""
    assert vertices.dim() == 3
    bs = vertices.shape[0]
    at = as_tensor([0, 0, 0] if at is None else at, vertices)
    up = as_tensor([0, 1, 0] if up is None else up, vertices)
    eye = as_tensor(eye, vertices)
    eye, at, up = (t.unsqueeze(0).expand(bs, 3) if t.dim() == 1 else t for t in (eye, at, up))
    r = camera_rotation(at - eye, up)
    return rotate(vertices, eye, r)

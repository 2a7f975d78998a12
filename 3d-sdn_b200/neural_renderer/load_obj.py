import numpy as np


def load_obj(filename_obj, normalization=True, filename_texture=None, texture_size=4):
    """Wavefront .obj reader: `v x y z` and `f a b c ...` (fan-triangulated, vertex index before any '/').
    Reference: load_obj.py:95-141.  Returns (vertices float32 [nv,3], faces int32 [nf,3]) as numpy arrays, like the
    reference.  Texture baking (load_obj.py:10-92, never used by 3D-SDN) is out of scope."""
    if filename_texture is not None:
        raise NotImplementedError('texture baking (reference load_obj.py:10-92) is outside the B200 hot path')
    verts, tris = [], []
    with open(filename_obj) as fh:
        for line in fh:
            tok = line.split()
            if not tok:
                continue
            if tok[0] == 'v':
                verts.append((float(tok[1]), float(tok[2]), float(tok[3])))
            elif tok[0] == 'f':
                idx = [int(t.split('/')[0]) for t in tok[1:]]
                for k in range(1, len(idx) - 1):
                    tris.append((idx[0], idx[k], idx[k + 1]))
    vertices = np.asarray(verts, dtype=np.float32).reshape(-1, 3)
    faces = np.asarray(tris, dtype=np.int32).reshape(-1, 3) - 1
    if normalization:       # into a cube of side 2 about the origin (load_obj.py:132-136)
        vertices -= vertices.min(0)[None, :]
        vertices /= np.abs(vertices).max()
        vertices *= 2
        vertices -= vertices.max(0)[None, :] / 2
    return vertices, faces

import os


def save_obj(filename, vertices, faces):
    """Write vertices [nv,3] / faces [nf,3] as a Wavefront .obj (reference: save_obj.py:4-18)."""
    assert vertices.ndim == 2 and faces.ndim == 2
    lines = ['# %s' % os.path.basename(filename), '#', '', 'g mesh', '']
    lines += ['v  %.4f %.4f %.4f' % (v[0], v[1], v[2]) for v in vertices]
    lines.append('')
    lines += ['f  %d %d %d' % (f[0] + 1, f[1] + 1, f[2] + 1) for f in faces]
    with open(filename, 'w') as fh:
        fh.write('\n'.join(lines) + '\n')

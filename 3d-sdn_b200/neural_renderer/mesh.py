"""Learnable mesh (reference: neural_renderer/mesh.py:7-38, a chainer.Link there): vertex positions and one texture
cube per face are torch Parameters, the face table is a buffer that follows the module across devices."""
import torch

from .load_obj import load_obj


def _tile(t, n):
    """[...] -> [n, ...] without copying."""
    return t.unsqueeze(0).expand(n, *t.shape)


class Mesh(torch.nn.Module):
    def __init__(self, filename_obj, texture_size=4, normalization=True):
        super(Mesh, self).__init__()
        v, f = load_obj(filename_obj, normalization)
        self.num_vertices, self.num_faces, self.texture_size = len(v), len(f), texture_size
        self.vertices = torch.nn.Parameter(torch.from_numpy(v))
        self.register_buffer('faces', torch.from_numpy(f))
        cube = (self.num_faces,) + (texture_size,) * 3 + (3,)
        self.textures = torch.nn.Parameter(0.05 * torch.randn(cube))        # chainer.initializers.Normal(): std 0.05

    def get_batch(self, batch_size):
        """-> vertices [bs,nv,3], faces [bs,nf,3], textures [bs,nf,ts,ts,ts,3] squashed to (0,1)."""
        return _tile(self.vertices, batch_size), _tile(self.faces, batch_size), \
            torch.sigmoid(_tile(self.textures, batch_size))

    def set_lr(self, lr_vertices, lr_textures):
        """Per-parameter learning rates (mesh.py:36-38), as torch optimizer parameter groups."""
        return [{'params': [self.vertices], 'lr': lr_vertices}, {'params': [self.textures], 'lr': lr_textures}]

import torch

from .load_obj import load_obj


class Mesh(torch.nn.Module):
    """Learnable mesh: vertices + per-face texture cubes (reference: mesh.py:7-38, a chainer.Link there)."""

    def __init__(self, filename_obj, texture_size=4, normalization=True):
        super(Mesh, self).__init__()
        vertices, faces = load_obj(filename_obj, normalization)
        self.vertices = torch.nn.Parameter(torch.from_numpy(vertices))
        self.register_buffer('faces', torch.from_numpy(faces))
        self.num_vertices, self.num_faces = self.vertices.shape[0], self.faces.shape[0]
        self.textures = torch.nn.Parameter(
            0.05 * torch.randn(self.num_faces, texture_size, texture_size, texture_size, 3))
        self.texture_size = texture_size

    def get_batch(self, batch_size):
        vertices = self.vertices.unsqueeze(0).expand(batch_size, *self.vertices.shape)
        faces = self.faces.unsqueeze(0).expand(batch_size, *self.faces.shape)
        textures = torch.sigmoid(self.textures.unsqueeze(0).expand(batch_size, *self.textures.shape))
        return vertices, faces, textures

"""Mesh deformation and camera-space transform of the geometric branch (reference:
derender3d/models/transforms.py:10-158), batched and device-resident.

Differences from the reference, none of them numerical:
  * FFD keeps its Bernstein tensor as a registered buffer (moved once with the module) instead of re-uploading
    ~4.8 MB on every call (transforms.py:97), and accepts a batch of coefficient vectors [N,192] -> [N,nv,3];
  * the deformation is evaluated as one [nv,64] x [64,3] product per object instead of a broadcast multiply + sum.
"""
import numpy as np
import scipy.special
import torch
from torch.nn.modules import Module


class _Constraint(object):
    """A linear constraint on the lattice offsets: `symmetry` mirrors the lattice across one axis (the mirrored
    component changes sign), `homogeneity` ties the listed lattice planes of one axis together.  Exposed as
    FFD.Constraint with the reference's factory names (transforms.py:11-36)."""

    class Type:
        symmetry, homogeneity = 0, 1

    class Axis:
        x, y, z = 0, 1, 2

    def __init__(self, type, axis=None, index=None):
        self.type, self.axis = type, axis
        if index is not None:
            self.index = index

    @classmethod
    def symmetry(cls, axis):
        return cls(cls.Type.symmetry, axis=axis)

    @classmethod
    def homogeneity(cls, axis, index):
        return cls(cls.Type.homogeneity, axis=axis, index=index)


class FFD(Module):
    """Free-form deformation of a fixed vertex set by a 4x4x4 Bezier lattice (transforms.py:10-99)."""
    Constraint = _Constraint

    def __init__(self, vertices, num_grids=4, constraints=None):
        super(FFD, self).__init__()
        assert num_grids % 2 == 0
        self.num_grids = num_grids
        self.constraints = constraints or []
        vertices = torch.as_tensor(vertices, dtype=torch.float32)
        grids = np.arange(num_grids)
        binoms = torch.tensor(scipy.special.binom(num_grids - 1, grids), dtype=torch.float32)
        g1 = torch.tensor(grids, dtype=torch.float32)
        # Bernstein basis per axis: C(n,i) (0.5+v)^i (0.5-v)^(n-i)        (transforms.py:57-61)
        coeff = (binoms * torch.pow((0.5 + vertices).unsqueeze(2), g1) *
                 torch.pow((0.5 - vertices).unsqueeze(2), num_grids - 1 - g1))
        cx, cy, cz = coeff.unbind(dim=1)
        B = torch.einsum('ni,nj,nk->nijk', cx, cy, cz)                      # [nv,4,4,4]   (transforms.py:63)
        g3 = torch.tensor(np.stack(np.meshgrid(grids, grids, grids, indexing='ij')), dtype=torch.float32)
        self.register_buffer('B', B.unsqueeze(1), persistent=False)          # [nv,1,4,4,4] like the reference
        self.register_buffer('P0', g3 / (num_grids - 1) - 0.5, persistent=False)   # [3,4,4,4]   (transforms.py:66)

    def constrain(self, dP):
        """dP [...,3,n,n,n] -> constrained offsets (transforms.py:71-95)."""
        n = self.num_grids
        for c in self.constraints:
            dim = dP.dim() - 3 + c.axis
            if c.type == FFD.Constraint.Type.symmetry:
                sign = dP.new_tensor([1.0, 1.0, -1.0]).view(3, 1, 1, 1)
                dP = (dP + dP.flip(dim) * sign) / 2
            elif c.type == FFD.Constraint.Type.homogeneity:
                slices = list(dP.unbind(dim))
                mean = sum(slices[i] for i in c.index) / len(c.index)
                comp = dP.dim() - 4                      # dimension of the xyz component in a slice
                for i in c.index:
                    keep = slices[i].select(comp, c.axis)
                    merged = mean.clone()
                    merged.select(comp, c.axis).copy_(keep)
                    slices[i] = merged
                dP = torch.stack(slices, dim=dim)
        return dP

    def constraint_matrix(self):
        """The constraint step is linear in the coefficients: constrain(x) == x @ C for this [3n^3, 3n^3] matrix
        (entries are 0, +-1/2, +-1/4: products exact, only the order of a 2-4 term sum differs)."""
        m = 3 * self.num_grids ** 3
        eye = torch.eye(m, dtype=torch.float32)
        n = self.num_grids
        return self.constrain(eye.reshape(m, 3, n, n, n)).reshape(m, m)

    def forward(self, ffd_coeff):
        """ffd_coeff [3*n^3] -> [nv,3]   or   [N,3*n^3] -> [N,nv,3]."""
        n = self.num_grids
        single = ffd_coeff.dim() == 1
        with torch.autocast(ffd_coeff.device.type, enabled=False):     # vertex positions stay float32 under autocast
            dP = ffd_coeff.float().reshape(-1, 3, n, n, n)
            P = (self.P0.to(dP.device) + self.constrain(dP)).reshape(-1, 3, n * n * n)        # [N,3,64]
            V = torch.matmul(self.B.to(dP.device).reshape(-1, n * n * n), P.transpose(1, 2))  # [N,nv,3]
        return V[0] if single else V


def quaternion_matrix(q):
    """[N,4] (a,b,c,d) -> [N,3,3].  Products of two components are formed once; 2*(b*c) == (2*b)*c exactly, so the
    entries equal the reference's `2 * b * c - 2 * a * d` style expressions (transforms.py:114-124) bit for bit."""
    a, b, c, d = q.unbind(dim=1)
    aa, bb, cc, dd = a * a, b * b, c * c, d * d
    ab, ac, ad, bc, bd, cd = a * b, a * c, a * d, b * c, b * d, c * d
    rows = ((aa + bb - cc - dd, 2 * bc - 2 * ad, 2 * bd + 2 * ac),
            (2 * bc + 2 * ad, aa - bb + cc - dd, 2 * cd - 2 * ab),
            (2 * bd - 2 * ac, 2 * cd + 2 * ab, aa - bb - cc + dd))
    return torch.stack([torch.stack(r, dim=1) for r in rows], dim=1)


class PerspectiveTransform(Module):
    """Scale, rotate (quaternion), translate, shear along the principal ray, zoom (transforms.py:102-158).
    Training mode: `zooms` given.  Test mode: `zoom_tos` given, the zoom is derived from the vertices and returned."""

    def forward(self, vertices, scales=None, rotations=None, translations=None, perspective_translations=None,
                zooms=None, zoom_tos=None):
        v = vertices
        if scales is not None:
            v = v * scales[:, None, :]
        if rotations is not None:
            v = torch.matmul(v, quaternion_matrix(rotations).transpose(1, 2))
        shift = None if translations is None else translations[:, None, :]
        if shift is not None:
            v = v + shift
        ray = shift if perspective_translations is None else perspective_translations[:, None, :]
        x, y, z = v.unbind(dim=2)
        rx, ry, rz = ray.unbind(dim=2)
        x, y = x - rx / rz * z, y - ry / rz * z                      # shear: the principal ray becomes the z axis
        derived = zoom_tos is not None
        if derived:                                                   # the tightest |z| / max(|x|,|y|) over the mesh
            fit = (z.abs() / torch.max(x.abs(), y.abs())).min(dim=1, keepdim=True)[0]
            zooms = fit * zoom_tos
        out = torch.stack((x, y, z / zooms), dim=2)
        return (out, zooms) if derived else out

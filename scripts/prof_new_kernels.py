"""Driver for the ncu capture of the round's new kernels: k_linear_tc at a GEMM-bound size and the fused pre-raster
stage at BASELINE config-3 size (16 objects x 36k-vertex templates)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, '3d-sdn_b200'))
import sdn_b200 as S
from sdn_b200 import workloads as wl
from derender3d.models.transforms import FFD

dev = torch.device('cuda:0')
torch.manual_seed(0)
x = torch.randn(65536, 256, device=dev); w = torch.randn(1552, 256, device=dev); b = torch.randn(1552, device=dev)
for _ in range(2):
    S.linear_tc(x, w, b, True)
cons = [FFD.Constraint.symmetry(axis=FFD.Constraint.Axis.z), FFD.Constraint.homogeneity(axis=FFD.Constraint.Axis.y, index=[0, 1])]
bases, faces = [], []
for k in range(2):
    v, f = wl.superellipsoid_mesh(192, 190, 0.5 + 0.04 * k, 0.6)
    v = v / np.ptp(v, axis=0) * 0.999
    bases.append(FFD(torch.tensor(v), constraints=cons).B)
    faces.append(torch.tensor(f))
bank = S.MeshBank(bases, faces, dev)
N = 16
g = torch.Generator().manual_seed(0)
classes = torch.randint(0, 2, (N,), generator=g).to(dev)
ctrl = torch.randn(N, 3, 64, generator=g).to(dev).requires_grad_(True)
sc = torch.rand(N, 3, generator=g).add(0.5).to(dev)
rot = torch.tensor([1.0, 0, 0, 0]).expand(N, 4).contiguous().to(dev)
tr = torch.tensor([0.0, 0, -10.0]).expand(N, 3).contiguous().to(dev)
zo = torch.ones(N, 1, device=dev)
for _ in range(2):
    v = S.object_vertices(bank, classes, ctrl, sc, rot, tr, tr, zooms=zo)
    v.sum().backward()
torch.cuda.synchronize()
print('ok', v.shape)

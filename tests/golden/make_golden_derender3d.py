#!/usr/bin/env python
"""Golden vectors for the host-side rows of the hot path (a2, a3, a4), produced by IMPORTING THE
REFERENCE'S OWN PyTorch code in the build container:

  /root/reference/geometric/derender3d/models/transforms.py   (FFD, PerspectiveTransform)      -- imported as is
  /root/reference/geometric/derender3d/models/__init__.py     (Derenderer3d.render pose algebra) -- imported with
      `neural_renderer` / `derender3d.models.renderer` / torchvision's download stubbed out (they are not on this
      path: the stub renderer records the vertices it is handed and returns zeros).

`Tensor.cuda()` is patched to a no-op so the code runs on the CPU.  Output: tests/golden/derender3d_*.npz.
Run:  python tests/golden/make_golden_derender3d.py   (needs /root/reference)"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference/geometric'
sys.path.insert(0, os.path.join(ROOT, '3d-sdn_b200'))

torch.Tensor.cuda = lambda self, *a, **k: self
torch.cuda.current_device = lambda: 0


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def golden_transforms():
    T = load(os.path.join(REF, 'derender3d/models/transforms.py'), 'ref_transforms')
    g = torch.Generator().manual_seed(0)
    verts = torch.rand(200, 3, generator=g) - 0.5
    cons = [T.FFD.Constraint.symmetry(axis=T.FFD.Constraint.Axis.z),
            T.FFD.Constraint.homogeneity(axis=T.FFD.Constraint.Axis.y, index=[0, 1])]
    ffd = T.FFD(verts, constraints=cons)
    coeffs = 0.2 * torch.randn(5, 192, generator=g)
    ffd_out = torch.stack([ffd(c) for c in coeffs])
    ffd_free = torch.stack([T.FFD(verts, constraints=[])(c) for c in coeffs])
    pt = T.PerspectiveTransform()
    n = 5
    scales = torch.exp(0.3 * torch.randn(n, 3, generator=g))
    th = torch.rand(n, generator=g) * 6.28
    rot = torch.stack([torch.cos(th / 2), torch.zeros(n), torch.sin(th / 2), torch.zeros(n)], dim=1)
    trans = torch.stack([torch.randn(n, generator=g), 0.3 * torch.randn(n, generator=g),
                         -(5 + 10 * torch.rand(n, generator=g))], dim=1)
    ptrans = trans + 0.1 * torch.randn(n, 3, generator=g)
    zooms = 0.5 + torch.rand(n, 1, generator=g)
    zoom_tos = 0.1 + 0.3 * torch.rand(n, 1, generator=g)
    train = pt(ffd_out, scales=scales, rotations=rot, translations=trans, perspective_translations=ptrans, zooms=zooms)
    test_v, test_z = pt(ffd_out, scales=scales, rotations=rot, translations=trans, perspective_translations=trans,
                        zoom_tos=zoom_tos)
    np.savez_compressed(os.path.join(HERE, 'derender3d_transforms.npz'), vertices=verts.numpy(),
                        coeffs=coeffs.numpy(), ffd_out=ffd_out.numpy(), ffd_free=ffd_free.numpy(),
                        scales=scales.numpy(), rotations=rot.numpy(), translations=trans.numpy(),
                        perspective_translations=ptrans.numpy(), zooms=zooms.numpy(), zoom_tos=zoom_tos.numpy(),
                        train_out=train.numpy(), test_out=test_v.numpy(), test_zooms=test_z.numpy())
    print('transforms ok', ffd_out.shape, train.shape)


def golden_pose():
    import neural_renderer as my_nr            # only load_obj is used (pure numpy)
    calls = []

    class StubRenderer(torch.nn.Module):
        def __init__(self, image_size=256, viewing_angle=30):
            super().__init__()
            self.image_size, self.viewing_angle = image_size, viewing_angle

        def forward(self, vertices, faces, textures=None, render_type=None):
            calls.append((vertices.detach().clone(), float(self.viewing_angle), int(render_type)))
            c = 3 if render_type == 3 else 1
            return torch.zeros(len(vertices), c, self.image_size, self.image_size)

    stub_nr = types.ModuleType('neural_renderer')
    stub_nr.load_obj = my_nr.load_obj
    sys.modules['neural_renderer'] = stub_nr
    pkg = types.ModuleType('derender3d')
    pkg.__path__ = [os.path.join(REF, 'derender3d')]
    sys.modules['derender3d'] = pkg
    load(os.path.join(REF, 'derender3d/__init__.py'), 'derender3d')
    sys.modules['derender3d'].__path__ = [os.path.join(REF, 'derender3d')]
    stub_r = types.ModuleType('derender3d.models.renderer')

    class RenderType:
        RGB, Silhouette, Depth, Normal = 0, 1, 2, 3
    stub_r.RenderType, stub_r.Renderer = RenderType, StubRenderer
    sys.modules['derender3d.models.renderer'] = stub_r
    import torchvision
    real = torchvision.models.resnet18
    torchvision.models.resnet18 = lambda pretrained=False, **kw: real(weights=None)
    os.environ['SHAPENET_ROOT_DIR'] = os.path.join(REF, 'assets')
    mpkg = types.ModuleType('derender3d.models')
    mpkg.__path__ = [os.path.join(REF, 'derender3d/models')]
    sys.modules['derender3d.models'] = mpkg
    load(os.path.join(REF, 'derender3d/models/derenderer.py'), 'derender3d.models.derenderer')
    load(os.path.join(REF, 'derender3d/models/transforms.py'), 'derender3d.models.transforms')
    M = load(os.path.join(REF, 'derender3d/models/__init__.py'), 'derender3d.models')
    # two template meshes are enough to pin the algebra; keep the fixture small
    M.Derenderer3d.__init__.__globals__['ShapenetObj'] = M.ShapenetObj
    model = M.Derenderer3d(mode=3 | 12, image_size=256, render_size=64)      # extend: geometry|reproject|normal|depth
    g = torch.Generator().manual_seed(1)
    B = 6
    c = 0.4 * (torch.rand(B, 2, generator=g) - 0.5)
    d = 0.1 + 0.3 * torch.rand(B, 2, generator=g)
    roi = torch.cat([c - d / 2, c + d / 2], dim=1)
    theta = torch.randn(B, 2, generator=g)
    probs = torch.softmax(3 * torch.randn(B, 8, generator=g), dim=1)
    blob = {
        '_mroi_norms': torch.stack([roi[:, 2] + roi[:, 0], roi[:, 3] + roi[:, 1]], dim=1) / 2.0,
        '_droi_norms': torch.stack([roi[:, 2] - roi[:, 0], roi[:, 3] - roi[:, 1]], dim=1),
        '_focals': torch.full((B, 1), 725.0),
        '_theta_deltas': theta / theta.norm(dim=1, keepdim=True),
        '_translation2ds': 0.1 * torch.randn(B, 2, generator=g),
        '_log_scales': 0.2 * torch.randn(B, 3, generator=g) + torch.tensor([1.3, 0.4, 0.5]),
        '_log_depths': 0.3 * torch.randn(B, 1, generator=g) + 3.0,
        '_class_probs': probs,
        '_ffd_coeffs': 0.05 * torch.randn(B, 8, 192, generator=g),
    }
    out = {'roi_norms': roi.numpy()}
    out.update({'in' + k: v.numpy() for k, v in blob.items()})
    for mode_name, training in (('test', False), ('train', True)):
        calls.clear()
        model.train(training)
        model._force_no_sample = True
        with torch.no_grad():
            res = model.render(blob)
        for k, v in res.items():
            if isinstance(v, torch.Tensor) and k not in ('_masks', '_normals', '_depth_maps'):
                out['%s%s' % (mode_name, k)] = v.numpy()
        sil = [c for c in calls if c[2] == 1]
        out[mode_name + '_angles'] = np.array([c[1] for c in sil], np.float64)
        out[mode_name + '_vertex_sample'] = np.stack([c[0][0, ::97][:64].numpy() for c in sil])
        out[mode_name + '_nv'] = np.array([c[0].shape[1] for c in sil], np.int64)
    np.savez_compressed(os.path.join(HERE, 'derender3d_pose.npz'), **out)
    print('pose ok', sorted(k for k in out if k.startswith('test'))[:6])


if __name__ == '__main__':
    golden_transforms()
    golden_pose()

"""The fused pre-raster stage (sdn_objects_forward / _backward: FFD + camera-space transform of a ragged batch of
objects in one launch) against (1) golden vectors produced by the reference's own transforms.py
(tests/golden/derender3d_transforms.npz) and (2) the torch restatement of the per-object path
(derender3d/models/transforms.py here, itself pinned by the same goldens on the CPU).  float32; the 64-term FFD sum and
the per-object gradient reductions are ordered differently: 1e-5 / 1e-4 rel-L2."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available()
    import sdn_b200
    sdn_b200.lib()
    return torch.device('cuda:0')


def _rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _constraints(FFD):
    return [FFD.Constraint.symmetry(axis=FFD.Constraint.Axis.z),
            FFD.Constraint.homogeneity(axis=FFD.Constraint.Axis.y, index=[0, 1])]


def _setup(dev, nvs=(150, 333, 64), N=9, seed=0):
    import sdn_b200
    from derender3d.models.transforms import FFD
    g = torch.Generator().manual_seed(seed)
    ffds = [FFD(torch.rand(nv, 3, generator=g) - 0.5, constraints=_constraints(FFD)).to(dev) for nv in nvs]
    faces = [torch.randint(0, nv, (2 * nv, 3), generator=g, dtype=torch.int32) for nv in nvs]
    bank = sdn_b200.MeshBank([f.B for f in ffds], faces, dev)
    bank.C = ffds[0].constraint_matrix().to(dev)
    bank.P0 = ffds[0].P0.reshape(1, -1).to(dev)
    classes = torch.randint(0, len(nvs), (N,), generator=g).to(dev)
    th = torch.rand(N, generator=g) * 6.28
    par = {
        'coeffs': 0.2 * torch.randn(N, 192, generator=g),
        'scales': torch.exp(0.3 * torch.randn(N, 3, generator=g)),
        'rotations': torch.stack([torch.cos(th / 2), 0.1 * torch.randn(N, generator=g), torch.sin(th / 2),
                                  0.1 * torch.randn(N, generator=g)], dim=1),
        'translations': torch.stack([torch.randn(N, generator=g), 0.3 * torch.randn(N, generator=g),
                                     -(5 + 10 * torch.rand(N, generator=g))], dim=1),
        'ptrans': torch.stack([torch.randn(N, generator=g), 0.3 * torch.randn(N, generator=g),
                               -(5 + 10 * torch.rand(N, generator=g))], dim=1),
        'zooms': 0.5 + torch.rand(N, 1, generator=g),
        'zoom_tos': 0.1 + 0.3 * torch.rand(N, 1, generator=g),
    }
    par = {k: v.to(dev).requires_grad_(True) for k, v in par.items()}
    return bank, ffds, classes, par


def _torch_path(ffds, classes, par, train):
    """One object at a time through the torch modules, like the reference's loop (models/__init__.py:161-200)."""
    from derender3d.models.transforms import PerspectiveTransform
    pt = PerspectiveTransform()
    outs, zs = [], []
    for i, k in enumerate(classes.tolist()):
        v = ffds[k](par['coeffs'][i]).unsqueeze(0)
        kw = dict(scales=par['scales'][i:i + 1], rotations=par['rotations'][i:i + 1],
                  translations=par['translations'][i:i + 1])
        if train:
            outs.append(pt(v, perspective_translations=par['ptrans'][i:i + 1], zooms=par['zooms'][i:i + 1], **kw)[0])
        else:
            o, z = pt(v, perspective_translations=par['translations'][i:i + 1], zoom_tos=par['zoom_tos'][i:i + 1], **kw)
            outs.append(o[0])
            zs.append(z)
    return outs, (torch.cat(zs) if zs else None)


def _fused(bank, classes, par, train):
    import sdn_b200
    control = (bank.P0 + par['coeffs'] @ bank.C).reshape(-1, 3, 64)
    kw = dict(scales=par['scales'], rotations=par['rotations'], translations=par['translations'])
    if train:
        return sdn_b200.object_vertices(bank, classes, control, ptrans=par['ptrans'], zooms=par['zooms'], **kw), None
    return sdn_b200.object_vertices(bank, classes, control, ptrans=par['translations'], zoom_tos=par['zoom_tos'], **kw)


@pytest.mark.parametrize('train', [True, False], ids=['train', 'test'])
def test_fused_objects_match_per_object_torch_path(dev, train):
    bank, ffds, classes, par = _setup(dev)
    ref, ref_z = _torch_path(ffds, classes, par, train)
    out, z = _fused(bank, classes, par, train)
    assert out.shape == (len(classes), bank.nv_max, 3)
    g = torch.Generator().manual_seed(5)
    ups = [torch.randn(r.shape, generator=g).to(dev) for r in ref]
    for i, r in enumerate(ref):
        assert _rel(out[i, :r.shape[0]], r) <= 1e-5
        pad = out[i, r.shape[0]:]
        assert torch.equal(pad, torch.tensor([0.0, 0.0, 1.0], device=dev).expand_as(pad))
    if not train:
        assert _rel(z, ref_z) <= 1e-6
    names = ['coeffs', 'scales', 'rotations', 'translations'] + (['ptrans', 'zooms'] if train else ['zoom_tos'])
    leaves = [par[n] for n in names]
    loss_ref = sum((r * u).sum() for r, u in zip(ref, ups))
    g_ref = torch.autograd.grad(loss_ref, leaves)
    up_full = torch.zeros_like(out)
    for i, u in enumerate(ups):
        up_full[i, :u.shape[0]] = u
    g_out = torch.autograd.grad((out * up_full).sum(), leaves)
    for n, a, b in zip(names, g_out, g_ref):
        assert _rel(a, b) <= 1e-4, n


def test_fused_objects_against_reference_goldens(dev):
    """tests/golden/derender3d_transforms.npz was written by the reference's transforms.py (make_golden_derender3d.py)."""
    import sdn_b200
    from derender3d.models.transforms import FFD
    z = np.load(os.path.join(HERE, 'golden', 'derender3d_transforms.npz'))
    t = {k: torch.tensor(z[k]).to(dev) for k in z.files}
    ffd = FFD(t['vertices'].cpu(), constraints=_constraints(FFD)).to(dev)
    bank = sdn_b200.MeshBank([ffd.B], [torch.zeros(1, 3, dtype=torch.int32)], dev)
    n = t['coeffs'].shape[0]
    classes = torch.zeros(n, dtype=torch.int64, device=dev)
    control = (ffd.P0.reshape(1, -1) + t['coeffs'] @ ffd.constraint_matrix().to(dev)).reshape(n, 3, 64)
    kw = dict(scales=t['scales'], rotations=t['rotations'], translations=t['translations'])
    train = sdn_b200.object_vertices(bank, classes, control, ptrans=t['perspective_translations'], zooms=t['zooms'], **kw)
    assert _rel(train, t['train_out']) <= 1e-5
    test_v, test_z = sdn_b200.object_vertices(bank, classes, control, ptrans=t['translations'], zoom_tos=t['zoom_tos'], **kw)
    assert _rel(test_v, t['test_out']) <= 1e-5
    assert _rel(test_z, t['test_zooms']) <= 1e-6
    # identity transform: the FFD alone
    one = torch.ones(n, 3, device=dev)
    ident = sdn_b200.object_vertices(bank, classes, control, scales=one,
                                     rotations=torch.tensor([1.0, 0, 0, 0], device=dev).expand(n, 4).contiguous(),
                                     translations=torch.tensor([0.0, 0, 1.0], device=dev).expand(n, 3).contiguous(),
                                     ptrans=torch.tensor([0.0, 0, 1.0], device=dev).expand(n, 3).contiguous(),
                                     zooms=torch.ones(n, 1, device=dev))
    ffd_out = t['ffd_out'].clone()
    ffd_out[..., 2] += 1.0
    assert _rel(ident, ffd_out) <= 1e-5


@pytest.mark.parametrize('training', [True, False], ids=['train', 'eval'])
def test_derenderer3d_fused_front_end_equals_class_grouped_path(dev, training):
    """Derenderer3d.render: ragged single-launch path vs the class-grouped torch path (same pose, same classes)."""
    from derender3d import TargetType
    from derender3d.models import Derenderer3d, ShapenetObj
    from sdn_b200 import workloads as wl
    objs = []
    for k in range(8):
        v, f = wl.superellipsoid_mesh(12 + 2 * k, 8 + k, 0.5 + 0.05 * k, 0.6)
        objs.append(ShapenetObj(vertices=v * np.array([1.0, 0.35, 0.45], np.float32), faces=f))
    torch.manual_seed(0)
    model = Derenderer3d(TargetType.extend, image_size=256, render_size=64,
                         objs=objs).to(dev)
    model.train(training)
    model._force_no_sample = True
    B = 10
    g = torch.Generator().manual_seed(3)
    c = 0.3 * (torch.rand(B, 2, generator=g) - 0.5)
    d = 0.25 + 0.15 * torch.rand(B, 2, generator=g)
    th = torch.rand(B, generator=g) * 6.28
    blob = {'_mroi_norms': c.to(dev), '_droi_norms': d.to(dev), '_focals': torch.full((B, 1), 725.0, device=dev),
            '_theta_deltas': torch.stack([torch.cos(th), torch.sin(th)], dim=1).to(dev),
            '_translation2ds': 0.05 * torch.randn(B, 2, generator=g).to(dev),
            '_log_scales': torch.tensor([1.3, 0.4, 0.5]).expand(B, 3).contiguous().to(dev),
            '_log_depths': torch.full((B, 1), 2.2, device=dev),
            '_class_probs': torch.softmax(3 * torch.randn(B, 8, generator=g), dim=1).to(dev),
            '_ffd_coeffs': (0.03 * torch.randn(B, 8, 192, generator=g)).to(dev).requires_grad_(True)}
    res = {}
    for fused in (True, False):
        model.fused_front_end = fused
        out = model.render(blob)
        up = torch.Generator().manual_seed(7)
        loss = sum((out[k] * torch.randn(out[k].shape, generator=up).to(dev)).sum()
                   for k in ('_masks', '_normals', '_depth_maps'))
        gr, = torch.autograd.grad(loss, blob['_ffd_coeffs'])
        res[fused] = (out, gr)
    a, b = res[True], res[False]
    for k in ('_masks', '_normals', '_depth_maps', '_zooms'):
        assert a[0][k].shape == b[0][k].shape, k
        assert _rel(a[0][k].float(), b[0][k].float()) <= 2e-3, k      # a 1e-7 vertex difference may move a pixel
    assert float((a[0]['_masks'] > 0).float().mean()) > 0.02
    assert _rel(a[1], b[1]) <= 2e-2


def test_whole_train_step_as_cuda_graph_matches_eager(dev):
    """GraphedTrainStep: encoder -> pose -> fused front end -> render -> losses -> backward -> Adam in ONE graph launch;
    same trajectory as the eager step: the warm-up steps the capture needs are undone (parameters, BatchNorm buffers
    and Adam state are restored), so graph replay k equals eager step k (float atomics reorder sums: 1e-3 on the loss)."""
    import copy
    from derender3d import TargetType
    from derender3d.models import Derenderer3d, ShapenetObj
    from derender3d.train_step import GraphedTrainStep, train_step
    from sdn_b200 import workloads as wl
    objs = []
    for k in range(8):
        v, f = wl.superellipsoid_mesh(12 + 2 * k, 8 + k, 0.5 + 0.05 * k, 0.6)
        objs.append(ShapenetObj(vertices=v * np.array([1.0, 0.35, 0.45], np.float32), faces=f))
    torch.manual_seed(0)
    B, S = 6, 64
    model_a = Derenderer3d(TargetType.full, image_size=256, render_size=S, objs=objs).to(dev).train()
    with torch.no_grad():
        model_a.derenderer._fc3.weight.mul_(0.01)
        model_a.derenderer._fc3.bias.zero_()
        model_a.derenderer._fc3.bias[0] = 1.0
        model_a.derenderer._fc3.bias[4:7] = torch.tensor([1.3, 0.4, 0.5])
        model_a.derenderer._fc3.bias[7] = 2.2
    model_a._force_no_sample = True
    model_b = copy.deepcopy(model_a)
    g = torch.Generator().manual_seed(1)
    c = 0.3 * (torch.rand(B, 2, generator=g) - 0.5)
    d = 0.25 + 0.15 * torch.rand(B, 2, generator=g)
    batch = {'images': torch.randn(B, 3, 224, 224, generator=g).to(dev),
             'roi_norms': torch.cat([c - d / 2, c + d / 2], dim=1).to(dev),
             'focals': torch.full((B, 1), 725.0, device=dev),
             'targets': torch.tensor([3, 3, 1, 2, 3, 3], device=dev),
             'thetas': torch.rand(B, 1, generator=g).to(dev), 'translation2ds': torch.zeros(B, 2, device=dev),
             'log_scales': torch.tensor([1.3, 0.4, 0.5], device=dev).expand(B, 3).contiguous(),
             'log_depths': torch.full((B, 1), 2.2, device=dev),
             'masks': (torch.rand(B, 1, 40, 40, generator=g) > 0.5).float().to(dev),
             'ignores': torch.zeros(B, 1, 40, 40, device=dev)}
    opt_a = torch.optim.Adam(model_a.parameters(), lr=1e-4, weight_decay=1e-3, capturable=True)
    opt_b = torch.optim.Adam(model_b.parameters(), lr=1e-4, weight_decay=1e-3, capturable=True)
    w0 = model_b.derenderer._fc3.weight.detach().clone()
    gs = GraphedTrainStep(model_b, opt_b, batch, TargetType.full, warmup=3)
    assert torch.equal(w0, model_b.derenderer._fc3.weight.detach()), 'building the graph must not train the model'
    assert int(model_b.derenderer.net.bn1.num_batches_tracked) == 0
    for _ in range(4):
        loss_a, _, blob_a = train_step(model_a, opt_a, batch, TargetType.full)
        loss_b = gs.step(batch)
    torch.cuda.synchronize()
    assert torch.isfinite(loss_b)
    assert abs(float(loss_a) - float(loss_b)) <= 1e-3 * abs(float(loss_a))
    wa, wb = model_a.derenderer._fc3.weight.detach(), model_b.derenderer._fc3.weight.detach()
    assert float((wa - wb).norm() / wa.norm()) <= 1e-3
    assert int(model_b.derenderer.net.bn1.num_batches_tracked) == 4
    assert float((gs.blob['_masks'] > 0).float().mean()) > 0.01
    assert _rel(gs.blob['_masks'].float(), blob_a['_masks'].float()) <= 5e-2
    wa, wb = model_a.derenderer._fc3.weight, model_b.derenderer._fc3.weight
    assert _rel(wb, wa) <= 1e-3

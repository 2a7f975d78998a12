"""Host-side rows of the hot path (a2 pose algebra, a3 FFD, a4 PerspectiveTransform) against golden
vectors produced by the REFERENCE'S OWN torch code (tests/golden/make_golden_derender3d.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def _t(a):
    return torch.tensor(a)


def test_ffd_and_perspective_transform_match_reference():
    from derender3d.models.transforms import FFD, PerspectiveTransform
    g = np.load(os.path.join(HERE, 'golden', 'derender3d_transforms.npz'))
    cons = [FFD.Constraint.symmetry(axis=FFD.Constraint.Axis.z),
            FFD.Constraint.homogeneity(axis=FFD.Constraint.Axis.y, index=[0, 1])]
    ffd = FFD(g['vertices'], constraints=cons)
    coeffs = _t(g['coeffs'])
    batched = ffd(coeffs)                                           # [5,200,3] in one call
    assert torch.allclose(batched, _t(g['ffd_out']), atol=2e-6)
    assert torch.allclose(ffd(coeffs[2]), _t(g['ffd_out'][2]), atol=2e-6)       # single-vector form of the reference
    assert torch.allclose(FFD(g['vertices'], constraints=[])(coeffs), _t(g['ffd_free']), atol=2e-6)
    # zero coefficients reproduce the template (Bernstein partition of unity)
    assert torch.allclose(ffd(torch.zeros(192)), _t(g['vertices']), atol=2e-6)
    pt = PerspectiveTransform()
    kw = dict(scales=_t(g['scales']), rotations=_t(g['rotations']), translations=_t(g['translations']))
    train = pt(batched, perspective_translations=_t(g['perspective_translations']), zooms=_t(g['zooms']), **kw)
    assert torch.allclose(train, _t(g['train_out']), rtol=1e-5, atol=1e-5)
    test_v, test_z = pt(batched, perspective_translations=_t(g['translations']), zoom_tos=_t(g['zoom_tos']), **kw)
    assert torch.allclose(test_v, _t(g['test_out']), rtol=1e-5, atol=1e-5)
    assert torch.allclose(test_z, _t(g['test_zooms']), rtol=1e-5)


def test_ffd_gradient_flows():
    from derender3d.models.transforms import FFD
    v = torch.rand(30, 3) - 0.5
    ffd = FFD(v, constraints=[FFD.Constraint.symmetry(axis=FFD.Constraint.Axis.z)])
    c = torch.zeros(4, 192, requires_grad=True)
    ffd(c).sum().backward()
    assert c.grad.shape == (4, 192) and float(c.grad.abs().sum()) > 0


@pytest.mark.parametrize('mode', ['test', 'train'])
def test_pose_algebra_matches_reference(mode):
    """Derenderer3d.pose == the part of the reference's render() before the object loop, for eval and train mode."""
    from derender3d import TargetType
    from derender3d.models import Derenderer3d, ShapenetObj
    g = np.load(os.path.join(HERE, 'golden', 'derender3d_pose.npz'))
    objs = [ShapenetObj(vertices=np.random.RandomState(k).rand(20, 3), faces=np.zeros((4, 3), np.int32))
            for k in range(8)]
    model = Derenderer3d(TargetType.extend, image_size=256, render_size=64, objs=objs)
    model.train(mode == 'train')
    model._force_no_sample = True
    blob = {k[2:]: _t(g[k]) for k in g.files if k.startswith('in_')}
    p = model.pose(blob)
    for key in ('_thetas', '_alphas', '_rotations', '_scales', '_depths', '_center2ds', '_translations',
                '_class_log_probs'):
        assert torch.allclose(p[key], _t(g[mode + key]), rtol=1e-5, atol=1e-6), key
    if mode == 'train':
        assert torch.allclose(p['_zooms'], _t(g['train_zooms']), rtol=1e-5)
    # the per-object viewing angle the reference sets on its renderer (derender3d/models/__init__.py:202)
    ang = torch.atan(64 / (2.0 * blob['_focals'].double().reshape(-1))) / np.pi * 180
    assert np.allclose(ang.numpy(), g[mode + '_angles'], rtol=1e-12)


def test_shapenet_obj_normalisation():
    from derender3d.models import ShapenetObj
    v = np.array([[0, 0, 0], [2, 1, 4], [1, 0.5, 2]], np.float32)
    o = ShapenetObj(vertices=v, faces=np.array([[0, 1, 2]], np.int32))
    # unit extent per axis, x<->z swapped, new x negated (reference :30-31)
    assert torch.allclose(o.vertices[1], torch.tensor([-1.0, 1.0, 1.0]))
    assert o.faces.dtype == torch.int32


def test_derenderer_heads_and_checkpoint_keys():
    from derender3d.models.derenderer import Derenderer
    m = Derenderer().eval()
    out = m(torch.randn(2, 3, 64, 64), torch.zeros(2, 2), torch.ones(2, 2))
    assert out['_ffd_coeffs'].shape == (2, 8, 192) and out['_class_probs'].shape == (2, 8)
    assert torch.allclose(out['_class_probs'].sum(1), torch.ones(2), atol=1e-5)
    assert torch.allclose(out['_theta_deltas'].norm(dim=1), torch.ones(2), atol=1e-5)
    keys = set(m.state_dict().keys())
    # names a reference checkpoint carries (derenderer.py:25-32)
    assert {'net.conv1.weight', 'net.fc.weight', 'fc1.weight', 'fc2.bias', '_fc3.weight'} <= keys
    assert m._fc3.weight.shape == (2 + 2 + 3 + 1 + 8 + 1536, 256)


@pytest.mark.parametrize('mode', ['test', 'train'])
def test_object_vertices_match_reference_on_real_meshes(mode):
    """FFD + PerspectiveTransform on the 8 ShapeNet templates, grouped by class, against the vertices the reference
    handed to its renderer object by object (needs the OBJ assets of /root/reference; skipped on the GPU box)."""
    root = '/root/reference/geometric/assets'
    if not os.path.isdir(root):
        pytest.skip('reference assets not present on this machine')
    from derender3d import TargetType
    from derender3d.models import Derenderer3d, ShapenetObj, SHAPENET_OBJS
    g = np.load(os.path.join(HERE, 'golden', 'derender3d_pose.npz'))
    objs = [ShapenetObj(c, o, root_dir=root) for c, o in SHAPENET_OBJS]
    model = Derenderer3d(TargetType.extend, image_size=256, render_size=64, objs=objs)
    model.train(mode == 'train')
    model._force_no_sample = True
    blob = {k[2:]: _t(g[k]) for k in g.files if k.startswith('in_')}
    p = model.pose(blob)
    classes = p['_class_samples'].tolist()
    seen = 0
    for k, sel, verts, z in model.object_vertices(blob, p, classes):
        for j, i in enumerate(sel.tolist()):
            assert verts.shape[1] == int(g[mode + '_nv'][i])
            ref = _t(g[mode + '_vertex_sample'][i])
            assert torch.allclose(verts[j, ::97][:64], ref, rtol=2e-5, atol=2e-5), (k, i)
            if z is not None:
                assert torch.allclose(z[j], _t(g['test_zooms'][i]), rtol=1e-5)
            seen += 1
    assert seen == len(classes)


def test_sync_free_losses_equal_the_nonzero_formulation():
    """main.py:97-112 selects rows with torch.nonzero (a host sync); train_step uses masks.  Same values, including
    the 'nothing selected -> 0' branch."""
    import torch.nn.functional as F
    from derender3d import TargetType
    from derender3d.train_step import _masked_mse, _masked_mean

    def partial(f, mask):
        def g(*args, **kw):
            index = torch.nonzero(mask)
            if index.numel():
                index = index.squeeze(dim=1)
                return f(*[a[index] for a in args], **kw)
            return torch.zeros(())
        return g

    g = torch.Generator().manual_seed(0)
    for targets in (torch.tensor([3, 1, 2, 0, 3, 2]), torch.zeros(6, dtype=torch.long), torch.full((6,), 3)):
        for bit in (TargetType.pretrain, TargetType.finetune):
            mask = targets & bit
            a, b = torch.randn(6, 3, generator=g), torch.randn(6, 3, generator=g)
            x = torch.randn(6, generator=g)
            assert torch.allclose(_masked_mse(a, b, mask), partial(F.mse_loss, mask)(a, b), atol=1e-6)
            assert torch.allclose(_masked_mean(x, mask), partial(torch.mean, mask)(x), atol=1e-6)


def test_ffd_constraint_matrix_is_the_constraint_step():
    """transforms.py:71-95 is linear: constrain(x) == x @ C (what the fused front end applies as one matmul)."""
    from derender3d.models.transforms import FFD
    v = torch.rand(30, 3, generator=torch.Generator().manual_seed(0)) - 0.5
    for cons in ([], [FFD.Constraint.symmetry(axis=FFD.Constraint.Axis.z)],
                 [FFD.Constraint.symmetry(axis=FFD.Constraint.Axis.z),
                  FFD.Constraint.homogeneity(axis=FFD.Constraint.Axis.y, index=[0, 1])]):
        f = FFD(v, constraints=cons)
        C = f.constraint_matrix()
        x = torch.randn(5, 192, generator=torch.Generator().manual_seed(1))
        want = f.constrain(x.reshape(5, 3, 4, 4, 4)).reshape(5, 192)
        assert torch.allclose(x @ C, want, atol=1e-6)
        assert torch.allclose(C @ C, C, atol=1e-6)          # a projection: constraining twice changes nothing


def test_mesh_bank_layout():
    """sdn_b200.MeshBank: concatenated Bernstein tables, row offsets, faces padded with (0,0,0) triangles."""
    import sdn_b200
    from derender3d.models.transforms import FFD
    g = torch.Generator().manual_seed(0)
    ffds = [FFD(torch.rand(nv, 3, generator=g) - 0.5) for nv in (11, 29, 5)]
    faces = [torch.randint(0, nv, (nf, 3), generator=g, dtype=torch.int32) for nv, nf in ((11, 20), (29, 7), (5, 9))]
    bank = sdn_b200.MeshBank([f.B for f in ffds], faces, 'cpu')
    assert (bank.K, bank.nv_max, bank.nf_max) == (3, 29, 20)
    assert bank.basis.shape == (45, 64) and bank.basis_off.tolist() == [0, 11, 40] and bank.nv.tolist() == [11, 29, 5]
    for k, f in enumerate(ffds):
        lo = int(bank.basis_off[k])
        assert torch.equal(bank.basis[lo:lo + f.B.shape[0]], f.B.reshape(-1, 64))
        assert torch.equal(bank.faces[k, :faces[k].shape[0]], faces[k])
        assert int(bank.faces[k, faces[k].shape[0]:].abs().sum()) == 0
    # each Bernstein row is a partition of unity
    assert torch.allclose(bank.basis.sum(1), torch.ones(45), atol=1e-5)

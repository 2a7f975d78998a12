"""Host-side glue of the product (3d-sdn_b200/neural_renderer) against the oracle's restatement, on CPU tensors:
these are pure elementwise float32 ops in a fixed order, so they must agree bit for bit."""
import os

import numpy as np
import pytest
import torch

import neural_renderer as nr


def _verts(seed=0, B=2, nv=50):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, nv, 3, generator=g)


def test_camera_glue_matches_oracle(oracle):
    v = _verts()
    eye = [1.0, 2.0, -5.0]
    assert torch.equal(nr.look_at(v, eye), oracle.look_at(v, eye))
    d = torch.tensor([[0.1, -0.2, 1.0], [0.0, 0.0, -1.0]])
    assert torch.equal(nr.look(v, eye, d), oracle.look(v, eye, d))
    a = nr.look_at(v, eye)
    assert torch.equal(nr.perspective(a, 30.), oracle.perspective(a, 30.))
    assert torch.equal(nr.perspective(a, 14.5, 1242 / 375), oracle.perspective(a, 14.5, 1242 / 375))
    assert nr.perspective_width(30.) == oracle.perspective_width(30.)
    # the literal 3.1416 of perspective.py:10, not math.pi
    assert abs(float(nr.perspective_width(30.)) - np.tan(30 / 180 * 3.1416)) < 1e-6


def test_gather_cross_lighting_match_oracle(oracle):
    v = _verts(1)
    f = torch.randint(0, 50, (2, 30, 3), generator=torch.Generator().manual_seed(2), dtype=torch.int32)
    fv = nr.vertices_to_faces(v, f)
    assert fv.shape == (2, 30, 3, 3) and torch.equal(fv, oracle.vertices_to_faces(v, f))
    a, b = v[0, :20], v[1, :20]
    assert torch.equal(nr.cross(a, b), oracle.cross(a, b))
    assert torch.allclose(nr.cross(a, b), torch.cross(a, b, dim=1), atol=1e-6)
    tex = torch.rand(2, 30, 2, 2, 2, 3, generator=torch.Generator().manual_seed(3))
    assert torch.equal(nr.lighting(fv, tex), oracle.lighting(fv, tex))


def test_get_points_from_angles():
    p = nr.get_points_from_angles(2.732, 20, 30)
    assert np.allclose(p, (2.732 * np.cos(np.radians(20)) * np.sin(np.radians(30)), 2.732 * np.sin(np.radians(20)),
                           -2.732 * np.cos(np.radians(20)) * np.cos(np.radians(30))))
    t = nr.get_points_from_angles(torch.tensor([2.732]), torch.tensor([20.]), torch.tensor([30.]))
    assert torch.allclose(t[0], torch.tensor(p, dtype=torch.float32), atol=1e-6)


def test_load_save_obj_roundtrip(tmp_path):
    v = np.array([[0, 0, 0], [2, 0, 0], [2, 1, 0], [0, 1, 0.5]], np.float32)
    path = str(tmp_path / 'q.obj')
    with open(path, 'w') as fh:
        fh.write('# quad\n')
        for p in v:
            fh.write('v %f %f %f\n' % tuple(p))
        fh.write('vn 0 0 1\nf 1/1/1 2/2/1 3/3/1 4/4/1\n')
    vv, ff = nr.load_obj(path, normalization=False)
    assert np.allclose(vv, v) and ff.tolist() == [[0, 1, 2], [0, 2, 3]] and ff.dtype == np.int32
    vn, _ = nr.load_obj(path)
    assert np.isclose(np.abs(vn).max(), 1.0) and np.allclose(vn.max(0) + vn.min(0), 0, atol=1e-6)
    out = str(tmp_path / 'o.obj')
    nr.save_obj(out, vv, ff)
    v2, f2 = nr.load_obj(out, normalization=False)
    assert np.allclose(v2, vv, atol=1e-4) and np.array_equal(f2, ff)


def test_load_obj_on_reference_assets():
    root = '/root/reference/geometric/assets/02958343'
    if not os.path.isdir(root):
        pytest.skip('reference assets not present on this machine')
    path = os.path.join(root, sorted(os.listdir(root))[0], 'models', 'model_normalized.obj')
    v, f = nr.load_obj(path)
    assert v.dtype == np.float32 and f.dtype == np.int32 and f.min() == 0 and f.max() < len(v)
    assert 9000 < len(v) < 20000 and 29000 < len(f) < 73000


def test_mesh_module_batches_and_parameter_groups(tmp_path):
    """neural_renderer.Mesh (mesh.py:7-38): parameters, broadcast batch, per-parameter learning rates."""
    import neural_renderer as nr
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], np.float32)
    f = np.array([[0, 1, 2], [0, 2, 3], [0, 3, 1], [1, 3, 2]], np.int32)
    path = str(tmp_path / 'tetra.obj')
    nr.save_obj(path, v, f)
    mesh = nr.Mesh(path, texture_size=2)
    assert (mesh.num_vertices, mesh.num_faces, mesh.texture_size) == (4, 4, 2)
    vb, fb, tb = mesh.get_batch(3)
    assert vb.shape == (3, 4, 3) and fb.shape == (3, 4, 3) and tb.shape == (3, 4, 2, 2, 2, 3)
    assert float(tb.min()) > 0 and float(tb.max()) < 1
    assert {id(p) for p in mesh.parameters()} == {id(mesh.vertices), id(mesh.textures)}
    opt = torch.optim.Adam(mesh.set_lr(0.01, 0.1))
    assert [g['lr'] for g in opt.param_groups] == [0.01, 0.1]


def test_pack_conv_weight_layouts_agree():
    """pack_conv_weight: the one-cast fast path (no padding needed) and the padded path give the same [Cout,KH,KW,Cin] bf16
    operand for contiguous and channels-last weights (host logic of the convolution boundary, no GPU)."""
    import torch
    import sdn_b200
    g = torch.Generator().manual_seed(0)
    w = torch.randn(128, 64, 3, 3, generator=g)
    a = sdn_b200.pack_conv_weight(w)
    b = sdn_b200.pack_conv_weight(w.contiguous(memory_format=torch.channels_last))
    want = w.permute(0, 2, 3, 1).to(torch.bfloat16)
    assert a.shape == (128, 3, 3, 64) and a.is_contiguous() and b.is_contiguous()
    assert torch.equal(a, want) and torch.equal(b, want)
    # channel padding: Cin 48 -> 64, Cout 3 -> 16 (the generator's stem / head), zero rows and channels
    w2 = torch.randn(3, 48, 7, 7, generator=g)
    p = sdn_b200.pack_conv_weight(w2)
    assert p.shape == (16, 7, 7, 64)
    assert torch.equal(p[:3, :, :, :48], w2.permute(0, 2, 3, 1).to(torch.bfloat16))
    assert float(p[3:].abs().sum()) == 0.0 and float(p[:, :, :, 48:].abs().sum()) == 0.0


def test_mlp_head_backward_formulas_match_autograd():
    """sdn_b200.mlp_head_backward (the torch backward of the fused encoder head) against autograd through the same head
    written with torch.nn.functional.linear (no GPU: the forward kernel is tested in tests/test_gpu_linear_tc.py)."""
    import torch
    import torch.nn.functional as F
    import sdn_b200
    g = torch.Generator().manual_seed(1)
    M, K0, N0, E, N1, N2, N3 = 7, 32, 50, 4, 24, 24, 11
    mk = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64).requires_grad_(True)
    x, extra = mk(M, K0), mk(M, E)
    W0, b0, W1, b1, W2, b2, W3, b3 = mk(N0, K0), mk(N0), mk(N1, N0 + E), mk(N1), mk(N2, N1), mk(N2), mk(N3, N2), mk(N3)
    cat = torch.cat([F.relu(F.linear(x, W0, b0)), extra], dim=1)
    h1 = F.relu(F.linear(cat, W1, b1))
    h2 = F.relu(F.linear(h1, W2, b2))
    y = F.linear(h2, W3, b3)
    gy = torch.randn(M, N3, generator=g, dtype=torch.float64)
    y.backward(gy)
    got = sdn_b200.mlp_head_backward(gy.float(), x.detach().float(), cat.detach().float(), h1.detach().float(), h2.detach().float(),
                                     W0.detach().float(), W1.detach().float(), W2.detach().float(), W3.detach().float(), N0)
    want = {'x': x.grad, 'extra': extra.grad, 'W0': W0.grad, 'b0': b0.grad, 'W1': W1.grad, 'b1': b1.grad, 'W2': W2.grad,
            'b2': b2.grad, 'W3': W3.grad, 'b3': b3.grad}
    for k, v in want.items():
        assert got[k].shape == v.shape, k
        assert float((got[k].double() - v).norm() / (v.norm() + 1e-30)) < 1e-5, k

"""Parity of the sm_100a path (through the C ABI) against the oracle and the reference-generated golden vectors.
Bars (BASELINE.json north_star): int32 face-index buffer bit-exact; RGB / depth / alpha <= 1e-4 rel-L2 (alpha is in
fact exact); gradients <= 1e-4 rel-L2 (a float sum in a different order)."""
import numpy as np
import pytest
import torch

from util import GOLDEN, load_golden, random_faces, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available()
    import sdn_b200
    sdn_b200.lib()
    return torch.device('cuda:0')


def _sdn_raster(faces, textures, H, W, ss, near, far, eps, bg, flags, dev, g=None):
    """Run sdn_rasterize_forward (+ backward) on raw NDC faces; returns numpy maps in RASTER orientation at
    sample resolution where that makes sense (ss == 1)."""
    import sdn_b200
    rgb, alpha, depth = [bool(x) for x in flags]
    cfg = sdn_b200.RasterConfig((H, W), ss == 2, near, far, eps, bg, rgb, alpha, depth)
    f = torch.tensor(faces, device=dev, requires_grad=True)
    t = torch.tensor(textures, device=dev, requires_grad=True) if rgb else None
    o_rgb, o_alpha, o_depth = sdn_b200.RasterizeFunction.apply(f, t, cfg)
    res = {'face_index': cfg.last_face_index.cpu().numpy(), 'rgb': o_rgb, 'alpha': o_alpha, 'depth': o_depth,
           'faces': f, 'textures': t}
    return res


@pytest.mark.parametrize('math', ['fast', 'exact'])
@pytest.mark.parametrize('path', GOLDEN, ids=lambda p: p.split('/')[-1][:-4])
def test_golden_vectors(dev, path, math, monkeypatch):
    """Vectors produced by the reference's own kernel bodies (tests/golden/make_golden.py), in the library's default
    backward arithmetic (fast: line sweeps) and in the IEEE-division mode."""
    import sdn_b200
    assert sdn_b200.DEFAULT_FAST_MATH == 1, 'the default backward mode is the one bench.py times'
    monkeypatch.setattr(sdn_b200, 'DEFAULT_FAST_MATH', 1 if math == 'fast' else 0)
    g = load_golden(path)
    S = int(g['S'])
    flags = g['flags']
    rgb, alpha, depth = [bool(x) for x in flags]
    r = _sdn_raster(g['faces'], g.get('textures'), S, S, 1, float(g['near']), float(g['far']), float(g['eps']),
                    tuple(float(x) for x in g['background']), flags, dev)
    assert np.array_equal(r['face_index'], g['face_index_map'])
    loss = 0
    # outputs are row-flipped and channel-first; undo for comparison with the raster-orientation maps
    if alpha:
        a = r['alpha'].flip(1)
        assert torch.equal(a.cpu(), torch.tensor(g['alpha_map']))
        loss = loss + (a * torch.tensor(g['grad_alpha_map'], device=dev)).sum()
    if depth:
        d = r['depth'].flip(1)
        assert torch.equal(d.cpu(), torch.tensor(g['depth_map']))
        loss = loss + (d * torch.tensor(g['grad_depth_map'], device=dev)).sum()
    if rgb:
        c = r['rgb'].flip(2).permute(0, 2, 3, 1)
        assert rel_l2(c, g['rgb_map']) <= 1e-6
        loss = loss + (c * torch.tensor(g['grad_rgb_map'], device=dev)).sum()
    loss.backward()
    assert rel_l2(r['faces'].grad, g['grad_faces']) <= TOL
    if rgb:
        assert rel_l2(r['textures'].grad, g['grad_textures']) <= TOL


@pytest.mark.parametrize('path', GOLDEN, ids=lambda p: p.split('/')[-1][:-4])
def test_rasterize_class_returns_the_reference_maps(dev, path):
    """nr.Rasterize(...)(faces, textures) has the return contract of rasterize.py:19-64 / :464-510: channel-last maps in
    raster orientation -- exactly what the golden files hold."""
    import neural_renderer as nr
    g = load_golden(path)
    S = int(g['S'])
    rgb, alpha, depth = [bool(x) for x in g['flags']]
    fn = nr.Rasterize(S, float(g['near']), float(g['far']), float(g['eps']), tuple(float(x) for x in g['background']),
                      rgb, alpha, depth)
    f = torch.tensor(g['faces'], device=dev, requires_grad=True)
    t = torch.tensor(g['textures'], device=dev, requires_grad=True) if rgb else None
    o_rgb, o_alpha, o_depth = fn(f, t) if rgb else fn(f)
    assert np.array_equal(fn.face_index_map.cpu().numpy(), g['face_index_map'])
    loss = 0
    if alpha:
        assert torch.equal(o_alpha.cpu(), torch.tensor(g['alpha_map']))
        loss = loss + (o_alpha * torch.tensor(g['grad_alpha_map'], device=dev)).sum()
    else:
        assert o_alpha is None
    if depth:
        assert torch.equal(o_depth.cpu(), torch.tensor(g['depth_map']))
        loss = loss + (o_depth * torch.tensor(g['grad_depth_map'], device=dev)).sum()
    else:
        assert o_depth is None
    if rgb:
        assert o_rgb.shape == g['rgb_map'].shape and rel_l2(o_rgb, g['rgb_map']) <= 1e-6
        loss = loss + (o_rgb * torch.tensor(g['grad_rgb_map'], device=dev)).sum()
    else:
        assert o_rgb is None
    loss.backward()
    assert rel_l2(f.grad, g['grad_faces']) <= TOL


@pytest.mark.parametrize('aa', [True, False])
@pytest.mark.parametrize('size', [64, (48, 80)])
def test_config1_cube(dev, oracle, aa, size):
    """BASELINE config 1: 12-face cube, nr.Renderer defaults, eye from get_points_from_angles(2.732, 20, 30)."""
    import neural_renderer as nr
    from sdn_b200 import workloads as wl
    v, f = wl.cube_mesh()
    V, Fc = torch.tensor(v)[None], torch.tensor(f)[None]
    eye = nr.get_points_from_angles(2.732, 20, 30)
    ro, rg = oracle.NRRenderer(), nr.Renderer()
    for r in (ro, rg):
        r.image_size, r.eye, r.anti_aliasing = size, eye, aa
    tex = torch.rand(1, 12, 2, 2, 2, 3, generator=torch.Generator().manual_seed(0))
    Vo, To = V.clone().requires_grad_(True), tex.clone().requires_grad_(True)
    Vg, Tg = V.clone().to(dev).requires_grad_(True), tex.clone().to(dev).requires_grad_(True)
    so, do, co = ro.render_silhouettes(Vo, Fc), ro.render_depth(Vo, Fc), ro.render(Vo, Fc, To)
    sg, dg, cg = rg.render_silhouettes(Vg, Fc.to(dev)), rg.render_depth(Vg, Fc.to(dev)), rg.render(Vg, Fc.to(dev), Tg)
    assert torch.equal(so, sg.cpu()) and so.sum() > 100
    assert rel_l2(dg, do) <= TOL and rel_l2(cg, co) <= TOL
    H, W = so.shape[-2:]
    wgt = torch.linspace(-1, 1, W)[None, None, :] * torch.linspace(0.5, 1.5, H)[None, :, None]
    ((so * wgt).sum() + (do * 0.01 * wgt).sum() + (co * wgt[:, None]).sum()).backward()
    wg = wgt.to(dev)
    ((sg * wg).sum() + (dg * 0.01 * wg).sum() + (cg * wg[:, None]).sum()).backward()
    assert rel_l2(Vg.grad, Vo.grad) <= TOL and rel_l2(Tg.grad, To.grad) <= TOL


@pytest.mark.parametrize('seed,B,F,H,W,ss', [(0, 1, 64, 40, 40, 1), (1, 2, 200, 33, 70, 2), (2, 3, 1500, 64, 96, 1),
                                             (3, 1, 3000, 70, 130, 2)])
def test_random_triangle_soup_face_index_exact(dev, oracle, seed, B, F, H, W, ss):
    """Ragged sizes (not multiples of the 64x32 tile), B > 1, more faces than the per-tile queue holds."""
    rng = np.random.RandomState(seed)
    faces = random_faces(rng, B, F, size=0.25 if F > 1000 else 0.5)
    tex = rng.uniform(0, 1, (B, F, 2, 2, 2, 3)).astype(np.float32)
    st = oracle.raster_forward(faces, tex, H * ss, W * ss, 0.1, 100, 1e-3, (0, 0, 0), True, True, True)
    r = _sdn_raster(faces, tex, H, W, ss, 0.1, 100, 1e-3, (0, 0, 0), (1, 1, 1), dev)
    assert np.array_equal(r['face_index'], st.face_index_map)
    assert (st.face_index_map >= 0).mean() > 0.3
    fo, to = torch.tensor(faces, requires_grad=True), torch.tensor(tex, requires_grad=True)
    out = oracle.rasterize_rgbad(fo, to, (H, W), ss == 2, 0.1, 100, 1e-3, (0, 0, 0))
    assert torch.equal(out['alpha'], r['alpha'].cpu())
    assert rel_l2(r['depth'], out['depth']) <= TOL and rel_l2(r['rgb'], out['rgb']) <= TOL
    gen = torch.Generator().manual_seed(seed)
    ga, gd, gc = (torch.randn(out[k].shape, generator=gen) for k in ('alpha', 'depth', 'rgb'))
    ((out['alpha'] * ga).sum() + (out['depth'] * gd).sum() + (out['rgb'] * gc).sum()).backward()
    ((r['alpha'] * ga.to(dev)).sum() + (r['depth'] * gd.to(dev)).sum() + (r['rgb'] * gc.to(dev)).sum()).backward()
    assert rel_l2(r['faces'].grad, fo.grad) <= TOL
    assert rel_l2(r['textures'].grad, to.grad) <= TOL


@pytest.mark.parametrize('rt', ['Silhouette', 'Depth', 'Normal'])
def test_derender3d_bridge(dev, oracle, rt):
    """derender3d.models.renderer.Renderer (x mirror, fixed camera, normal-x negation) at the reference's call shape."""
    from sdn_b200 import workloads as wl
    from derender3d.models.renderer import Renderer, RenderType
    H, W, focal = 72, 200, 150.0
    verts, faces = wl.vkitti_scene_batch(3, seed=11, H=H, W=W, focal=focal, z_range=(4.0, 9.0))
    fb = faces[None].expand(3, -1, -1).contiguous()
    angle = wl.vkitti_viewing_angle(H, focal)
    t = getattr(RenderType, rt)
    vo, vg = verts.clone().requires_grad_(True), verts.clone().to(dev).requires_grad_(True)
    io = oracle.derender_render(vo, fb, (H, W), angle, t)
    ig = Renderer(image_size=(H, W), viewing_angle=angle)(vg, fb.to(dev).int(), render_type=t)
    assert ig.shape == io.shape and (io != io[0, :, 0, 0].view(1, -1, 1, 1)).float().mean() > 0.02
    if rt == 'Silhouette':
        assert torch.equal(ig.cpu(), io)
    assert rel_l2(ig, io) <= TOL
    g = torch.randn(io.shape, generator=torch.Generator().manual_seed(1))
    (io * g).sum().backward()
    (ig * g.to(dev)).sum().backward()
    assert rel_l2(vg.grad, vo.grad) <= TOL


def test_multi_output_equals_three_reference_passes(dev, oracle):
    """forward_multi: one rasterisation; gradient = sum of the three reference graphs (eps 1e-4 / 1e-4 / 1e-3)."""
    from sdn_b200 import workloads as wl
    from derender3d.models.renderer import Renderer, RenderType
    H, W, focal = 64, 64, 80.0
    verts, faces = wl.vkitti_scene_batch(2, seed=4, H=H, W=W, focal=focal, z_range=(5.0, 8.0))
    fb = faces[None].expand(2, -1, -1).contiguous()
    angle = wl.vkitti_viewing_angle(H, focal)
    vg = verts.clone().to(dev).requires_grad_(True)
    out = Renderer(image_size=H, viewing_angle=angle).forward_multi(vg, fb.to(dev).int())
    gen = torch.Generator().manual_seed(1)
    gs, gd, gn = (torch.randn(out[k].shape, generator=gen) for k in ('silhouette', 'depth', 'normal'))
    ((out['silhouette'] * gs.to(dev)).sum() + (out['depth'] * gd.to(dev)).sum() +
     (out['normal'] * gn.to(dev)).sum()).backward()
    vo = verts.clone().requires_grad_(True)
    so = oracle.derender_render(vo, fb, H, angle, RenderType.Silhouette)
    do = oracle.derender_render(vo, fb, H, angle, RenderType.Depth)
    no = oracle.derender_render(vo, fb, H, angle, RenderType.Normal)
    ((so * gs).sum() + (do * gd).sum() + (no * gn).sum()).backward()
    assert torch.equal(out['silhouette'].cpu(), so)
    assert rel_l2(out['depth'], do) <= TOL and rel_l2(out['normal'], no) <= TOL
    assert rel_l2(vg.grad, vo.grad) <= TOL
    # fast_math trades the IEEE divisions of the distance term for reciprocals: still inside the bar
    vf = verts.clone().to(dev).requires_grad_(True)
    of = Renderer(image_size=H, viewing_angle=angle).forward_multi(vf, fb.to(dev).int(), fast_math=True)
    ((of['silhouette'] * gs.to(dev)).sum() + (of['depth'] * gd.to(dev)).sum() +
     (of['normal'] * gn.to(dev)).sum()).backward()
    assert rel_l2(vf.grad, vo.grad) <= TOL


def test_full_size_properties_config2(dev):
    """BASELINE config 2 shape (375x1242, 2x SSAA): size-independent properties instead of the O(S^2 F) oracle."""
    from sdn_b200 import workloads as wl
    from derender3d.models.renderer import Renderer
    B = 4
    verts, faces = wl.vkitti_scene_batch(B, seed=0)
    angle = wl.vkitti_viewing_angle()
    r = Renderer(image_size=(wl.VKITTI_H, wl.VKITTI_W), viewing_angle=angle)
    vg = verts.to(dev).requires_grad_(True)
    fg = faces.to(dev).int()
    a = r.forward_multi(vg, fg)
    fi1 = r.last_face_index.clone()
    b = r.forward_multi(vg, fg)
    assert torch.equal(fi1, r.last_face_index)                                   # deterministic
    for k in ('silhouette', 'depth', 'normal'):
        assert torch.equal(a[k], b[k])
    sil = a['silhouette']
    assert sil.shape == (B, 1, 375, 1242)
    assert set(torch.unique(sil).tolist()) <= {0.0, 0.25, 0.5, 0.75, 1.0}
    cov = (fi1 >= 0).view(B, 375, 2, 1242, 2).float().mean(dim=(2, 4))            # raster rows: flip to compare
    assert torch.equal(cov.flip(1), sil[:, 0])
    assert int(fi1.max()) < 2 * faces.shape[0] and int(fi1.min()) == -1
    d = a['depth']
    assert float(d.max()) == 100.0 and float(d.min()) > 5.0
    bgmask = sil[:, 0] == 0
    assert torch.all(d[:, 0][bgmask] == 100.0) and torch.all(a['normal'].abs().sum(1)[bgmask] == 0)
    n = a['normal']
    full = sil[:, 0] == 1
    assert torch.all((n[:, :, :, :].pow(2).sum(1)[full]) <= 1.0 + 1e-4)
    (sil.sum() + d.sum() * 1e-3 + n.sum()).backward()
    assert torch.isfinite(vg.grad).all() and float(vg.grad.abs().sum()) > 0


@pytest.mark.parametrize('fast', [0, 1, 2])
def test_full_size_single_image_against_oracle(dev, oracle, fast, monkeypatch):
    """One 375x1242 image end to end against the oracle (a few seconds of host time).  fast=1 is the packed direct line
    sweep, fast=2 the silhouette-only far-field sweep (block moments) on full-length rows and columns."""
    import sdn_b200
    monkeypatch.setattr(sdn_b200, 'DEFAULT_FAST_MATH', fast)
    from sdn_b200 import workloads as wl
    from derender3d.models.renderer import Renderer, RenderType
    verts, faces = wl.vkitti_scene_batch(1, seed=9, z_range=(7.0, 9.0))
    fb = faces[None].contiguous()
    angle = wl.vkitti_viewing_angle()
    size = (wl.VKITTI_H, wl.VKITTI_W)
    vg = verts.clone().to(dev).requires_grad_(True)
    ig = Renderer(image_size=size, viewing_angle=angle)(vg, fb.to(dev).int(), render_type=RenderType.Silhouette)
    vo = verts.clone().requires_grad_(True)
    io = oracle.derender_render(vo, fb, size, angle, RenderType.Silhouette)
    assert torch.equal(ig.cpu(), io) and io.sum() > 1000
    g = torch.randn(io.shape, generator=torch.Generator().manual_seed(2))
    (io * g).sum().backward()
    (ig * g.to(dev)).sum().backward()
    assert rel_l2(vg.grad, vo.grad) <= TOL


@pytest.mark.parametrize('outputs', ['sdn', 's'], ids=['config2', 'config2b'])
def test_full_size_default_mode_against_oracle(dev, oracle, outputs):
    """The benchmarked path IS the tested path: one 375x1242 image, 2x SSAA, in the library's default backward mode
    (what bench.py times), through forward_multi exactly as bench.py calls it.  config2 = silhouette + depth + normal
    (the sum of the reference's three graphs, eps 1e-4 / 1e-4 / 1e-3); config2b = silhouette only (the reference's
    training path, derender3d/models/__init__.py:203-208).  Face index bit-exact, maps and vertex gradients <= 1e-4."""
    import sdn_b200
    assert sdn_b200.DEFAULT_FAST_MATH == 1
    from sdn_b200 import workloads as wl
    from derender3d.models.renderer import Renderer, RenderType
    verts, faces = wl.vkitti_scene_batch(1, seed=13, z_range=(6.5, 8.0))
    fb = faces[None].contiguous()
    angle = wl.vkitti_viewing_angle()
    size = (wl.VKITTI_H, wl.VKITTI_W)
    want = {'silhouette': 's' in outputs, 'depth': 'd' in outputs, 'normal': 'n' in outputs}
    vg = verts.clone().to(dev).requires_grad_(True)
    out = Renderer(image_size=size, viewing_angle=angle).forward_multi(vg, fb.to(dev).int(), **want)
    vo = verts.clone().requires_grad_(True)
    gen = torch.Generator().manual_seed(5)
    lo = lg = 0
    for key, rt in (('silhouette', RenderType.Silhouette), ('depth', RenderType.Depth), ('normal', RenderType.Normal)):
        if not want[key]:
            continue
        io = oracle.derender_render(vo, fb, size, angle, rt)
        g = torch.randn(io.shape, generator=gen)
        if key == 'silhouette':
            assert torch.equal(out[key].cpu(), io) and float(io.sum()) > 1000
        assert rel_l2(out[key], io) <= TOL, key
        lo = lo + (io * g).sum()
        lg = lg + (out[key] * g.to(dev)).sum()
    lo.backward()
    lg.backward()
    assert float(vo.grad.abs().sum()) > 0
    assert rel_l2(vg.grad, vo.grad) <= TOL


def test_large_mesh_default_mode_against_oracle(dev, oracle):
    """BASELINE config 2c shape: a 58k-triangle closed mesh (116k effective faces with fill_back -- the size of the
    reference's ShapeNet templates) at the reference's own 384x384 render, default backward mode, three outputs."""
    from sdn_b200 import workloads as wl
    from derender3d.models.renderer import Renderer, RenderType
    v, f = wl.superellipsoid_mesh(172, 170, 0.6, 0.8)
    assert 57000 < f.shape[0] < 60000
    S, focal = 384, 500.0
    verts = torch.tensor(v * np.array([2.0, 0.8, 0.9], np.float32))[None].clone()
    c, s_ = np.cos(0.6), np.sin(0.6)
    R = torch.tensor([[c, 0, s_], [0, 1, 0], [-s_, 0, c]], dtype=torch.float32)
    verts = verts @ R.T + torch.tensor([0.2, -0.1, -5.0])
    fb = torch.tensor(f)[None].contiguous()
    angle = wl.vkitti_viewing_angle(S, focal)
    vg = verts.clone().to(dev).requires_grad_(True)
    out = Renderer(image_size=S, viewing_angle=angle).forward_multi(vg, fb.to(dev).int())
    vo = verts.clone().requires_grad_(True)
    gen = torch.Generator().manual_seed(6)
    lo = lg = 0
    for key, rt in (('silhouette', RenderType.Silhouette), ('depth', RenderType.Depth), ('normal', RenderType.Normal)):
        io = oracle.derender_render(vo, fb, S, angle, rt)
        g = torch.randn(io.shape, generator=gen)
        if key == 'silhouette':
            assert torch.equal(out[key].cpu(), io) and float(io.sum()) > 2000
        assert rel_l2(out[key], io) <= TOL, key
        lo = lo + (io * g).sum()
        lg = lg + (out[key] * g.to(dev)).sum()
    lo.backward()
    lg.backward()
    assert rel_l2(vg.grad, vo.grad) <= TOL


def test_errors(dev):
    import neural_renderer as nr
    import sdn_b200
    f = torch.zeros(1, 2, 3, 3, device=dev)
    with pytest.raises(Exception):
        nr.rasterize_rgbad(f, None, 8, True, return_rgb=False, return_alpha=False, return_depth=False)
    with pytest.raises(sdn_b200.SdnError):
        nr.rasterize_silhouettes(f.cpu(), 8)
    out = nr.rasterize_silhouettes(f, 8)          # degenerate faces draw nothing
    assert out.shape == (1, 8, 8) and float(out.abs().sum()) == 0


@pytest.mark.parametrize('flags', [(0, 1, 0), (1, 0, 0), (1, 1, 0), (1, 1, 1), (0, 1, 1)])
@pytest.mark.parametrize('ss', [1, 2])
def test_fast_math_line_sweep_backward(dev, oracle, flags, ss, monkeypatch):
    """The line-sweep backward (fast path: deferred out-sweeps, one reciprocal per pixel/query pair) against the
    oracle in every mode: silhouette only, rgb only, one combined pass, with depth."""
    import sdn_b200
    monkeypatch.setattr(sdn_b200, 'DEFAULT_FAST_MATH', True)
    rng = np.random.RandomState(17)
    B, F, H, W = 2, 300, 41, 90
    faces = random_faces(rng, B, F, size=0.3)
    tex = rng.uniform(0, 1, (B, F, 2, 2, 2, 3)).astype(np.float32)
    rgb, alpha, depth = [bool(x) for x in flags]
    r = _sdn_raster(faces, tex, H, W, ss, 0.1, 100, 1e-3, (0.2, 0.1, 0.3), flags, dev)
    fo, to = torch.tensor(faces, requires_grad=True), torch.tensor(tex, requires_grad=True)
    out = oracle.rasterize_rgbad(fo, to if rgb else None, (H, W), ss == 2, 0.1, 100, 1e-3, (0.2, 0.1, 0.3),
                                 rgb, alpha, depth)
    gen = torch.Generator().manual_seed(3)
    lo = lg = 0
    for k, on in (('alpha', alpha), ('depth', depth), ('rgb', rgb)):
        if on:
            g = torch.randn(out[k].shape, generator=gen)
            lo = lo + (out[k] * g).sum()
            lg = lg + (r[k] * g.to(dev)).sum()
    lo.backward()
    lg.backward()
    assert rel_l2(r['faces'].grad, fo.grad) <= TOL
    assert float(fo.grad.abs().sum()) > 0


def _synthetic_objs():
    from sdn_b200 import workloads as wl
    from derender3d.models import ShapenetObj
    objs = []
    for k in range(8):
        v, f = wl.superellipsoid_mesh(10 + k, 8 + k, 0.4 + 0.1 * k, 0.5) if k % 2 else wl.cube_mesh()
        objs.append(ShapenetObj(vertices=v * np.array([1.0, 0.6, 0.5], np.float32), faces=f))
    return objs


def _synthetic_blob(B, dev, seed=0):
    g = torch.Generator().manual_seed(seed)
    c = 0.4 * (torch.rand(B, 2, generator=g) - 0.5)
    d = 0.25 + 0.2 * torch.rand(B, 2, generator=g)
    roi = torch.cat([c - d / 2, c + d / 2], dim=1)
    theta = torch.randn(B, 2, generator=g)
    blob = {
        '_mroi_norms': torch.stack([roi[:, 2] + roi[:, 0], roi[:, 3] + roi[:, 1]], dim=1) / 2.0,
        '_droi_norms': torch.stack([roi[:, 2] - roi[:, 0], roi[:, 3] - roi[:, 1]], dim=1),
        '_focals': torch.full((B, 1), 725.0),
        '_theta_deltas': theta / theta.norm(dim=1, keepdim=True),
        '_translation2ds': 0.05 * torch.randn(B, 2, generator=g),
        '_log_scales': 0.1 * torch.randn(B, 3, generator=g) + torch.tensor([1.3, 0.4, 0.5]),
        '_log_depths': 0.2 * torch.randn(B, 1, generator=g) + 2.0,
        '_class_probs': torch.softmax(3 * torch.randn(B, 8, generator=g), dim=1),
        '_ffd_coeffs': 0.03 * torch.randn(B, 8, 192, generator=g),
    }
    return {k: v.to(dev) for k, v in blob.items()}, roi.to(dev)


@pytest.mark.parametrize('training', [False, True])
def test_derenderer3d_batched_render_matches_per_object_oracle(dev, oracle, training):
    """Derenderer3d.render groups objects by mesh class and renders each group in one native call; the reference
    renders them one by one (derender3d/models/__init__.py:161-224).  Same masks / normals / depth maps."""
    from derender3d import TargetType
    from derender3d.models import Derenderer3d
    from derender3d.models.renderer import RenderType
    S, B = 48, 7
    model = Derenderer3d(TargetType.extend, image_size=256, render_size=S, objs=_synthetic_objs()).to(dev)
    model.train(training)
    model._force_no_sample = True
    blob, _ = _synthetic_blob(B, dev)
    blob['_ffd_coeffs'].requires_grad_(True)
    out = model.render(blob)
    assert out['_masks'].shape == (B, 1, S, S) and out['_normals'].shape == (B, 3, S, S)
    assert float(out['_masks'].sum()) > 20
    p = model.pose(blob)
    classes = p['_class_samples'].tolist()
    ang = (torch.atan(S / (2.0 * blob['_focals'].double().reshape(-1))) / np.pi * 180).cpu()
    for k, sel, verts, z in model.object_vertices(blob, p, classes):
        faces = model.objs[k].faces[None]
        for j, i in enumerate(sel.tolist()):
            v = verts[j:j + 1].detach().cpu()
            so = oracle.derender_render(v, faces, S, float(ang[i]), RenderType.Silhouette)
            assert torch.equal(out['_masks'][i:i + 1].cpu(), so), (k, i)
            do = oracle.derender_render(v, faces, S, float(ang[i]), RenderType.Depth)
            no = oracle.derender_render(v, faces, S, float(ang[i]), RenderType.Normal)
            assert rel_l2(out['_depth_maps'][i:i + 1], do) <= TOL and rel_l2(out['_normals'][i:i + 1], no) <= TOL
    (out['_masks'].sum() + out['_normals'].sum()).backward()
    assert torch.isfinite(blob['_ffd_coeffs'].grad).all() and float(blob['_ffd_coeffs'].grad.abs().sum()) > 0


def test_full_train_step_config3_shape(dev):
    """BASELINE config 3 in miniature: encoder -> render -> MSE + REINFORCE -> Adam, on synthetic targets."""
    from derender3d import TargetType
    from derender3d.models import Derenderer3d
    from derender3d.train_step import train_step
    torch.manual_seed(0)
    B, S = 6, 64
    model = Derenderer3d(TargetType.full, image_size=256, render_size=S, objs=_synthetic_objs()).to(dev).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=1e-3)
    blob, roi = _synthetic_blob(B, dev, seed=2)
    batch = {'images': torch.randn(B, 3, 224, 224, device=dev), 'roi_norms': roi,
             'focals': torch.full((B, 1), 725.0, device=dev),
             'targets': torch.full((B,), TargetType.full, device=dev, dtype=torch.long),
             'thetas': torch.rand(B, 1, device=dev), 'translation2ds': torch.zeros(B, 2, device=dev),
             'log_scales': torch.zeros(B, 3, device=dev), 'log_depths': torch.full((B, 1), 2.0, device=dev),
             'masks': (torch.rand(B, 1, 40, 40, device=dev) > 0.5).float(), 'ignores': torch.zeros(B, 1, 40, 40, device=dev)}
    before = model.derenderer._fc3.weight.detach().clone()
    loss, losses, out = train_step(model, opt, batch, TargetType.full)
    assert torch.isfinite(loss) and set(losses) >= {'mask_loss', 'class_reward', 'ffd_coeff_reg', 'scale_loss'}
    assert out['_masks'].shape == (B, 1, S, S)
    assert not torch.equal(before, model.derenderer._fc3.weight)


@pytest.mark.timeout(300)
def test_many_faces_per_tile_queue_flush_is_stable(dev, oracle):
    """Thousands of faces overlapping one 64x32 tile force repeated queue flushes inside k_raster_tiles; the flush
    decision guards barriers and must be CTA-uniform (regression test for a divergent-barrier hang)."""
    import sdn_b200
    rng = np.random.RandomState(5)
    B, F, H, W = 2, 6000, 48, 96
    faces = random_faces(rng, B, F, spread=0.5, size=0.35)
    st = oracle.raster_forward(faces, None, H, W, 0.1, 100, 1e-4, None, False, True, True)
    cfg = sdn_b200.RasterConfig((H, W), False, return_alpha=True, return_depth=True)
    f = torch.tensor(faces, device=dev)
    for _ in range(8):
        sdn_b200.RasterizeFunction.apply(f, None, cfg)
        assert np.array_equal(cfg.last_face_index.cpu().numpy(), st.face_index_map)


def test_analysis_by_synthesis_reduces_silhouette_error(dev):
    """main.py:405-459 in miniature: Adam on pose / FFD latents through render fwd+bwd drives the silhouettes towards
    targets rendered from perturbed latents."""
    import sdn_b200
    from derender3d import TargetType
    from derender3d.models import Derenderer3d
    from derender3d.optimise import refine
    S, B = 64, 4
    model = Derenderer3d(TargetType.full, image_size=256, render_size=S, objs=_synthetic_objs()).to(dev)
    model.train()
    model._force_no_sample = True
    blob, _ = _synthetic_blob(B, dev, seed=3)
    with torch.no_grad():
        target = model.render(blob)['_masks'].clone()
    start = dict(blob)
    start['_translation2ds'] = blob['_translation2ds'] + 0.08
    start['_log_scales'] = blob['_log_scales'] + 0.15
    old = sdn_b200.DEFAULT_FAST_MATH
    sdn_b200.DEFAULT_FAST_MATH = True
    try:
        refined, hist = refine(model, start, target, num_opts=25)
    finally:
        sdn_b200.DEFAULT_FAST_MATH = old
    assert float(hist[-1]) < 0.6 * float(hist[0]), hist.tolist()
    assert refined['_masks'].shape == target.shape


def _refine_setup(dev, seed=3, S=64, B=4):
    from derender3d import TargetType
    from derender3d.models import Derenderer3d
    model = Derenderer3d(TargetType.full, image_size=256, render_size=S, objs=_synthetic_objs()).to(dev)
    blob, _ = _synthetic_blob(B, dev, seed=seed)
    model.train()
    model._force_no_sample = True
    with torch.no_grad():
        target = model.render(blob)['_masks'].clone()
    start = dict(blob)
    start['_translation2ds'] = blob['_translation2ds'] + 0.08
    start['_log_scales'] = blob['_log_scales'] + 0.15
    return model, start, target


def test_analysis_by_synthesis_equals_the_reference_loop(dev):
    """optimise.refine against a literal restatement of geometric/scripts/main.py:405-459 on fixed latents: train mode
    with _force_no_sample, detached clones of the encoder outputs, Adam(lr 3e-2) over theta / translation2d / log-scale /
    FFD coefficients, loss = mean((mse(_masks, pad_like(masks), reduce=False) + 100 * mean(ffd^2)) * (1 - pad_like(ignores,
    replicate))).  Same renderer underneath, so the loss trajectories agree to float-sum reordering (the vertex scatter
    of sdn_faces_backward uses float atomics)."""
    from derender3d.optimise import refine
    from derender3d.train_step import pad_like
    model, start, target = _refine_setup(dev)
    g = torch.Generator().manual_seed(4)
    masks = torch.nn.functional.avg_pool2d(target, 2).round()             # GT masks arrive smaller than the render
    ignores = (torch.rand(masks.shape, generator=g) > 0.9).float().to(dev)
    n = 8
    # ---- the reference loop, restated (main.py:424-456) ----
    model.train()
    model._force_no_sample = True
    _blob = {k: (v.clone().detach() if isinstance(v, torch.Tensor) else v) for k, v in start.items()}
    opt_vars = [_blob[k].requires_grad_() for k in ('_theta_deltas', '_translation2ds', '_log_scales', '_ffd_coeffs')]
    optimizer = torch.optim.Adam(opt_vars, lr=3e-2)
    ref_hist = []
    for _ in range(n):
        optimizer.zero_grad()
        _blob.update(model.render(_blob))
        _masks = _blob['_masks']
        loss = torch.nn.functional.mse_loss(_masks, pad_like(masks, _masks), reduction='none') + \
            100 * torch.mean(_blob['_ffd_coeffs'] ** 2)
        loss = loss * (1 - pad_like(ignores, _masks, mode='replicate'))
        loss = torch.mean(loss)
        loss.backward()
        optimizer.step()
        ref_hist.append(float(loss))
    out_blob, hist = refine(model, start, masks, ignores, num_opts=n)
    hist = hist.cpu().tolist()
    assert abs(hist[0] - ref_hist[0]) <= 1e-6 * abs(ref_hist[0])          # identical arithmetic before any update
    assert rel_l2(torch.tensor(hist[:3]), torch.tensor(ref_hist[:3])) <= 1e-4
    assert rel_l2(torch.tensor(hist), torch.tensor(ref_hist)) <= 2e-2
    assert hist[-1] < hist[0]
    for k in ('_translation2ds', '_log_scales', '_theta_deltas'):
        assert rel_l2(out_blob[k], _blob[k].detach()) <= 2e-2, k


def test_analysis_by_synthesis_as_cuda_graph_matches_eager(dev):
    """refine(graph=True): iteration 0 eager, the rest replays of one captured iteration; same loss trajectory.

    Why the trajectory bound is loose while the first steps are tight: forward and the per-face backward are
    deterministic, but the face -> vertex scatter (sdn_faces_backward) and the per-object reductions of
    sdn_objects_backward add floats with atomics, so two runs differ by ~1e-7 in the latent gradients.  Adam divides by
    sqrt(v): components whose gradient is ~0 amplify that noise into the update, and a silhouette is a step function of
    the latents -- once a boundary sample lands on the other side of an edge the loss moves by O(1 / pixels).  The first
    iterations (no such flip yet) agree to 1e-4; 12 iterations to 5e-2."""
    from derender3d.optimise import refine
    model, start, target = _refine_setup(dev)
    ref_blob, ref_hist = refine(model, start, target, num_opts=12)
    out_blob, out_hist = refine(model, start, target, num_opts=12, graph=True)
    assert out_hist.shape == ref_hist.shape == (12,)
    assert rel_l2(out_hist[:3].cpu().numpy(), ref_hist[:3].cpu().numpy()) <= 1e-4, (out_hist.tolist(), ref_hist.tolist())
    assert rel_l2(out_hist.cpu().numpy(), ref_hist.cpu().numpy()) <= 5e-2, (out_hist.tolist(), ref_hist.tolist())
    assert float(out_hist[-1]) < float(out_hist[0])
    for k in ('_translation2ds', '_log_scales'):
        assert rel_l2(out_blob[k].cpu().numpy(), ref_blob[k].cpu().numpy()) <= 5e-2, k


def test_pathological_faces_match_oracle(dev, oracle):
    """Edge cases the reference's kernels meet in the wild: faces far outside the frustum (|NDC| > 64: no bbox
    culling), faces crossing the camera plane (negative / tiny z), NaN and infinite vertices, zero-area faces,
    faces clipped by near / far.  Face-index map bit-exact, nothing crashes, gradients finite where the oracle's are."""
    import sdn_b200
    rng = np.random.RandomState(3)
    B, F, H, W = 1, 64, 40, 56
    faces = random_faces(rng, B, F, size=0.5)
    faces[0, 0] *= np.array([300.0, 300.0, 1.0], np.float32)          # huge coordinates
    faces[0, 1, :, 2] = np.array([-0.5, 1.0, 1.0], np.float32)         # one vertex behind the camera
    faces[0, 2, :, 2] = 1e-3                                           # nearer than near
    faces[0, 3, :, 2] = 500.0                                          # beyond far
    faces[0, 4, 1, 0] = np.nan
    faces[0, 5, 2, 1] = np.inf
    faces[0, 6, 1] = faces[0, 6, 0]                                    # zero area (two coincident vertices)
    faces[0, 7] = faces[0, 7, 0]                                       # a point
    faces[0, 8, :, :2] = np.array([[-3, -3], [3, -3], [0, 3]], np.float32)   # covers the whole frame
    st = oracle.raster_forward(faces, None, H, W, 0.1, 100, 1e-4, None, False, True, True)
    cfg = sdn_b200.RasterConfig((H, W), False, return_alpha=True, return_depth=True)
    f = torch.tensor(faces, device=dev, requires_grad=True)
    _, alpha, depth = sdn_b200.RasterizeFunction.apply(f, None, cfg)
    assert np.array_equal(cfg.last_face_index.cpu().numpy(), st.face_index_map)
    assert (st.face_index_map >= 0).mean() > 0.5
    g_a = torch.randn(alpha.shape, generator=torch.Generator().manual_seed(0))
    (alpha * g_a.to(dev)).sum().backward()
    gf, _ = oracle.raster_backward(st, None, g_a.flip(1).numpy().copy(), np.zeros((B, H, W), np.float32))
    fin = np.isfinite(gf).all(axis=(2, 3))[0]                          # faces whose oracle gradient is finite
    got = f.grad.cpu().numpy()[0]
    assert np.isfinite(got[fin]).all()
    assert rel_l2(torch.tensor(got[fin]), torch.tensor(gf[0][fin])) <= TOL


def test_batch_of_one_tiny_raster_and_single_face(dev, oracle):
    import sdn_b200
    faces = np.array([[[[-0.9, -0.8, 1.0], [0.8, -0.7, 2.0], [0.1, 0.9, 1.5]]]], np.float32)
    for (H, W, ss) in ((1, 1, 1), (2, 3, 2), (5, 1, 1)):
        st = oracle.raster_forward(faces, None, H * ss, W * ss, return_alpha=True)
        cfg = sdn_b200.RasterConfig((H, W), ss == 2, return_alpha=True)
        sdn_b200.RasterizeFunction.apply(torch.tensor(faces, device=dev), None, cfg)
        assert np.array_equal(cfg.last_face_index.cpu().numpy(), st.face_index_map), (H, W, ss)


@pytest.mark.parametrize('H,W,ss,bg', [(70, 131, 1, (0.2, 0.5, 0.9)), (66, 130, 2, (0.3, 0.3, 0.3)), (37, 129, 2, 'per_image'),
                                       (96, 200, 1, (0.0, 0.0, 0.0))])
def test_background_outside_the_tile_rectangle_and_resolve_inside(dev, oracle, H, W, ss, bg):
    """Every output element is written exactly by one of two kernels: k_fill_background outside each image's tile rectangle
    (rows / columns not multiples of the vector width, flipped output rows, three-channel colours), k_raster_tiles' resolve
    inside it (pairs of output pixels, 8- / 16-byte stores).  Small objects in a corner, one image without a visible face,
    non-uniform and per-image backgrounds; the allocator is primed with NaN-filled blocks of the output sizes so that an
    element nobody writes cannot pass by luck."""
    import sdn_b200
    rng = np.random.RandomState(H + W)
    B, F = 3, 40
    faces = random_faces(rng, B, F, size=0.2)
    faces[..., 0] = faces[..., 0] * 0.35 - 0.55        # a small cluster left of centre
    faces[..., 1] = faces[..., 1] * 0.35 + 0.45
    faces[1, :, :, 2] = -5.0                           # image 1: everything behind the camera
    tex = rng.uniform(0, 1, (B, F, 2, 2, 2, 3)).astype(np.float32)
    if bg == 'per_image':
        bg_t = torch.tensor([[0.1, 0.2, 0.3], [0.9, 0.8, 0.7], [0.4, 0.0, 0.6]])
        bg_o = bg_t.numpy()
    else:
        bg_t, bg_o = bg, bg
    for shape in ((B, 3, H, W), (B, H, W), (B, H, W), (B, H * ss, W * ss, 3)):     # poison the caching allocator's free blocks
        x = torch.full(shape, float('nan'), device=dev)
        del x
    poison = torch.full((B, H * ss, W * ss), 12345, device=dev, dtype=torch.int32)
    del poison
    st = oracle.raster_forward(faces, tex, H * ss, W * ss, 0.1, 100, 1e-3, bg_o if bg != 'per_image' else (0, 0, 0), True, True, True)
    r = _sdn_raster(faces, tex, H, W, ss, 0.1, 100, 1e-3, bg_t, (1, 1, 1), dev)
    assert np.array_equal(r['face_index'], st.face_index_map)
    assert (st.face_index_map[1] < 0).all() and 0.005 < (st.face_index_map[0] >= 0).mean() < 0.3
    for k in ('rgb', 'alpha', 'depth'):
        assert torch.isfinite(r[k]).all(), k
    if bg != 'per_image':
        out = oracle.rasterize_rgbad(torch.tensor(faces), torch.tensor(tex), (H, W), ss == 2, 0.1, 100, 1e-3, bg_o)
        assert torch.equal(out['alpha'], r['alpha'].cpu())
        assert rel_l2(r['depth'], out['depth']) <= TOL and rel_l2(r['rgb'], out['rgb']) <= TOL
    else:       # per-image background: one oracle call per image with that image's colour
        for b in range(B):
            out = oracle.rasterize_rgbad(torch.tensor(faces[b:b + 1]), torch.tensor(tex[b:b + 1]), (H, W), ss == 2, 0.1, 100, 1e-3,
                                         tuple(float(v) for v in bg_o[b]))
            assert torch.equal(out['alpha'][0], r['alpha'][b].cpu())
            assert rel_l2(r['rgb'][b], out['rgb'][0]) <= TOL and rel_l2(r['depth'][b], out['depth'][0]) <= TOL
        assert torch.allclose(r['rgb'][1].cpu(), bg_t[1][:, None, None].expand(3, H, W))

"""sdn_composite_scene (two launches per frame) against the reference's per-object torchvision + Pillow loop
(oracle/compositor.py `composite_reference`, a literal restatement of geometric/scripts/main.py:547-606).
Integer / byte work: the instance map and the 8-bit normal map are bit-exact; the float depth map is bit-exact too
(double accumulation in the same tap order)."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available()
    import sdn_b200
    sdn_b200.lib()
    return torch.device('cuda:0')


def _scene(N, S, seed, H, W, focal):
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, S), torch.linspace(-1, 1, S), indexing='ij')
    masks, normals, dmaps = [], [], []
    for _ in range(N):
        a, b = 0.4 + 0.5 * torch.rand(2, generator=g)
        r = (xx / a) ** 2 + (yy / b) ** 2
        masks.append(torch.clamp(1.5 - r, 0, 1).unsqueeze(0))                  # soft-edged ellipse in [0,1]
        normals.append(torch.rand(3, S, S, generator=g) * 2 - 1)
        dmaps.append(5 + 95 * torch.rand(1, S, S, generator=g))
    masks, normals, dmaps = torch.stack(masks), torch.stack(normals), torch.stack(dmaps)
    zooms = 0.4 + 2.6 * torch.rand(N, 1, generator=g)
    center2ds = torch.stack([(torch.rand(N, generator=g) - 0.5) * H / focal, (torch.rand(N, generator=g) - 0.5) * W / focal], dim=1)
    depths = 5 + 30 * torch.rand(N, 1, generator=g)
    interests = torch.ones(N, dtype=torch.uint8)
    if N > 2:
        interests[1] = 0
    return masks, normals, dmaps, zooms, center2ds, depths, interests


@pytest.mark.parametrize('N,S,H,W,seed', [(6, 64, 75, 200, 0), (9, 96, 120, 310, 1), (1, 32, 40, 40, 2), (3, 48, 375, 1242, 3)])
def test_composite_equals_reference_loop(dev, N, S, H, W, seed):
    from derender3d.compositor import compose_scene
    from oracle.compositor import composite_reference
    focal, u0, v0 = 100.0 if W < 1000 else 725.0, W / 2.0 - 0.5, H / 2.0 - 0.5
    sc = _scene(N, S, seed, H, W, focal)
    ref_i, ref_n, ref_d = composite_reference(*sc, H, W, focal, u0, v0, S)
    masks, normals, dmaps, zooms, center2ds, depths, interests = sc
    out_i, out_n, out_d = compose_scene(masks.to(dev), normals.to(dev), dmaps.to(dev), zooms.to(dev), center2ds.to(dev),
                                        depths.to(dev), interests, H, W, focal, u0, v0, S)
    assert float((ref_i > 0).float().mean()) > 0.002                     # something was drawn
    assert torch.equal(out_i.cpu(), ref_i)
    assert torch.equal(out_n.cpu(), ref_n)
    assert torch.equal(out_d.cpu(), ref_d)
    # without the normal plane
    o_i, o_n, o_d = compose_scene(masks.to(dev), None, dmaps.to(dev), zooms.to(dev), center2ds.to(dev), depths.to(dev),
                                  interests, H, W, focal, u0, v0, S)
    assert o_n is None and torch.equal(o_i.cpu(), ref_i) and torch.equal(o_d.cpu(), ref_d)


def test_composite_nothing_to_draw(dev):
    from derender3d.compositor import compose_scene
    sc = _scene(3, 32, 5, 30, 50, 100.0)
    masks, normals, dmaps, zooms, center2ds, depths, _ = sc
    none = torch.zeros(3, dtype=torch.uint8)
    i, n, d = compose_scene(masks.to(dev), normals.to(dev), dmaps.to(dev), zooms.to(dev), center2ds.to(dev),
                            depths.to(dev), none, 30, 50, 100.0, 24.5, 14.5, 32)
    assert float(i.abs().max()) == 0 and torch.all(n == 0.5) and torch.all(d == 1.0)
    with pytest.raises(Exception):
        compose_scene(masks, normals, dmaps, zooms, center2ds, depths, none, 30, 50, 100.0, 24.5, 14.5, 32)   # CPU tensors


@pytest.mark.parametrize('N,S,H,W,seed', [(7, 64, 75, 200, 11), (5, 48, 375, 1242, 12)])
def test_composite_default_test_mode_paints_uninteresting_objects(dev, N, S, H, W, seed):
    """`operations is None` (the reference's default test mode, main.py:597-604): an object that is not interesting is
    still blended into the instance map from its frame-resolution input mask, in far-to-near order -- it gets an
    instance id and occludes what lies behind it.  Normal / depth maps are unaffected by such objects."""
    from derender3d.compositor import compose_scene
    from oracle.compositor import composite_reference
    focal, u0, v0 = 100.0 if W < 1000 else 725.0, W / 2.0 - 0.5, H / 2.0 - 0.5
    masks, normals, dmaps, zooms, center2ds, depths, interests = _scene(N, S, seed, H, W, focal)
    g = torch.Generator().manual_seed(seed + 100)
    interests = (torch.rand(N, generator=g) > 0.4).to(torch.uint8)
    interests[0], interests[1] = 1, 0
    yy, xx = torch.meshgrid(torch.arange(H).float(), torch.arange(W).float(), indexing='ij')
    image_masks = []
    for i in range(N):          # ground-truth style 0/1 masks around each object's centre
        cy, cx = float(center2ds[i, 0]) * focal + v0, float(center2ds[i, 1]) * focal + u0
        r = 0.25 * S / float(zooms[i])
        image_masks.append((((yy - cy) / r) ** 2 + ((xx - cx) / (1.5 * r)) ** 2 <= 1).float()[None])
    image_masks = torch.stack(image_masks)
    sc = (masks, normals, dmaps, zooms, center2ds, depths, interests)
    ref_i, ref_n, ref_d = composite_reference(*sc, H, W, focal, u0, v0, S, image_masks=image_masks,
                                              operations_given=False)
    edit_i, edit_n, edit_d = composite_reference(*sc, H, W, focal, u0, v0, S)
    assert not torch.equal(ref_i, edit_i)                                 # the two modes differ on this scene
    assert torch.equal(ref_n, edit_n) and torch.equal(ref_d, edit_d)
    out_i, out_n, out_d = compose_scene(masks.to(dev), normals.to(dev), dmaps.to(dev), zooms.to(dev), center2ds.to(dev),
                                        depths.to(dev), interests, H, W, focal, u0, v0, S,
                                        image_masks=image_masks.to(dev), operations_given=False)
    assert torch.equal(out_i.cpu(), ref_i)
    assert torch.equal(out_n.cpu(), ref_n) and torch.equal(out_d.cpu(), ref_d)
    ids = set(torch.unique(out_i).tolist())
    assert any((i + 1.0) in ids for i in range(N) if not interests[i])    # an uninteresting object owns pixels
    with pytest.raises(Exception):
        compose_scene(masks.to(dev), normals.to(dev), dmaps.to(dev), zooms.to(dev), center2ds.to(dev), depths.to(dev),
                      interests, H, W, focal, u0, v0, S, operations_given=False)      # needs image_masks

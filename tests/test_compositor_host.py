"""Pins the third-party arithmetic of the test-time compositor (geometric/scripts/main.py:547-606 -> torchvision ->
Pillow's bilinear resize): the numpy restatement in oracle/compositor.py equals Pillow bit for bit (uint8 and 'F')."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.mark.parametrize('S,out', [(64, 23), (64, 64), (64, 100), (96, 17), (50, 49), (48, 7), (32, 200), (384, 131)])
def test_numpy_restatement_equals_pillow(S, out):
    from PIL import Image
    from oracle.compositor import pil_resize_f32, pil_resize_u8
    rng = np.random.default_rng(S * 1000 + out)
    a = rng.integers(0, 256, (S, S), dtype=np.uint8)
    ref = np.asarray(Image.fromarray(a).resize((out, out), Image.BILINEAR))
    assert np.array_equal(ref, pil_resize_u8(a, out))
    f = rng.random((S, S), dtype=np.float32)
    reff = np.asarray(Image.fromarray(f).resize((out, out), Image.BILINEAR))
    assert np.array_equal(reff, pil_resize_f32(f, out))


@pytest.mark.parametrize('S,out', [(64, 23), (64, 64), (64, 100), (384, 131), (48, 7), (384, 1)])
def test_vectorised_tables_equal_the_loop_restatement(S, out):
    sys.path.insert(0, os.path.join(ROOT, '3d-sdn_b200'))
    from derender3d.compositor import resample_tables
    from oracle.compositor import pil_coeffs
    b0, kd0, k80 = pil_coeffs(S, out)
    b1, kd1, k81 = resample_tables(S, out)
    assert np.array_equal(b0, b1) and np.array_equal(kd0, kd1) and np.array_equal(k80, k81)


def test_scene_files_round_trip(tmp_path):
    import torch
    from PIL import Image
    sys.path.insert(0, os.path.join(ROOT, '3d-sdn_b200'))
    from derender3d.compositor import save_scene, scene_json
    inst = torch.zeros(1, 12, 20); inst[0, 2:5, 3:9] = 3
    nrm = torch.full((3, 12, 20), 0.5); dep = torch.full((1, 12, 20), 1.0); dep[0, 2:5, 3:9] = 0.25
    js = scene_json([2], [7, 7, 5], [9.0, 8.0, 7.5], [0.1, 0.2, 0.3])
    save_scene(str(tmp_path), '00001', inst, nrm, dep, js)
    assert np.asarray(Image.open(tmp_path / '00001.png')).max() == 3
    assert np.asarray(Image.open(tmp_path / '00001-normal.png')).shape == (12, 20, 3)
    d = np.asarray(Image.open(tmp_path / '00001-depth.png'))
    assert d.dtype in (np.uint16, np.int32) and int(d[3, 4]) == int(0.25 * 65535) and int(d[0, 0]) == 65535
    import json
    assert json.load(open(tmp_path / '00001.json'))['3'] == {'class_id': 5, 'depth': 7.5, 'alpha': 0.3}


def test_placement_matches_the_reference_expressions():
    """Draw order, on-screen size and paste box of main.py:538,556-565 (float32 tensor arithmetic, int() truncation)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, '3d-sdn_b200'))
    from derender3d.compositor import placement
    g = torch.Generator().manual_seed(0)
    N = 40
    zooms = 0.3 + 3 * torch.rand(N, 1, generator=g)
    c2d = (torch.rand(N, 2, generator=g) - 0.5) * torch.tensor([0.5, 1.7])
    depths = 5 + 30 * torch.rand(N, 1, generator=g)
    interests = torch.rand(N, generator=g) > 0.2
    focal, u0, v0, rs = 725.0, 620.5, 187.0, 384
    got = placement(zooms, c2d, depths, interests, focal, u0, v0, rs)
    want = []
    for i in torch.sort(depths[:, 0], dim=0, descending=True)[1].tolist():        # the reference's loop, verbatim
        if interests[i]:
            size = int(rs / zooms[i])
            want.append((i, size, int(c2d[i, 1] * focal + u0 - size // 2), int(c2d[i, 0] * focal + v0 - size // 2)))
    assert got == want and len(got) == int(interests.sum())
    assert [depths[i, 0] for i, *_ in got] == sorted([depths[i, 0] for i, *_ in got], reverse=True)


def test_random_sizes_against_pillow_through_the_product_tables():
    """Property check over random (in, out) sizes: resizing with the tables the product uploads to the GPU
    (derender3d.compositor.resample_tables) reproduces Pillow for the 8-bit and the float plane."""
    from PIL import Image
    sys.path.insert(0, os.path.join(ROOT, '3d-sdn_b200'))
    from derender3d.compositor import PRECISION_BITS, resample_tables
    rng = np.random.default_rng(7)
    for _ in range(25):
        S, out = int(rng.integers(8, 97)), int(rng.integers(1, 161))
        b, kd, k8 = resample_tables(S, out)
        a = rng.integers(0, 256, (S, S), dtype=np.uint8)
        t = np.zeros((S, out), np.uint8)
        for xx in range(out):
            x0, n = b[xx]
            t[:, xx] = np.clip(((1 << (PRECISION_BITS - 1)) + (a[:, x0:x0 + n].astype(np.int64) * k8[xx, :n]).sum(1)) >> PRECISION_BITS, 0, 255)
        o = np.zeros((out, out), np.uint8)
        for yy in range(out):
            y0, n = b[yy]
            o[yy] = np.clip(((1 << (PRECISION_BITS - 1)) + (t[y0:y0 + n].astype(np.int64) * k8[yy, :n, None]).sum(0)) >> PRECISION_BITS, 0, 255)
        assert np.array_equal(o, np.asarray(Image.fromarray(a).resize((out, out), Image.BILINEAR))), (S, out)
        f = rng.random((S, S), dtype=np.float32)
        tf = np.zeros((S, out), np.float32)
        for xx in range(out):
            x0, n = b[xx]
            acc = np.zeros(S, np.float64)
            for j in range(n):
                acc += f[:, x0 + j].astype(np.float64) * kd[xx, j]
            tf[:, xx] = acc.astype(np.float32)
        of = np.zeros((out, out), np.float32)
        for yy in range(out):
            y0, n = b[yy]
            acc = np.zeros(out, np.float64)
            for j in range(n):
                acc += tf[y0 + j].astype(np.float64) * kd[yy, j]
            of[yy] = acc.astype(np.float32)
        assert np.array_equal(of, np.asarray(Image.fromarray(f).resize((out, out), Image.BILINEAR))), (S, out)

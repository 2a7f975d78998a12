"""Known-answer tests for the oracle (the reference has none for this path): analytic cases whose
answers do not depend on any implementation."""
import numpy as np
import torch


def quad(z=2.0, x0=-0.5, x1=0.5, y0=-0.5, y1=0.5):
    """Two CCW triangles forming an axis-aligned NDC rectangle at constant depth."""
    a, b, c, d = (x0, y0, z), (x1, y0, z), (x1, y1, z), (x0, y1, z)
    return np.array([[[a, b, c], [a, c, d]]], np.float32)


def test_fronto_parallel_quad_depth_and_count(oracle):
    S = 64
    st = oracle.raster_forward(quad(), None, S, S, return_alpha=True, return_depth=True)
    cov = st.face_index_map >= 0
    # sample centres (2i+1-S)/S in [-0.5, 0.5]: i in 16..47 -> 32 x 32 samples
    assert cov.sum() == 32 * 32
    assert np.all(st.depth_map[cov] == np.float32(2.0))
    assert np.all(st.depth_map[~cov] == np.float32(100))
    w = st.weight_map[cov]
    assert np.allclose(w.sum(-1), 1.0, atol=2e-7) and w.min() >= 0


def test_row_zero_is_bottom_and_output_is_flipped(oracle):
    f = torch.tensor(quad(y0=-1.0, y1=-0.5))
    st = oracle.raster_forward(f.numpy(), None, 32, 32, return_alpha=True)
    assert st.alpha_map[0, :8].sum() > 0 and st.alpha_map[0, 8:].sum() == 0        # raster row 0 = NDC y=-1
    out = oracle.rasterize_silhouettes(f, 32, anti_aliasing=False)
    assert out[0, -8:].sum() > 0 and out[0, :-8].sum() == 0                         # image: bottom rows


def test_ssaa_values_and_strict_depth_tie(oracle):
    f = torch.tensor(quad(x0=-0.51, x1=0.37, y0=-0.33, y1=0.61))
    a = oracle.rasterize_silhouettes(f, 32, anti_aliasing=True)
    assert set(np.unique(a.numpy()).tolist()) <= {0.0, 0.25, 0.5, 0.75, 1.0}
    # two coincident quads: strict '<' keeps the lower face index
    ff = np.concatenate([quad(), quad()], axis=1)
    st = oracle.raster_forward(ff, None, 32, 32, return_alpha=True)
    assert set(np.unique(st.face_index_map).tolist()) <= {-1, 0, 1}


def test_near_far_and_backface(oracle):
    st = oracle.raster_forward(quad(z=0.05), None, 16, 16, return_alpha=True)
    assert (st.face_index_map >= 0).sum() == 0                       # nearer than near=0.1
    st = oracle.raster_forward(quad(z=150.0), None, 16, 16, return_alpha=True)
    assert (st.face_index_map >= 0).sum() == 0                       # beyond far=100
    back = quad()[:, :, ::-1].copy()                                 # clockwise
    st = oracle.raster_forward(back, None, 16, 16, return_alpha=True)
    assert (st.face_index_map >= 0).sum() == 0
    both = torch.tensor(np.concatenate([back, back[:, :, ::-1]], axis=1))   # fill_back makes it visible again
    assert oracle.rasterize_silhouettes(both, 16, anti_aliasing=False).sum() == 64


def test_backward_zero_sets_and_determinism(oracle):
    f = quad()
    S = 32
    st = oracle.raster_forward(f, None, S, S, return_alpha=True)
    g = np.ones((1, S, S), np.float32)
    gf1, _ = oracle.raster_backward(st, None, g, None)
    gf2, _ = oracle.raster_backward(st, None, g, None)
    assert np.array_equal(gf1, gf2) and np.abs(gf1).sum() > 0
    assert np.all(gf1[..., 2] == 0)              # K6 never touches z (rasterize.py:648,653 write x/y only)
    gz, _ = oracle.raster_backward(st, None, np.zeros_like(g), None)
    assert np.all(gz == 0)                       # diff_grad <= 0 contributes nothing (rasterize.py:644,714)
    # a gradient that is non-zero only where alpha == alpha_in/out leaves (alpha - alpha_ref) = 0 everywhere
    # except across the silhouette: a quad rendered into a fully covered raster has no silhouette => zero
    full = quad(x0=-2, x1=2, y0=-2, y1=2)
    stf = oracle.raster_forward(full, None, 16, 16, return_alpha=True)
    assert (stf.face_index_map >= 0).all()
    gff, _ = oracle.raster_backward(stf, None, np.random.RandomState(0).randn(1, 16, 16).astype(np.float32), None)
    assert np.all(gff == 0)


def test_rectangular_reduces_to_square(oracle):
    rng = np.random.RandomState(0)
    from util import random_faces
    f = random_faces(rng, 1, 20)
    a = oracle.raster_forward(f, None, 24, 24, return_alpha=True)
    b = oracle.raster_forward(f, None, 24, 48, return_alpha=True)     # twice as many columns
    assert a.face_index_map.shape == (1, 24, 24) and b.face_index_map.shape == (1, 24, 48)
    assert abs((a.face_index_map >= 0).mean() - (b.face_index_map >= 0).mean()) < 0.05

"""CPU oracle for the 3D-SDN geometric render path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product (3d-sdn_b200/) never does.

It restates, on the CPU in IEEE float32 without FMA contraction,
  * the rasteriser kernels (C, oracle/nmr_oracle.c; each function cites rasterize.py),
  * the Chainer-level glue of geometric/neural_renderer: look.py:7-45, look_at.py:7-46,
    perspective.py:5-19, vertices_to_faces.py:4-21, cross.py:25-35, lighting.py:8-52,
    rasterize.py:897-974 (SSAA, transpose, row flip, 2x2 mean), renderer.py:38-110,
  * the torch bridge geometric/derender3d/models/renderer.py:19-127,216-272.

Third-party arithmetic that is NOT under /root/reference (chainer==4.1.0, cupy==4.1.0,
environment.yml:8-9) is restated from its published definition:
  cf.normalize(x, eps=1e-5)      y = x / (sqrt(sum_j x_j^2) + eps)          (axis=1)
  cf.average_pooling_2d(x,2,2)   mean of the 2x2 block (sum order is cuDNN's; the
                                 oracle uses ((a+b)+c)+d) * 0.25 -- tolerance only)
  cf.matmul(v, r, transb=True)   v'_i = sum_j v_j r_ij, left-to-right, no FMA
  cf.tan                         float32(tan(float64(x))) -- the reference's is CUDA tanf
Parity status: kernels PINNED against the reference's own kernel bodies compiled for
the host (oracle/build_ref.py -> oracle/_ref/, tests/test_oracle_vs_ref.py and the
vectors under tests/golden/).  The Chainer/CuPy glue above is UNPINNED (no copy of
those libraries exists here); DESIGN.md says so.

Rectangular rasters are this repo's extension: image_size may be
(H, W); viewing_angle is then the half-angle of the VERTICAL axis and
width_x = width_y * (W / H).  With H == W everything reduces to the reference.
"""
import ctypes
import math
import os
import subprocess

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

DEFAULT_NEAR = 0.1      # rasterize.py:9
DEFAULT_FAR = 100       # rasterize.py:10
DEFAULT_EPS = 1e-4      # rasterize.py:11

_f = ctypes.POINTER(ctypes.c_float)
_i = ctypes.POINTER(ctypes.c_int32)


def build(force=False):
    """gcc the C restatement into oracle/liboracle.so (building the checker is not using it)."""
    so = os.path.join(HERE, 'liboracle.so')
    src = os.path.join(HERE, 'nmr_oracle.c')
    if force or not os.path.isfile(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(['gcc', '-O2', '-ffp-contract=off', '-fno-fast-math', '-fopenmp', '-fPIC',
                               '-shared', '-o', so, src, '-lm'])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.orc_num_threads.restype = ctypes.c_int
    return _LIB


def ref_lib():
    """The reference's own kernel bodies compiled for the host, or None."""
    global _REF
    if _REF is None:
        so = os.path.join(HERE, '_ref', 'libnmr_ref.so')
        _REF = ctypes.CDLL(so) if os.path.isfile(so) else False
    return _REF or None


def num_threads():
    return lib().orc_num_threads()


def _fp(a):
    assert a.dtype == np.float32 and a.flags['C_CONTIGUOUS']
    return a.ctypes.data_as(_f)


def _ip(a):
    assert a.dtype == np.int32 and a.flags['C_CONTIGUOUS']
    return a.ctypes.data_as(_i)


def _hw(image_size):
    if isinstance(image_size, (tuple, list)):
        return int(image_size[0]), int(image_size[1])
    return int(image_size), int(image_size)


# ----------------------------------------------------------------------------------------------
# Rasterize.forward_gpu / backward_gpu  (rasterize.py:464-510, 846-886) on numpy arrays
# ----------------------------------------------------------------------------------------------
class RasterState(object):
    """The buffers a reference `Rasterize` instance owns (rasterize.py:39-58)."""
    pass


SCANLINE = False        # bench.py's reference arm: use the per-face scanline kernel 3D-SDN actually runs (K1u)


def raster_forward(faces, textures, H, W, near=DEFAULT_NEAR, far=DEFAULT_FAR, eps=DEFAULT_EPS,
                   background_color=(0, 0, 0), return_rgb=False, return_alpha=False, return_depth=False,
                   batch_bug=False, use_ref=False, scanline=None):
    """faces [B,F,3,3] f32, textures [B,F,ts,ts,ts,3] f32 or None; raster is H rows x W cols.
    use_ref=True runs the reference's own kernel bodies (square rasters only)."""
    if not any((return_rgb, return_alpha, return_depth)):
        raise Exception  # rasterize.py:25-27
    faces = np.ascontiguousarray(faces, np.float32)
    B, F = faces.shape[:2]
    st = RasterState()
    st.B, st.F, st.H, st.W = B, F, H, W
    st.near, st.far, st.eps = float(near), float(far), float(eps)
    st.return_rgb, st.return_alpha, st.return_depth = return_rgb, return_alpha, return_depth
    st.faces = faces
    st.face_index_map = -np.ones((B, H, W), np.int32)
    st.weight_map = np.zeros((B, H, W, 3), np.float32)
    st.depth_map = np.zeros((B, H, W), np.float32) + np.float32(far)
    st.face_inv_map = np.zeros((B, H, W, 3, 3), np.float32) if return_depth else np.zeros(1, np.float32)
    if return_rgb:
        textures = np.ascontiguousarray(textures, np.float32)
        st.textures = textures
        st.ts = ts = textures.shape[2]
        st.rgb_map = np.zeros((B, H, W, 3), np.float32)
        st.sampling_index_map = np.zeros((B, H, W, 8), np.int32)
        st.sampling_weight_map = np.zeros((B, H, W, 8), np.float32)
    else:
        st.ts = ts = 0
    faces_inv = np.zeros_like(faces)
    if use_ref:
        r = ref_lib()
        assert r is not None and H == W, 'reference kernels: square rasters only'
        cfg = (ctypes.c_int(F), ctypes.c_int(H), ctypes.c_int(ts), ctypes.c_double(st.near),
               ctypes.c_double(st.far), ctypes.c_double(st.eps), ctypes.c_int(int(return_rgb)),
               ctypes.c_int(int(return_alpha)), ctypes.c_int(int(return_depth)))
        st._cfg = cfg
        r.ref_k1a_faces_inv(ctypes.c_int(B * F), _fp(faces), _fp(faces_inv), *cfg)
        r.ref_k1b_face_index(ctypes.c_int(B * H * W), _fp(faces), _fp(faces_inv), _ip(st.face_index_map),
                             _fp(st.weight_map), _fp(st.depth_map), _fp(st.face_inv_map), *cfg)
        if return_rgb:
            r.ref_k2_texture_sampling(ctypes.c_int(B * H * W), _fp(faces), _fp(textures), _ip(st.face_index_map),
                                      _fp(st.weight_map), _fp(st.depth_map), _fp(st.rgb_map),
                                      _ip(st.sampling_index_map), _fp(st.sampling_weight_map), *cfg)
    elif (SCANLINE if scanline is None else scanline):
        L = lib()      # rasterize.py:102-236, the NEURAL_RENDERER_UNSAFE=1 kernel (comparator / timing only)
        L.orc_visibility_scanline(_fp(faces), ctypes.c_int(B), ctypes.c_int(F), ctypes.c_int(H), ctypes.c_int(W),
                                  ctypes.c_double(st.near), ctypes.c_double(st.far), ctypes.c_int(int(return_depth)),
                                  _ip(st.face_index_map), _fp(st.weight_map), _fp(st.depth_map), _fp(st.face_inv_map))
        if return_rgb:
            L.orc_texture_sampling(_fp(faces), _fp(textures), _ip(st.face_index_map), _fp(st.weight_map),
                                   _fp(st.depth_map), ctypes.c_int(B), ctypes.c_int(F), ctypes.c_int(H),
                                   ctypes.c_int(W), ctypes.c_int(ts), ctypes.c_double(st.eps),
                                   ctypes.c_int(int(batch_bug)), _fp(st.rgb_map), _ip(st.sampling_index_map),
                                   _fp(st.sampling_weight_map))
    else:
        L = lib()
        L.orc_face_inv(_fp(faces), ctypes.c_int(B * F), ctypes.c_int(H), ctypes.c_int(W), _fp(faces_inv))
        L.orc_visibility(_fp(faces), _fp(faces_inv), ctypes.c_int(B), ctypes.c_int(F), ctypes.c_int(H),
                         ctypes.c_int(W), ctypes.c_double(st.near), ctypes.c_double(st.far),
                         ctypes.c_int(int(return_depth)), _ip(st.face_index_map), _fp(st.weight_map),
                         _fp(st.depth_map), _fp(st.face_inv_map))
        if return_rgb:
            L.orc_texture_sampling(_fp(faces), _fp(textures), _ip(st.face_index_map), _fp(st.weight_map),
                                   _fp(st.depth_map), ctypes.c_int(B), ctypes.c_int(F), ctypes.c_int(H),
                                   ctypes.c_int(W), ctypes.c_int(ts), ctypes.c_double(st.eps),
                                   ctypes.c_int(int(batch_bug)), _fp(st.rgb_map), _ip(st.sampling_index_map),
                                   _fp(st.sampling_weight_map))
    st.faces_inv = faces_inv
    covered = st.face_index_map >= 0
    if return_rgb:      # rasterize.py:448-462
        bg = np.asarray(background_color, np.float32)
        mask = covered.astype(np.float32)[:, :, :, None]
        if bg.ndim == 1:
            st.rgb_map = st.rgb_map * mask + (1 - mask) * bg[None, None, None, :]
        else:
            st.rgb_map = st.rgb_map * mask + (1 - mask) * bg[:, None, None, :]
        st.rgb_map = np.ascontiguousarray(st.rgb_map, np.float32)
    else:
        st.rgb_map = np.zeros(1, np.float32)
    if return_alpha:    # rasterize.py:437-446
        st.alpha_map = covered.astype(np.float32)
    else:
        st.alpha_map = np.zeros(1, np.float32)
    st.use_ref = use_ref
    return st


def raster_backward(st, grad_rgb_map=None, grad_alpha_map=None, grad_depth_map=None):
    """rasterize.py:846-886 -> (grad_faces [B,F,3,3], grad_textures or None)."""
    B, F, H, W = st.B, st.F, st.H, st.W
    grad_faces = np.zeros((B, F, 3, 3), np.float32)
    dummy = np.zeros(1, np.float32)
    g_rgb = dummy
    if st.return_rgb:
        g_rgb = np.zeros_like(st.rgb_map) if grad_rgb_map is None else np.ascontiguousarray(grad_rgb_map, np.float32)
    g_alpha = dummy
    if st.return_alpha:
        g_alpha = (np.zeros_like(st.alpha_map) if grad_alpha_map is None
                   else np.ascontiguousarray(grad_alpha_map, np.float32))
    g_depth = dummy
    if st.return_depth:
        g_depth = (np.zeros_like(st.depth_map) if grad_depth_map is None
                   else np.ascontiguousarray(grad_depth_map, np.float32))
    grad_textures = np.zeros_like(st.textures) if st.return_rgb else None
    cI = ctypes.c_int
    if st.use_ref:
        r = ref_lib()
        cfg = st._cfg
        if st.return_rgb or st.return_alpha:
            r.ref_k6_backward_pixel(cI(B * F), _fp(st.faces), _ip(st.face_index_map), _fp(st.rgb_map),
                                    _fp(st.alpha_map), _fp(g_rgb), _fp(g_alpha), _fp(grad_faces), *cfg)
        if st.return_rgb:
            r.ref_k7_backward_textures(cI(B * H * W), _ip(st.face_index_map), _fp(st.sampling_weight_map),
                                       _ip(st.sampling_index_map), _fp(g_rgb), _fp(grad_textures), *cfg)
        if st.return_depth:
            r.ref_k8_backward_depth(cI(B * H * W), _fp(st.faces), _fp(st.depth_map), _ip(st.face_index_map),
                                    _fp(st.face_inv_map), _fp(st.weight_map), _fp(g_depth), _fp(grad_faces), *cfg)
    else:
        L = lib()
        L.orc_backward_pixel(_fp(st.faces), _ip(st.face_index_map), _fp(st.rgb_map), _fp(st.alpha_map),
                             _fp(g_rgb), _fp(g_alpha), cI(B), cI(F), cI(H), cI(W), ctypes.c_double(st.eps),
                             cI(int(st.return_rgb)), cI(int(st.return_alpha)), _fp(grad_faces))
        if st.return_rgb:
            L.orc_backward_textures(_ip(st.face_index_map), _fp(st.sampling_weight_map),
                                    _ip(st.sampling_index_map), _fp(g_rgb), cI(B), cI(F), cI(H), cI(W),
                                    cI(st.ts), _fp(grad_textures))
        if st.return_depth:
            L.orc_backward_depth(_fp(st.faces), _fp(st.depth_map), _ip(st.face_index_map), _fp(st.face_inv_map),
                                 _fp(st.weight_map), _fp(g_depth), cI(B), cI(F), cI(H), cI(W), _fp(grad_faces))
    return grad_faces, grad_textures


def raster_forward_unsafe_ref(faces, H, near=DEFAULT_NEAR, far=DEFAULT_FAR, return_depth=False):
    """The reference's per-face scanline kernel (rasterize.py:102-236) run serially on the host.
    Comparator / timing only: it is what 3D-SDN runs (scripts/env.sh:11), and it differs from the
    safe path on which samples count as inside (ceil/floor scanline bounds vs half-plane tests)."""
    r = ref_lib()
    assert r is not None
    faces = np.ascontiguousarray(faces, np.float32)
    B, F = faces.shape[:2]
    W = H
    face_index_map = -np.ones((B, H, W), np.int32)
    weight_map = np.zeros((B, H, W, 3), np.float32)
    depth_map = np.zeros((B, H, W), np.float32) + np.float32(far)
    face_inv_map = np.zeros((B, H, W, 3, 3), np.float32) if return_depth else np.zeros(1, np.float32)
    lock = np.zeros((B, H, W), np.int32)
    cfg = (ctypes.c_int(F), ctypes.c_int(H), ctypes.c_int(0), ctypes.c_double(float(near)),
           ctypes.c_double(float(far)), ctypes.c_double(0.0), ctypes.c_int(0), ctypes.c_int(1),
           ctypes.c_int(int(return_depth)))
    r.ref_k1u_unsafe_face_index(ctypes.c_int(B * F), _fp(faces), _ip(face_index_map), _fp(weight_map),
                                _fp(depth_map), _fp(face_inv_map), _ip(lock), *cfg)
    return face_index_map, weight_map, depth_map


# ----------------------------------------------------------------------------------------------
# torch-CPU autograd wrapper: rasterize_rgbad  (rasterize.py:897-974)
# ----------------------------------------------------------------------------------------------
class _Rasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, faces, textures, H, W, near, far, eps, background_color, return_rgb, return_alpha,
                return_depth, use_ref):
        st = raster_forward(faces.detach().numpy(), None if textures is None else textures.detach().numpy(),
                            H, W, near, far, eps, background_color, return_rgb, return_alpha, return_depth,
                            use_ref=use_ref)
        ctx.st = st
        ctx.has_tex = textures is not None
        rgb = torch.from_numpy(st.rgb_map) if return_rgb else torch.zeros(1)
        alpha = torch.from_numpy(st.alpha_map.copy()) if return_alpha else torch.zeros(1)
        depth = torch.from_numpy(st.depth_map.copy()) if return_depth else torch.zeros(1)
        ctx.mark_non_differentiable(*[t for t, r in ((rgb, return_rgb), (alpha, return_alpha),
                                                      (depth, return_depth)) if not r])
        return rgb, alpha, depth

    @staticmethod
    def backward(ctx, g_rgb, g_alpha, g_depth):
        st = ctx.st
        gf, gt = raster_backward(
            st,
            g_rgb.contiguous().numpy() if (st.return_rgb and g_rgb is not None) else None,
            g_alpha.contiguous().numpy() if (st.return_alpha and g_alpha is not None) else None,
            g_depth.contiguous().numpy() if (st.return_depth and g_depth is not None) else None)
        gt_t = torch.from_numpy(gt) if (ctx.has_tex and gt is not None) else None
        return (torch.from_numpy(gf), gt_t) + (None,) * 10


def _avg_pool_2x2(x):
    """cf.average_pooling_2d(x, 2, 2) on [..., 2h, 2w]."""
    s = x[..., 0::2, 0::2] + x[..., 0::2, 1::2]
    s = s + x[..., 1::2, 0::2]
    s = s + x[..., 1::2, 1::2]
    return s * 0.25


def rasterize_rgbad(faces, textures=None, image_size=256, anti_aliasing=True, near=DEFAULT_NEAR, far=DEFAULT_FAR,
                    eps=DEFAULT_EPS, background_color=(0, 0, 0), return_rgb=True, return_alpha=True,
                    return_depth=True, use_ref=False, return_state=False):
    """rasterize.py:897-974 on torch CPU tensors (differentiable w.r.t. faces / textures)."""
    H, W = _hw(image_size)
    s = 2 if anti_aliasing else 1
    fn = _Rasterize.apply
    rgb, alpha, depth = fn(faces, textures, H * s, W * s, near, far, eps, background_color, return_rgb,
                           return_alpha, return_depth, use_ref)
    if return_rgb:
        rgb = rgb.permute(0, 3, 1, 2).flip(2)
    if return_alpha:
        alpha = alpha.flip(1)
    if return_depth:
        depth = depth.flip(1)
    if anti_aliasing:
        if return_rgb:
            rgb = _avg_pool_2x2(rgb)
        if return_alpha:
            alpha = _avg_pool_2x2(alpha)
        if return_depth:
            depth = _avg_pool_2x2(depth)
    return {'rgb': rgb if return_rgb else None, 'alpha': alpha if return_alpha else None,
            'depth': depth if return_depth else None}


def rasterize(faces, textures, image_size=256, anti_aliasing=True, near=DEFAULT_NEAR, far=DEFAULT_FAR,
              eps=DEFAULT_EPS, background_color=(0, 0, 0), **kw):
    return rasterize_rgbad(faces, textures, image_size, anti_aliasing, near, far, eps, background_color,
                           True, False, False, **kw)['rgb']


def rasterize_silhouettes(faces, image_size=256, anti_aliasing=True, near=DEFAULT_NEAR, far=DEFAULT_FAR,
                          eps=DEFAULT_EPS, **kw):
    return rasterize_rgbad(faces, None, image_size, anti_aliasing, near, far, eps, None, False, True, False,
                           **kw)['alpha']


def rasterize_depth(faces, image_size=256, anti_aliasing=True, near=DEFAULT_NEAR, far=DEFAULT_FAR,
                    eps=DEFAULT_EPS, **kw):
    return rasterize_rgbad(faces, None, image_size, anti_aliasing, near, far, eps, None, False, False, True,
                           **kw)['depth']


# ----------------------------------------------------------------------------------------------
# Chainer-level glue in torch CPU float32, explicit op order (no matmul/cross library calls)
# ----------------------------------------------------------------------------------------------
def cross(a, b):
    """cross.py:25-35."""
    c0 = a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1]
    c1 = a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2]
    c2 = a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]
    return torch.stack([c0, c1, c2], dim=1)


def normalize(x, eps=1e-5):
    """chainer.functions.normalize (v4.1): x / (||x||_2 + eps), axis=1, sum left-to-right."""
    n2 = x[:, 0] * x[:, 0] + x[:, 1] * x[:, 1]
    n2 = n2 + x[:, 2] * x[:, 2]
    norm = torch.sqrt(n2) + np.float32(eps)
    return x / norm[:, None]


def _as_f32(x, like=None):
    if isinstance(x, torch.Tensor):
        return x.float()
    return torch.tensor(np.asarray(x, np.float32))


def camera_rotation(z_dir, up):
    """look.py:28-34 / look_at.py:29-35: rows x_axis, y_axis, z_axis -> [bs,3,3]."""
    z_axis = normalize(z_dir)
    x_axis = normalize(cross(up, z_axis))
    y_axis = normalize(cross(z_axis, x_axis))
    return torch.stack([x_axis, y_axis, z_axis], dim=1)


def apply_camera(vertices, eye, r):
    """look.py:38-43: (v - eye) @ r^T with a left-to-right 3-term sum."""
    v = vertices - eye[:, None, :]
    out = []
    for i in range(3):
        s = v[:, :, 0] * r[:, i, 0][:, None] + v[:, :, 1] * r[:, i, 1][:, None]
        s = s + v[:, :, 2] * r[:, i, 2][:, None]
        out.append(s)
    return torch.stack(out, dim=2)


def look(vertices, eye, direction=None, up=None):
    """look.py:7-45."""
    assert vertices.dim() == 3
    bs = vertices.shape[0]
    direction = _as_f32([0, 0, 1] if direction is None else direction)
    up = _as_f32([0, 1, 0] if up is None else up)
    eye = _as_f32(eye)
    if eye.dim() == 1:
        eye = eye[None, :]
    if direction.dim() == 1:
        direction = direction[None, :]
    if up.dim() == 1:
        up = up[None, :]
    r = camera_rotation(direction, up)
    return apply_camera(vertices, eye.expand(bs, 3), r.expand(bs, 3, 3))


def look_at(vertices, eye, at=None, up=None):
    """look_at.py:7-46."""
    assert vertices.dim() == 3
    bs = vertices.shape[0]
    at = _as_f32([0, 0, 0] if at is None else at)
    up = _as_f32([0, 1, 0] if up is None else up)
    eye = _as_f32(eye)
    if eye.dim() == 1:
        eye = eye[None, :].expand(bs, 3)
    if at.dim() == 1:
        at = at[None, :].expand(bs, 3)
    if up.dim() == 1:
        up = up[None, :].expand(bs, 3)
    r = camera_rotation(at - eye, up)
    return apply_camera(vertices, eye, r)


def perspective_width(angle):
    """perspective.py:10-13: tan(angle / 180. * 3.1416) in float32 (note the literal 3.1416).
    tan itself: float32(tan(float64(x))), shared verbatim with the product's host code."""
    a = np.float32(angle)
    a = np.float32(a / np.float32(180.))
    a = np.float32(a * np.float32(3.1416))
    return np.float32(math.tan(float(a)))


def perspective(vertices, angle=30., aspect=None):
    """perspective.py:5-19.  aspect = W/H (this repo's rectangular extension): width_x = width*aspect."""
    assert vertices.dim() == 3
    width_y = perspective_width(angle)
    width_x = width_y if aspect is None else np.float32(width_y * np.float32(aspect))
    z = vertices[:, :, 2]
    x = vertices[:, :, 0] / z / float(width_x)
    y = vertices[:, :, 1] / z / float(width_y)
    return torch.stack([x, y, z], dim=2)


def vertices_to_faces(vertices, faces):
    """vertices_to_faces.py:4-21 -> [bs, nf, 3, 3]."""
    assert vertices.dim() == 3 and faces.dim() == 3
    assert vertices.shape[0] == faces.shape[0] and vertices.shape[2] == 3 and faces.shape[2] == 3
    bs, nv = vertices.shape[:2]
    idx = faces.long() + (torch.arange(bs) * nv)[:, None, None]
    return vertices.reshape(bs * nv, 3)[idx]


def fill_back_faces(faces):
    """renderer.py:41: concat(faces, faces[:, :, ::-1]) along the face axis."""
    return torch.cat([faces, faces.flip(2)], dim=1)


def face_normals(faces_v):
    """derender3d/models/renderer.py:66-73 (and lighting.py:37-41): normalize(cross(v0-v1, v2-v1))."""
    bs, nf = faces_v.shape[:2]
    f = faces_v.reshape(bs * nf, 3, 3)
    v10 = f[:, 0] - f[:, 1]
    v12 = f[:, 2] - f[:, 1]
    return normalize(cross(v10, v12)).reshape(bs, nf, 3)


def lighting(faces, textures, intensity_ambient=0.5, intensity_directional=0.5, color_ambient=(1, 1, 1),
             color_directional=(1, 1, 1), direction=(0, 1, 0)):
    """lighting.py:8-52."""
    bs, nf = faces.shape[:2]
    color_ambient = _as_f32(color_ambient)
    color_directional = _as_f32(color_directional)
    direction = _as_f32(direction)
    if color_ambient.dim() == 1:
        color_ambient = color_ambient[None, :].expand(bs, 3)
    if color_directional.dim() == 1:
        color_directional = color_directional[None, :].expand(bs, 3)
    if direction.dim() == 1:
        direction = direction[None, :].expand(bs, 3)
    light = torch.zeros(bs, nf, 3)
    if intensity_ambient != 0:
        light = light + np.float32(intensity_ambient) * color_ambient[:, None, :]
    if intensity_directional != 0:
        normals = face_normals(faces)
        d = direction[:, None, :]
        cos = normals[:, :, 0] * d[:, :, 0] + normals[:, :, 1] * d[:, :, 1]
        cos = torch.relu(cos + normals[:, :, 2] * d[:, :, 2])
        light = light + np.float32(intensity_directional) * (color_directional[:, None, :] * cos[:, :, None])
    return textures * light[:, :, None, None, None, :]


class NRRenderer(object):
    """neural_renderer/renderer.py:11-110 (+ render_normal of derender3d/models/renderer.py:60-93)."""

    def __init__(self):
        self.image_size = 256
        self.anti_aliasing = True
        self.background_color = [0, 0, 0]
        self.fill_back = True
        self.perspective = True
        self.viewing_angle = 30
        self.eye = [0, 0, -(1. / math.tan(math.radians(self.viewing_angle)) + 1)]
        self.camera_mode = 'look_at'
        self.camera_direction = [0, 0, 1]
        self.up = None
        self.near = 0.1
        self.far = 100
        self.light_intensity_ambient = 0.5
        self.light_intensity_directional = 0.5
        self.light_color_ambient = [1, 1, 1]
        self.light_color_directional = [1, 1, 1]
        self.light_direction = [0, 1, 0]
        self.rasterizer_eps = 1e-3
        self.use_ref = False

    def _aspect(self):
        H, W = _hw(self.image_size)
        return None if H == W else W / H

    def _camera(self, vertices):
        if self.camera_mode == 'look_at':
            vertices = look_at(vertices, self.eye)
        elif self.camera_mode == 'look':
            vertices = look(vertices, self.eye, self.camera_direction, self.up)
        if self.perspective:
            vertices = perspective(vertices, angle=self.viewing_angle, aspect=self._aspect())
        return vertices

    def render_silhouettes(self, vertices, faces):
        if self.fill_back:
            faces = fill_back_faces(faces)
        f = vertices_to_faces(self._camera(vertices), faces)
        return rasterize_silhouettes(f, self.image_size, self.anti_aliasing, use_ref=self.use_ref)

    def render_depth(self, vertices, faces):
        if self.fill_back:
            faces = fill_back_faces(faces)
        f = vertices_to_faces(self._camera(vertices), faces)
        return rasterize_depth(f, self.image_size, self.anti_aliasing, use_ref=self.use_ref)

    def render_normal(self, vertices, faces):
        if self.fill_back:
            faces = fill_back_faces(faces)
        normals = face_normals(vertices_to_faces(vertices, faces))
        textures = normals[:, :, None, None, None, :].repeat(1, 1, 2, 2, 2, 1)
        f = vertices_to_faces(self._camera(vertices), faces)
        return rasterize(f, textures, self.image_size, self.anti_aliasing, self.near, self.far,
                         self.rasterizer_eps, self.background_color, use_ref=self.use_ref)

    def render(self, vertices, faces, textures):
        if self.fill_back:
            faces = fill_back_faces(faces)
            textures = torch.cat([textures, textures.permute(0, 1, 4, 3, 2, 5)], dim=1)
        faces_lighting = vertices_to_faces(vertices, faces)
        textures = lighting(faces_lighting, textures, self.light_intensity_ambient,
                            self.light_intensity_directional, self.light_color_ambient,
                            self.light_color_directional, self.light_direction)
        f = vertices_to_faces(self._camera(vertices), faces)
        return rasterize(f, textures, self.image_size, self.anti_aliasing, self.near, self.far,
                         self.rasterizer_eps, self.background_color, use_ref=self.use_ref)


class RenderType:       # derender3d/models/renderer.py:12-16
    RGB = 0
    Silhouette = 1
    Depth = 2
    Normal = 3


def derender_render(vertices, faces, image_size, viewing_angle, render_type, textures=None, use_ref=False):
    """derender3d/models/renderer.py:231-272 (Renderer.forward) + :153-196 (RenderFunction.forward):
    x-flip, look with eye 0 / direction (0,0,-1) / up (0,1,0), dispatch, normal x-negate."""
    r = NRRenderer()
    r.image_size = image_size
    r.viewing_angle = viewing_angle
    r.use_ref = use_ref
    vertices = vertices * torch.tensor([-1., 1., 1.])
    bs = vertices.shape[0]
    r.eye = torch.zeros(bs, 3)
    r.camera_mode = 'look'
    r.camera_direction = torch.tensor([0., 0., -1.]).expand(bs, 3)
    r.up = torch.tensor([0., 1., 0.]).expand(bs, 3)
    if render_type == RenderType.RGB:
        images = r.render(vertices, faces, textures)
    elif render_type == RenderType.Silhouette:
        images = r.render_silhouettes(vertices, faces)[:, None]
    elif render_type == RenderType.Depth:
        images = r.render_depth(vertices, faces)[:, None]
    elif render_type == RenderType.Normal:
        images = r.render_normal(vertices, faces)
        images = torch.stack([-images[:, 0], images[:, 1], images[:, 2]], dim=1)
    return images

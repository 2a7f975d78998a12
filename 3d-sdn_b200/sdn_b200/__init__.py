"""sdn_b200 -- thin host layer over the C ABI of include/sdn_b200.h (libsdn_b200.so, sm_100a).

PyTorch is used for device memory, streams and autograd plumbing only; every pixel is produced by the
hand-written CUDA library.  There is NO CPU or eager fallback: if the shared library is missing, or a tensor is
not on a CUDA device, the call raises.
"""
import ctypes
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libsdn_b200.so')

SDN_RGB, SDN_ALPHA, SDN_DEPTH = 1, 2, 4

# Backward distance arithmetic when a RasterConfig does not say.  1 (default, the mode bench.py times and the full-size
# parity tests check): line-sweep kernels, identical inclusion tests, distances through hardware reciprocals and
# well-separated block moments (<= 1e-5 relative on the gradient, inside north_star's 1e-4 bar).  0: IEEE divisions in
# the op order of rasterize.py:646-653 (one warp per (face, edge, axis) item; ~10x slower).  SDN_FAST_MATH=0 in the
# environment selects the IEEE path for a whole process.
DEFAULT_FAST_MATH = int(os.environ.get('SDN_FAST_MATH', '1') or 0)

_lib = None


class SdnError(RuntimeError):
    pass


class _Camera(ctypes.Structure):
    _fields_ = [('rot', ctypes.c_void_p), ('eye', ctypes.c_void_p), ('width_x', ctypes.c_void_p),
                ('width_y', ctypes.c_void_p), ('flip_x', ctypes.c_float)]


class _Raster(ctypes.Structure):
    _fields_ = [('B', ctypes.c_int), ('F', ctypes.c_int), ('H', ctypes.c_int), ('W', ctypes.c_int),
                ('ss', ctypes.c_int), ('flags', ctypes.c_int), ('texture_size', ctypes.c_int),
                ('batch_bug', ctypes.c_int), ('near', ctypes.c_double), ('far', ctypes.c_double),
                ('eps', ctypes.c_double), ('background', ctypes.c_float * 3), ('background_b', ctypes.c_void_p),
                ('faces', ctypes.c_void_p), ('textures', ctypes.c_void_p), ('rgb', ctypes.c_void_p),
                ('alpha', ctypes.c_void_p), ('depth', ctypes.c_void_p), ('face_index', ctypes.c_void_p),
                ('rgb_ss', ctypes.c_void_p), ('workspace', ctypes.c_void_p), ('workspace_bytes', ctypes.c_size_t)]


class _RasterGrad(ctypes.Structure):
    _fields_ = [('grad_rgb', ctypes.c_void_p), ('grad_alpha', ctypes.c_void_p), ('grad_depth', ctypes.c_void_p),
                ('grad_faces', ctypes.c_void_p), ('grad_textures', ctypes.c_void_p), ('fast_math', ctypes.c_int)]


class _Objects(ctypes.Structure):
    _fields_ = [('N', ctypes.c_int), ('K', ctypes.c_int), ('nv_max', ctypes.c_int), ('derive_zoom', ctypes.c_int),
                ('basis', ctypes.c_void_p), ('basis_off', ctypes.c_void_p), ('nv', ctypes.c_void_p),
                ('classes', ctypes.c_void_p), ('control', ctypes.c_void_p), ('scales', ctypes.c_void_p),
                ('rotations', ctypes.c_void_p), ('translations', ctypes.c_void_p), ('ptrans', ctypes.c_void_p),
                ('zooms', ctypes.c_void_p), ('zoom_tos', ctypes.c_void_p), ('zoom_key', ctypes.c_void_p),
                ('vertices', ctypes.c_void_p)]


class _Conv2d(ctypes.Structure):
    _fields_ = [('N', ctypes.c_int), ('H', ctypes.c_int), ('W', ctypes.c_int), ('Cin', ctypes.c_int),
                ('Cout', ctypes.c_int), ('KH', ctypes.c_int), ('KW', ctypes.c_int), ('stride', ctypes.c_int),
                ('pad', ctypes.c_int), ('relu', ctypes.c_int), ('input', ctypes.c_void_p), ('weight', ctypes.c_void_p),
                ('scale', ctypes.c_void_p), ('shift', ctypes.c_void_p), ('residual', ctypes.c_void_p),
                ('output', ctypes.c_void_p), ('out_h', ctypes.c_int), ('out_w', ctypes.c_int),
                ('out_stride_n', ctypes.c_longlong), ('out_stride_h', ctypes.c_longlong),
                ('out_stride_w', ctypes.c_longlong), ('stats', ctypes.c_void_p), ('stats_keep', ctypes.c_int)]


def lib():
    """Load libsdn_b200.so (built by __graft_entry__.build() / csrc/build.py).  Fails loudly when absent."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise SdnError('%s is missing: run `python __graft_entry__.py build` (nvcc, sm_100a). '
                           'There is no CPU fallback.' % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        L.sdn_version.restype = ctypes.c_int
        L.sdn_error_string.restype = ctypes.c_char_p
        L.sdn_error_string.argtypes = [ctypes.c_int]
        L.sdn_workspace_bytes.restype = ctypes.c_size_t
        L.sdn_workspace_bytes.argtypes = [ctypes.c_int] * 4
        L.sdn_launch_count.restype = ctypes.c_longlong
        L.sdn_timing_enable.argtypes = [ctypes.c_int]
        L.sdn_timing_read.argtypes = [ctypes.c_char_p, ctypes.c_int]
        vp, ci = ctypes.c_void_p, ctypes.c_int
        L.sdn_faces_forward.argtypes = [vp, vp, ci, ci, ci, ci, ci, ctypes.POINTER(_Camera), vp, vp, vp]
        L.sdn_faces_backward.argtypes = [vp, vp, vp, vp, ci, ci, ci, ci, ci, ctypes.POINTER(_Camera), vp, vp,
                                         ctypes.c_size_t, vp]
        L.sdn_rasterize_forward.argtypes = [ctypes.POINTER(_Raster), vp]
        L.sdn_linear_forward.argtypes = [vp, vp, vp, vp, ci, ci, ci, ci, vp]
        L.sdn_conv2d_forward.argtypes = [ctypes.POINTER(_Conv2d), vp]
        L.sdn_mlp_head_forward.argtypes = [vp, vp] + [ci] * 7 + [vp] * 8 + [vp] * 4 + [vp, vp]
        L.sdn_im2col_nhwc.argtypes = [vp, vp, ci, ci, ci, ci, ci, ci, ci, ci, ci, vp]
        L.sdn_conv2d_wgrad.argtypes = [vp, vp, vp, ci, ci, ci, ci, ci, ci, ci, ci, ci, ci, vp]
        L.sdn_affine_act_nhwc.argtypes = [vp, vp, vp, vp, vp, ctypes.c_longlong, ci, ci, vp]
        L.sdn_batch_norm_finalize.argtypes = [vp, ci, ci, ctypes.c_longlong, ctypes.c_float, ctypes.c_float, vp, vp, vp, vp, vp,
                                              vp, vp]
        L.sdn_pack_conv_weight_transposed.argtypes = [vp, vp, ci, ci, ci, ci, ci, vp]
        L.sdn_max_pool_nhwc.argtypes = [vp, vp, vp, ci, ci, ci, ci, ci, ci, ci, vp]
        L.sdn_max_pool_backward_nhwc.argtypes = [vp, vp, vp, ci, ci, ci, ci, ci, ci, ci, vp]
        L.sdn_batch_norm_backward_nhwc.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, ctypes.c_longlong, ci, ci, vp]
        L.sdn_instance_norm_nhwc.argtypes = [vp, vp, vp, vp, ci, ci, ci, ctypes.c_float, ci, vp]
        L.sdn_reflect_pad_nhwc.argtypes = [vp, vp, ci, ci, ci, ci, ci, ci, vp]
        L.sdn_instance_norm_apply_nhwc.argtypes = [vp, vp, vp, vp, ci, ci, ci, ctypes.c_float, ci, vp]
        L.sdn_instance_norm_apply_pad_nhwc.argtypes = [vp, vp, vp, vp, ci, ci, ci, ci, ctypes.c_float, ci, ci, vp]
        L.sdn_objects_forward.argtypes = [ctypes.POINTER(_Objects), vp]
        L.sdn_objects_backward.argtypes = [ctypes.POINTER(_Objects), vp, vp, vp, vp, vp, vp, vp, vp, ctypes.c_size_t, vp]
        L.sdn_rasterize_backward.argtypes = [ctypes.POINTER(_Raster), ctypes.POINTER(_RasterGrad), vp]
        L.sdn_rasterize_backward_split.argtypes = [ctypes.POINTER(_Raster), ctypes.POINTER(_RasterGrad),
                                                   ctypes.c_double, ctypes.c_double, vp]
        _lib = L
    return _lib


def launch_count():
    return int(lib().sdn_launch_count())


def timing_enable(on=True):
    """Bracket every kernel launch of the library with CUDA events (diagnostics; off by default)."""
    lib().sdn_timing_enable(int(bool(on)))


def timing_read():
    """-> {kernel_name: (total_ms, launches)} since the last read."""
    buf = ctypes.create_string_buffer(1 << 16)
    n = lib().sdn_timing_read(buf, len(buf))
    out = {}
    for line in buf.raw[:n].decode().splitlines():
        name, ms, cnt = line.split()
        out[name] = (float(ms), int(cnt))
    return out


def _check(rc):
    if rc != 0:
        raise SdnError('libsdn_b200: %s (code %d)' % (lib().sdn_error_string(rc).decode(), rc))


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise SdnError('sdn_b200 runs on CUDA tensors only (no CPU fallback); got a %s tensor' % t.device)


def _f32c(t):
    return None if t is None else t.detach().to(torch.float32).contiguous()


def hw(image_size):
    if isinstance(image_size, (tuple, list)):
        return int(image_size[0]), int(image_size[1])
    return int(image_size), int(image_size)


class Camera(object):
    """Per-image camera handed to the face kernels (device tensors or None)."""

    def __init__(self, rot=None, eye=None, width_x=None, width_y=None, flip_x=1.0):
        self.rot, self.eye = _f32c(rot), _f32c(eye)
        self.width_x, self.width_y = _f32c(width_x), _f32c(width_y)
        self.flip_x = float(flip_x)

    def c(self):
        return _Camera(_ptr(self.rot), _ptr(self.eye), _ptr(self.width_x), _ptr(self.width_y), self.flip_x)


# --------------------------------------------------------------------------------------------------------------
# vertices -> NDC faces (+ normals)
# --------------------------------------------------------------------------------------------------------------
class FacesFunction(torch.autograd.Function):
    """fill_back + camera + perspective + vertices_to_faces in one launch (sdn_faces_forward / _backward).

    forward(vertices [B,nv,3] f32, faces [B,nf,3] or [nf,3] i32, camera, fill_back, want_normals)
      -> face_ndc [B,F,3,3], face_normal [B,F,3] (zeros-size tensor when not requested)"""

    @staticmethod
    def forward(ctx, vertices, faces, camera, fill_back, want_normals):
        _need_cuda(vertices, faces)
        v = _f32c(vertices)
        f = faces.detach().to(torch.int32).contiguous()
        B, nv = v.shape[0], v.shape[1]
        batched = 1 if f.dim() == 3 else 0
        nf = f.shape[-2]
        if batched and f.shape[0] != B:
            raise SdnError('faces batch %d != vertices batch %d' % (f.shape[0], B))
        F = 2 * nf if fill_back else nf
        face_ndc = torch.empty(B, F, 3, 3, device=v.device, dtype=torch.float32)
        face_normal = torch.empty(B, F, 3, device=v.device, dtype=torch.float32) if want_normals else None
        cam = camera.c()
        with torch.cuda.device(v.device):
            _check(lib().sdn_faces_forward(_ptr(v), _ptr(f), B, nv, nf, batched, int(bool(fill_back)),
                                           ctypes.byref(cam), _ptr(face_ndc), _ptr(face_normal), _stream()))
        ctx.save_for_backward(v, f)
        ctx.camera, ctx.fill_back, ctx.dims = camera, int(bool(fill_back)), (B, nv, nf, batched)
        if face_normal is None:
            face_normal = torch.empty(0, device=v.device)
            ctx.mark_non_differentiable(face_normal)
        return face_ndc, face_normal

    @staticmethod
    def backward(ctx, g_ndc, g_nrm):
        v, f = ctx.saved_tensors
        B, nv, nf, batched = ctx.dims
        g_ndc = _f32c(g_ndc)
        g_nrm = _f32c(g_nrm) if (g_nrm is not None and g_nrm.numel() > 0) else None
        grad_v = torch.empty_like(v)
        scratch = torch.empty(2 * B * nv * 3, device=v.device, dtype=torch.float32)
        cam = ctx.camera.c()
        with torch.cuda.device(v.device):
            _check(lib().sdn_faces_backward(_ptr(g_ndc), _ptr(g_nrm), _ptr(v), _ptr(f), B, nv, nf, batched,
                                            ctx.fill_back, ctypes.byref(cam), _ptr(grad_v), _ptr(scratch),
                                            scratch.numel() * 4, _stream()))
        return grad_v, None, None, None, None


# --------------------------------------------------------------------------------------------------------------
# NDC faces (+ textures) -> images
# --------------------------------------------------------------------------------------------------------------
class RasterConfig(object):
    """Everything a reference `Rasterize` instance is constructed with (rasterize.py:20-37) plus SSAA."""

    def __init__(self, image_size, anti_aliasing=True, near=0.1, far=100, eps=1e-4, background_color=(0, 0, 0),
                 return_rgb=False, return_alpha=False, return_depth=False, batch_bug=False, split_eps=None,
                 fast_math=None):
        self.H, self.W = hw(image_size)
        self.ss = 2 if anti_aliasing else 1
        self.near, self.far, self.eps = float(near), float(far), float(eps)
        self.background_color = background_color
        self.return_rgb, self.return_alpha, self.return_depth = bool(return_rgb), bool(return_alpha), bool(return_depth)
        self.batch_bug = bool(batch_bug)
        self.split_eps = split_eps      # (eps_alpha, eps_rgb): three-Rasterize semantics of the derender3d bridge
        fm = DEFAULT_FAST_MATH if fast_math is None else fast_math
        self.fast_math = int(fm) if not isinstance(fm, bool) else int(fm)     # 0 exact, 1 fast, 2 fast + far field

    @property
    def flags(self):
        return (SDN_RGB if self.return_rgb else 0) | (SDN_ALPHA if self.return_alpha else 0) | \
               (SDN_DEPTH if self.return_depth else 0)


class RasterizeFunction(torch.autograd.Function):
    """sdn_rasterize_forward / _backward.

    forward(faces [B,F,3,3], textures, cfg) -> rgb [B,3,H,W], alpha [B,H,W], depth [B,H,W]
    textures: None, [B,F,ts,ts,ts,3] (ts >= 2) or [B,F,3] (one colour per face).
    The int32 face-index map of the last call is kept on `cfg.last_face_index` for inspection / tests."""

    @staticmethod
    def forward(ctx, faces, textures, cfg):
        if cfg.flags == 0:
            raise Exception('nothing to draw')      # rasterize.py:25-27
        _need_cuda(faces, textures)
        fc = _f32c(faces)
        assert fc.dim() == 4 and fc.shape[2] == 3 and fc.shape[3] == 3
        B, F = fc.shape[0], fc.shape[1]
        dev = fc.device
        ts = 0
        tex = None
        if cfg.return_rgb:
            tex = _f32c(textures)
            if tex.dim() == 3:
                assert tex.shape == (B, F, 3)
                ts = 1
            else:
                assert tex.dim() == 6 and tex.shape[0] == B and tex.shape[1] == F and tex.shape[5] == 3
                ts = tex.shape[2]
                assert ts >= 2 and tex.shape[3] == ts and tex.shape[4] == ts
        H, W, ss = cfg.H, cfg.W, cfg.ss
        rgb = torch.empty(B, 3, H, W, device=dev) if cfg.return_rgb else None
        alpha = torch.empty(B, H, W, device=dev) if cfg.return_alpha else None
        depth = torch.empty(B, H, W, device=dev) if cfg.return_depth else None
        face_index = torch.empty(B, H * ss, W * ss, device=dev, dtype=torch.int32)
        rgb_ss = torch.empty(B, H * ss, W * ss, 3, device=dev) if (cfg.return_rgb and ts >= 2) else None
        nbytes = lib().sdn_workspace_bytes(B, F, H * ss, W * ss)
        workspace = torch.empty(nbytes, device=dev, dtype=torch.uint8)
        bg = cfg.background_color
        bg_b = None
        bg3 = (0.0, 0.0, 0.0)
        if cfg.return_rgb and bg is not None:
            if isinstance(bg, torch.Tensor) and bg.dim() == 2:
                bg_b = _f32c(bg.to(dev))
            else:
                bg3 = tuple(float(x) for x in (bg.tolist() if isinstance(bg, torch.Tensor) else bg))
        r = _Raster(B, F, H, W, ss, cfg.flags, ts, int(cfg.batch_bug), cfg.near, cfg.far, cfg.eps,
                    (ctypes.c_float * 3)(*bg3), _ptr(bg_b), _ptr(fc), _ptr(tex), _ptr(rgb), _ptr(alpha), _ptr(depth),
                    _ptr(face_index), _ptr(rgb_ss), _ptr(workspace), nbytes)
        with torch.cuda.device(dev):
            _check(lib().sdn_rasterize_forward(ctypes.byref(r), _stream()))
        ctx.cfg, ctx.ts = cfg, ts
        ctx.keep = (fc, tex, face_index, rgb_ss, workspace, bg_b)
        ctx.bg3 = bg3
        ctx.tex_shape = None if textures is None else tuple(textures.shape)
        cfg.last_face_index = face_index
        outs = []
        for t in (rgb, alpha, depth):
            if t is None:
                t = torch.empty(0, device=dev)
                ctx.mark_non_differentiable(t)
            outs.append(t)
        return tuple(outs)

    @staticmethod
    def backward(ctx, g_rgb, g_alpha, g_depth):
        cfg, ts = ctx.cfg, ctx.ts
        fc, tex, face_index, rgb_ss, workspace, bg_b = ctx.keep
        B, F = fc.shape[0], fc.shape[1]
        dev = fc.device
        g_rgb = _f32c(g_rgb) if cfg.return_rgb else None
        g_alpha = _f32c(g_alpha) if cfg.return_alpha else None
        g_depth = _f32c(g_depth) if cfg.return_depth else None
        grad_faces = torch.empty_like(fc)
        grad_tex = torch.empty_like(tex) if (tex is not None and ctx.needs_input_grad[1]) else None
        r = _Raster(B, F, cfg.H, cfg.W, cfg.ss, cfg.flags, ts, int(cfg.batch_bug), cfg.near, cfg.far, cfg.eps,
                    (ctypes.c_float * 3)(*ctx.bg3), _ptr(bg_b), _ptr(fc), _ptr(tex), None, None,
                    None, _ptr(face_index), _ptr(rgb_ss), _ptr(workspace), workspace.numel())
        g = _RasterGrad(_ptr(g_rgb), _ptr(g_alpha), _ptr(g_depth), _ptr(grad_faces), _ptr(grad_tex),
                        int(cfg.fast_math))
        with torch.cuda.device(dev):
            if cfg.split_eps is not None:
                _check(lib().sdn_rasterize_backward_split(ctypes.byref(r), ctypes.byref(g),
                                                          float(cfg.split_eps[0]), float(cfg.split_eps[1]),
                                                          _stream()))
            else:
                _check(lib().sdn_rasterize_backward(ctypes.byref(r), ctypes.byref(g), _stream()))
        return grad_faces, grad_tex, None


def rasterize_faces(faces, textures, cfg):
    """-> dict(rgb, alpha, depth) with None for what was not requested."""
    rgb, alpha, depth = RasterizeFunction.apply(faces, textures, cfg)
    return {'rgb': rgb if cfg.return_rgb else None, 'alpha': alpha if cfg.return_alpha else None,
            'depth': depth if cfg.return_depth else None}


class LinearTCFunction(torch.autograd.Function):
    """y = act(x . weight^T + bias) with the contraction on the tcgen05 tensor cores (csrc/sdn_linear_tc.cu: bf16
    operands, fp32 accumulation in TMEM).  Stands where the reference's encoder head calls torch.nn.Linear
    (derender3d/models/derenderer.py:26-32,35-43).  The backward is the plain fp32 matmul pair on cuBLAS."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        _need_cuda(x, weight)
        x2 = _f32c(x.reshape(-1, x.shape[-1]))
        w = _f32c(weight)
        b = _f32c(bias) if bias is not None else None
        M, K = x2.shape
        N = w.shape[0]
        if w.shape[1] != K:
            raise SdnError('linear_tc: weight is [%d,%d], input has %d features' % (N, w.shape[1], K))
        _need_cuda(b)
        if w.device != x2.device or (b is not None and b.device != x2.device):
            raise SdnError('linear_tc: input on %s, weight on %s' % (x2.device, w.device))
        y = torch.empty((M, N), dtype=torch.float32, device=x.device)
        if M > 0:
            with torch.cuda.device(x2.device):
                _check(lib().sdn_linear_forward(_ptr(x2), _ptr(w), _ptr(b) if b is not None else None, _ptr(y),
                                                M, N, K, 1 if relu else 0, _stream()))
        ctx.save_for_backward(x2, w, y if relu else None)
        ctx.has_bias = bias is not None
        ctx.in_shape = x.shape
        return y.reshape(x.shape[:-1] + (N,))

    @staticmethod
    def backward(ctx, gy):
        x2, w, y = ctx.saved_tensors
        g = gy.reshape(-1, gy.shape[-1]).float()
        if y is not None:
            g = g * (y > 0).to(g.dtype)
        gx = (g @ w).reshape(ctx.in_shape) if ctx.needs_input_grad[0] else None
        gw = g.t() @ x2 if ctx.needs_input_grad[1] else None
        gb = g.sum(0) if ctx.has_bias and ctx.needs_input_grad[2] else None
        return gx, gw, gb, None


def mlp_head_backward(grad_y, x, cat, h1, h2, W0, W1, W2, W3, n0):
    """Gradients of the fused encoder head (pure torch: five small matmul pairs).  cat = [relu(x W0^T + b0) | extra],
    h1 = relu(cat W1^T + b1), h2 = relu(h1 W2^T + b2), y = h2 W3^T + b3.
    -> dict(x, extra, W0, b0, W1, b1, W2, b2, W3, b3)."""
    g3 = grad_y.float()
    out = {'W3': g3.t() @ h2, 'b3': g3.sum(0)}
    g2 = (g3 @ W3) * (h2 > 0).to(g3.dtype)
    out['W2'], out['b2'] = g2.t() @ h1, g2.sum(0)
    g1 = (g2 @ W2) * (h1 > 0).to(g3.dtype)
    out['W1'], out['b1'] = g1.t() @ cat, g1.sum(0)
    gcat = g1 @ W1
    out['extra'] = gcat[:, n0:]
    g0 = gcat[:, :n0] * (cat[:, :n0] > 0).to(g3.dtype)
    out['W0'], out['b0'] = g0.t() @ x, g0.sum(0)
    out['x'] = g0 @ W0
    return out


class MlpHeadFunction(torch.autograd.Function):
    """The encoder head (net.fc, fc1, fc2, fc3 of derender3d/models/derenderer.py:26-32,35-43) as ONE launch
    (sdn_mlp_head_forward: fp32 FMA, a warp per output feature, grid-wide barriers between the layers).  Backward: torch."""

    @staticmethod
    def forward(ctx, x, extra, W0, b0, W1, b1, W2, b2, W3, b3):
        _need_cuda(x, extra, W0, b0, W1, b1, W2, b2, W3, b3)
        ts = [_f32c(t) for t in (x, extra, W0, b0, W1, b1, W2, b2, W3, b3)]
        x2, ex, w0, c0, w1, c1, w2, c2, w3, c3 = ts
        if any(t.device != x2.device for t in ts):
            raise SdnError('mlp_head: all operands must be on the input\'s device')
        M, K0 = x2.shape
        N0, N1, N2, N3, E = w0.shape[0], w1.shape[0], w2.shape[0], w3.shape[0], ex.shape[1]
        if w0.shape[1] != K0 or w1.shape[1] != N0 + E or w2.shape[1] != N1 or w3.shape[1] != N2 or ex.shape[0] != M:
            raise SdnError('mlp_head: layer shapes do not chain')
        dev = x2.device
        cat = torch.empty(M, N0 + E, device=dev)
        h1, h2, y = torch.empty(M, N1, device=dev), torch.empty(M, N2, device=dev), torch.empty(M, N3, device=dev)
        bar = torch.empty(1, device=dev, dtype=torch.int32)
        with torch.cuda.device(dev):
            _check(lib().sdn_mlp_head_forward(_ptr(x2), _ptr(ex), M, K0, N0, E, N1, N2, N3, _ptr(w0), _ptr(c0), _ptr(w1), _ptr(c1),
                                              _ptr(w2), _ptr(c2), _ptr(w3), _ptr(c3), _ptr(cat), _ptr(h1), _ptr(h2), _ptr(y),
                                              _ptr(bar), _stream()))
        ctx.save_for_backward(x2, cat, h1, h2, w0, w1, w2, w3)
        ctx.n0 = N0
        return y

    @staticmethod
    def backward(ctx, gy):
        x2, cat, h1, h2, w0, w1, w2, w3 = ctx.saved_tensors
        g = mlp_head_backward(gy, x2, cat, h1, h2, w0, w1, w2, w3, ctx.n0)
        return (g['x'], g['extra'], g['W0'], g['b0'], g['W1'], g['b1'], g['W2'], g['b2'], g['W3'], g['b3'])


def mlp_head(x, extra, fc0, fc1, fc2, fc3):
    """relu(fc0(x)) | extra -> relu(fc1) -> relu(fc2) -> fc3 with torch.nn.Linear modules fc0..fc3, one kernel launch."""
    return MlpHeadFunction.apply(x, extra, fc0.weight, fc0.bias, fc1.weight, fc1.bias, fc2.weight, fc2.bias, fc3.weight, fc3.bias)


def linear_tc(x, weight, bias=None, relu=False):
    return LinearTCFunction.apply(x, weight, bias, relu)


# --------------------------------------------------------------------------------------------------------------
# convolution on the tcgen05 tensor cores (csrc/sdn_conv_tc.cu)
# --------------------------------------------------------------------------------------------------------------
def pack_conv_weight(weight, cin_pad=None, cout_pad=None):
    """torch Conv2d weight [Cout, Cin, KH, KW] -> the kernel's K-major operand [Cout_pad, KH, KW, Cin_pad] bf16
    (zero rows / channels for the padding; Cin_pad a multiple of 64; Cout_pad a multiple of 64, or 16 for <= 16 outputs)."""
    co, ci, kh, kw = weight.shape
    cin_pad = cin_pad or (ci + 63) // 64 * 64
    cout_pad = cout_pad or (16 if co <= 16 else (co + 63) // 64 * 64)
    if cin_pad == ci and cout_pad == co:        # nothing to pad: one cast (a channels-last weight is already [Cout,KH,KW,Cin])
        return weight.detach().permute(0, 2, 3, 1).to(torch.bfloat16).contiguous()
    w = torch.zeros(cout_pad, kh, kw, cin_pad, device=weight.device, dtype=torch.bfloat16)
    w[:co, :, :, :ci] = weight.detach().permute(0, 2, 3, 1).to(torch.bfloat16)
    return w.contiguous()


def conv2d_tc(x, w_packed, stride=1, pad=0, scale=None, shift=None, residual=None, relu=False, out_hw=None, out=None,
              stats=None, stats_keep=False):
    """x [N, H, W, Cin] bf16 NHWC (Cin a multiple of 64), w_packed [Cout, KH, KW, Cin] bf16 (pack_conv_weight) ->
    [N, Ho, Wo, Cout] bf16:  act(scale * conv(x, w) + shift (+ residual)), act = relu in (False, True / 'relu', 'tanh').
    sdn_conv2d_forward: implicit GEMM, TMA loads, tcgen05.mma with fp32 accumulation in tensor memory.
    out_hw = (Ho, Wo) overrides the output extent (reads beyond the input are zeros); out = a [N, Ho, Wo, Cout] bf16 VIEW
    (channels contiguous, strides multiples of 8) to write into -- e.g. one parity phase of a transposed convolution.
    stats = fp32 [N, Cout, 2]: receives per (image, channel) sum / sum of squares of the result before the activation
    (cleared first unless stats_keep) -- the statistics of a following InstanceNorm2d, straight from the accumulators.
    Inference-path operator (no autograd)."""
    _need_cuda(x, w_packed, scale, shift, residual, out, stats)
    if x.dtype != torch.bfloat16 or w_packed.dtype != torch.bfloat16 or x.dim() != 4 or w_packed.dim() != 4:
        raise SdnError('conv2d_tc: x [N,H,W,Cin] and w [Cout,KH,KW,Cin] must be bf16')
    x, w_packed = x.contiguous(), w_packed.contiguous()
    N, H, W, Cin = x.shape
    Cout, KH, KW, Cw = w_packed.shape
    if Cw != Cin or Cin % 64 or (Cout % 64 and Cout != 16):
        raise SdnError('conv2d_tc: channels must match; Cin a multiple of 64, Cout a multiple of 64 or 16 '
                       '(got Cin %d / %d, Cout %d)' % (Cin, Cw, Cout))
    if out_hw is None:
        Ho, Wo = (H + 2 * pad - KH) // stride + 1, (W + 2 * pad - KW) // stride + 1
    else:
        Ho, Wo = int(out_hw[0]), int(out_hw[1])
    sn = sh = sw = 0
    if out is None:
        out = torch.empty(N, Ho, Wo, Cout, device=x.device, dtype=torch.bfloat16)
    else:
        if tuple(out.shape) != (N, Ho, Wo, Cout) or out.dtype != torch.bfloat16 or out.stride(3) != 1:
            raise SdnError('conv2d_tc: out must be a bf16 [%d,%d,%d,%d] view with contiguous channels' % (N, Ho, Wo, Cout))
        sn, sh, sw = out.stride(0), out.stride(1), out.stride(2)
    f32 = lambda t: None if t is None else t.detach().float().contiguous()      # noqa: E731
    scale, shift = f32(scale), f32(shift)
    if residual is not None:
        residual = residual.contiguous()
        if tuple(residual.shape) != (N, Ho, Wo, Cout) or residual.dtype != torch.bfloat16:
            raise SdnError('conv2d_tc: residual must be bf16 with the output shape %s' % ((N, Ho, Wo, Cout),))
    act = {False: 0, True: 1, None: 0, 'relu': 1, 'tanh': 2, 'none': 0}[relu]
    c = _Conv2d(N, H, W, Cin, Cout, KH, KW, int(stride), int(pad), act, x.data_ptr(), w_packed.data_ptr(),
                None if scale is None else scale.data_ptr(), None if shift is None else shift.data_ptr(),
                None if residual is None else residual.data_ptr(), out.data_ptr(),
                0 if out_hw is None else Ho, 0 if out_hw is None else Wo, sn, sh, sw,
                None if stats is None else stats.data_ptr(), 1 if stats_keep else 0)
    if stats is not None and (stats.dtype != torch.float32 or stats.numel() != N * Cout * 2 or not stats.is_contiguous()):
        raise SdnError('conv2d_tc: stats must be a contiguous fp32 [N, Cout, 2] tensor')
    with torch.cuda.device(x.device):
        _check(lib().sdn_conv2d_forward(ctypes.byref(c), _stream()))
    return out


def conv2d_wgrad_tc(grad_out, x, KH, KW, stride=1, pad=0, accumulate_into=None):
    """Weight gradient of conv2d_tc: grad_out [N,Ho,Wo,Cout] bf16, x [N,H,W,Cin] bf16 -> fp32 [Cout,KH,KW,Cin]
    (sdn_conv2d_wgrad: tcgen05 GEMM over pixels, MN-major operands straight from the NHWC tensors, split over pixel
    ranges with fp32 atomics).  accumulate_into: a contiguous fp32 [Cout,KH,KW,Cin] tensor the result is ADDED to (e.g. the
    channels-last storage of a parameter's .grad); it is returned."""
    _need_cuda(grad_out, x, accumulate_into)
    if grad_out.dtype != torch.bfloat16 or x.dtype != torch.bfloat16:
        raise SdnError('conv2d_wgrad_tc: bf16 NHWC tensors expected')
    grad_out, x = grad_out.contiguous(), x.contiguous()
    N, H, W, Cin = x.shape
    Cout = grad_out.shape[-1]
    Ho, Wo = (H + 2 * pad - KH) // stride + 1, (W + 2 * pad - KW) // stride + 1
    if tuple(grad_out.shape) != (N, Ho, Wo, Cout) or Cin % 64 or Cout % 64:
        raise SdnError('conv2d_wgrad_tc: grad_out %s does not match x %s (channels must be multiples of 64)'
                       % (tuple(grad_out.shape), tuple(x.shape)))
    if accumulate_into is not None:
        gw = accumulate_into
        if tuple(gw.shape) != (Cout, KH, KW, Cin) or gw.dtype != torch.float32 or not gw.is_contiguous():
            raise SdnError('conv2d_wgrad_tc: accumulate_into must be a contiguous fp32 [Cout,KH,KW,Cin] tensor')
    else:
        gw = torch.empty(Cout, KH, KW, Cin, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _check(lib().sdn_conv2d_wgrad(_ptr(grad_out), _ptr(x), _ptr(gw), N, H, W, Cin, Cout, KH, KW, int(stride), int(pad),
                                      1 if accumulate_into is not None else 0, _stream()))
    return gw


def pack_conv_weight_transposed(weight, stride=1):
    """Operands of the data gradient from a Conv2d weight [Cout, Cin, KH, KW] fp32 (sdn_pack_conv_weight_transposed: one
    launch, shared-memory transpose).  stride 1 / 1x1 -> [packed [Cin,KH,KW,Cout] bf16, taps flipped]; stride 2, 3x3 ->
    the four parity-phase operands [(0,0), (0,1), (1,0), (1,1)], views of one buffer."""
    _need_cuda(weight)
    co, ci, kh, kw = weight.shape
    if ci % 64 or co % 64:
        raise SdnError('pack_conv_weight_transposed: channels must be multiples of 64')
    src = weight.detach().float().contiguous(memory_format=torch.channels_last)    # physical [Cout, KH, KW, Cin]
    buf = torch.empty(ci * kh * kw * co, device=weight.device, dtype=torch.bfloat16)
    with torch.cuda.device(weight.device):
        _check(lib().sdn_pack_conv_weight_transposed(_ptr(src), _ptr(buf), co, ci, kh, kw, int(stride), _stream()))
    if stride == 2 and kh == 3:
        unit, out, off = ci * co, [], 0
        for ta, tb in ((1, 1), (1, 2), (2, 1), (2, 2)):
            out.append(buf[off:off + unit * ta * tb].view(ci, ta, tb, co))
            off += unit * ta * tb
        return out
    return [buf.view(ci, kh, kw, co)]


def batch_norm_finalize(stats, pixels, bn):
    """Batch statistics [N,C,2] (from conv2d_tc(stats=...)) -> (mean, rstd, scale, shift) fp32 [C] each, and the running
    statistics / num_batches_tracked of the torch.nn.BatchNorm2d `bn` updated in place (sdn_batch_norm_finalize)."""
    _need_cuda(stats)
    N, C, _ = stats.shape
    out = torch.empty(4, C, device=stats.device, dtype=torch.float32)
    track = bn.track_running_stats and bn.running_mean is not None
    mom = 0.1 if bn.momentum is None else float(bn.momentum)
    gamma, beta = bn.weight.detach(), bn.bias.detach()
    if gamma.dtype != torch.float32 or not gamma.is_contiguous() or beta.dtype != torch.float32:
        raise SdnError('batch_norm_finalize: fp32 contiguous BatchNorm parameters expected')
    with torch.cuda.device(stats.device):
        _check(lib().sdn_batch_norm_finalize(_ptr(stats.contiguous()), N, C, int(pixels), float(bn.eps), mom, _ptr(gamma),
                                             _ptr(beta), _ptr(bn.running_mean) if track else None,
                                             _ptr(bn.running_var) if track else None,
                                             _ptr(bn.num_batches_tracked) if track else None, _ptr(out), _stream()))
    return out[0], out[1], out[2], out[3]


def conv2d_dgrad_tc(grad_out, weight, in_hw, stride=1, pad=0, residual=None):
    """Data gradient of a convolution with `weight` [Cout, Cin, KH, KW] (fp32 / bf16 torch layout): grad_out
    [N,Ho,Wo,Cout] bf16 -> [N,H,W,Cin] bf16, on the SAME tcgen05 kernel as the forward: stride 1 = a convolution with the
    spatially flipped, channel-transposed weights and padding K-1-pad; stride 2 (3x3 pad 1, or 1x1) = the parity phases
    of the transposed convolution, each written into its interleaved positions.  `residual` (dense [N,H,W,Cin] bf16) is
    added in the epilogue (the skip connection's gradient)."""
    Cout, Cin, KH, KW = weight.shape
    H, W = in_hw
    N = grad_out.shape[0]
    if stride == 1:
        wt, = pack_conv_weight_transposed(weight, 1)
        return conv2d_tc(grad_out, wt, 1, KH - 1 - pad, None, None, residual, False, out_hw=(H, W))
    if stride != 2 or H % 2 or W % 2:
        raise SdnError('conv2d_dgrad_tc: stride %d with input %dx%d is not supported' % (stride, H, W))
    out = torch.empty(N, H, W, Cin, device=grad_out.device, dtype=torch.bfloat16)
    if KH == 3 and pad == 1:
        ops = pack_conv_weight_transposed(weight, 2)
        for k, (ph, pw) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
            conv2d_tc(grad_out, ops[k], 1, 0, None, None, None, False, out_hw=(H // 2, W // 2), out=out[:, ph::2, pw::2])
    elif KH == 1 and pad == 0:
        out.zero_()
        wt, = pack_conv_weight_transposed(weight, 2)
        conv2d_tc(grad_out, wt, 1, 0, None, None, None, False, out_hw=(H // 2, W // 2), out=out[:, 0::2, 0::2])
    else:
        raise SdnError('conv2d_dgrad_tc: stride-2 kernel %dx%d pad %d is not supported' % (KH, KW, pad))
    if residual is not None:
        out = out + residual
    return out


def affine_act_nhwc(z, scale, shift, residual=None, relu=False):
    """y = relu?(z * scale[c] + shift[c] (+ residual)) on NHWC bf16 (sdn_affine_act_nhwc): BatchNorm's normalising pass."""
    _need_cuda(z, scale, shift, residual)
    z = z.contiguous()
    C = z.shape[-1]
    y = torch.empty_like(z)
    scale, shift = scale.detach().float().contiguous(), shift.detach().float().contiguous()
    if residual is not None:
        residual = residual.contiguous()
    with torch.cuda.device(z.device):
        _check(lib().sdn_affine_act_nhwc(_ptr(z), _ptr(scale), _ptr(shift), _ptr(residual), _ptr(y), z.numel() // C, C,
                                         1 if relu else 0, _stream()))
    return y


def batch_norm_backward_nhwc(grad_y, y, z, mean, rstd, gamma, relu, want_skip=False, sums=None, accumulate_into=None):
    """-> (grad_z bf16, grad_gamma [C], grad_beta [C], grad_skip bf16 | None); see include/sdn_b200.h.
    sums: an already CLEARED fp32 [C, 2] buffer to accumulate into (saves the per-call clearing pass).
    accumulate_into = (grad_gamma, grad_beta): contiguous fp32 [C] tensors the parameter gradients are ADDED to in place."""
    _need_cuda(grad_y, y, z, mean, rstd, gamma)
    grad_y, z = grad_y.contiguous(), z.contiguous()
    y = None if y is None else y.contiguous()
    C = z.shape[-1]
    precleared = sums is not None
    if sums is None:
        sums = torch.empty(C, 2, device=z.device, dtype=torch.float32)
    elif tuple(sums.shape) != (C, 2) or sums.dtype != torch.float32 or not sums.is_contiguous():
        raise SdnError('batch_norm_backward_nhwc: sums must be a contiguous fp32 [C, 2] tensor')
    gz = torch.empty_like(z)
    gs = torch.empty_like(z) if want_skip else None
    mean, rstd, gamma = mean.float().contiguous(), rstd.float().contiguous(), gamma.detach().float().contiguous()
    acc_g, acc_b = accumulate_into if accumulate_into is not None else (None, None)
    for a in (acc_g, acc_b):
        if a is not None and (a.dtype != torch.float32 or not a.is_contiguous() or a.numel() != C):
            raise SdnError('batch_norm_backward_nhwc: accumulate_into needs contiguous fp32 [C] tensors')
    with torch.cuda.device(z.device):
        _check(lib().sdn_batch_norm_backward_nhwc(_ptr(grad_y), _ptr(y), _ptr(z), _ptr(mean), _ptr(rstd), _ptr(gamma),
                                                  _ptr(sums), _ptr(gz), _ptr(gs), _ptr(acc_g), _ptr(acc_b), z.numel() // C, C,
                                                  (1 if relu else 0) | (2 if precleared else 0), _stream()))
    return gz, sums[:, 1], sums[:, 0], gs


def max_pool_nhwc(x, k=3, stride=2, pad=1, return_argmax=False):
    """nn.MaxPool2d(k, stride, pad) on x [N,H,W,C] bf16 NHWC (sdn_max_pool_nhwc).  return_argmax: also the uint8 tensor of
    window positions that max_pool_backward_nhwc needs."""
    _need_cuda(x)
    x = x.contiguous()
    N, H, W, C = x.shape
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    y = torch.empty(N, Ho, Wo, C, device=x.device, dtype=torch.bfloat16)
    am = torch.empty(N, Ho, Wo, C, device=x.device, dtype=torch.uint8) if return_argmax else None
    with torch.cuda.device(x.device):
        _check(lib().sdn_max_pool_nhwc(_ptr(x), _ptr(y), _ptr(am), N, H, W, C, int(k), int(stride), int(pad), _stream()))
    return (y, am) if return_argmax else y


def max_pool_backward_nhwc(argmax, grad_y, in_hw, k=3, stride=2, pad=1):
    """Gradient of max_pool_nhwc with respect to its input [N, H, W, C] (sdn_max_pool_backward_nhwc: gather form)."""
    _need_cuda(argmax, grad_y)
    argmax, grad_y = argmax.contiguous(), grad_y.contiguous()
    N, _, _, C = grad_y.shape
    H, W = in_hw
    gx = torch.empty(N, H, W, C, device=grad_y.device, dtype=torch.bfloat16)
    with torch.cuda.device(grad_y.device):
        _check(lib().sdn_max_pool_backward_nhwc(_ptr(argmax), _ptr(grad_y), _ptr(gx), N, H, W, C, int(k), int(stride), int(pad),
                                                _stream()))
    return gx


def reflect_pad_nhwc(x, pad, cpad=None):
    """nn.ReflectionPad2d(pad) on x [N,H,W,C] bf16 NHWC, channels zero-padded to cpad (sdn_reflect_pad_nhwc)."""
    _need_cuda(x)
    x = x.contiguous()
    N, H, W, C = x.shape
    cp = C if cpad is None else int(cpad)
    y = torch.empty(N, H + 2 * pad, W + 2 * pad, cp, device=x.device, dtype=torch.bfloat16)
    with torch.cuda.device(x.device):
        _check(lib().sdn_reflect_pad_nhwc(_ptr(x), _ptr(y), N, H, W, C, int(pad), cp, _stream()))
    return y


def instance_norm_pad_nhwc(x, stats, pad, residual=None, relu=False, eps=1e-5):
    """reflect_pad(relu?(instance_norm(x) (+ residual)), pad) in ONE pass, from statistics a convolution produced
    (sdn_instance_norm_apply_pad_nhwc): [N, H, W, C] -> [N, H + 2 pad, W + 2 pad, C] bf16."""
    _need_cuda(x, residual, stats)
    x = x.contiguous()
    N, H, W, C = x.shape
    y = torch.empty(N, H + 2 * pad, W + 2 * pad, C, device=x.device, dtype=torch.bfloat16)
    if residual is not None:
        residual = residual.contiguous()
    with torch.cuda.device(x.device):
        _check(lib().sdn_instance_norm_apply_pad_nhwc(_ptr(x), _ptr(residual), _ptr(y), _ptr(stats), N, H, W, C, float(eps),
                                                      1 if relu else 0, int(pad), _stream()))
    return y


def instance_norm_nhwc(x, residual=None, relu=False, eps=1e-5, out=None, stats=None):
    """x [N, H, W, C] bf16 NHWC -> relu?((x - mean) / sqrt(var + eps) (+ residual)) per (n, c) over H*W
    (sdn_instance_norm_nhwc; torch.nn.InstanceNorm2d(affine=False) semantics).  stats = the fp32 [N, C, 2] sums a
    convolution already produced (conv2d_tc(stats=...)): then only the normalising pass runs."""
    _need_cuda(x, residual, stats)
    x = x.contiguous()
    N, H, W, C = x.shape
    y = torch.empty_like(x) if out is None else out
    if residual is not None:
        residual = residual.contiguous()
    with torch.cuda.device(x.device):
        if stats is None:
            scratch = torch.empty(N, C, 2, device=x.device, dtype=torch.float32)
            _check(lib().sdn_instance_norm_nhwc(_ptr(x), _ptr(residual), _ptr(y), _ptr(scratch), N, H * W, C, float(eps),
                                                1 if relu else 0, _stream()))
        else:
            _check(lib().sdn_instance_norm_apply_nhwc(_ptr(x), _ptr(residual), _ptr(y), _ptr(stats), N, H * W, C,
                                                      float(eps), 1 if relu else 0, _stream()))
    return y


def im2col_nhwc(x, KH, KW, stride, pad, Kpad):
    """x [N,H,W,Cin] bf16 -> [N,Ho,Wo,Kpad] bf16 patches (sdn_im2col_nhwc): turns a few-channel stem convolution into a
    1x1 convolution over Kpad channels."""
    _need_cuda(x)
    x = x.contiguous()
    N, H, W, Cin = x.shape
    Ho, Wo = (H + 2 * pad - KH) // stride + 1, (W + 2 * pad - KW) // stride + 1
    out = torch.empty(N, Ho, Wo, Kpad, device=x.device, dtype=torch.bfloat16)
    with torch.cuda.device(x.device):
        _check(lib().sdn_im2col_nhwc(_ptr(x), _ptr(out), N, H, W, Cin, KH, KW, int(stride), int(pad), int(Kpad), _stream()))
    return out


class MeshBank(object):
    """The template meshes of a Derenderer3d on one device: Bernstein tables concatenated [sum nv, 64], face tables
    padded to [K, nf_max, 3] with (0,0,0) triangles (three coincident corners: zero area, never drawn, no gradient).
    Built once; reference re-uploads B per object per call (derender3d/models/transforms.py:97)."""

    def __init__(self, bases, faces, device):
        self.K = len(bases)
        nv = [int(b.shape[0]) for b in bases]
        self.nv_list, self.nf_list = nv, [int(f.shape[0]) for f in faces]
        self.nv_max, self.nf_max = max(nv), max(self.nf_list)
        self.basis = torch.cat([b.reshape(b.shape[0], 64).float() for b in bases]).contiguous().to(device)
        off = [0]
        for n in nv[:-1]:
            off.append(off[-1] + n)
        self.basis_off = torch.tensor(off, dtype=torch.int64, device=device)
        self.nv = torch.tensor(nv, dtype=torch.int32, device=device)
        fp = torch.zeros(self.K, self.nf_max, 3, dtype=torch.int32)
        for k, f in enumerate(faces):
            fp[k, :f.shape[0]] = f.to(torch.int32)
        self.faces = fp.to(device)
        self.device = torch.device(device)


class ObjectVerticesFunction(torch.autograd.Function):
    """(classes [N] int64, control [N,3,64], scales [N,3], rotations [N,4], translations [N,3], ptrans [N,3],
    zoom [N,1]) -> camera-space vertices [N,nv_max,3] (+ zooms [N,1] in test mode, where `zoom` holds zoom_tos).
    One launch for the batch; see include/sdn_b200.h `sdn_objects`."""

    @staticmethod
    def forward(ctx, bank, classes, control, scales, rotations, translations, ptrans, zoom, derive):
        _need_cuda(classes, control, scales, rotations, translations, ptrans, zoom)
        N = int(classes.shape[0])
        dev = control.device
        cl = classes.detach().to(torch.int64).contiguous()
        t = [_f32c(x) for x in (control.reshape(N, 3, 64), scales, rotations, translations, ptrans, zoom.reshape(N))]
        verts = torch.empty(N, bank.nv_max, 3, device=dev)
        zooms = torch.empty(N, device=dev) if derive else t[5]
        key = torch.empty(N, dtype=torch.int64, device=dev) if derive else None
        o = _Objects(N, bank.K, bank.nv_max, 1 if derive else 0, bank.basis.data_ptr(), bank.basis_off.data_ptr(),
                     bank.nv.data_ptr(), cl.data_ptr(), t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(),
                     t[3].data_ptr(), t[4].data_ptr(), zooms.data_ptr(), t[5].data_ptr() if derive else None,
                     key.data_ptr() if derive else None, verts.data_ptr())
        with torch.cuda.device(dev):
            _check(lib().sdn_objects_forward(ctypes.byref(o), _stream()))
        ctx.bank, ctx.derive, ctx.keep = bank, derive, (cl, t, zooms, key)
        ctx.zoom_shape = zoom.shape
        out_z = zooms.reshape(N, 1)
        ctx.mark_non_differentiable(out_z)
        return verts, out_z

    @staticmethod
    def backward(ctx, gv, _gz):
        bank, derive = ctx.bank, ctx.derive
        cl, t, zooms, key = ctx.keep
        N = int(cl.shape[0])
        dev = gv.device
        gv = _f32c(gv)
        g_control = torch.empty(N, 3, 64, device=dev)
        g_s, g_r, g_t, g_p = (torch.empty(N, k, device=dev) for k in (3, 4, 3, 3))
        g_z = torch.empty(N, device=dev)
        scratch = torch.empty(N * 24, device=dev)
        verts_dummy = gv                     # `vertices` is not read by the backward
        o = _Objects(N, bank.K, bank.nv_max, 1 if derive else 0, bank.basis.data_ptr(), bank.basis_off.data_ptr(),
                     bank.nv.data_ptr(), cl.data_ptr(), t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(),
                     t[3].data_ptr(), t[4].data_ptr(), zooms.data_ptr(), t[5].data_ptr() if derive else None,
                     key.data_ptr() if derive else None, verts_dummy.data_ptr())
        with torch.cuda.device(dev):
            _check(lib().sdn_objects_backward(ctypes.byref(o), _ptr(gv), _ptr(g_control), _ptr(g_s), _ptr(g_r),
                                              _ptr(g_t), _ptr(g_p), _ptr(g_z), _ptr(scratch), scratch.numel() * 4,
                                              _stream()))
        return None, None, g_control, g_s, g_r, g_t, g_p, g_z.reshape(ctx.zoom_shape), None


def object_vertices(bank, classes, control, scales, rotations, translations, ptrans, zooms=None, zoom_tos=None):
    """Training mode: pass zooms [N,1] -> vertices.  Test mode: pass zoom_tos [N,1] -> (vertices, zooms [N,1])."""
    if (zooms is None) == (zoom_tos is None):
        raise SdnError('object_vertices: give exactly one of zooms / zoom_tos')
    if zoom_tos is None:
        return ObjectVerticesFunction.apply(bank, classes, control, scales, rotations, translations, ptrans, zooms,
                                            False)[0]
    return ObjectVerticesFunction.apply(bank, classes, control, scales, rotations, translations, ptrans, zoom_tos, True)

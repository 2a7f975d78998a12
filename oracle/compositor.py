"""ORACLE (test infrastructure -- only tests/, smoke() and bench.py's cpu_baseline may import this).

CPU restatement of the test-time scene compositor of the geometric branch:
    geometric/scripts/main.py:530-606   far-to-near paste of every object's rendered silhouette / normal / depth crop
                                        into the full-frame instance, normal and depth maps
The reference does this with torchvision.transforms.functional (to_pil_image / resize / to_tensor,
derender3d/datasets.py:18-24) on top of Pillow.  Third-party arithmetic on this path: Pillow's `Image.resize(BILINEAR)`
(libImaging/Resample.c: separable, support scaled by the down-scale factor, 8-bit images in 22-bit fixed point with a
uint8 intermediate between the horizontal and the vertical pass, 'F' images accumulated in double with a float32
intermediate).  `composite_reference` below calls Pillow itself (this image ships Pillow 12.2.0; the reference pinned
pillow 5.1 -- the resampler has not changed since 3.x); `pil_coeffs` / `pil_resize_*` restate the algorithm in numpy and
are pinned bit-for-bit against Pillow by tests/test_compositor_host.py -- they are what the CUDA kernel mirrors.
"""
import numpy as np
import torch

PRECISION_BITS = 32 - 8 - 2          # Resample.c: PRECISION_BITS


def pil_coeffs(in_size, out_size):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the bilinear filter, box = the whole image.
    -> bounds [out,2] int32 (first tap, tap count), kk [out,ksize] float64, kk8 [out,ksize] int32."""
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 1.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.float64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        ww = 0.0
        for x in range(xmax):
            a = abs((x + xmin - center + 0.5) * ss)
            w = 1.0 - a if a < 1.0 else 0.0
            kk[xx, x] = w
            ww += w
        if ww != 0.0:
            kk[xx, :xmax] /= ww
        bounds[xx] = (xmin, xmax)
    kk8 = np.where(kk < 0, (-0.5 + kk * (1 << PRECISION_BITS)).astype(np.int64),
                   (0.5 + kk * (1 << PRECISION_BITS)).astype(np.int64)).astype(np.int32)
    return bounds, kk, kk8


def pil_resize_u8(a, out):
    """[H,W] uint8 -> [out,out] uint8, horizontal pass then vertical pass (ImagingResampleHorizontal/Vertical_8bpc)."""
    H, W = a.shape
    bx, _, kx = pil_coeffs(W, out)
    by, _, ky = pil_coeffs(H, out)
    t = np.zeros((H, out), np.uint8)
    for xx in range(out):
        x0, n = bx[xx]
        s = (1 << (PRECISION_BITS - 1)) + (a[:, x0:x0 + n].astype(np.int64) * kx[xx, :n]).sum(1)
        t[:, xx] = np.clip(s >> PRECISION_BITS, 0, 255)
    o = np.zeros((out, out), np.uint8)
    for yy in range(out):
        y0, n = by[yy]
        s = (1 << (PRECISION_BITS - 1)) + (t[y0:y0 + n].astype(np.int64) * ky[yy, :n, None]).sum(0)
        o[yy] = np.clip(s >> PRECISION_BITS, 0, 255)
    return o


def pil_resize_f32(a, out):
    """[H,W] float32 -> [out,out] float32 (ImagingResampleHorizontal/Vertical_32bpc: double accumulation, taps in order)."""
    H, W = a.shape
    bx, kx, _ = pil_coeffs(W, out)
    by, ky, _ = pil_coeffs(H, out)
    t = np.zeros((H, out), np.float32)
    for xx in range(out):
        x0, n = bx[xx]
        s = np.zeros(H, np.float64)
        for j in range(n):
            s += a[:, x0 + j].astype(np.float64) * kx[xx, j]
        t[:, xx] = s.astype(np.float32)
    o = np.zeros((out, out), np.float32)
    for yy in range(out):
        y0, n = by[yy]
        s = np.zeros(out, np.float64)
        for j in range(n):
            s += t[y0 + j].astype(np.float64) * ky[yy, j]
        o[yy] = s.astype(np.float32)
    return o


def composite_reference(masks, normals, depth_maps, zooms, center2ds, depths, interests, height, width, focal, u0, v0,
                        render_size, image_masks=None, operations_given=True):
    """main.py:547-606 through torchvision + Pillow exactly as the reference calls them.  With operations given only
    interesting objects are drawn; with `operations_given=False` (the reference's default test mode, `operations is
    None`) the others are blended into the instance map from their frame masks `image_masks` [N,1,H,W] (:601-604).
    CPU tensors in; -> instance map [1,H,W], normal map [3,H,W], depth map [1,H,W]."""
    import PIL.Image
    import torchvision.transforms.functional as TF
    index_objs = torch.sort(depths[:, 0], dim=0, descending=True)[1]
    inst = torch.zeros(1, height, width)
    nrm = torch.full((3, height, width), 0.5)
    dep = torch.full((1, height, width), 1.0)
    for index_obj in index_objs.tolist():
        if not bool(interests[index_obj]):
            if not operations_given:                                      # main.py:601-604
                image_mask = image_masks[index_obj].float()
                inst = (1 - image_mask) * inst + image_mask * (1 + index_obj)
            continue
        size = int(render_size / zooms[index_obj])
        box = (int(center2ds[index_obj, 1] * focal + u0 - size // 2), int(center2ds[index_obj, 0] * focal + v0 - size // 2))
        m_pil = TF.resize(TF.to_pil_image(masks[index_obj].detach().cpu()), (size, size))
        frame = PIL.Image.new(mode='L', size=(width, height))
        frame.paste(m_pil, box=box)
        m = torch.round(TF.to_tensor(frame))
        inst = (1 - m) * inst + m * (1 + index_obj)
        n_pil = TF.resize(TF.to_pil_image((normals[index_obj] / 2 + 0.5).detach().cpu()), (size, size))
        frame = PIL.Image.new(mode='RGB', size=(width, height))
        frame.paste(n_pil, box=box)
        nrm = (1 - m) * nrm + m * TF.to_tensor(frame)
        d = torch.min(depth_maps[index_obj] * zooms[index_obj] / 100.0, torch.tensor(1.0))
        # torchvision 0.2.1 (environment.yml:25) maps a float32 HxWx1 ndarray to a mode-'F' image without scaling;
        # current torchvision would quantise it to 8 bits, so the mode-'F' image is built directly
        d_pil = TF.resize(PIL.Image.fromarray(d.detach().cpu().numpy()[0], mode='F'), (size, size))
        frame = PIL.Image.new(mode='F', size=(width, height))
        frame.paste(d_pil, box=box)
        dep = (1 - m) * dep + m * TF.to_tensor(frame)
    return inst, nrm, dep

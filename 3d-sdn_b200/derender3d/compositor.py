"""Test-time scene compositor (reference: geometric/scripts/main.py:530-622): paste every object's rendered
silhouette / normal / depth crop, far to near, into the 375x1242 instance, normal and depth maps and write the on-disk
contract consumed by textural/edit_vkitti.py:43,63-95 (`<name>.png` instance ids, `<name>.json`, `<name>-normal.png`,
`<name>-depth.png` 16-bit).

The reference does the resize / paste per object on the CPU through torchvision + Pillow (GPU -> CPU -> GPU for each
object); here the whole frame is two native launches (`sdn_composite_scene`), bit-identical to Pillow's bilinear resize.
Only the few per-object scalars the reference also brings to the host (zoom, centre, depth: main.py:556-565 `int(...)`)
cross the bus."""
import ctypes
import json
import os

import numpy as np
import torch

import sdn_b200

PRECISION_BITS = 32 - 8 - 2


def resample_tables(in_size, out_size):
    """Pillow's precompute_coeffs + normalize_coeffs_8bpc (libImaging/Resample.c) for the bilinear filter, vectorised:
    -> bounds [out,2] int32, kkd [out,ksize] float64, kk8 [out,ksize] int32."""
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 1.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    ss = 1.0 / filterscale
    center = (np.arange(out_size, dtype=np.float64) + 0.5) * scale
    xmin = np.maximum((center - support + 0.5).astype(np.int64), 0)
    xmax = np.minimum((center + support + 0.5).astype(np.int64), in_size) - xmin
    x = np.arange(ksize, dtype=np.float64)[None, :]
    a = np.abs((x + xmin[:, None] - center[:, None] + 0.5) * ss)
    w = np.where(a < 1.0, 1.0 - a, 0.0)
    w = np.where(np.arange(ksize)[None, :] < xmax[:, None], w, 0.0)
    ww = np.zeros(out_size, np.float64)
    for j in range(ksize):              # left-to-right sum, like the C loop
        ww = ww + w[:, j]
    kkd = np.where(ww[:, None] != 0.0, w / np.where(ww == 0.0, 1.0, ww)[:, None], w)
    kkd = np.where(np.arange(ksize)[None, :] < xmax[:, None], kkd, 0.0)
    kk8 = np.where(kkd < 0, (-0.5 + kkd * (1 << PRECISION_BITS)).astype(np.int64),
                   (0.5 + kkd * (1 << PRECISION_BITS)).astype(np.int64)).astype(np.int32)
    bounds = np.stack([xmin, xmax], axis=1).astype(np.int32)
    return bounds, kkd, kk8


class _Obj(ctypes.Structure):
    _fields_ = [('index', ctypes.c_int), ('size', ctypes.c_int), ('box_x', ctypes.c_int), ('box_y', ctypes.c_int),
                ('ksize', ctypes.c_int), ('id', ctypes.c_int), ('row_off', ctypes.c_longlong),
                ('k_off', ctypes.c_longlong), ('tmp_off', ctypes.c_longlong)]


def placement(zooms, center2ds, depths, interests, focal, u0, v0, render_size, keep_uninteresting=False):
    """Draw order (far to near, main.py:538) and the integer size / paste box of every interesting object, with the
    reference's own float32 expressions (main.py:556-565).  Host tensors in -> list of (index, size, box_x, box_y);
    with keep_uninteresting the objects that are not interesting stay in the list as (index, 0, 0, 0): they are
    painted from their frame-resolution mask (main.py:601-604, the `operations is None` branch)."""
    zooms, center2ds, depths = zooms.detach().float().cpu(), center2ds.detach().float().cpu(), depths.detach().float().cpu()
    order = torch.sort(depths[:, 0], dim=0, descending=True)[1].tolist()
    out = []
    for i in order:
        if not bool(interests[i]):
            if keep_uninteresting:
                out.append((i, 0, 0, 0))
            continue
        size = int(render_size / zooms[i])
        bx = int(center2ds[i, 1] * focal + u0 - size // 2)
        by = int(center2ds[i, 0] * focal + v0 - size // 2)
        out.append((i, size, bx, by))
    return out


def compose_scene(masks, normals, depth_maps, zooms, center2ds, depths, interests, height, width, focal, u0, v0,
                  render_size=None, image_masks=None, operations_given=True):
    """masks [N,1,S,S], normals [N,3,S,S] or None, depth_maps [N,1,S,S], zooms [N,1], center2ds [N,2], depths [N,1]
    (CUDA tensors), interests [N] -> instance [1,H,W], normal [3,H,W] (or None), depth [1,H,W] on the same device.

    `operations_given=False` is the reference's default test mode (`operations is None`, main.py:597-604): objects
    that are not interesting are still painted into the INSTANCE map, in the same far-to-near order, from their
    frame-resolution input masks `image_masks` [N,1,H,W] -- they get an instance id and occlude farther objects."""
    if not masks.is_cuda:
        raise sdn_b200.SdnError('compose_scene runs on CUDA tensors only (no CPU fallback)')
    dev = masks.device
    S = int(masks.shape[-1])
    render_size = S if render_size is None else render_size
    interests = torch.as_tensor(interests).cpu()
    keep = not operations_given
    if keep and image_masks is None:
        raise sdn_b200.SdnError('compose_scene: operations_given=False needs image_masks [N,1,H,W] (main.py:601-604)')
    place = [p for p in placement(zooms, center2ds, depths, interests, focal, u0, v0, render_size, keep)
             if p[1] > 0 or (keep and not bool(interests[p[0]]))]
    objs = (_Obj * max(len(place), 1))()
    bounds, kkd, kk8 = [], [], []
    row_off = k_off = tmp_off = 0
    cache = {}
    for m, (i, size, bx, by) in enumerate(place):
        if size == 0:               # mask-only entry: reads image_masks[i]
            objs[m] = _Obj(i, 0, 0, 0, 0, i + 1, 0, 0, 0)
            continue
        if size not in cache:
            cache[size] = resample_tables(S, size)
        b, kd, k8 = cache[size]
        objs[m] = _Obj(i, size, bx, by, kd.shape[1], i + 1, row_off, k_off, tmp_off)
        bounds.append(b); kkd.append(kd.reshape(-1)); kk8.append(k8.reshape(-1))
        row_off += size
        k_off += kd.size
        tmp_off += S * size
    M = len(place)
    n_rendered = len(bounds)
    t_bounds = torch.from_numpy(np.concatenate(bounds) if n_rendered else np.zeros((1, 2), np.int32)).to(dev)
    t_kkd = torch.from_numpy(np.concatenate(kkd) if n_rendered else np.zeros(1)).to(dev)
    t_kk8 = torch.from_numpy(np.concatenate(kk8) if n_rendered else np.zeros(1, np.int32)).to(dev)
    frame_masks = None
    if n_rendered < M:
        frame_masks = image_masks.detach().to(dev).float().reshape(-1, height, width).contiguous()
    t_objs = torch.frombuffer(bytearray(bytes(objs)), dtype=torch.uint8).to(dev)
    tmp8 = torch.empty(max(tmp_off, 1) * 4, dtype=torch.uint8, device=dev)
    tmpf = torch.empty(max(tmp_off, 1), dtype=torch.float32, device=dev)
    f32 = lambda t: None if t is None else t.detach().float().contiguous()      # noqa: E731
    masks, normals, depth_maps, zooms = f32(masks), f32(normals), f32(depth_maps), f32(zooms.reshape(-1)).to(dev)
    instance = torch.empty(1, height, width, device=dev)
    normal = torch.empty(3, height, width, device=dev) if normals is not None else None
    depth = torch.empty(1, height, width, device=dev)
    p = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())           # noqa: E731
    lib = sdn_b200.lib()
    lib.sdn_composite_scene.argtypes = [ctypes.c_void_p] * 1 + [ctypes.c_int] + [ctypes.c_void_p] * 4 + [ctypes.c_int] + \
        [ctypes.c_void_p] * 5 + [ctypes.c_int] * 3 + [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 4
    with torch.cuda.device(dev):
        rc = lib.sdn_composite_scene(p(t_objs), M, p(masks), p(normals), p(depth_maps), p(zooms), S, p(t_bounds),
                                     p(t_kk8), p(t_kkd), p(tmp8), p(tmpf), max([q[1] for q in place] + [1]), height,
                                     width, p(frame_masks), n_rendered, p(instance), p(normal), p(depth),
                                     ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc != 0:
        raise sdn_b200.SdnError('sdn_composite_scene: %s (code %d)' % (lib.sdn_error_string(rc).decode(), rc))
    return instance, normal, depth


def save_scene(image_dir, name, instance, normal, depth, json_obj):
    """The files of main.py:608-621 that the textural branch reads (the -visualize.png overlay is not part of the
    contract): instance ids as 8-bit PNG, per-object json, normals as RGB PNG, depth as 16-bit PNG."""
    import PIL.Image
    os.makedirs(image_dir, exist_ok=True)
    inst = np.uint8(instance.detach().cpu().numpy()[0])
    PIL.Image.fromarray(inst, mode='L').save(os.path.join(image_dir, '%s.png' % name))
    with open(os.path.join(image_dir, '%s.json' % name), 'w') as f:
        json.dump(json_obj, f, indent=4)
    if normal is not None:
        n = normal.detach().cpu().mul(255).byte().numpy().transpose(1, 2, 0)       # to_pil_image of a float tensor
        PIL.Image.fromarray(n, mode='RGB').save(os.path.join(image_dir, '%s-normal.png' % name))
    d16 = np.uint16(depth.detach().cpu().numpy()[0] * 65535)
    PIL.Image.fromarray(d16).save(os.path.join(image_dir, '%s-depth.png' % name))


def scene_json(place_indices, class_ids, depths, alphas, metas=None):
    """main.py:547-554: one entry per drawn object, keyed by instance id."""
    out = {}
    for i in place_indices:
        out[i + 1] = {'class_id': int(class_ids[i]), 'depth': float(depths[i]), 'alpha': float(alphas[i])}
        if metas is not None:
            out[i + 1].update(metas[i])
    return out

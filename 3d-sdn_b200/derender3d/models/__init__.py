"""Derenderer3d: encoder -> pose algebra -> class sampling -> FFD -> camera-space transform -> render
(reference: derender3d/models/__init__.py:18-250).

Same constructor, forward(images, roi_norms, focals) and render(blob) contract, same dictionary keys.  The reference
renders objects one at a time in a Python loop with >= 7 host round trips each; here the objects of
a batch are deformed and moved to camera space by ONE native launch over the ragged batch (sdn_objects_forward, the
mesh class is an index into a device-resident bank of templates) and rendered by ONE rasterisation; tensors never
leave the GPU and nothing syncs with the host, so the whole step can be captured in a CUDA graph.  On the CPU (host
tests) and with `fused_front_end = False` the objects are grouped by class and each group goes through the torch
FFD / PerspectiveTransform modules (<= 8 render calls, one host sync for the class ids)."""
import os

import numpy as np
import torch
from torch.distributions import Categorical
from torch.nn.modules import Module

import neural_renderer as nr

from derender3d import TargetType
from derender3d.models.derenderer import Derenderer
from derender3d.models.renderer import RenderType, Renderer
from derender3d.models.transforms import FFD, PerspectiveTransform

SHAPENET_OBJS = [           # derender3d/models/__init__.py:50-59
    ('02958343', '137f67657cdc9da5f985cd98f7d73e9a'),
    ('02958343', '5343e944a7753108aa69dfdc5532bb13'),
    ('02958343', '3776e4d1e2587fd3253c03b7df20edd5'),
    ('02958343', '3ba5bce1b29f0be725f689444c7effe2'),
    ('02958343', '53a031dd120e81dc3aa562f24645e326'),
    ('02924116', '7905d83af08a0ca6dafc1d33c05cbcf8'),
    ('02958343', 'a0fe4aac120d5f8a5145cad7315443b3'),
    ('02958343', 'cd7feedd6041209131ac5fb37e6c8324'),
]


class ShapenetObj(object):
    """One normalised template mesh (reference :18-35): unit extent per axis, x<->z swapped, x negated."""
    root_dir = os.getenv('SHAPENET_ROOT_DIR')

    def __init__(self, class_id=None, obj_id=None, root_dir=None, vertices=None, faces=None):
        if vertices is None:
            root_dir = root_dir or ShapenetObj.root_dir or os.getenv('SHAPENET_ROOT_DIR')
            path = os.path.join(root_dir, class_id, obj_id, 'models', 'model_normalized.obj')
            vertices, faces = nr.load_obj(path)
        vertices = np.asarray(vertices, np.float32)
        vertices = vertices / np.ptp(vertices, axis=0)
        vertices = vertices[:, [2, 1, 0]] * np.asarray([-1, 1, 1], dtype=np.float32)
        self.vertices = torch.tensor(vertices, dtype=torch.float32)
        self.faces = torch.tensor(np.asarray(faces), dtype=torch.int32)


class Derenderer3d(Module):
    def __init__(self, mode, image_size, render_size, objs=None):
        super(Derenderer3d, self).__init__()
        self.mode = mode
        self.image_size = image_size
        self.render_size = render_size
        self.derenderer = Derenderer()
        self._force_no_sample = False
        if mode & TargetType.reproject:
            self.objs = objs if objs is not None else [ShapenetObj(c, o) for c, o in SHAPENET_OBJS]
            self.ffds = torch.nn.ModuleList([FFD(obj.vertices, constraints=[
                FFD.Constraint.symmetry(axis=FFD.Constraint.Axis.z),
                FFD.Constraint.homogeneity(axis=FFD.Constraint.Axis.y, index=[0, 1]),
            ]) for obj in self.objs])
            self.perspective_transform = PerspectiveTransform()
            self.renderer = Renderer(image_size=render_size)
            self._faces_dev = {}
            self._banks = {}
            self.fused_front_end = os.environ.get('SDN_FUSED_FRONT_END', '1') != '0'

    def forward(self, images, roi_norms, focals):
        _mroi_norms = torch.stack([roi_norms[:, 2] + roi_norms[:, 0], roi_norms[:, 3] + roi_norms[:, 1]], dim=1) / 2.0
        _droi_norms = torch.stack([roi_norms[:, 2] - roi_norms[:, 0], roi_norms[:, 3] - roi_norms[:, 1]], dim=1)
        _blob = {'_roi_norms': roi_norms, '_mroi_norms': _mroi_norms, '_droi_norms': _droi_norms, '_focals': focals}
        _blob.update(self.derenderer(images, _mroi_norms, _droi_norms))
        if not (self.mode & TargetType.reproject):
            return _blob
        _blob.update(self.render(_blob))
        return _blob

    def _faces(self, k, dev):
        key = (k, str(dev))
        if key not in self._faces_dev:
            self._faces_dev[key] = self.objs[k].faces.to(dev)
        return self._faces_dev[key]

    def pose(self, blob):
        """Pose algebra of render() (reference :107-150): everything up to the per-object loop."""
        _mroi_norms, _droi_norms, _focals = blob['_mroi_norms'], blob['_droi_norms'], blob['_focals']
        _theta_deltas, _translation2ds = blob['_theta_deltas'], blob['_translation2ds']
        _log_scales, _log_depths, _class_probs = blob['_log_scales'], blob['_log_depths'], blob['_class_probs']
        bs = len(_focals)
        dev = _focals.device
        zeros = torch.zeros(bs, 1, device=dev)
        _thetas = torch.unsqueeze(torch.atan2(_theta_deltas[:, 1], _theta_deltas[:, 0]), dim=1)
        _rotations = torch.cat([torch.cos(_thetas / 2), zeros, torch.sin(_thetas / 2), zeros], dim=1)
        _areas = torch.unsqueeze(_droi_norms[:, 0] * _droi_norms[:, 1], dim=1)
        _scales = torch.exp(_log_scales)
        _depths = torch.sqrt(torch.exp(_log_depths) / _areas)
        _center2ds = _mroi_norms + _translation2ds * _droi_norms
        ones = torch.ones(bs, device=dev)
        _tu = torch.stack([_center2ds[:, 1], -_center2ds[:, 0], -ones], dim=1)
        _tu = _tu / torch.norm(_tu, p=2, dim=1, keepdim=True)
        _translations = _depths * _tu
        _alphas = -(_thetas - torch.atan(_translations[:, 0:1] / _translations[:, 2:3]))
        _alphas = torch.remainder(_alphas + np.pi, 2 * np.pi) - np.pi
        if self.training and not self._force_no_sample:
            dist = Categorical(_class_probs, validate_args=False)       # the simplex check would sync with the host
            _class_samples = dist.sample()
            _class_log_probs = dist.log_prob(_class_samples)
        else:
            _class_max_probs, _class_samples = torch.max(_class_probs, dim=1)
            _class_log_probs = torch.log(_class_max_probs)
        out = {'_thetas': _thetas, '_alphas': _alphas, '_rotations': _rotations, '_scales': _scales,
               '_depths': _depths, '_center2ds': _center2ds, '_translations': _translations,
               '_class_log_probs': _class_log_probs, '_class_samples': _class_samples}
        if self.training:
            _ptu = torch.stack([_mroi_norms[:, 1], -_mroi_norms[:, 0], -ones], dim=1)
            _ptu = _ptu / torch.norm(_ptu, p=2, dim=1, keepdim=True)
            out['_perspective_translations'] = _depths * _ptu
            out['_zooms'] = (self.image_size / _focals) / torch.max(_droi_norms, dim=1, keepdim=True)[0]
        else:
            out['_zoom_tos'] = self.render_size / (2.0 * _focals)
        return out

    def object_vertices(self, blob, p, classes):
        """Camera-space vertices of every object, grouped by mesh class (reference :161-200, one object at a time
        there): yields (class id, batch indices, vertices [n,nv,3], zooms [n,1] or None)."""
        dev = blob['_focals'].device
        for k in sorted(set(classes)):
            sel = torch.tensor([i for i, c in enumerate(classes) if c == k], device=dev)
            verts = self.ffds[k](blob['_ffd_coeffs'][sel, k])                   # [n,nv,3]
            kw = dict(scales=p['_scales'][sel], rotations=p['_rotations'][sel], translations=p['_translations'][sel])
            if self.training:
                verts = self.perspective_transform(verts, perspective_translations=p['_perspective_translations'][sel],
                                                   zooms=p['_zooms'][sel], **kw)
                yield k, sel, verts, None
            else:
                verts, z = self.perspective_transform(verts, perspective_translations=p['_translations'][sel],
                                                      zoom_tos=p['_zoom_tos'][sel], **kw)
                yield k, sel, verts, z

    def _bank(self, dev):
        """Device-resident template bank + the FFD's constant tensors (built once per device)."""
        key = str(dev)
        if key not in self._banks:
            import sdn_b200
            bank = sdn_b200.MeshBank([f.B for f in self.ffds], [o.faces for o in self.objs], dev)
            f0 = self.ffds[0]
            sig = lambda f: [(c.type, c.axis, getattr(c, 'index', None)) for c in f.constraints]  # noqa: E731
            assert all(sig(f) == sig(f0) and f.num_grids == 4 for f in self.ffds), 'templates share one FFD lattice'
            bank.C = f0.constraint_matrix().to(dev)
            bank.P0 = f0.P0.reshape(1, -1).to(dev)
            self._banks[key] = bank
        return self._banks[key]

    def render_fused(self, blob):
        """render() with the per-object loop (reference :161-224) replaced by one ragged launch + one rasterisation."""
        import sdn_b200
        p = self.pose(blob)
        _focals = blob['_focals']
        bs, dev = len(_focals), _focals.device
        bank = self._bank(dev)
        classes = p['_class_samples']
        with torch.autocast('cuda', enabled=False):
            dP = blob['_ffd_coeffs'][torch.arange(bs, device=dev), classes].float()     # [bs,192]
            control = (bank.P0 + dP @ bank.C).reshape(bs, 3, 64)
            kw = dict(scales=p['_scales'].float(), rotations=p['_rotations'].float(),
                      translations=p['_translations'].float())
            if self.training:
                verts = sdn_b200.object_vertices(bank, classes, control, ptrans=p['_perspective_translations'].float(),
                                                 zooms=p['_zooms'].float(), **kw)
                _zooms = p['_zooms']
            else:
                verts, _zooms = sdn_b200.object_vertices(bank, classes, control, ptrans=kw['translations'],
                                                         zoom_tos=p['_zoom_tos'].float(), **kw)
        faces = bank.faces.index_select(0, classes)                                      # [bs,nf_max,3]
        want_normal = bool(self.mode & TargetType.normal)
        want_depth = bool(self.mode & TargetType.depth)
        self.renderer.viewing_angle = torch.atan(self.render_size / (2.0 * _focals.double().reshape(-1))) / np.pi * 180
        if want_normal or want_depth:
            out = self.renderer.forward_multi(verts, faces, silhouette=True, depth=want_depth, normal=want_normal)
            _masks = out['silhouette']
            _normals = out['normal'] if want_normal else []
            _depth_maps = out['depth'] if want_depth else []
        else:
            _masks = self.renderer(vertices=verts, faces=faces, render_type=RenderType.Silhouette)
            _normals, _depth_maps = [], []
        return {
            '_thetas': p['_thetas'], '_alphas': p['_alphas'], '_rotations': p['_rotations'], '_scales': p['_scales'],
            '_depths': p['_depths'], '_center2ds': p['_center2ds'], '_translations': p['_translations'],
            '_class_log_probs': p['_class_log_probs'], '_zooms': _zooms, '_masks': _masks, '_normals': _normals,
            '_depth_maps': _depth_maps,
        }

    def render(self, blob):
        if self.fused_front_end and blob['_focals'].is_cuda:
            return self.render_fused(blob)
        p = self.pose(blob)
        _focals, _ffd_coeffs = blob['_focals'], blob['_ffd_coeffs']
        bs, dev = len(_focals), _focals.device
        classes = p['_class_samples'].tolist()                          # the one host sync of the step
        want_normal = bool(self.mode & TargetType.normal)
        want_depth = bool(self.mode & TargetType.depth)
        S = self.render_size
        _masks = torch.empty(bs, 1, S, S, device=dev)
        _normals = torch.empty(bs, 3, S, S, device=dev) if want_normal else []
        _depth_maps = torch.empty(bs, 1, S, S, device=dev) if want_depth else []
        _zooms = p['_zooms'] if self.training else torch.empty(bs, 1, device=dev)
        # viewing angle per object: atan(render_size / (2 focal)) in degrees (reference :202), kept on the device
        angles = torch.atan(self.render_size / (2.0 * _focals.double().reshape(-1))) / np.pi * 180
        masks, normals, depths, order = [], [], [], []
        for k, sel, verts, z in self.object_vertices(blob, p, classes):
            if z is not None:
                _zooms = _zooms.index_copy(0, sel, z)
            self.renderer.viewing_angle = angles[sel]
            faces = self._faces(k, dev)
            if want_normal or want_depth:
                out = self.renderer.forward_multi(verts, faces, silhouette=True, depth=want_depth, normal=want_normal)
                masks.append(out['silhouette'])
                if want_normal:
                    normals.append(out['normal'])
                if want_depth:
                    depths.append(out['depth'])
            else:
                masks.append(self.renderer(vertices=verts, faces=faces, render_type=RenderType.Silhouette))
            order.append(sel)
        order = torch.cat(order)
        _masks = _masks.index_copy(0, order, torch.cat(masks))
        if want_normal:
            _normals = _normals.index_copy(0, order, torch.cat(normals))
        if want_depth:
            _depth_maps = _depth_maps.index_copy(0, order, torch.cat(depths))
        return {
            '_thetas': p['_thetas'], '_alphas': p['_alphas'], '_rotations': p['_rotations'], '_scales': p['_scales'],
            '_depths': p['_depths'], '_center2ds': p['_center2ds'], '_translations': p['_translations'],
            '_class_log_probs': p['_class_log_probs'], '_zooms': _zooms, '_masks': _masks, '_normals': _normals,
            '_depth_maps': _depth_maps,
        }

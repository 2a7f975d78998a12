"""torch boundary of the renderer (reference: derender3d/models/renderer.py:12-272).

Same names and call surface -- RenderType, RenderFunction.apply(...), Renderer(image_size, viewing_angle)(vertices,
faces, textures=None, render_type=RenderType.RGB) -- but tensors never leave the GPU: the reference's
torch -> numpy -> cupy -> chainer round trips (renderer.py:130-150,164-172,195-207) are replaced by direct calls into
libsdn_b200.so on the current CUDA stream."""
import torch
from torch.nn.modules import Module

import neural_renderer as nr


class RenderType:
    RGB = 0
    Silhouette = 1
    Depth = 2
    Normal = 3


class _Renderer(nr.Renderer):
    """nr.Renderer + render_normal (renderer.py:19-127); both live in neural_renderer.renderer here."""
    pass


class RenderFunction(object):
    """Call-compatible stand-in for the reference's autograd.Function (renderer.py:130-213).  `apply` composes the
    two native autograd nodes (faces, rasterise); gradients reach `vertices` and `textures` only, as in :213."""

    @staticmethod
    def apply(vertices, faces, textures, renderer, render_type, eye, camera_mode, camera_direction, camera_up):
        renderer.eye = eye
        renderer.camera_mode = camera_mode
        renderer.camera_direction = camera_direction
        renderer.up = camera_up
        if render_type == RenderType.RGB:
            return renderer.render(vertices, faces, textures)
        if render_type == RenderType.Silhouette:
            return renderer.render_silhouettes(vertices, faces).unsqueeze(1)
        if render_type == RenderType.Depth:
            return renderer.render_depth(vertices, faces).unsqueeze(1)
        if render_type == RenderType.Normal:
            return renderer.render_normal(vertices, faces)
        raise ValueError('unknown render_type %r' % (render_type,))


class Renderer(Module):
    """Module with the reference's attributes (renderer.py:216-229): image_size, viewing_angle (re-assigned per
    object by derender3d/models/__init__.py:202), eye, camera_mode, camera_direction, camera_up.

    forward(vertices [B,nv,3] f32 cuda, faces [B,nf,3] i32 cuda, textures=None, render_type) -> [B,C,S,S].
    The x mirror of renderer.py:243 and the final normal-x negation of :268-270 are folded into the face kernel."""

    def __init__(self, image_size=256, viewing_angle=30):
        super(Renderer, self).__init__()
        self.image_size = image_size
        self.viewing_angle = viewing_angle
        self.eye = torch.Tensor([0, 0, 0])
        self.camera_mode = 'look'
        self.camera_direction = torch.Tensor([0, 0, -1])
        self.camera_up = torch.Tensor([0, 1, 0])

    def _native(self):
        r = _Renderer()
        r.image_size = self.image_size
        r.viewing_angle = self.viewing_angle
        r.flip_x = -1.0
        return r

    def _camera_args(self, vertices):
        # the module's camera tensors live on the host (renderer.py:224-229); they stay there so that the native
        # renderer can evaluate and cache the (constant) view rotation once per (batch size, device)
        bs = len(vertices)
        expand = lambda t: t[None, :].expand(bs, -1)
        return expand(self.eye), self.camera_mode, expand(self.camera_direction), expand(self.camera_up)

    def forward(self, vertices, faces, textures=None, render_type=RenderType.RGB):
        eye, mode, direction, up = self._camera_args(vertices)
        return RenderFunction.apply(vertices, faces, textures, self._native(), render_type, eye, mode, direction, up)

    def forward_multi(self, vertices, faces, silhouette=True, depth=True, normal=True, fast_math=None):
        """Silhouette [B,1,S,S], depth [B,1,S,S] and normal [B,3,S,S] of the same objects in ONE rasterisation."""
        r = self._native()
        r.eye, r.camera_mode, r.camera_direction, r.up = self._camera_args(vertices)
        out = r.render_multi(vertices, faces, silhouette, depth, normal, fast_math=fast_math)
        self.last_face_index = r.last_face_index
        return {'silhouette': None if out['silhouette'] is None else out['silhouette'].unsqueeze(1),
                'depth': None if out['depth'] is None else out['depth'].unsqueeze(1),
                'normal': out['normal']}

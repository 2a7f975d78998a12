"""Functional rasteriser API (reference: rasterize.py:897-1062), backed by libsdn_b200.so."""
import os

from sdn_b200 import RasterConfig, RasterizeFunction, hw

DEFAULT_IMAGE_SIZE = 256
DEFAULT_ANTI_ALIASING = True
DEFAULT_NEAR = 0.1
DEFAULT_FAR = 100
DEFAULT_EPS = 1e-4
DEFAULT_BACKGROUND_COLOR = (0, 0, 0)
# The reference picks a racy per-face kernel when NEURAL_RENDERER_UNSAFE=1 (rasterize.py:13-16, scripts/env.sh:11).
# Here there is one deterministic kernel; the switch is honoured as a no-op.
USE_UNSAFE_IMPLEMENTATION = bool(int(os.environ.get('NEURAL_RENDERER_UNSAFE', '0') or 0))


class Rasterize(object):
    """The reference's chainer.Function as a callable (rasterize.py:19-64, forward_gpu :464-510): same constructor,
    same return contract -- `Rasterize(image_size, near, far, eps, background_color, return_rgb, return_alpha,
    return_depth)(faces[, textures])` -> (rgb_map [B,is,is,3] | None, alpha_map [B,is,is] | None,
    depth_map [B,is,is] | None) at raster resolution `image_size`, channel-LAST and in RASTER orientation (row 0 is the
    bottom of the image: the flip / transpose / 2x2 mean belong to rasterize_rgbad, :950-966).  Differentiable in
    faces and textures.  `image_size` may also be (H, W)."""

    def __init__(self, image_size, near, far, eps, background_color, return_rgb=False, return_alpha=False,
                 return_depth=False):
        if not any((return_rgb, return_alpha, return_depth)):
            raise Exception     # nothing to draw, rasterize.py:25-27
        self.image_size, self.near, self.far, self.eps = image_size, near, far, eps
        self.background_color = background_color
        self.return_rgb, self.return_alpha, self.return_depth = return_rgb, return_alpha, return_depth
        self.cfg = RasterConfig(image_size, False, near, far, eps, background_color, return_rgb, return_alpha,
                                return_depth)

    def __call__(self, faces, textures=None):
        rgb, alpha, depth = RasterizeFunction.apply(faces, textures if self.return_rgb else None, self.cfg)
        self.face_index_map = self.cfg.last_face_index      # int32 [B,is,is], raster orientation (rasterize.py:475)
        return (rgb.flip(2).permute(0, 2, 3, 1) if self.return_rgb else None,
                alpha.flip(1) if self.return_alpha else None,
                depth.flip(1) if self.return_depth else None)


def rasterize_rgbad(faces, textures=None, image_size=DEFAULT_IMAGE_SIZE, anti_aliasing=DEFAULT_ANTI_ALIASING,
                    near=DEFAULT_NEAR, far=DEFAULT_FAR, eps=DEFAULT_EPS, background_color=DEFAULT_BACKGROUND_COLOR,
                    return_rgb=True, return_alpha=True, return_depth=True):
    """RGB / alpha / depth images from NDC faces [B,F,3,3] and textures [B,F,ts,ts,ts,3] (rasterize.py:897-974).
    image_size may be an int (the reference) or (H, W).  Returns {'rgb': [B,3,H,W], 'alpha': [B,H,W],
    'depth': [B,H,W]} with None for outputs not requested."""
    cfg = RasterConfig(image_size, anti_aliasing, near, far, eps, background_color, return_rgb, return_alpha,
                       return_depth)
    rgb, alpha, depth = RasterizeFunction.apply(faces, textures if return_rgb else None, cfg)
    return {'rgb': rgb if return_rgb else None, 'alpha': alpha if return_alpha else None,
            'depth': depth if return_depth else None}


def rasterize(faces, textures, image_size=DEFAULT_IMAGE_SIZE, anti_aliasing=DEFAULT_ANTI_ALIASING,
              near=DEFAULT_NEAR, far=DEFAULT_FAR, eps=DEFAULT_EPS, background_color=DEFAULT_BACKGROUND_COLOR):
    """RGB images [B,3,H,W] (rasterize.py:977-1005)."""
    return rasterize_rgbad(faces, textures, image_size, anti_aliasing, near, far, eps, background_color,
                           True, False, False)['rgb']


def rasterize_silhouettes(faces, image_size=DEFAULT_IMAGE_SIZE, anti_aliasing=DEFAULT_ANTI_ALIASING,
                          near=DEFAULT_NEAR, far=DEFAULT_FAR, eps=DEFAULT_EPS):
    """Alpha channels [B,H,W] (rasterize.py:1008-1031)."""
    return rasterize_rgbad(faces, None, image_size, anti_aliasing, near, far, eps, None, False, True, False)['alpha']


def rasterize_depth(faces, image_size=DEFAULT_IMAGE_SIZE, anti_aliasing=DEFAULT_ANTI_ALIASING,
                    near=DEFAULT_NEAR, far=DEFAULT_FAR, eps=DEFAULT_EPS):
    """Depth images [B,H,W] (rasterize.py:1034-1057)."""
    return rasterize_rgbad(faces, None, image_size, anti_aliasing, near, far, eps, None, False, False, True)['depth']


def use_unsafe_rasterizer(flag):
    """Kept for API compatibility (rasterize.py:1060-1062); there is only the deterministic kernel."""
    global USE_UNSAFE_IMPLEMENTATION
    USE_UNSAFE_IMPLEMENTATION = flag

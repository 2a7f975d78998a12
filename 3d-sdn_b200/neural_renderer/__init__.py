"""neural_renderer -- torch-typed, B200-native drop-in for /root/reference/geometric/neural_renderer (v1.1.3 API,
__init__.py:2-15).  Same names, same argument meaning; tensors are torch CUDA tensors instead of chainer Variables.
The rasteriser and its gradient run in libsdn_b200.so (hand-written sm_100a CUDA); there is no CPU path, exactly as in
the reference (rasterize.py:890-894).

Not mirrored: `optimizers.Adam` (a Chainer update rule; torch.optim.Adam takes its place)."""
# the renderer object and the functional rasteriser API
from .renderer import Renderer
from .rasterize import (Rasterize, rasterize, rasterize_depth, rasterize_rgbad, rasterize_silhouettes,
                        use_unsafe_rasterizer)
# camera and geometry helpers
from .look import look
from .look_at import look_at
from .perspective import perspective, perspective_width
from .get_points_from_angles import get_points_from_angles
from .vertices_to_faces import vertices_to_faces
from .cross import cross
from .lighting import lighting
# meshes on disk
from .load_obj import load_obj
from .save_obj import save_obj
from .mesh import Mesh

__version__ = '1.1.3+b200'
__all__ = ['Renderer', 'Rasterize', 'rasterize', 'rasterize_depth', 'rasterize_rgbad', 'rasterize_silhouettes',
           'use_unsafe_rasterizer', 'look', 'look_at', 'perspective', 'perspective_width', 'get_points_from_angles',
           'vertices_to_faces', 'cross', 'lighting', 'load_obj', 'save_obj', 'Mesh']

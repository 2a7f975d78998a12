"""neural_renderer -- torch-typed, B200-native drop-in for /root/reference/geometric/neural_renderer (v1.1.3 API,
__init__.py:2-15).  Same names, same argument meaning; tensors are torch CUDA tensors instead of chainer Variables.
The rasteriser and its gradient run in libsdn_b200.so (hand-written sm_100a CUDA); there is no CPU path, exactly as in
the reference (rasterize.py:890-894)."""
from .cross import cross
from .get_points_from_angles import get_points_from_angles
from .lighting import lighting
from .load_obj import load_obj
from .look import look
from .look_at import look_at
from .mesh import Mesh
from .perspective import perspective, perspective_width
from .rasterize import (
    rasterize_rgbad, rasterize, rasterize_silhouettes, rasterize_depth, use_unsafe_rasterizer, Rasterize)
from .renderer import Renderer
from .save_obj import save_obj
from .vertices_to_faces import vertices_to_faces

__version__ = '1.1.3+b200'

"""The C-ABI shared library loads on a machine without a GPU and exports every symbol include/sdn_b200.h declares;
argument errors come back as codes, never as exceptions or exits."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, 'include', 'sdn_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(sdn_[a-z_]+)\s*\(', text)))


def test_header_symbols_exported():
    import sdn_b200
    lib = sdn_b200.lib()
    names = declared_functions()
    assert len(names) >= 9
    for n in names:
        assert hasattr(lib, n), 'libsdn_b200.so does not export %s' % n
    assert lib.sdn_version() == 201
    assert sdn_b200.launch_count() >= 0


def test_error_codes_without_gpu():
    import sdn_b200
    lib = sdn_b200.lib()
    assert lib.sdn_workspace_bytes(0, 10, 8, 8) == 0
    n1, n2 = lib.sdn_workspace_bytes(1, 100, 64, 64), lib.sdn_workspace_bytes(2, 100, 64, 64)
    assert 0 < n1 < n2
    r = sdn_b200._Raster()                      # all zeros: invalid
    assert lib.sdn_rasterize_forward(ctypes.byref(r), None) == -1
    r.B, r.F, r.H, r.W, r.ss, r.flags = 1, 4, 8, 8, 2, 0
    assert lib.sdn_rasterize_forward(ctypes.byref(r), None) == -3        # nothing to draw (rasterize.py:25-27)
    assert b'nothing' in lib.sdn_error_string(-3)
    assert lib.sdn_faces_forward(None, None, 1, 1, 1, 0, 1, None, None, None, None) == -1


def test_no_cpu_fallback():
    import torch
    import sdn_b200
    import neural_renderer as nr
    v = torch.zeros(1, 3, 3)
    f = torch.zeros(1, 1, 3, dtype=torch.int32)
    with pytest.raises(sdn_b200.SdnError):
        nr.Renderer().render_silhouettes(v, f)
    with pytest.raises(Exception):
        nr.rasterize_rgbad(torch.zeros(1, 1, 3, 3), None, 8, True, return_rgb=False, return_alpha=False,
                           return_depth=False)

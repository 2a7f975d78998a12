"""Compatibility shims that let the reference's 2018 driver (`geometric/scripts/main.py`, Python 3.6 / torch 0.4 /
chainer 4) import and run UNCHANGED on Python 3.12 / torch 2.x with this repo standing in for `neural_renderer` and
`derender3d.models`.

Nothing here is on the render hot path: every shim stands for a module or attribute the driver imports around it.

    import sdn_b200.compat as compat
    compat.install('/path/to/3D-SDN/geometric')     # shims + sys.path (this repo first, then the reference)
    main = compat.load_main()                        # the reference's scripts/main.py, executed as a module

What is provided, and why (reference file:line):
  chainer                 `import chainer` at scripts/main.py:3 -- never used there; an empty module
  tensorboardX            scripts/main.py:14 `SummaryWriter` -> torch.utils.tensorboard.SummaryWriter (no-op writer if
                          tensorboard is not installed)
  collections.Mapping...  derender3d/data_loader.py:33 (moved to collections.abc in Python 3.10)
  np.str / np.int / ...   derender3d/datasets.py:343,388, maskrcnn/vkitti.py:101 (removed in numpy 1.24)
  matplotlib(.cm/.pyplot/.patches/.lines), skimage(.color/.io/.measure/.transform), scipy.misc.imresize/imread/imsave
                          derender3d/datasets.py:2,109 and maskrcnn/{utils,visualize}.py -- import-time presence; `cm.jet`
                          is implemented (it colours a debug visualisation), the rest raise when called
  torch.utils.ffi         maskrcnn/{nms,roialign}/build.py:3 (removed in torch 1.0)
  nms._ext / roi_align._ext
                          the compiled torch-0.4 FFI extensions of Mask R-CNN (scripts/main.py:28-29 imports the model;
                          `--source gt` never calls them): modules whose functions raise when called
  torchvision resnet18(pretrained=True)
                          derender3d/models/derenderer.py:25 downloads ImageNet weights; this repo's Derenderer builds
                          the same architecture offline (RESNET18_WEIGHTS=<file> loads them when available)
"""
import importlib
import importlib.util
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.dirname(_HERE)                       # .../3d-sdn_b200: neural_renderer/, derender3d/, sdn_b200/
REPO = os.path.dirname(PKG_DIR)

_installed = {}


def find_reference_geometric(path=None):
    """The reference's `geometric/` directory: the argument, $SDN_REFERENCE_GEOMETRIC, /root/reference/geometric, or
    the git-ignored travelling copy baseline/_ref/geometric (made by __graft_entry__.build()).  None if absent."""
    cands = [path, os.environ.get('SDN_REFERENCE_GEOMETRIC'), '/root/reference/geometric',
             os.path.join(REPO, 'baseline', '_ref', 'geometric')]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, 'scripts', 'main.py')):
            return os.path.abspath(c)
    return None


def _module(name, **attrs):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        m.__sdn_compat_stub__ = True
        sys.modules[name] = m
        if '.' in name:
            parent, leaf = name.rsplit('.', 1)
            setattr(_module(parent), leaf, m)
    for k, v in attrs.items():
        if not hasattr(m, k):
            setattr(m, k, v)
    return m


def _absent(name):
    if name in sys.modules:
        return False
    try:
        return importlib.util.find_spec(name) is None
    except (ImportError, ValueError):
        return True


def _raiser(what):
    def f(*a, **k):
        raise NotImplementedError('%s is a sdn_b200.compat import-time stub (not on the render path); install the real '
                                  'package to use it' % what)
    f.__name__ = what.split('.')[-1]
    return f


def _jet(x):
    """matplotlib.cm.jet on an array in [0, 1] -> RGBA float64 (piecewise-linear jet; datasets.py:109 uses it to colour
    a debug overlay)."""
    import numpy as np
    x = np.clip(np.asarray(x, dtype=np.float64), 0.0, 1.0)
    r = np.clip(np.minimum(4 * x - 1.5, -4 * x + 4.5), 0, 1)
    g = np.clip(np.minimum(4 * x - 0.5, -4 * x + 3.5), 0, 1)
    b = np.clip(np.minimum(4 * x + 0.5, -4 * x + 2.5), 0, 1)
    return np.stack([r, g, b, np.ones_like(x)], axis=-1)


class _NullWriter(object):
    """tensorboardX.SummaryWriter stand-in when tensorboard is not installed: accepts every call, writes nothing."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        return lambda *a, **k: None


def install_python_shims():
    """Interpreter / library renames.  Idempotent."""
    import collections
    import collections.abc
    for n in ('Mapping', 'MutableMapping', 'Sequence', 'Iterable', 'Callable', 'Set', 'Hashable', 'Sized', 'Container'):
        if not hasattr(collections, n):
            setattr(collections, n, getattr(collections.abc, n))
    import numpy as np
    for n, t in (('str', str), ('int', int), ('float', float), ('bool', bool), ('object', object), ('long', int),
                 ('unicode', str), ('complex', complex)):
        if n not in np.__dict__:
            setattr(np, n, t)
    _installed['python'] = True


def install_module_shims():
    """Stub modules for packages the driver imports but the hot path never calls.  Real packages win when present."""
    if _absent('chainer'):
        _module('chainer', __version__='0.0-sdn-compat-stub')
    if _absent('tensorboardX'):
        try:
            from torch.utils.tensorboard import SummaryWriter
        except Exception:       # noqa: BLE001 -- tensorboard not installed
            SummaryWriter = _NullWriter
        _module('tensorboardX', SummaryWriter=SummaryWriter)
    if _absent('matplotlib'):
        _module('matplotlib', use=lambda *a, **k: None)
        cm = _module('matplotlib.cm', jet=_jet)
        _module('matplotlib.pyplot', cm=cm, switch_backend=lambda *a, **k: None,
                **{n: _raiser('matplotlib.pyplot.' + n) for n in
                   ('figure', 'subplots', 'imshow', 'show', 'plot', 'savefig', 'axis')})
        _module('matplotlib.patches', Polygon=_raiser('matplotlib.patches.Polygon'),
                Rectangle=_raiser('matplotlib.patches.Rectangle'))
        _module('matplotlib.lines', Line2D=_raiser('matplotlib.lines.Line2D'))
    if _absent('skimage'):
        _module('skimage')
        _module('skimage.color', gray2rgb=_raiser('skimage.color.gray2rgb'))
        _module('skimage.io', imread=_raiser('skimage.io.imread'), imsave=_raiser('skimage.io.imsave'))
        _module('skimage.measure', find_contours=_raiser('skimage.measure.find_contours'))
        _module('skimage.transform', resize=_raiser('skimage.transform.resize'))
    try:
        import scipy.misc as sm
        for n in ('imresize', 'imread', 'imsave'):
            if not hasattr(sm, n):
                setattr(sm, n, _raiser('scipy.misc.' + n))
    except Exception:           # noqa: BLE001
        pass
    import torch.utils
    if _absent('torch.utils.ffi'):
        _module('torch.utils.ffi', create_extension=_raiser('torch.utils.ffi.create_extension'))
    _installed['modules'] = True


_EXT_STUBS = {}


def install_maskrcnn_ext_stubs():
    """maskrcnn/model.py:30-31 imports `nms.nms_wrapper` and `roialign.roi_align.crop_and_resize`, which import torch-0.4
    FFI extensions (`._ext`) built by torch.utils.ffi.  Stand-ins that import and raise on use (source 'gt' never calls
    Mask R-CNN).  They are served by a meta-path finder, because their parent packages (`nms`, `roialign.roi_align`)
    are real directories of the reference on disk."""
    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__sdn_compat_stub__ = True
        m.__path__ = []
        for k, v in attrs.items():
            setattr(m, k, v)
        _EXT_STUBS[name] = m
        return m
    nms = stub('nms._ext.nms', gpu_nms=_raiser('nms._ext.nms.gpu_nms'), cpu_nms=_raiser('nms._ext.nms.cpu_nms'))
    stub('nms._ext', nms=nms)
    car = stub('roialign.roi_align._ext.crop_and_resize',
               **{fn: _raiser('roi_align._ext.' + fn) for fn in
                  ('crop_and_resize_gpu_forward', 'crop_and_resize_gpu_backward', 'crop_and_resize_forward',
                   'crop_and_resize_backward')})
    stub('roialign.roi_align._ext', crop_and_resize=car)
    if not any(isinstance(f, _ExtStubFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _ExtStubFinder())
    _installed['maskrcnn'] = True


class _ExtStubFinder(object):
    def find_spec(self, fullname, path=None, target=None):
        m = _EXT_STUBS.get(fullname)
        if m is None:
            return None
        return importlib.util.spec_from_loader(fullname, _ExtStubLoader(m), is_package=True)


class _ExtStubLoader(object):
    def __init__(self, module):
        self.module = module

    def create_module(self, spec):
        return self.module

    def exec_module(self, module):
        pass


def install_paths(reference_geometric=None):
    """sys.path: this repo's packages FIRST (so `neural_renderer`, `derender3d.models` are the B200 ones), then the
    reference's `geometric/` (maskrcnn, the rest of derender3d via its __path__ extension) and `geometric/bulb`."""
    ref = find_reference_geometric(reference_geometric)
    if ref:
        os.environ.setdefault('SDN_REFERENCE_GEOMETRIC', ref)
        os.environ.setdefault('SHAPENET_ROOT_DIR', os.path.join(ref, 'assets'))
    want = [PKG_DIR] + ([ref, os.path.join(ref, 'bulb')] if ref else [])
    for p in reversed(want):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    # a `derender3d` imported before the reference was known has a one-entry __path__: extend it now
    d = sys.modules.get('derender3d')
    if d is not None and ref and getattr(d, '__file__', '').startswith(PKG_DIR):
        rd = os.path.join(ref, 'derender3d')
        if rd not in d.__path__:
            d.__path__.append(rd)
    _installed['paths'] = ref
    return ref


def install(reference_geometric=None):
    """All shims + paths.  Returns the reference's `geometric/` directory (None when it cannot be found)."""
    install_python_shims()
    install_module_shims()
    install_maskrcnn_ext_stubs()
    return install_paths(reference_geometric)


def load_main(reference_geometric=None, name='sdn_reference_main'):
    """Execute the reference's UNMODIFIED scripts/main.py as module `name` (its `if __name__ == '__main__'` block does
    not run) and return it.  absl flags are defined at import, so a second load returns the first module."""
    if name in sys.modules:
        return sys.modules[name]
    ref = install(reference_geometric)
    if not ref:
        raise FileNotFoundError('reference geometric/ directory not found (set SDN_REFERENCE_GEOMETRIC)')
    spec = importlib.util.spec_from_file_location(name, os.path.join(ref, 'scripts', 'main.py'))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    try:
        spec.loader.exec_module(mod)
    except BaseException:
        sys.modules.pop(name, None)
        raise
    return mod


def run_main(argv, reference_geometric=None):
    """`python scripts/main.py <argv>` of the reference, against this repo's renderer (scripts/run_reference_main.py)."""
    main = load_main(reference_geometric)
    flags = main.FLAGS
    flags([os.path.join(_installed.get('paths') or '', 'scripts', 'main.py')] + list(argv))
    if flags.mode is not None and not isinstance(flags.mode, int):
        flags.mode = main.TargetType.__dict__[flags.mode]      # scripts/main.py:846-847
    return getattr(main, flags.do)()

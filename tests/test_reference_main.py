"""The drop-in boundary at the Python level: the reference's UNMODIFIED geometric/scripts/main.py imports and runs with
this repo standing in for `neural_renderer` and `derender3d.models`.

The reference tree is read from /root/reference/geometric (build container) or from the git-ignored travelling copy
baseline/_ref/geometric that __graft_entry__.build() stages (GPU box); without either the tests skip."""
import os

import numpy as np
import pytest
import torch


def _main():
    import sdn_b200.compat as compat
    ref = compat.find_reference_geometric()
    if ref is None:
        pytest.skip('reference geometric/ tree not available (neither /root/reference nor baseline/_ref)')
    main = compat.load_main(ref)
    return compat, ref, main


def test_compat_shims_are_idempotent_and_local():
    import sdn_b200.compat as compat
    compat.install_python_shims()
    compat.install_python_shims()
    import collections
    assert collections.Mapping is collections.abc.Mapping
    assert np.str is str and np.int is int
    compat.install_module_shims()
    import chainer  # noqa: F401
    from tensorboardX import SummaryWriter  # noqa: F401
    import matplotlib.cm
    rgba = matplotlib.cm.jet(np.linspace(0, 1, 5))
    assert rgba.shape == (5, 4) and rgba[0, 2] > rgba[0, 0] and rgba[-1, 0] > rgba[-1, 2]      # blue -> red


def test_reference_main_imports_against_this_package():
    compat, ref, main = _main()
    import derender3d
    import derender3d.datasets
    import derender3d.models
    import derender3d.models.renderer
    import neural_renderer
    pkg = compat.PKG_DIR
    # the hot path resolves HERE ...
    assert os.path.abspath(neural_renderer.__file__).startswith(pkg)
    assert os.path.abspath(derender3d.models.__file__).startswith(pkg)
    assert os.path.abspath(derender3d.models.renderer.__file__).startswith(pkg)
    # ... everything else of derender3d, bulb and maskrcnn is the reference's own file
    assert os.path.abspath(derender3d.datasets.__file__).startswith(ref)
    assert main.Transforms is derender3d.datasets.Transforms
    import bulb.net
    assert os.path.abspath(bulb.net.__file__).startswith(ref) and main.Net is bulb.net.Net
    assert issubclass(main.Model, derender3d.models.Derenderer3d)
    assert main.TargetType.full == 3 and main.TargetType.extend == 15
    for name in ('train', 'test', '_test', '_test_2d', 'BaseNet', 'TrainNet', 'TestNet', 'TrainLoader'):
        assert hasattr(main, name)


def test_reference_model_builds_from_the_real_templates():
    """Model() = Derenderer3d(mode, image_size, render_size) with the eight ShapeNet templates of geometric/assets."""
    compat, ref, main = _main()
    main.FLAGS(['main.py', '--do', 'train', '--mode', 'full', '--dataset', 'vkitti'])
    main.FLAGS.mode = main.TargetType.__dict__[main.FLAGS.mode] if isinstance(main.FLAGS.mode, str) else main.FLAGS.mode
    model = main.Model()
    assert len(model.objs) == 8 and len(model.ffds) == 8
    nf = [int(o.faces.shape[0]) for o in model.objs]
    assert min(nf) > 20000 and max(nf) < 80000, nf          # SURVEY: F = 2 nf in 58k..145k
    keys = set(model.state_dict().keys())
    assert {'derenderer.net.conv1.weight', 'derenderer.fc1.weight', 'derenderer._fc3.bias'} <= keys
    assert sum(p.numel() for p in model.parameters()) == 11839312      # 47.4 MB of fp32 gradients
    # the encoder alone (mode pretrain needs no renderer) runs on the CPU through the reference's call
    blob = model.derenderer(torch.randn(2, 3, 224, 224), torch.zeros(2, 2), torch.ones(2, 2))
    assert blob['_ffd_coeffs'].shape == (2, 8, 192) and blob['_class_probs'].shape == (2, 8)


@pytest.mark.gpu
def test_reference_train_net_steps_on_a_synthetic_batch():
    """One optimisation step of the reference's own TrainNet (bulb Net.step_epoch -> BaseNet.step_batch,
    scripts/main.py:114-154) over DataParallel(Model()) with a synthetic VKITTI-shaped batch: forward through this
    repo's encoder + fused front end + rasteriser, the reference's losses, backward through the native kernels, Adam."""
    compat, ref, main = _main()
    assert torch.cuda.is_available()
    main.FLAGS(['main.py', '--do', 'train', '--mode', 'full', '--dataset', 'vkitti', '--render_size', '128',
                '--image_steps', '1000000'])
    main.FLAGS.mode = main.TargetType.__dict__[main.FLAGS.mode] if isinstance(main.FLAGS.mode, str) else main.FLAGS.mode
    torch.manual_seed(0)
    model = torch.nn.DataParallel(main.Model(), device_ids=[0]).cuda()
    B = 6
    g = torch.Generator().manual_seed(1)
    c = 0.3 * (torch.rand(B, 2, generator=g) - 0.5)
    d = 0.25 + 0.2 * torch.rand(B, 2, generator=g)
    batch = {
        'images': torch.randn(B, 3, 224, 224, generator=g).numpy(),
        'roi_norms': torch.cat([c - d / 2, c + d / 2], dim=1).numpy(),
        'focals': np.full((B, 1), 725.0, np.float32),
        'targets': np.full((B,), main.TargetType.full, np.int64),
        'thetas': torch.rand(B, 1, generator=g).numpy(),
        'translation2ds': np.zeros((B, 2), np.float32),
        'log_scales': np.tile(np.array([[1.3, 0.4, 0.5]], np.float32), (B, 1)),
        'log_depths': np.full((B, 1), 2.5, np.float32),
        'masks': (torch.rand(B, 1, 100, 100, generator=g) > 0.5).float().numpy(),
        'ignores': np.zeros((B, 1, 100, 100), np.float32),
    }
    net = main.TrainNet(optimizer=torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=1e-3), model=model,
                        writer=compat._NullWriter(), data_loader=[batch], lr_decay_epochs=None, summarize_steps=1, save_steps=None,
                        saver=None)
    before = model.module.derenderer._fc3.weight.detach().clone()
    import sdn_b200
    launches0 = sdn_b200.launch_count()
    net.step_epoch()
    assert sdn_b200.launch_count() > launches0, 'the native library was not on the path of the reference step'
    for k in ('loss', 'mask_loss', 'class_reward', 'ffd_coeff_reg', 'theta_delta_loss', 'scale_loss', 'depth_loss'):
        assert np.isfinite(getattr(net, k)), k
    assert net._masks.shape == (B, 1, 128, 128) and float(net._masks.sum()) > 0
    assert not torch.equal(before, model.module.derenderer._fc3.weight)

"""derender3d -- B200-native mirror of /root/reference/geometric/derender3d (the parts on the render hot path).

Only `derender3d.models` (and the additive train_step / optimise / compositor modules) live here.  The reference's
`derender3d.datasets`, `derender3d.data_loader` and `derender3d.utils` -- which `scripts/main.py:22-25` imports -- are NOT
re-implemented: when the reference's `geometric/` directory is known (SDN_REFERENCE_GEOMETRIC, or the locations
sdn_b200.compat.find_reference_geometric() probes) its `derender3d/` directory is appended to this package's
`__path__`, so those sub-modules resolve to the reference's own files while `derender3d.models` resolves here."""
import os as _os


class TargetType:
    """Mode bit-mask (reference: derender3d/__init__.py:1-10): which heads are supervised / which maps are rendered."""
    geometry, reproject, normal, depth = (1 << bit for bit in range(4))
    pretrain, finetune = geometry, reproject        # stage names used by scripts/main.py
    full = pretrain | finetune
    extend = full | normal | depth


def _extend_path():
    here = _os.path.dirname(_os.path.abspath(__file__))
    repo = _os.path.dirname(_os.path.dirname(here))
    for cand in (_os.environ.get('SDN_REFERENCE_GEOMETRIC'), '/root/reference/geometric',
                 _os.path.join(repo, 'baseline', '_ref', 'geometric')):
        if cand and _os.path.isfile(_os.path.join(cand, 'derender3d', 'datasets.py')):
            d = _os.path.join(_os.path.abspath(cand), 'derender3d')
            if d not in __path__:
                __path__.append(d)
            return


_extend_path()

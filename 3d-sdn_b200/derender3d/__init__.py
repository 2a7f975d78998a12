"""derender3d -- B200-native mirror of /root/reference/geometric/derender3d (the parts on the render hot path)."""


class TargetType:
    """Mode bit-mask (reference: derender3d/__init__.py:1-10): which heads are supervised / which maps are rendered."""
    geometry, reproject, normal, depth = (1 << bit for bit in range(4))
    pretrain, finetune = geometry, reproject        # stage names used by scripts/main.py
    full = pretrain | finetune
    extend = full | normal | depth

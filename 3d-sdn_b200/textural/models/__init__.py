from .networks import GlobalGenerator, ResnetBlock, get_norm_layer  # noqa: F401

"""Drop-in mirror of the reference's `deeplens` package for the dual-pixel ray-tracing hot path (flat re-exports,
like deeplens/__init__.py:1-8 of the reference)."""
from .basics import *          # noqa: F401,F403
from .basics import DeepObj, Material, Ray  # noqa: F401
from .surfaces import Aspheric, Surface     # noqa: F401
from .monte_carlo import forward_integral, forward_integral_lr  # noqa: F401
from .render_psf import local_dp_psf_render, local_psf_render, local_psf_render_fast  # noqa: F401
from .psfnet_arch import MLP, initialize_weights  # noqa: F401
from .optics import Lensgroup   # noqa: F401
from .psfnet import PSFNet      # noqa: F401

"""sdirt_b200 — B200-native engine for Sdirt's dual-pixel ray-tracing hot path.

`sdirt_b200.deeplens` mirrors the reference's Python API (Lensgroup / PSFNet / Ray / Aspheric / forward_integral /
local_psf_render_fast); `sdirt_b200._engine` is the ctypes binding of the C ABI in include/sdirt_engine.h;
`sdirt_b200.sharding` partitions PSF banks and render batches over the GPUs of one node."""
import os

from . import _engine  # noqa: F401

LENS_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lenses")


def lens_file(name):
    """Path of a bundled prescription ('rf50mm' or 'rf35mm')."""
    return os.path.join(LENS_DIR, name + ".json")

"""Import the UNMODIFIED reference (`/root/reference/deeplens`) in this container.

Test infrastructure only (used by `make_golden.py`); never imported by the product or by the
`-m gpu` tests (the GPU box has no /root/reference).  The reference imports plotting/metric modules
that are not installed and not on the hot path; they are replaced by empty stubs.
"""
import sys
import types

REF_ROOT = "/root/reference"


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def import_reference():
    if "deeplens" in sys.modules and getattr(sys.modules["deeplens"], "__file__", "").startswith(REF_ROOT):
        return sys.modules["deeplens"]
    for n in ("matplotlib", "matplotlib.pyplot", "lpips", "imageio", "skimage"):
        if n not in sys.modules:
            try:
                __import__(n)
            except Exception:
                _stub(n)
    if "skimage.metrics" not in sys.modules:
        _stub("skimage.metrics", peak_signal_noise_ratio=None, structural_similarity=None)
    mp = sys.modules["matplotlib"]
    if not hasattr(mp, "pyplot"):
        mp.pyplot = sys.modules["matplotlib.pyplot"]
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import deeplens  # noqa: E402
    return deeplens

"""Run the UNMODIFIED reference (LinYark/Sdirt `deeplens`, staged byte for byte under baseline/_ref/ by
tools/stage_reference.py) through its own public API, for timing only: `bench.py --impl reference`, bench's `cpu_baseline`
leg and its eager-GPU baseline.  Nothing of the product is on this path and the product never imports this file.

The reference imports plotting / metric modules at module top that are not installed and not on the hot path
(matplotlib, lpips, imageio, skimage): they are replaced by empty stub modules."""
import os
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available():
    return os.path.isfile(os.path.join(REF, "deeplens", "optics.py"))


class _Stub(types.ModuleType):
    """An absent plotting / metric module: any attribute is a callable that must not be reached on the timed path."""

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)

        def _absent(*a, **k):
            raise RuntimeError(f"{self.__name__}.{item} is a stub: this module is not installed and not on the hot path")
        return _absent


def _stub(name):
    m = _Stub(name)
    m.__path__ = []                     # lets `import a.b` resolve sub-modules registered the same way
    sys.modules[name] = m
    parent, _, leaf = name.rpartition(".")
    if parent in sys.modules:
        setattr(sys.modules[parent], leaf, m)
    return m


def import_reference():
    """The reference's `deeplens` package from baseline/_ref (never the product's mirror of the same name)."""
    mod = sys.modules.get("deeplens")
    if mod is not None and os.path.abspath(getattr(mod, "__file__", "")).startswith(REF):
        return mod
    for n in ("matplotlib", "matplotlib.pyplot", "lpips", "imageio", "skimage", "skimage.io", "skimage.filters", "skimage.morphology",
              "skimage.metrics"):
        if n not in sys.modules:
            try:
                __import__(n)
            except Exception:
                _stub(n)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import deeplens
    assert os.path.abspath(deeplens.__file__).startswith(REF), deeplens.__file__
    return deeplens


def make_basenet(device):
    """The reference's DfDP network (dfdp/basenet.py:9-21), random weights: the consumer of the generated focal stacks."""
    import_reference()
    from dfdp.basenet import Basenet
    return Basenet("dfdp").to(device)


def lens_json(name):
    return os.path.join(REF, "lenses", name, "lens_web.json")


def make_psfnet(name, sensor_res, ks, device):
    dl = import_reference()
    return dl.PSFNet(filename=lens_json(name), sensor_res=tuple(sensor_res), kernel_size=ks, device=device)


def time_psf(lens, points_norm, ks, spp, repeats=1, seed=0, sync=None):
    """Seconds per call of the reference's Lensgroup.psf_diff (optics.py:934-996) on `points_norm` [N,3] (normalised x, y and
    depth in mm), `spp` rays per point.  One call traces spp x N rays and splats them into N left PSFs (the reference's R needs
    a second, mirrored call, psfnet.py:540-544)."""
    import torch
    times = []
    for r in range(repeats):
        torch.manual_seed(seed + r)
        if sync:
            sync()
        t0 = time.perf_counter()
        psf = lens.psf_diff(points=points_norm.clone(), ks=ks, spp=spp)
        if sync:
            sync()
        times.append(time.perf_counter() - t0)
    return times, psf


def time_render(lens, img, depth, foc, repeats=1, sync=None):
    """Seconds per call of the reference's PSFNet.render (psfnet.py:645-714)."""
    import torch
    times = []
    out = None
    for r in range(repeats):
        if sync:
            sync()
        t0 = time.perf_counter()
        with torch.no_grad():
            out = lens.render(img, depth, foc)
        if sync:
            sync()
        times.append(time.perf_counter() - t0)
    return times, out

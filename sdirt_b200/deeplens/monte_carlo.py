"""forward_integral of the reference (deeplens/monte_carlo.py:9-68) on the CUDA engine."""
import torch

from .. import _engine as E

DP_DEFAULT = (0.78, 1.44, 0.3, 0.5, "l")          # monte_carlo.py:157-162


def forward_integral(ray, ps, ks, pointc_ref=None, interpolate=False, param_list=None):
    """Monte-Carlo PSF integral of traced rays: ray.o / ray.d / ray.ra are [spp, N, ...] on the sensor plane.

    Returns the [N, ks, ks] grid the reference returns: the LEFT sub-pixel PSF, or the RIGHT one when
    `param_list = (h, f, w, radius, direct)` has direct != 'l' (monte_carlo.py:64, 237-240).  Raw sums, un-normalised.
    `pointc_ref=None` centres every PSF on the ra-weighted centroid of its hits (monte_carlo.py:28-31)."""
    L, R = forward_integral_lr(ray, ps, ks, pointc_ref, param_list)
    direct = "l" if param_list is None else param_list[4]
    return L if direct == "l" else R


def forward_integral_lr(ray, ps, ks, pointc_ref=None, param_list=None):
    """Both sub-pixel grids of the same trace (the engine always fills L and R)."""
    if ray.o.dim() == 2:
        raise NotImplementedError("forward_integral expects [spp, N, 3] rays, as Lensgroup.psf_diff produces")
    dp = None if param_list is None else tuple(float(v) for v in param_list[:4])
    centre = None
    if pointc_ref is not None:
        centre = pointc_ref.to(ray.o.device, torch.float32).contiguous()
    return E.splat_rays(ray.o, ray.d, ray.ra, centre, ks, ps, dp)

"""CPU oracle for the Sdirt dual-pixel ray-tracing hot path.  TEST INFRASTRUCTURE ONLY.

This is a numpy restatement of the reference's algorithm, written from the arithmetic contract in
SURVEY.md Appendix A and checked against golden vectors produced by running the unmodified reference
(`tests/golden/make_golden.py`, fixtures in `tests/golden/*.npz`).  Parity status: PINNED by those
reference-generated fixtures (the reference ships no tests of its own).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this file, and only as the checker / reported CPU baseline.  The product (`sdirt_b200/`) never
imports it.

Layout differs from the reference on purpose: rays are a structure of arrays (`RayBundle` with
separate x/y/z planes) and every stage is a free function; all arithmetic is float32 in the same
operation order as the reference so that IEEE add/mul/div/sqrt reproduce its roundings:

  * vector norms use fused multiply-add accumulation, as torch's CPU `norm` kernel does (measured);
  * `scalar / tensor` is `reciprocal(tensor) * scalar` (torch `Tensor.__rtruediv__`);
  * python scalars are rounded to float32 before meeting a float32 array (torch and numpy>=2 agree).

Reference sites restated (paths relative to /root/reference):
  deeplens/basics.py:299-380      Material (Cauchy "n/V" glasses)          -> cauchy_ab, refractive_index
  deeplens/optics.py:2173-2198    read_lens_json                            -> load_lens
  deeplens/optics.py:193-201      find_aperture                             -> Lens.aper_idx
  deeplens/optics.py:476-494      sample_from_points                        -> pupil_points, rays_from_points
  deeplens/surfaces.py:391-520    Aspheric.ray_reaction                     -> surface_step
  deeplens/surfaces.py:523-586    _newtons_method                           -> newton_intersect
  deeplens/surfaces.py:589-679    _normal / _refract                        -> surface_normal, refract
  deeplens/surfaces.py:724-743    _valid / _valid_loose                     -> _strict_mask, _loose_mask
  deeplens/surfaces.py:787-830    _g / _dgd                                 -> sag, dsag_dr2
  deeplens/optics.py:601-689      trace / _forward_tracing / _backward      -> trace
  deeplens/basics.py:256-264      Ray.propagate_to                          -> propagate_to_z
  deeplens/optics.py:889-904      psf_center (chief ray)                    -> chief_ray_centre
  deeplens/monte_carlo.py:9-68    forward_integral                          -> splat_points
  deeplens/monte_carlo.py:135-372 assign_points_to_pixels_small_r / big_r   -> dp_weights_small_r / _big_r, _bilinear_splat
  deeplens/optics.py:934-996      psf_diff                                  -> psf_bank
  deeplens/optics.py:1203-1233, 1335-1396, 1170-1196  calc_fov / entrance_pupil / refocus -> calc_hfov, pupil_paraxial, refocus
  deeplens/render_psf.py:120-155  local_psf_render_fast                     -> render_local_psf
  deeplens/psfnet.py:589-620      degamma / gamma                           -> degamma, gamma
  deeplens/psfnet.py:317-336, 681-713; psfnet_arch.py:32-56  pred / render under CUDA autocast -> mlp_*_half, psf_pack_half, psfnet_render_half
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

F32 = np.float32
EPSILON = 1e-9                 # basics.py:35
NEWTON_MAXITER = 10            # surfaces.py:25
NEWTON_TOL_TIGHT = 10e-6       # surfaces.py:26
NEWTON_TOL_LOOSE = 50e-6       # surfaces.py:27
NEWTON_STEP_BOUND = 5          # surfaces.py:28
MAXT = 1e5                     # basics.py:33
DEFAULT_WAVE = 0.589           # basics.py:20
GEO_SPP = 2048                 # basics.py:29
DP_DEFAULT = (0.78, 1.44, 0.3, 0.5, "l")   # (h, f, w, r, direct)  monte_carlo.py:157-162


# ------------------------------------------------------------------------------------------------
# float32 helpers
# ------------------------------------------------------------------------------------------------
class precision:
    """`with precision(np.float64):` re-runs the same restatement in float64 (the arbiter used to show which of
    two float32 results is closer to the exact geometry).  Not thread-safe; tests only."""

    def __init__(self, dtype):
        self.dtype = dtype

    def __enter__(self):
        global F32
        self.prev, F32 = F32, self.dtype

    def __exit__(self, *exc):
        global F32
        F32 = self.prev


def _f(x):
    return np.asarray(x, dtype=F32)


def _fma(a, b, c):
    """float32 fused multiply-add emulated through float64 (exact product, one final rounding)."""
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(F32)



def _norm3(x, y, z):
    """sqrt(x^2+y^2+z^2) the way torch's CPU norm kernel accumulates it (FMA chain, measured)."""
    acc = _fma(x, x, np.zeros_like(x))
    acc = _fma(y, y, acc)
    acc = _fma(z, z, acc)
    return np.sqrt(acc)


def _normalize3(x, y, z):
    """F.normalize(p=2, eps=1e-12) (basics.py:245, surfaces.py:628)."""
    n = np.maximum(_norm3(x, y, z), F32(1e-12))
    return x / n, y / n, z / n


def _recip(x):
    return F32(1.0) / x


# ------------------------------------------------------------------------------------------------
# Lens prescription
# ------------------------------------------------------------------------------------------------
def cauchy_ab(name: str) -> Tuple[float, float]:
    """(A, B) of Cauchy's n = A + B / lambda_nm^2 for an "n/V" glass string or air (basics.py:354-376)."""
    nm = name.lower()
    if nm in ("air", "vacuum", "occluder"):
        n, v = 1.0, float("inf")                          # MATERIAL_TABLE, basics.py:43-46
    else:
        a, b = nm.split("/")
        n, v = float(a), float(b)

    def ivs(a):
        return 1.0 / a ** 2

    lam = (656.3, 589.3, 486.1)
    B = (n - 1) / v / (ivs(lam[2]) - ivs(lam[0]))
    A = n - B * ivs(lam[1])
    return A, B


def refractive_index(ab: Tuple[float, float], wvln_um: float) -> float:
    """float64 index at a wavelength in micrometres (basics.py:316-340, 'naive' branch)."""
    wv = wvln_um if wvln_um < 10 else wvln_um * 1e-3
    return ab[0] + ab[1] / (wv * 1e3) ** 2


@dataclass
class SurfaceSpec:
    r: float                     # semi-diameter (python float, surfaces.py:16)
    d: np.float32                # vertex z
    c: np.float32                # curvature
    k: np.float32                # conic
    ai: Optional[np.ndarray]     # even-asphere coefficients a2, a4, ... (float32) or None
    mat1: Tuple[float, float]
    mat2: Tuple[float, float]
    square: bool = False

    @property
    def is_flat(self):
        return float(self.c) == 0.0

    @property
    def is_sphere(self):
        return (not self.is_flat) and self.ai is None and float(self.k) == 0.0


@dataclass
class Lens:
    surfaces: List[SurfaceSpec]
    d_sensor: float
    r_last: float
    sensor_size: Tuple[float, float] = (24.0, 36.0)
    sensor_res: Tuple[int, int] = (512, 768)
    aper_idx: Optional[int] = None
    hfov: float = 0.0
    pixel_size: float = field(init=False, default=0.0)

    def __post_init__(self):
        # optics.py:154-178 (sensor size is forced to 24x36 mm; r_last recomputed)
        H, W = self.sensor_res
        self.r_last = float(np.sqrt(self.sensor_size[0] ** 2 + self.sensor_size[1] ** 2) / 2)
        assert self.sensor_size[0] / self.sensor_size[1] == H / W, "Pixel is not square."
        self.pixel_size = self.sensor_size[0] / H
        self.aper_idx = None
        for i, s in enumerate(self.surfaces[:-1]):          # optics.py:193-201
            if s.mat1[0] < 1.0003 and s.mat2[0] < 1.0003:
                self.aper_idx = i
                break


def load_lens(path: str, sensor_res=(512, 768), d_sensor: Optional[float] = None) -> Lens:
    with open(path) as fh:
        data = json.load(fh)
    surfs = []
    for sd in data["surfaces"]:
        ai = None
        k = 0.0
        if sd["type"] == "Aspheric":
            ai = np.asarray(sd["ai"], dtype=F32)
            k = sd["k"]
        elif sd["type"] not in ("Stop", "Spheric"):
            raise ValueError("Surface type not implemented.")
        surfs.append(SurfaceSpec(r=float(sd["r"]), d=F32(sd["d"]), c=F32(sd["c"]), k=F32(k), ai=ai,
                                 mat1=cauchy_ab(sd["mat1"]), mat2=cauchy_ab(sd["mat2"])))
    return Lens(surfaces=surfs, d_sensor=float(data["d_sensor"] if d_sensor is None else d_sensor),
                r_last=float(data["r_last"]), sensor_res=tuple(sensor_res))


# ------------------------------------------------------------------------------------------------
# Rays (structure of arrays)
# ------------------------------------------------------------------------------------------------
@dataclass
class RayBundle:
    ox: np.ndarray
    oy: np.ndarray
    oz: np.ndarray
    dx: np.ndarray
    dy: np.ndarray
    dz: np.ndarray
    ra: np.ndarray
    wvln: float = DEFAULT_WAVE

    @staticmethod
    def from_od(o, d, wvln=DEFAULT_WAVE, normalize=True):
        o = _f(o)
        d = _f(d)
        dx, dy, dz = d[..., 0], d[..., 1], d[..., 2]
        if normalize:
            dx, dy, dz = _normalize3(dx, dy, dz)
        return RayBundle(o[..., 0].copy(), o[..., 1].copy(), o[..., 2].copy(), dx.copy(), dy.copy(), dz.copy(),
                         np.ones(o.shape[:-1], F32), wvln)

    def copy(self):
        return RayBundle(*(a.copy() for a in (self.ox, self.oy, self.oz, self.dx, self.dy, self.dz, self.ra)),
                         wvln=self.wvln)

    def o(self):
        return np.stack((self.ox, self.oy, self.oz), -1)

    def d(self):
        return np.stack((self.dx, self.dy, self.dz), -1)


def pupil_points(u_theta, u_rho, pupil_r):
    """Entrance-pupil disk points from two uniform [0,1) streams (optics.py:483-487)."""
    theta = _f(u_theta) * (2 * np.pi)
    rho = np.sqrt(_f(u_rho) * (float(pupil_r) ** 2))     # numpy cos/sin differ from torch's by <=1 ulp
    return rho * np.cos(theta), rho * np.sin(theta)


def rays_from_points(points, px, py, pupil_z, wvln=DEFAULT_WAVE) -> RayBundle:
    """[spp, N] bundle from N object points towards spp shared pupil points (optics.py:476-494)."""
    points = _f(points)
    spp, n = px.shape[0], points.shape[0]
    o = np.broadcast_to(points[None], (spp, n, 3))
    o2 = np.stack((_f(px), _f(py), np.full(spp, float(pupil_z), F32)), -1)
    d = o2[:, None, :] - o
    return RayBundle.from_od(o, d, wvln)


# ------------------------------------------------------------------------------------------------
# Surface functions
# ------------------------------------------------------------------------------------------------
def _poly_terms(s: SurfaceSpec):
    return [] if s.ai is None else [F32(a) for a in s.ai]


def _ipow(x, p):
    """x**p for small integer p.  torch evaluates p<=3 as float32 products and larger p with a (nearly)
    correctly rounded pow; the latter is restated as the float64 product rounded once, which equals torch's
    result for 98 % of inputs (measured) and is exactly reproducible on the GPU."""
    if p == 1:
        return x
    if p == 2:
        return x * x
    if p == 3:
        return (x * x) * x
    xd = np.asarray(x, np.float64)
    q = xd * xd * xd
    for _ in range(p - 3):
        q = q * xd
    return q.astype(F32)


def sag(s: SurfaceSpec, r2):
    """Conic + even-asphere sag as a function of rho^2 (surfaces.py:787-808)."""
    c2 = s.c * s.c
    total = (r2 * s.c) / (F32(1) + np.sqrt(F32(1) - ((F32(1) + s.k) * r2) * c2))
    ai = _poly_terms(s)
    n = len(ai)
    if n in (0,):
        return total
    if n in (7, 8):
        h = ai[-1] * r2
        for a in ai[-2::-1]:
            h = (a + h) * r2
        return total + h
    for i, a in enumerate(ai, start=1):
        total = total + a * _ipow(r2, i)
    return total


def dsag_dr2(s: SurfaceSpec, r2):
    """d sag / d rho^2 (surfaces.py:811-830)."""
    c2 = s.c * s.c
    onek = F32(1) + s.k
    sf = np.sqrt(F32(1) - (onek * r2) * c2)
    out = (((F32(1) + sf) + (((onek * r2) * c2) / F32(2)) / sf) * s.c) / ((F32(1) + sf) * (F32(1) + sf))
    ai = _poly_terms(s)
    n = len(ai)
    if n == 0:
        return out
    if n == 8:
        h = (F32(8) * ai[7]) * r2
        for i in range(6, 0, -1):
            h = ((F32(i + 1) * ai[i]) + h) * r2
        return (out + ai[0]) + h
    out = out + ai[0]
    for i in range(2, n + 1):
        out = out + (F32(i) * ai[i - 1]) * _ipow(r2, i - 1)
    return out


def _loose_mask(s: SurfaceSpec, x, y):
    r2 = x * x + y * y
    if float(s.k) > -1:
        bound = (_recip(s.c * s.c) * F32(1 - EPSILON)) / (F32(1) + s.k)
        return r2 < bound
    return r2 > 0


def _strict_mask(s: SurfaceSpec, x, y):
    r2 = x * x + y * y
    m = r2 < F32(s.r ** 2)
    if float(s.k) > -1:
        bound = (_recip(s.c * s.c) * F32(1 - EPSILON)) / (F32(1) + s.k)
        m = m & (r2 < bound)
    return m


def _newton_eval(s, ray, t, strict):
    nx = ray.ox + ray.dx * t
    ny = ray.oy + ray.dy * t
    nz = ray.oz + ray.dz * t
    m = (_strict_mask(s, nx, ny) if strict else _loose_mask(s, nx, ny)) & (ray.ra > 0)
    mf = m.astype(F32)
    x = nx * mf
    y = ny * mf
    r2 = x * x + y * y
    ft = (sag(s, r2) + s.d) - nz
    dr2dt = F32(2) * ((ray.dx * ray.dx + ray.dy * ray.dy) * t + (ray.dx * ray.ox + ray.dy * ray.oy))
    dfdt = dsag_dr2(s, r2) * dr2dt - ray.dz
    step = np.clip(ft / (dfdt + F32(EPSILON)), F32(-NEWTON_STEP_BOUND), F32(NEWTON_STEP_BOUND))
    return ft, t - step


def newton_intersect(s: SurfaceSpec, ray: RayBundle, max_iters=None):
    """Newton ray/sag intersection with the reference's GLOBAL loop (surfaces.py:523-586).

    Returns (valid_asphere_rule, t, iterations_run).  `max_iters`: None = the reference's bundle-global
    any()-driven count; an int = replay exactly that many loop evaluations; "per_ray" = every ray stops as
    soon as its own residual is within the loose tolerance (the engine's fast path).
    """
    t0 = (s.d - ray.oz) / ray.dz
    t = t0
    ft = np.full_like(ray.oz, MAXT)
    it = 0
    with np.errstate(all="ignore"):
        while True:
            active = np.abs(ft) > F32(NEWTON_TOL_LOOSE)
            if max_iters is None or max_iters == "per_ray":
                if not (active.any() and it < NEWTON_MAXITER):
                    break
            elif it >= max_iters:
                break
            it += 1
            ft_new, t_new = _newton_eval(s, ray, t, strict=False)
            if max_iters == "per_ray":
                ft, t = np.where(active, ft_new, ft), np.where(active, t_new, t)
            else:
                ft, t = ft_new, t_new
            if np.isnan(ft).any():
                raise FloatingPointError("nan in Newton residual")
        t = t0 + (t - t0)
        ft_last, t = _newton_eval(s, ray, t, strict=True)
        nx = ray.ox + ray.dx * t
        ny = ray.oy + ray.dy * t
        valid = _strict_mask(s, nx, ny) & (np.abs(ft_last) < F32(NEWTON_TOL_TIGHT)) & (ray.ra > 0) & (t > 0)
    return valid, t, it


def surface_normal(s: SurfaceSpec, ray: RayBundle):
    """Unit gradient of the implicit surface at ray.o (surfaces.py:589-630)."""
    x, y, z = ray.ox, ray.oy, ray.oz
    if s.is_flat:
        gx, gy, gz = np.zeros_like(x), np.zeros_like(y), np.full_like(z, -1)
    elif s.is_sphere:
        R = _recip(s.c)
        if float(s.c) > 0:
            gx, gy, gz = F32(2) * x, F32(2) * y, F32(2) * z - F32(2) * (s.d + R)
        else:
            gx, gy, gz = F32(-2) * x, F32(-2) * y, F32(-2) * z + F32(2) * (s.d + R)
    else:
        mf = (ray.ra > 0).astype(F32)
        xm, ym = x * mf, y * mf
        g = dsag_dr2(s, xm * xm + ym * ym)
        gx, gy, gz = (g * F32(2)) * xm, (g * F32(2)) * ym, np.full_like(x, -1)
    return _normalize3(gx, gy, gz)


def refract(s: SurfaceSpec, ray: RayBundle, eta: float, forward: bool):
    """Vector Snell refraction with the TIR and grazing cuts (surfaces.py:633-679).  In place."""
    nx, ny, nz = surface_normal(s, ray)
    if forward:
        nx, ny, nz = -nx, -ny, -nz
    cosi = (ray.dx * nx + ray.dy * ny) + ray.dz * nz
    eta2 = F32(eta ** 2)
    etaf = F32(eta)
    one_m = F32(1) - cosi * cosi
    valid = (cosi * cosi > F32(0.1)) & (eta2 * one_m < F32(1)) & (ray.ra > 0)
    with np.errstate(invalid="ignore"):
        sr = np.sqrt(F32(1) - (eta2 * one_m) * valid.astype(F32))
    ndx = sr * nx + etaf * (ray.dx - cosi * nx)
    ndy = sr * ny + etaf * (ray.dy - cosi * ny)
    ndz = sr * nz + etaf * (ray.dz - cosi * nz)
    ray.dx = np.where(valid, ndx, ray.dx)
    ray.dy = np.where(valid, ndy, ray.dy)
    ray.dz = np.where(valid, ndz, ray.dz)
    ray.ra = ray.ra * valid.astype(F32)


def surface_step(s: SurfaceSpec, ray: RayBundle, max_iters=None) -> int:
    """One surface: intersect, mask, refract (surfaces.py:391-520).  In place; returns Newton iterations."""
    forward = bool((ray.dz * ray.ra).sum() > 0)          # global direction test, surfaces.py:399
    n1, n2 = refractive_index(s.mat1, ray.wvln), refractive_index(s.mat2, ray.wvln)
    eta = n1 / n2 if forward else n2 / n1
    iters = 0
    if s.is_flat:
        t = (s.d - ray.oz) / ray.dz
        nx, ny, nz = ray.ox + t * ray.dx, ray.oy + t * ray.dy, ray.oz + t * ray.dz
        if s.square:
            valid = (np.abs(nx) <= F32(s.r)) & (np.abs(ny) <= F32(s.r)) & (ray.ra > 0)
        else:
            valid = (np.sqrt(nx * nx + ny * ny) <= F32(s.r)) & (ray.ra > 0)
    else:
        valid, t, iters = newton_intersect(s, ray, max_iters)
        nx, ny, nz = ray.ox + t * ray.dx, ray.oy + t * ray.dy, ray.oz + t * ray.dz
        if s.is_sphere:                                   # validity overridden, surfaces.py:464
            valid = (nx * nx + ny * ny <= F32(s.r ** 2)) & (t >= 0) & (ray.ra > 0)
    ray.ox = np.where(valid, nx, ray.ox)
    ray.oy = np.where(valid, ny, ray.oy)
    ray.oz = np.where(valid, nz, ray.oz)
    ray.ra = ray.ra * valid.astype(F32)
    if not (s.is_flat and eta == 1):
        refract(s, ray, eta, forward)
    return iters


def trace(lens: Lens, ray: RayBundle, lens_range: Optional[Sequence[int]] = None, record=None,
          newton_iters: Optional[Sequence[Optional[int]]] = None) -> List[int]:
    """Sequential trace, direction decided by the first ray's d_z (optics.py:601-689).  In place.

    `record`, if a list, receives a RayBundle copy after every surface.  `newton_iters`: None (global
    loop), "per_ray", or loop counts indexed by LENS surface index.  Returns the Newton loop counts per
    visited surface.
    """
    is_forward = bool(ray.dz.reshape(-1)[0] > 0)
    idx = list(range(len(lens.surfaces))) if lens_range is None else list(lens_range)
    if not is_forward:
        idx = idx[::-1]
    counts = []
    for j, i in enumerate(idx):
        mi = newton_iters if (newton_iters is None or isinstance(newton_iters, str)) else newton_iters[i]
        counts.append(surface_step(lens.surfaces[i], ray, mi))
        if record is not None:
            record.append(ray.copy())
    return counts


def propagate_to_z(ray: RayBundle, z: float):
    """basics.py:256-264."""
    t = (F32(z) - ray.oz) / ray.dz
    ray.ox = ray.ox + ray.dx * t
    ray.oy = ray.oy + ray.dy * t
    ray.oz = ray.oz + ray.dz * t


def trace_to_sensor(lens: Lens, ray: RayBundle, record=None, newton_iters=None):
    counts = trace(lens, ray, record=record, newton_iters=newton_iters)
    propagate_to_z(ray, lens.d_sensor)
    return counts


# ------------------------------------------------------------------------------------------------
# Setup geometry (tiny traces)
# ------------------------------------------------------------------------------------------------
def _pairwise_line_intersections(o2, d2):
    """Mean-of-both-solutions intersection of all 2-D line pairs (optics.py:1471-1514)."""
    n = o2.shape[0]
    ii, jj = np.triu_indices(n, 1)
    pts = []
    for i, j in zip(ii, jj):
        A = np.stack((d2[i], -d2[j]), -1).astype(F32)
        b = (o2[j] - o2[i]).astype(F32)
        sol = np.linalg.lstsq(A.astype(np.float64), b.astype(np.float64), rcond=None)[0].astype(F32)
        pi = o2[i] + sol[0] * d2[i]
        pj = o2[j] + sol[1] * d2[j]
        pts.append((pi + pj) / F32(2))
    return np.asarray(pts, F32)


def pupil_paraxial(lens: Lens, entrance=True, shrink=False) -> Tuple[float, float]:
    """Entrance / exit pupil (z, radius) from 16 paraxial rays off the stop edge (optics.py:1335-1396)."""
    if lens.aper_idx is None:
        s = lens.surfaces[0] if entrance else lens.surfaces[-1]
        return float(s.d), s.r
    ap = lens.surfaces[lens.aper_idx]
    delta_r = 1e-3
    o = np.tile(np.asarray([[delta_r, 0, float(ap.d)]], F32), (16, 1))
    phi = _torch_linspace(-0.1, 0.1, 16) / F32(180.0) * F32(np.pi)
    sgn = -1.0 if entrance else 1.0
    d = np.stack((np.sin(phi), np.zeros_like(phi), F32(sgn) * np.cos(phi)), -1)
    ray = RayBundle.from_od(o, d)
    rng = range(0, lens.aper_idx) if entrance else range(lens.aper_idx + 1, len(lens.surfaces))
    trace(lens, ray, rng)
    keep = ray.ra != 0
    o2 = np.stack((ray.ox[keep], ray.oz[keep]), -1)
    d2 = np.stack((ray.dx[keep], ray.dz[keep]), -1)
    pts = _pairwise_line_intersections(o2, d2)
    if len(pts) == 0:
        return float(lens.surfaces[0].d), lens.surfaces[0].r
    r = abs(float(pts[:, 0].mean(dtype=F32)) / delta_r * ap.r)
    z = float(pts[:, 1].mean(dtype=F32))
    return z, (r * 0.25 if shrink else r)


def _torch_linspace(a, b, n):
    """torch.linspace float32 semantics: symmetric fill from both ends with a float32 step."""
    step = (F32(b) - F32(a)) / F32(n - 1)
    i = np.arange(n)
    lo = F32(a) + step * i.astype(F32)
    hi = F32(b) - step * (n - 1 - i).astype(F32)
    return np.where(i < n // 2, lo, hi).astype(F32)


def calc_hfov(lens: Lens) -> float:
    """Half diagonal field of view from 100 backward rays off the sensor corner (optics.py:1203-1233)."""
    M = 100
    pz, pr = pupil_paraxial(lens, entrance=False, shrink=True)
    o1 = np.tile(np.asarray([[lens.r_last, 0, lens.d_sensor]], F32), (M, 1))
    x2 = _torch_linspace(-pr, pr, M)
    o2 = np.stack((x2, np.zeros_like(x2), np.full_like(x2, pz)), -1)
    ray = RayBundle.from_od(o1, o2 - o1)
    trace(lens, ray)
    tan_fov = ray.dx / ray.dz
    return float(np.arctan((tan_fov * ray.ra).sum(dtype=F32) / ray.ra.sum(dtype=F32)))


def refocus(lens: Lens, depth: float, u_theta, u_rho) -> float:
    """Least-squares best-focus sensor position for an on-axis point (optics.py:1170-1196)."""
    s0 = lens.surfaces[0]
    px, py = pupil_points(u_theta, u_rho, s0.r)
    o = np.stack((px, py, np.full_like(px, float(s0.d))), -1)
    d = o - np.asarray([0, 0, depth], F32)
    ray = RayBundle.from_od(o, d)
    trace(lens, ray)
    t = (ray.dx * ray.ox + ray.dy * ray.oy) / (ray.dx * ray.dx + ray.dy * ray.dy)
    t = t * ray.ra
    fd = ray.oz - ray.dz * t
    fd = fd[ray.ra > 0]
    fd = fd[~np.isnan(fd) & (fd > 0)]
    return float(np.mean(fd))


def object_points(lens: Lens, points_norm):
    """Normalised (x, y, depth) -> object-space mm (optics.py:956-960, 1302-1306)."""
    p = _f(points_norm).copy()
    scale = ((-p[:, 2]) * F32(np.tan(lens.hfov))) / F32(lens.r_last)
    out = p.copy()
    out[:, 0] = p[:, 0] * scale * F32(lens.sensor_size[1]) / F32(2)
    out[:, 1] = p[:, 1] * scale * F32(lens.sensor_size[0]) / F32(2)
    return out


# ------------------------------------------------------------------------------------------------
# PSF centre, DP weights, splat
# ------------------------------------------------------------------------------------------------
def chief_ray_centre(ray: RayBundle):
    """-(ra-weighted centroid of sensor hits), per point (optics.py:902-904).  ray is [spp, N]."""
    # float64 accumulation: the reference sums 2048 float32 terms in torch's (build-dependent) cascade
    # order, which carries ~1e-6 relative noise; the oracle takes the correctly rounded mean.
    den = (ray.ra.sum(0, dtype=np.float64) + EPSILON)
    cx = -((ray.ox * ray.ra).sum(0, dtype=np.float64) / den)
    cy = -((ray.oy * ray.ra).sum(0, dtype=np.float64) / den)
    return np.stack((cx, cy), -1).astype(F32)


def _seg(u):
    """A(u) = acos(u) - sin(2 acos(u)) / 2."""
    a = np.arccos(u)
    return a - F32(0.5) * np.sin(F32(2) * a)


def dp_weights_small_r(x_tan, params=DP_DEFAULT):
    """(d_l, d_r) sub-pixel areas for micro-lens radius <= 0.5 px (monte_carlo.py:157-206)."""
    h, f, w, r, _ = params
    assert r <= 0.5
    r = F32(r)
    kap_num, kap_den = F32(h), F32(f - h)
    fx = F32(f) * x_tan
    xr = np.clip(F32(w) - ((fx - F32(w)) * kap_num) / kap_den, -r, r)
    xm = np.clip(-((fx * kap_num) / kap_den), -r, r)
    xl = np.clip(F32(-w) - ((fx + F32(w)) * kap_num) / kap_den, -r, r)
    ar, am, al = _seg(xr / r), _seg(xm / r), _seg(xl / r)
    sr_ml = (r * r) * (am - ar)
    sl_ml = (r * r) * (al - am)
    hx = F32(h) * x_tan
    xr = np.clip(F32(w) - hx, F32(-0.5), F32(0.5))
    xm = np.clip(F32(0) - hx, F32(-0.5), F32(0.5))
    xl = np.clip(F32(-w) - hx, F32(-0.5), F32(0.5))
    ar, am, al = _seg(np.clip(xr, -r, r) / r), _seg(np.clip(xm, -r, r) / r), _seg(np.clip(xl, -r, r) / r)
    sr_mg = (xr - xm) * F32(1) - (r * r) * (am - ar)
    sl_mg = (xm - xl) * F32(1) - (r * r) * (al - am)
    return sl_ml + sl_mg, sr_ml + sr_mg


def dp_weights_big_r(x_tan, params):
    """(d_l, d_r) for micro-lens radius >= 0.5 px (monte_carlo.py:263-338)."""
    h, f, w, r, _ = params
    assert r >= 0.5
    r = F32(r)
    tr = np.arcsin(F32(0.5) / r)
    tl = F32(np.pi) - tr                                   # torch.pi - tensor -> float32

    def area_minus_overhang(xr, xm, xl):
        ur, um, ul = np.arccos(xr / r), np.arccos(xm / r), np.arccos(xl / r)

        def A(u):
            return u - F32(0.5) * np.sin(F32(2) * u)
        s_r = (r * r) * (A(um) - A(ur))
        s_l = (r * r) * (A(ul) - A(um))
        er, em, el = np.clip(ur, tr, tl), np.clip(um, tr, tl), np.clip(ul, tr, tl)
        xer, xem, xel = np.cos(er) * r, np.cos(em) * r, np.cos(el) * r
        s_r_ext = (r * r) * (A(em) - A(er)) - (xer - xem)
        s_l_ext = (r * r) * (A(el) - A(em)) - (xem - xel)
        return s_r - s_r_ext, s_l - s_l_ext

    kap_num, kap_den = F32(h), F32(f - h)
    fx = F32(f) * x_tan
    half = F32(0.5)
    xr = np.clip(F32(w) - ((fx - F32(w)) * kap_num) / kap_den, -half, half)
    xm = np.clip(-((fx * kap_num) / kap_den), -half, half)
    xl = np.clip(F32(-w) - ((fx + F32(w)) * kap_num) / kap_den, -half, half)
    sr_ml, sl_ml = area_minus_overhang(xr, xm, xl)
    hx = F32(h) * x_tan
    xr = np.clip(F32(w) - hx, -half, half)
    xm = np.clip(F32(0) - hx, -half, half)
    xl = np.clip(F32(-w) - hx, -half, half)
    sr_in, sl_in = area_minus_overhang(xr, xm, xl)
    sr_mg = (xr - xm) * F32(1) - sr_in
    sl_mg = (xm - xl) * F32(1) - sl_in
    return sl_ml + sl_mg, sr_ml + sr_mg


def dp_weights(x_tan, params=DP_DEFAULT):
    return dp_weights_small_r(x_tan, params) if params[3] <= 0.5 else dp_weights_big_r(x_tan, params)


def splat_indices(qx, qy, ks, ps):
    """Pixel rows/cols and bilinear weights of recentred sensor points (monte_carlo.py:209-222)."""
    lo, hi = (-ks / 2 + 0.5) * ps, (ks / 2 - 0.5) * ps
    row_f = ((qy - F32(hi)) / F32(lo - hi)) * F32(ks - 1)
    col_f = ((qx - F32(lo)) / F32(hi - lo)) * F32(ks - 1)
    r0, c0 = np.floor(row_f), np.floor(col_f)
    wb, wr = row_f - r0, col_f - c0
    r1 = np.floor(row_f + F32(1)).astype(np.int64)
    c1 = np.floor(col_f + F32(1)).astype(np.int64)
    r0, c0 = r0.astype(np.int64), c0.astype(np.int64)
    return r0, c0, r1, c1, wb, wr


def _bilinear_splat(qx, qy, wgt, ks, ps):
    """Four index_put_(accumulate=True) taps, in the reference's tap order (monte_carlo.py:224-228)."""
    r0, c0, r1, c1, wb, wr = splat_indices(qx, qy, ks, ps)
    grid = np.zeros((ks, ks), F32)
    one = F32(1)
    np.add.at(grid, (r0, c0), (((one - wb) * (one - wr)) * wgt[0]) * wgt[1])
    np.add.at(grid, (r0, c1), (((one - wb) * wr) * wgt[0]) * wgt[1])
    np.add.at(grid, (r1, c0), ((wb * (one - wr)) * wgt[0]) * wgt[1])
    np.add.at(grid, (r0 + 1, c0 + 1), ((wb * wr) * wgt[0]) * wgt[1])
    return grid


def crop_and_shift(ray: RayBundle, centre, ks, ps):
    """Flip, recentre and crop sensor hits (monte_carlo.py:24-38).  Returns qx, qy, weight [spp, N]."""
    hi = (ks / 2 - 0.5) * ps
    if centre is None:
        den = ray.ra.sum(0, dtype=np.float64) + EPSILON          # float64 mean, see chief_ray_centre
        centre = np.stack((((-ray.ox) * ray.ra).sum(0, dtype=np.float64) / den,
                           ((-ray.oy) * ray.ra).sum(0, dtype=np.float64) / den), -1)
    centre = _f(centre)
    qx = (-ray.ox) - centre[:, 0]
    qy = (-ray.oy) - centre[:, 1]
    lim = F32(hi - 0.01 * ps)
    w = (ray.ra * (np.abs(qx) < lim).astype(F32)) * (np.abs(qy) < lim).astype(F32)
    return qx * w, qy * w, w


def splat_points(ray: RayBundle, ps, ks, centre=None, params=None):
    """forward_integral restated; returns BOTH grids: (L [N,ks,ks], R [N,ks,ks]) (monte_carlo.py:9-68).

    With `params=None` the reference fills only L (and returns only L); here R is filled as well so a
    single call yields the same-sample pair (equal to the reference called with direct='l' and 'r').
    """
    qx, qy, w = crop_and_shift(ray, centre, ks, ps)
    prm = DP_DEFAULT if params is None else params
    n = ray.ox.shape[1]
    L = np.zeros((n, ks, ks), F32)
    R = np.zeros((n, ks, ks), F32)
    for i in range(n):
        x_tan = (-ray.dx[:, i]) / ray.dz[:, i]
        d_l, d_r = dp_weights(x_tan, prm)
        L[i] = _bilinear_splat(qx[:, i], qy[:, i], (w[:, i], d_l), ks, ps)
        R[i] = _bilinear_splat(qx[:, i], qy[:, i], (w[:, i], d_r), ks, ps)
    return L, R


def splat_points_f64(ray: RayBundle, ps, ks, centre, params=None):
    """The arbiter of a summation-order dispute: the SAME float32 addends forward_integral forms for every ray (tap
    weights and d_l / d_r exactly as splat_points computes them, monte_carlo.py:209-235), accumulated in float64 --
    the sum the reference's sequential float32 `index_put_(accumulate=True)` and the engine's run / tile / chunk
    summation both approximate.  Returns (L, R) as float64 [N, ks, ks]."""
    qx, qy, w = crop_and_shift(ray, centre, ks, ps)
    prm = DP_DEFAULT if params is None else params
    n = ray.ox.shape[1]
    out = np.zeros((2, n, ks, ks), np.float64)
    one = F32(1)
    for i in range(n):
        x_tan = (-ray.dx[:, i]) / ray.dz[:, i]
        r0, c0, r1, c1, wb, wr = splat_indices(qx[:, i], qy[:, i], ks, ps)
        for side, d in enumerate(dp_weights(x_tan, prm)):
            taps = ((r0, c0, (one - wb) * (one - wr)), (r0, c1, (one - wb) * wr), (r1, c0, wb * (one - wr)), (r0 + 1, c0 + 1, wb * wr))
            for rr, cc, tw in taps:
                add = ((tw * w[:, i]) * d).astype(np.float64)
                keep = (rr >= 0) & (rr < ks) & (cc >= 0) & (cc < ks)
                out[side, i] += np.bincount((rr * ks + cc)[keep], weights=add[keep], minlength=ks * ks).reshape(ks, ks)
    return out[0], out[1]


def max_normalise(psf):
    """optics.py:984-987."""
    m = psf.reshape(psf.shape[0], -1).max(-1)[:, None, None]
    return psf / (m + F32(1e-6))


def sum_normalise(psf):
    """psfnet.py:159-160 (eval-time normalisation; the parity metric)."""
    return psf / psf.sum((-1, -2), keepdims=True, dtype=F32)


def psf_bank(lens: Lens, points_obj, px, py, pupil_z, ks, wvln=DEFAULT_WAVE, centre=None,
             centre_samples=None, params=None, normalise=True, newton_iters=None):
    """psf_diff restated on explicit samples (optics.py:934-996).

    points_obj: [N,3] object-space mm.  (px, py): main pupil samples.  `centre` [N,2] or, if None,
    `centre_samples=(cx, cy)` pupil samples of the 0.25x chief-ray bundle.  Returns (L, R, centre).
    """
    ray = rays_from_points(points_obj, px, py, pupil_z, wvln)
    trace_to_sensor(lens, ray, newton_iters=newton_iters)
    if centre is None:
        cray = rays_from_points(points_obj, centre_samples[0], centre_samples[1], pupil_z, DEFAULT_WAVE)
        trace_to_sensor(lens, cray, newton_iters=newton_iters)
        centre = chief_ray_centre(cray)
    L, R = splat_points(ray, lens.pixel_size, ks, centre, params)
    if normalise:
        L, R = max_normalise(L), max_normalise(R)
    return L, R, centre


# ------------------------------------------------------------------------------------------------
# Render
# ------------------------------------------------------------------------------------------------
_TONE = (0.89129432, 0.27217316, -0.00246187, 5.94018909e-01, 1.20060450e+01, -5.24983855e-03)


def degamma(img):
    """psfnet.py:589-603."""
    a1, b1, c1, a2, b2, c2 = _TONE
    x = _f(img) * F32(255.0)
    l1 = _recip(_recip(F32(a1) * x + F32(b1)) + F32(c1))
    l2 = _recip(_recip(F32(a2) * x + F32(b2)) + F32(c2))
    ratio = np.minimum(x / F32(100), F32(1))
    return l2 * ratio + l1 * (F32(1) - ratio)


def gamma(lum):
    """psfnet.py:605-620."""
    a1, b1, c1, a2, b2, c2 = _TONE
    lum = _f(lum)
    x1 = (_recip(_recip(lum + F32(1e-9)) - F32(c1)) - F32(b1)) / F32(a1)
    x2 = (_recip(_recip(lum + F32(1e-9)) - F32(c2)) - F32(b2)) / F32(a2)
    ratio = ((x1 + x2) / F32(2)) / F32(100)
    ratio = np.where(ratio > 1, F32(1), ratio)
    return (x2 * ratio + x1 * (F32(1) - ratio)) / F32(255.0)


def render_local_psf(img, psf, ks):
    """Per-pixel L/R gather-convolution in fp16 (render_psf.py:120-155).

    img [B,C,H,W] float32, psf [B,H,W,2,ks,ks] float32 -> (rl, rr) [B,C,H,W] float32.  Products are
    rounded to fp16, summed in float32 and rounded to fp16 once (torch's CPU half sum).
    """
    img16 = _f(img).astype(np.float16)
    psf16 = _f(psf).astype(np.float16)
    b, c, h, w = img16.shape
    pad = (ks - 1) // 2
    ip = np.pad(img16, ((0, 0), (0, 0), (pad, pad), (pad, pad)), mode="edge")
    kf = psf16.reshape(b, h, w, 2, ks, ks)[..., ::-1, ::-1]
    out = np.zeros((2, b, c, h, w), np.float32)
    for u in range(ks):
        for v in range(ks):
            patch = ip[:, :, u:u + h, v:v + w]                              # [B,C,H,W]
            for s in range(2):
                k = kf[:, :, :, s, u, v][:, None]                           # [B,1,H,W]
                out[s] += (patch * k).astype(np.float16).astype(np.float32)
    out16 = out.astype(np.float16)
    return out16[0].astype(F32), out16[1].astype(F32)


# ------------------------------------------------------------------------------------------------
# PSFNet.pred / PSFNet.render under CUDA autocast (fp16 MLP), restated
# ------------------------------------------------------------------------------------------------
def _h(a):
    """Round to fp16 and come back to float32 (one fp16 rounding point)."""
    return np.asarray(a, dtype=F32).astype(np.float16).astype(F32)


def mlp_linear_relu_half(x16, w, b):
    """One Linear + ReLU of the PSF MLP under torch.autocast (psfnet_arch.py:40-56): activations, weights and bias in
    fp16, products accumulated in float32 (float64 here: the fp32 sum of fp16 products is order dependent in the last
    bit only), bias added before the single rounding to fp16, then ReLU."""
    acc = _h(x16).astype(np.float64) @ _h(w).astype(np.float64).T + _h(b).astype(np.float64)
    return np.maximum(_h(acc.astype(F32)), F32(0))


def mlp_input_rows(xs, ys, z, b0, nb, row0, n_rows):
    """MLP input rows for a window of pixels, pixel-major / side-minor: row 2p = (x, y, z), row 2p + 1 = (-x, y, z)
    (psfnet.py:683-694, 328).  xs [W], ys [H], z [B,H,W] float32 -> [2P, 3] float32."""
    zz = _f(z)[b0:b0 + nb, row0:row0 + n_rows]                              # [nb, n_rows, W]
    x = np.broadcast_to(_f(xs)[None, None, :], zz.shape)
    y = np.broadcast_to(_f(ys)[None, row0:row0 + n_rows, None], zz.shape)
    left = np.stack((x, y, zz), -1).reshape(-1, 3)
    right = left.copy()
    right[:, 0] = -right[:, 0]
    return np.stack((left, right), 1).reshape(-1, 3)


def mlp_forward_half(weights, inp):
    """The whole MLP (list of (W, b) pairs) on rows `inp` [M, 3] -> [M, out] with autocast's fp16 rounding points."""
    h = _h(inp)
    for w, b in weights:
        h = mlp_linear_relu_half(h, w, b)
    return h


def psf_pack_half(raw, ks):
    """PSFNet.pred's tail in torch's fp16 arithmetic (psfnet.py:326-333): raw [2P, >= ks*ks] (row 2p left, 2p + 1 right,
    unflipped) -> [P, 2, ks, ks]: flip the right kernels along the last axis, stack, divide by sum(-1).sum(-1) + 1e-9.
    sum(-1) accumulates in float32 and rounds to fp16 at each of the two reductions; the quotient is rounded once.
    All-zero kernels are returned as zeros (the reference's CUDA run yields NaN = 0/0 there, its CPU run 0)."""
    raw = _h(np.asarray(raw)[:, :ks * ks]).reshape(-1, 2, ks, ks).copy()
    raw[:, 1] = raw[:, 1, :, ::-1]
    rows = np.zeros(raw.shape[:3], F32)
    for v in range(ks):                                                     # sequential float32 accumulation
        rows = rows + raw[..., v]
    rows = _h(rows)
    tot = np.zeros(raw.shape[:2], F32)
    for u in range(ks):
        tot = tot + rows[..., u]
    den = _h(_h(tot) + F32(1e-9))[..., None, None]
    with np.errstate(divide="ignore", invalid="ignore"):
        out = np.where(den > 0, raw / den, F32(0))
    return _h(out)


def psfnet_render_half(weights, img, depth_z, ks, tone=3):
    """PSFNet.render(train=False) as the reference's CUDA run computes it (psfnet.py:681-713): coordinate grid, fp16 MLP
    for both sides, pack, degamma, fp16 gather-convolution, gamma, clip.  img [B,3,H,W], depth_z = depth2z(depth) [B,H,W]."""
    b, c, h, w = img.shape
    xs, ys = _torch_linspace(-1, 1, w), _torch_linspace(1, -1, h)
    rows = mlp_input_rows(xs, ys, depth_z, 0, b, 0, h)
    psf = psf_pack_half(mlp_forward_half(weights, rows), ks).reshape(b, h, w, 2, ks, ks)
    rl, rr = render_local_psf(degamma(img) if tone & 1 else img, psf, ks)
    out = np.concatenate((rl, rr), 1)
    return np.clip(gamma(out), 0, 1) if tone & 2 else out


def gamma_noise_clip(x, randn, noise_range, weight):
    """Tail of PSFNet.render(train=True) (psfnet.py:605-620, 629-642, 708-713): x [N,2C,H,W] linear image -> clip(gamma(x) +
    (randn * noise_range) * ramp, 0, 1), ramp = weight[n, col] on the left channels and mirrored on the right ones."""
    x, randn = _f(x), _f(randn)
    n, c2, h, w = x.shape
    wl = _f(weight)[:, None, None, :]
    ramp = np.concatenate((np.broadcast_to(wl, (n, c2 // 2, h, w)), np.broadcast_to(wl[..., ::-1], (n, c2 // 2, h, w))), 1)
    noise = (randn * _f(noise_range)[:, None, None, None]) * ramp
    return np.clip(gamma(x) + noise, F32(0), F32(1))

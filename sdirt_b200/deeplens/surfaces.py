"""Host-side mirror of the reference's `Aspheric` surface (deeplens/surfaces.py:281-830) for the hot path.

The object keeps the reference's attribute names (`r`, `d`, `c`, `k`, `ai`, `ai_degree`, `mat1`, `mat2`) so that
code written against the reference keeps working; `ray_reaction` hands the ray to the CUDA engine."""
import numpy as np
import torch

from .. import _engine as E
from .basics import DEVICE, DeepObj, Material


class Surface(DeepObj):
    def __init__(self, r, d, mat1, mat2, is_square=False, device=DEVICE):
        self.d = d.float() if torch.is_tensor(d) else torch.tensor([d]).float()
        self.r = float(r)
        self.is_square = is_square
        self.mat1 = Material(mat1)
        self.mat2 = Material(mat2)
        self.NEWTONS_MAXITER = 10
        self.NEWTONS_TOLERANCE_TIGHT = 10e-6
        self.NEWTONS_TOLERANCE_LOOSE = 50e-6
        self.NEWTONS_STEP_BOUND = 5
        self.device = device

    def surface_sample(self, N=1000):
        """Uniform points on the surface's aperture disc, CPU RNG order of the reference (surfaces.py:188-199)."""
        theta = torch.rand(N) * 2 * np.pi
        r = torch.sqrt(torch.rand(N) * self.r ** 2)
        x2, y2 = r * torch.cos(theta), r * torch.sin(theta)
        z2 = torch.full_like(x2, self.d.item())
        return torch.stack((x2, y2, z2), 1).to(self.device)


class Aspheric(Surface):
    """Plane / stop (c == 0), sphere (ai is None and k == 0) or conic + even asphere."""

    def __init__(self, r, d, c=0., k=0., ai=None, mat1=None, mat2=None, is_square=False, device=DEVICE, diff=False,
                 square=False):
        Surface.__init__(self, r, d, mat1, mat2, is_square, device)
        self.c = torch.Tensor([c])
        self.k = torch.Tensor([k])
        if ai is not None:
            self.ai = torch.Tensor(np.array(ai))
            self.ai_degree = len(ai)
            if self.ai_degree > E.MAX_AI:
                raise ValueError(f"the engine supports at most {E.MAX_AI} even-asphere coefficients")
            for i, a in enumerate(ai):
                setattr(self, f"ai{2 * i + 2}", torch.Tensor([a]))
        else:
            self.ai = None
            self.ai_degree = 0
        self.is_square = square
        self._handle = None
        self._handle_key = None

    # ---- engine plumbing -------------------------------------------------------------------------
    def kind(self):
        if float(self.c) == 0.0:
            return E.SURF_FLAT
        if self.ai is None and float(self.k) == 0.0:
            return E.SURF_SPHERE
        return E.SURF_ASPHERE

    def engine_record(self):
        kind = self.kind()
        ai = [float(a) for a in self.ai] if (kind == E.SURF_ASPHERE and self.ai is not None) else None
        return E.make_surface(kind, self.r, float(self.d), float(self.c), float(self.k), ai,
                              (self.mat1.A, self.mat1.B), (self.mat2.A, self.mat2.B), square=self.is_square)

    def _state_key(self):
        return (self.r, float(self.d), float(self.c), float(self.k), None if self.ai is None else tuple(self.ai.tolist()),
                self.mat1.name, self.mat2.name, self.is_square)

    def _single_surface_lens(self):
        key = self._state_key()
        if self._handle is None or self._handle_key != key:
            self._handle, self._handle_key = E.LensHandle([self.engine_record()], 0.0), key
        return self._handle

    # ---- reference API ---------------------------------------------------------------------------
    def ray_reaction(self, ray, numerics=None):
        """Intersect + refract at this surface (surfaces.py:391-520), in place on `ray`.  Direction is the sign of
        the summed d_z of the valid rays, as in the reference (surfaces.py:399)."""
        forward = bool((ray.d[..., 2] * ray.ra).sum() > 0)
        E.trace_rays(self._single_surface_lens(), ray.wvln, ray.o, ray.d, ray.ra, 0, 1, backward=not forward,
                     numerics=numerics)
        return ray

    def surf_dict(self):
        d = {"type": "Stop" if float(self.c) == 0 else ("Aspheric" if self.kind() == E.SURF_ASPHERE else "Spheric"),
             "r": self.r, "c": self.c.item(), "d": self.d.item(), "mat1": self.mat1.name, "mat2": self.mat2.name}
        if self.kind() == E.SURF_ASPHERE:
            d["k"] = self.k.item()
            d["ai"] = [] if self.ai is None else [float(a) for a in self.ai]
        return d

"""Host-side mirror of the reference's `deeplens/basics.py` for the dual-pixel hot path: constants, Material
(dispersion in float64 on the host) and the Ray bundle object.  Device work goes through libsdirt_engine."""
import copy
import math

import numpy as np
import torch

from .. import _engine as E
from ..prescription import cauchy_ab, ior as _cauchy_ior

# constants (reference deeplens/basics.py:18-35)
DEVICE = torch.device("cuda:0") if torch.cuda.is_available() else torch.device("cpu")
DEFAULT_WAVE = 0.589
WAVE_RGB = [0.656, 0.589, 0.486]
DEPTH = -20000
GEO_SPP = 2048
MINT = 1e-5
MAXT = 1e5
DELTA = 1e-6
EPSILON = 1e-9


class DeepObj:
    """Minimal counterpart of the reference's DeepObj (basics.py:165-213): `.to(device)` moves tensor attributes."""

    def to(self, device=DEVICE):
        self.device = torch.device(device) if not isinstance(device, torch.device) else device
        for key, val in list(vars(self).items()):
            if torch.is_tensor(val):
                setattr(self, key, val.to(self.device))
            elif isinstance(val, DeepObj) and val is not self:
                val.to(self.device)
            elif isinstance(val, (list, tuple)):
                moved = [v.to(self.device) if (torch.is_tensor(v) or isinstance(v, DeepObj)) else v for v in val]
                if isinstance(val, list):
                    val[:] = moved
        return self

    def clone(self):
        return copy.deepcopy(self)


class Material:
    """Refractive medium (basics.py:299-380).  Only air-like names and Cauchy "n/V" strings are supported, which
    is what the rf50mm / rf35mm prescriptions use; `ior` runs in float64 on the host exactly like the reference."""

    def __init__(self, name=None):
        self.name = "vacuum" if name is None else str(name).lower()
        self.A, self.B = cauchy_ab(self.name)
        self.dispersion = "naive"
        self.glassname = self.name
        if self.name in ("air", "vacuum", "occluder"):
            self.n, self.V = 1.0, math.inf
        else:
            n, v = self.name.split("/")
            self.n, self.V = float(n), float(v)

    def ior(self, wvln):
        return _cauchy_ior((self.A, self.B), wvln)


class Ray(DeepObj):
    """A bundle of rays of one wavelength (basics.py:216-296): `o`, `d` are [..., 3] float32 tensors, `ra` the
    0/1 validity.  `d` is normalised on construction with the reference's arithmetic (by the engine when the
    tensors live on the GPU)."""

    def __init__(self, o, d, wvln=DEFAULT_WAVE, normalized=True, ra=None, en=None, obliq=None, opl=None,
                 coherent=False, device=DEVICE):
        if coherent:
            raise NotImplementedError("coherent ray tracing is outside the dual-pixel hot path (SURVEY.md §2)")
        self.o = o if torch.is_tensor(o) else torch.tensor(o).type(torch.float32)
        self.d = d if torch.is_tensor(d) else torch.tensor(d).type(torch.float32)
        self.wvln = wvln if wvln < 10 else wvln * 1e-3
        self.coherent = False
        shape = self.o.shape[:-1]
        self.ra = ra if ra is not None else torch.full(shape, 1.0, dtype=torch.float32)
        self.en = en if en is not None else torch.full(shape, 1.0, dtype=torch.float32)
        self.obliq = obliq if obliq is not None else torch.full(shape, 1.0, dtype=torch.float32)
        self.opl = opl if opl is not None else torch.full(shape, 0.0, dtype=torch.float32)
        self.phi = torch.zeros_like(self.opl)
        self.to(device)
        self.o = self.o.float().contiguous()
        self.d = _normalize(self.d.float().contiguous())
        self.ra = self.ra.float().contiguous()

    @classmethod
    def _from_engine(cls, o, d, wvln):
        """Wrap tensors the engine produced (already normalised, on the device) without touching them."""
        self = cls.__new__(cls)
        self.o, self.d, self.wvln, self.coherent = o, d, (wvln if wvln < 10 else wvln * 1e-3), False
        shape = o.shape[:-1]
        self.ra = torch.ones(shape, device=o.device, dtype=torch.float32)
        self.en, self.obliq = torch.ones_like(self.ra), torch.ones_like(self.ra)
        self.opl, self.phi = torch.zeros_like(self.ra), torch.zeros_like(self.ra)
        self.device = o.device
        return self

    def prop_to(self, z, n=1):
        return self.propagate_to(z, n)

    def propagate_to(self, z, n=1):
        """o += d * (z - o_z) / d_z  (basics.py:256-264), in place on the device."""
        E.propagate_rays(self.o, self.d, z)
        return self

    def project_to(self, z):
        o = self.o.clone()
        E.propagate_rays(o, self.d, z)
        return o[..., :2]

    def clone(self, device=None):
        new = copy.copy(self)
        for k, v in vars(self).items():
            if torch.is_tensor(v):
                setattr(new, k, v.clone() if device is None else v.to(device).clone())
        if device is not None:
            new.device = torch.device(device)
        return new


def _normalize(d):
    """F.normalize(d, p=2, dim=-1) with the reference's CPU rounding (FMA-chained norm, IEEE divide), on the GPU."""
    if not d.is_cuda:
        raise RuntimeError("sdirt_b200: rays must live on a CUDA device (the engine has no CPU path)")
    d = d.clone() if d.is_contiguous() else d.contiguous()
    E.normalize_rays(d)
    return d

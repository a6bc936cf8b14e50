"""Lens prescription IO for the engine: JSON -> sdirt_surface records (host logic, no device work).

Mirrors Lensgroup.read_lens_json (deeplens/optics.py:2173-2198), Material (deeplens/basics.py:299-380, the
Cauchy "n/V" and air entries the two Sdirt prescriptions use) and find_aperture (optics.py:193-201).
"""
import json
import math

from . import _engine as E

AIR_LIKE = ("air", "vacuum", "occluder")          # MATERIAL_TABLE, basics.py:43-46


def cauchy_ab(name):
    """(A, B) with n(lambda) = A + B / lambda_nm^2 (Material.nV_to_AB, basics.py:354-362)."""
    nm = "vacuum" if name is None else str(name).lower()
    if nm in AIR_LIKE:
        n, v = 1.0, math.inf
    else:
        try:
            a, b = nm.split("/")
            n, v = float(a), float(b)
        except ValueError:
            raise ValueError(f"material '{name}': only air and 'n/V' glasses are supported by the engine") from None
    lam = (656.3, 589.3, 486.1)
    B = (n - 1) / v / (1.0 / lam[2] ** 2 - 1.0 / lam[0] ** 2)
    A = n - B * (1.0 / lam[1] ** 2)
    return A, B


def ior(ab, wvln):
    """Material.ior for the 'naive' (Cauchy) dispersion (basics.py:316-340); wvln in um (or nm if >= 10)."""
    wv = wvln if wvln < 10 else wvln * 1e-3
    return ab[0] + ab[1] / (wv * 1e3) ** 2


def surface_record(sd):
    """One JSON surface dict -> (sdirt_surface, python-side description)."""
    typ = sd["type"]
    c = float(sd.get("c", 0.0))
    k, ai = 0.0, None
    if typ == "Aspheric":
        k, ai = float(sd.get("k", 0.0)), sd.get("ai")
    elif typ not in ("Stop", "Spheric"):
        raise Exception("Surface type not implemented.")
    if c == 0.0:
        kind = E.SURF_FLAT
    elif ai is None and k == 0.0:
        kind = E.SURF_SPHERE
    else:
        kind = E.SURF_ASPHERE
    m1, m2 = cauchy_ab(sd["mat1"]), cauchy_ab(sd["mat2"])
    rec = E.make_surface(kind, sd["r"], sd["d"], c, k, ai if kind == E.SURF_ASPHERE else None, m1, m2)
    return rec, dict(type=typ, kind=kind, r=float(sd["r"]), d=float(sd["d"]), c=c, k=k, ai=ai,
                     mat1=sd["mat1"], mat2=sd["mat2"], ab1=m1, ab2=m2)


def load_lens_json(path):
    """Returns (list of sdirt_surface, list of dict descriptions, header dict)."""
    with open(path) as fh:
        data = json.load(fh)
    recs, descs = [], []
    for sd in data["surfaces"]:
        r, dsc = surface_record(sd)
        recs.append(r)
        descs.append(dsc)
    head = {k: data[k] for k in ("foclen", "fnum", "r_last", "d_sensor", "sensor_size") if k in data}
    return recs, descs, head


def find_aperture(descs):
    """First surface with air on both sides, excluding the last (optics.py:193-201)."""
    for i, d in enumerate(descs[:-1]):
        if d["ab1"][0] < 1.0003 and d["ab2"][0] < 1.0003:
            return i
    return None

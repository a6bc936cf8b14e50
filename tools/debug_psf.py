import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import dp_oracle as O
from test_oracle_golden import make_lens, psf_golden_samples, l1_sumnorm, D_SENSOR
from test_engine_gpu import engine_lens, cu
from sdirt_b200 import _engine as E
name = "rf50mm"
g = np.load(os.path.join(ROOT, "tests/golden/psf.npz"))
lens = make_lens(name, g[f"{name}_hfov"])
obj = g[f"{name}_points_obj"]; pz, pr = g[f"{name}_pupil"]
(px, py), (cx, cy) = psf_golden_samples(g, name)
h = engine_lens(name)
pts, pup = cu(obj), cu(np.stack([px, py], -1)); gc = cu(g[f"{name}_centre"])
res = {}
for num in ("strict", "hybrid", "fast"):
    L, R = E.psf_bank(h, 0.589, pts, pup, float(pz), gc, 21, lens.pixel_size, numerics=num, normalise=0)
    res[num] = L.cpu().numpy()
ref = g[f"{name}_chief_raw"]
for num in res:
    d = res[num] - ref
    i = np.unravel_index(np.abs(d).argmax(), d.shape)
    print(num, "max abs raw diff", d[i], "at", i, "ref value", ref[i], "sum diff per point", (res[num].sum((1, 2)) - ref.sum((1, 2))), "L1", l1_sumnorm(res[num], ref))
    p = i[0]
    print("  neighbourhood diff:\n", np.round(d[p][max(i[1]-2,0):i[1]+3, max(i[2]-2,0):i[2]+3], 2))
    print("  ref:\n", np.round(ref[p][max(i[1]-2,0):i[1]+3, max(i[2]-2,0):i[2]+3], 1))

import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from oracle import dp_oracle as O
from test_oracle_golden import make_lens, torch_pupil, D_SENSOR
from test_engine_gpu import engine_lens, cu
from sdirt_b200 import _engine as E
name = sys.argv[1] if len(sys.argv) > 1 else "rf50mm"
lens = make_lens(name, 0.40959781408309937)
ds = D_SENSOR[name]
ptsn = np.array([[0, 0, -2000 + ds], [0.4, 0.3, -700 + ds], [-0.7, 0.7, -1000.1 + ds], [0.98, -0.98, -20000 + ds], [0, 0.9, -300 + ds], [0.5, -0.2, -5000 + ds]], np.float32)
obj = O.object_points(lens, ptsn)
rng = np.random.default_rng(1)
spp = 20000
px, py = torch_pupil(rng.uniform(0, 1, (2, spp)).astype(np.float32), 6.019352912902832)
ray0 = O.rays_from_points(obj, px, py, 22.51324462890625)
h = engine_lens(name)
out = {}
for label, kw in (("strict", dict(newton="per_ray", numerics="strict")), ("fast", dict(numerics=sys.argv[2] if len(sys.argv) > 2 else "fast"))):
    o, d = cu(ray0.o().reshape(-1, 3)), cu(ray0.d().reshape(-1, 3)); ra = torch.ones(o.shape[0], device="cuda")
    rec = E.trace_rays(h, 0.589, o, d, ra, to_sensor=True, record=True, **kw).cpu().numpy().astype(np.float64)
    out[label] = (rec, o.cpu().numpy().astype(np.float64))
a, b = out["strict"][0], out["fast"][0]
for i in range(a.shape[0]):
    both = (a[i][:, 6] == 1) & (b[i][:, 6] == 1)
    do = b[i][:, :3] - a[i][:, :3]; dd = a[i][:, 3:6]
    along = (do * dd).sum(-1, keepdims=True) * dd
    perp = np.linalg.norm((do - along)[both], axis=-1)
    ddir = np.linalg.norm((b[i][:, 3:6] - a[i][:, 3:6])[both], axis=-1)
    print(f"surf {i:2d} validity mismatch {int((a[i][:,6]!=b[i][:,6]).sum()):5d}  |do_along| max {np.abs((do*dd).sum(-1))[both].max():.2e}  perp mean {perp.mean():.2e} max {perp.max():.2e}   |dd| mean {ddir.mean():.2e} max {ddir.max():.2e}")
so, fo = out["strict"][1], out["fast"][1]
both = (a[-1][:, 6] == 1) & (b[-1][:, 6] == 1)
print("sensor |dx| mean %.2e max %.2e" % (np.abs(so - fo)[both][:, :2].mean(), np.abs(so - fo)[both][:, :2].max()))

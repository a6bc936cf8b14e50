import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import dp_oracle as O
from test_oracle_golden import make_lens, torch_pupil, D_SENSOR
from test_engine_gpu import engine_lens, cu
from sdirt_b200 import _engine as E
name = sys.argv[1] if len(sys.argv) > 1 else "rf50mm"
hf = {"rf50mm": 0.40959781408309937, "rf35mm": 0.5514792203903198}[name]
pz, pr = {"rf50mm": (22.51324462890625, 6.019352912902832), "rf35mm": (14.338210105895996, 4.767455577850342)}[name]
lens = make_lens(name, hf)
ds = D_SENSOR[name]
ptsn = np.array([[0, 0, -2000 + ds], [0.4, 0.3, -700 + ds], [-0.7, 0.7, -1000.1 + ds], [0.98, -0.98, -20000 + ds], [0, 0.9, -300 + ds], [0.5, -0.2, -5000 + ds]], np.float32)
obj = O.object_points(lens, ptsn)
rng = np.random.default_rng(1)
spp = 20000
px, py = torch_pupil(rng.uniform(0, 1, (2, spp)).astype(np.float32), pr)
ray0 = O.rays_from_points(obj, px, py, pz)
with O.precision(np.float64):
    truth = O.RayBundle(*(a.astype(np.float64) for a in (ray0.ox, ray0.oy, ray0.oz, ray0.dx, ray0.dy, ray0.dz, ray0.ra)))
    O.trace_to_sensor(lens, truth, newton_iters="per_ray")
ref = ray0.copy(); O.trace_to_sensor(lens, ref)          # float32 oracle, reference's global loop
h = engine_lens(name)
print("errors vs float64 arbiter on the sensor plane [mm]; pixel = %.4f mm" % lens.pixel_size)
def report(label, ox, oy, ra):
    ok = (ra == 1) & (truth.ra == 1)
    ex, ey = (ox - truth.ox), (oy - truth.oy)
    for p in range(6):
        m = ok[:, p]
        print(f"  {label:10s} pt{p}: mean signed dx {ex[:,p][m].mean():+.2e} dy {ey[:,p][m].mean():+.2e}  mean|d| {np.hypot(ex[:,p][m], ey[:,p][m]).mean():.2e}  max {np.hypot(ex[:,p][m], ey[:,p][m]).max():.2e}  ra mismatch {(ra[:,p]!=truth.ra[:,p]).sum()}")
report("oracle32", ref.ox, ref.oy, ref.ra)
for label, kw in (("strict", dict(newton="per_ray", numerics="strict")), ("hybrid", dict(numerics="hybrid")), ("fast", dict(numerics="fast"))):
    o, d = cu(ray0.o().reshape(-1, 3)), cu(ray0.d().reshape(-1, 3)); ra = torch.ones(o.shape[0], device="cuda")
    E.trace_rays(h, 0.589, o, d, ra, to_sensor=True, **kw)
    on = o.cpu().numpy().reshape(spp, 6, 3)
    report(label, on[..., 0], on[..., 1], ra.cpu().numpy().reshape(spp, 6))

"""Differential check: hybrid PSFs with the packed strict first-surface step vs the ray-by-ray one (SDIRT_DEBUG_SCALAR_STRICT=1)."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
if len(sys.argv) > 1 and sys.argv[1] == "run":
    from sdirt_b200 import _engine as E
    from sdirt_b200.prescription import load_lens_json
    name = "rf50mm"
    recs, descs, head = load_lens_json(os.path.join(os.path.dirname(E.__file__), "lenses", name + ".json"))
    h = E.LensHandle(recs, 62.25)
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    npts, spp = 64, 200000
    pz, pr = 22.51324462890625, 6.019352912902832
    xy = torch.rand(npts, 2, generator=g) * 2 - 1
    depth = -(torch.rand(npts, generator=g) * 19800 + 200) + 62.25
    scale = -depth * np.tan(0.40959781408309937) / 21.633307652783937
    pts = torch.stack([xy[:, 0] * scale * 18, xy[:, 1] * scale * 12, depth], -1).float().to(dev)
    th = torch.rand(spp, generator=g) * 2 * np.pi
    rr = torch.sqrt(torch.rand(spp, generator=g) * pr ** 2)
    pup = E.pupil_sort(torch.stack([rr * torch.cos(th), rr * torch.sin(th)], -1).to(dev), pr)
    centre = E.psf_centre(h, 0.589, pts, (pup[:2048] * 0.25).contiguous(), pz)
    L, R, cnt = E.psf_bank(h, 0.589, pts, pup, pz, centre, 21, 0.046875, numerics=os.environ.get("DBG_NUMERICS", "hybrid"), want_counts=True)
    np.savez(sys.argv[2], L=L.cpu().numpy(), R=R.cpu().numpy(), cnt=cnt.cpu().numpy(), pts=pts.cpu().numpy())
else:
    env = dict(os.environ)
    if "DBG_BASE" in os.environ:
        env["SDIRT_DEBUG_SCALAR_STRICT"] = os.environ["DBG_BASE"]
    subprocess.run([sys.executable, __file__, "run", "/tmp/s2_packed.npz"], check=True, env=env)
    env["SDIRT_DEBUG_SCALAR_STRICT"] = os.environ.get("DBG_LEVEL", "1")
    subprocess.run([sys.executable, __file__, "run", "/tmp/s2_scalar.npz"], check=True, env=env)
    a, b = np.load("/tmp/s2_packed.npz"), np.load("/tmp/s2_scalar.npz")
    d = np.abs(a["L"] - b["L"]).reshape(len(a["L"]), -1).max(1)
    print("points with different L PSFs:", int((d > 0).sum()), "of", len(d), "max diff", d.max())
    print("hit counts differ at:", np.nonzero(a["cnt"] != b["cnt"])[0][:10], (a["cnt"] - b["cnt"])[a["cnt"] != b["cnt"]][:10])
    bad = np.nonzero(d > 0)[0][:8]
    for i in bad:
        print(i, a["pts"][i], d[i], int(a["cnt"][i]), int(b["cnt"][i]))

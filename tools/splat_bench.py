"""Scratch timing of the API-compat kernels on explicit AoS rays (sample_rays -> trace_rays -> splat_rays = the reference's
sample_from_points -> trace2sensor -> forward_integral sequence) against their HBM roofline."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sdirt_b200 import _engine as E
from sdirt_b200.prescription import load_lens_json

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def main():
    npts = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    spp = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hbm = 6650.0
    try:
        hbm = json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    recs, _, _ = load_lens_json(os.path.join(root, "sdirt_b200", "lenses", "rf50mm.json"))
    h = E.LensHandle(recs, 62.25)
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    pz, pr, hfov = 22.51324462890625, 6.019352912902832, 0.40959781408309937
    xy = torch.rand(npts, 2, generator=g) * 2 - 1
    depth = -(torch.rand(npts, generator=g) * 5000 + 300) + 62.25
    scale = -depth * np.tan(hfov) / 21.633307652783937
    pts = torch.stack([xy[:, 0] * scale * 18, xy[:, 1] * scale * 12, depth], -1).float().to(dev)
    th = torch.rand(spp, generator=g) * 2 * np.pi
    rr = torch.sqrt(torch.rand(spp, generator=g) * pr ** 2)
    pup = torch.stack([rr * torch.cos(th), rr * torch.sin(th)], -1).to(dev)
    n = npts * spp
    o, d = E.sample_rays(pts, pup, pz)
    ms = timed(lambda: E.sample_rays(pts, pup, pz))
    print(f"sample_rays   {n:.3e} rays: {ms:8.3f} ms  {n / ms * 1e3:.3e} rays/s  {n * 24 / ms / 1e6:7.0f} GB/s = {n * 24 / ms / 1e6 / hbm:.2f} of HBM peak (24 B/ray written)")
    for numerics in ("strict", "fast"):
        def tr():
            o2, d2 = o.clone(), d.clone()
            ra = torch.ones(spp, npts, device=dev)
            E.trace_rays(h, 0.589, o2.view(-1, 3), d2.view(-1, 3), ra.view(-1), to_sensor=True, numerics=numerics)
            return o2, d2, ra
        base = timed(lambda: (o.clone(), d.clone(), torch.ones(spp, npts, device=dev)))
        ms = timed(tr) - base
        print(f"trace_rays[{numerics:6s}] {n:.3e} rays: {ms:8.3f} ms  {n / ms * 1e3:.3e} rays/s  {n * 56 / ms / 1e6:7.0f} GB/s = {n * 56 / ms / 1e6 / hbm:.2f} of HBM peak (28 B/ray read + 28 written)")
    o2, d2, ra = tr()
    centre = E.psf_centre(h, 0.589, pts, (pup[:2048] * 0.25).contiguous(), pz)
    ms = timed(lambda: E.splat_rays(o2, d2, ra, centre, 21, 0.046875))
    print(f"splat_rays    {n:.3e} rays: {ms:8.3f} ms  {n / ms * 1e3:.3e} rays/s  {n * 24 / ms / 1e6:7.0f} GB/s = {n * 24 / ms / 1e6 / hbm:.2f} of HBM peak (24 B/ray read: o.xy, d, ra)")

main()

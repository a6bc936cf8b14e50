"""Scratch check of the specialised parity kernel (csrc/strict_path.cuh): fraction of rays whose sensor-plane state is
bit-identical to the generic strict trace, pixel agreement, and kernel time.  Not a test; used to compare build variants
(SDIRT_ENGINE_LIB=...).  Usage: python tools/strict_check.py [rf50mm|rf35mm] [npts] [spp]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sdirt_b200 import _engine as E
from sdirt_b200.prescription import load_lens_json


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "rf50mm"
    npts = int(sys.argv[2]) if len(sys.argv) > 2 else 592
    spp = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 20
    recs, _, _ = load_lens_json(os.path.join(os.path.dirname(E.__file__), "lenses", name + ".json"))
    ds = {"rf50mm": 62.25, "rf35mm": 80.447}[name]
    h = E.LensHandle(recs, ds)
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    hfov = {"rf50mm": 0.40959781408309937, "rf35mm": 0.5514792203903198}[name]
    pz, pr = {"rf50mm": (22.51324462890625, 6.019352912902832), "rf35mm": (14.338210105895996, 4.767455577850342)}[name]
    print("library:", E._LIB_PATH)
    m = 200001
    th = torch.rand(m, generator=g) * 2 * np.pi
    rr = torch.sqrt(torch.rand(m, generator=g) * pr ** 2)
    pup = torch.stack([rr * torch.cos(th), rr * torch.sin(th)], -1).to(dev).contiguous()
    tot = same = flags = 0
    for pt in ([0.0, 0.0, -2000 + ds], [-86.98888, 2320.7637, -12153.938], [-5909.853, -2731.2825, -17124.674],
               [250.0, -160.0, -999.0 + ds], [7000.0, 4500.0, -20000.0 + ds], [60.0, 95.0, -200.0 + ds]):
        p = torch.tensor(pt, device=dev)
        got = E.debug_trace_strict2(h, 0.589, p, pup, pz)
        o, d = E.sample_rays(p.reshape(1, 3), pup, pz)
        o, d = o.reshape(-1, 3).contiguous(), d.reshape(-1, 3).contiguous()
        ra = torch.ones(m, device=dev)
        E.trace_rays(h, 0.589, o, d, ra, to_sensor=True, newton="per_ray", numerics="strict")
        want = torch.cat([o, d, ra[:, None]], -1)
        flags += int((got[:, 6] != want[:, 6]).sum())
        keep = (want[:, 6] > 0) & (got[:, 6] > 0)
        if int(keep.sum()) == 0:
            print(f"  point {pt}: no surviving rays"); continue
        eq = (got[keep].view(torch.int32) == want[keep].view(torch.int32)).all(-1)
        tot += int(keep.sum()); same += int(eq.sum())
        ps = 0.046875
        pix = lambda t: torch.floor(t[:, :2] / ps)
        print(f"  point {pt}: alive {int(keep.sum())}, bit-identical {float(eq.float().mean()):.6f}, same pixel "
              f"{float((pix(got[keep]) == pix(want[keep])).all(-1).float().mean()):.6f}, max |dx| {float((got[keep][:, :2] - want[keep][:, :2]).abs().max()):.2e}")
    print(f"rays {tot}: bit-identical fraction {same / tot:.7f}, validity flags differing {flags}")
    # timing on the bank workload of quick_bench
    xy = torch.rand(npts, 2, generator=g) * 2 - 1
    depth = -(torch.rand(npts, generator=g) * 19800 + 200) + ds
    scale = -depth * np.tan(hfov) / 21.633307652783937
    pts = torch.stack([xy[:, 0] * scale * 18, xy[:, 1] * scale * 12, depth], -1).float().to(dev)
    th = torch.rand(spp, generator=g) * 2 * np.pi
    rr = torch.sqrt(torch.rand(spp, generator=g) * pr ** 2)
    pup = E.pupil_sort(torch.stack([rr * torch.cos(th), rr * torch.sin(th)], -1).to(dev), pr)
    centre = E.psf_centre(h, 0.589, pts, (pup[:2048] * 0.25).contiguous(), pz)
    for numerics in os.environ.get("SC_MODES", "strict").split(","):
        for _ in range(2):
            E.psf_bank(h, 0.589, pts, pup, pz, centre, 21, 0.046875, numerics=numerics)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            E.psf_bank(h, 0.589, pts, pup, pz, centre, 21, 0.046875, numerics=numerics)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"{name} N={npts} spp={spp} {numerics}: {ms:.2f} ms  {npts * spp / ms * 1e3:.3e} rays/s")


main()

"""Scratch timing of the fused PSF-bank kernel (not the contract bench; see bench.py)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sdirt_b200 import _engine as E
from sdirt_b200.prescription import load_lens_json

def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "rf50mm"
    npts = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    spp = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 20
    recs, descs, head = load_lens_json(os.path.join(os.path.dirname(E.__file__), "lenses", name + ".json"))
    ds = {"rf50mm": 62.25, "rf35mm": 80.447}[name]
    h = E.LensHandle(recs, ds)
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    hfov = {"rf50mm": 0.40959781408309937, "rf35mm": 0.5514792203903198}[name]
    pz, pr = {"rf50mm": (22.51324462890625, 6.019352912902832), "rf35mm": (14.338210105895996, 4.767455577850342)}[name]
    xy = torch.rand(npts, 2, generator=g) * 2 - 1
    depth = -(torch.rand(npts, generator=g) * 19800 + 200) + ds
    scale = -depth * np.tan(hfov) / 21.633307652783937
    pts = torch.stack([xy[:, 0] * scale * 18, xy[:, 1] * scale * 12, depth], -1).float().to(dev)
    th = torch.rand(spp, generator=g) * 2 * np.pi
    rr = torch.sqrt(torch.rand(spp, generator=g) * pr ** 2)
    pup = torch.stack([rr * torch.cos(th), rr * torch.sin(th)], -1).to(dev)
    cpup = (pup[:2048] * 0.25).contiguous()
    pup_sorted = E.pupil_sort(pup, pr)
    pup_raw = pup
    centre = E.psf_centre(h, 0.589, pts, cpup, pz)
    want = os.environ.get("QB_MODES", "strict,replay,hybrid,adaptive,fast").split(",")
    for mode, numerics in (("per_ray", "strict"), ([10, 3, 4, 3, 4, 0, 3, 3, 4, 5, 3, 3] if name == "rf50mm" else None, "strict"), ("per_ray", "hybrid"), ("per_ray", "adaptive"), ("per_ray", "fast")):
        if mode is None or (numerics if mode == "per_ray" else "replay") not in want:
            continue
        import functools
        pup = pup_raw if (numerics == 'strict' and not os.environ.get('QB_STRICT_SORTED')) else pup_sorted
        E.psf_bank = functools.partial(E.psf_bank.func if hasattr(E.psf_bank, "func") else E.psf_bank, numerics=numerics)
        for _ in range(2):
            L, R, cnt = E.psf_bank(h, 0.589, pts, pup, pz, centre, 21, 0.046875, newton=mode, want_counts=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record()
        for _ in range(reps):
            L, R, cnt = E.psf_bank(h, 0.589, pts, pup, pz, centre, 21, 0.046875, newton=mode, want_counts=True)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"{name} N={npts} spp={spp} newton={'per_ray' if mode == 'per_ray' else 'replay'} {numerics}: {ms:.2f} ms  "
              f"{npts * spp / ms * 1e3:.3e} rays/s  valid frac {cnt.float().mean().item() / spp:.3f}")
    # fp32 FMA probe
    blocks, threads, iters = 148 * 8, 256, 1 << 16
    E.fp32_peak_probe(dev, blocks, threads, 1024); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); E.fp32_peak_probe(dev, blocks, threads, iters); e1.record(); torch.cuda.synchronize()
    fl = blocks * threads * iters * 8 * 2
    print(f"fp32 FMA probe: {fl / e0.elapsed_time(e1) / 1e9:.1f} TFLOP/s")

main()

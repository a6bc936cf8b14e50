"""Scratch timing of the spatially varying DP render kernel (local_psf_render_fast) against its HBM roofline.
Usage: render_bench.py [H W B ks]   (explicit per-pixel PSFs, fp16 and fp32)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sdirt_b200 import _engine as E

def main():
    H = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 1536
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    ks = int(sys.argv[4]) if len(sys.argv) > 4 else 21
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    img = torch.rand((B, 3, H, W), device=dev, generator=g)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    for dt in (torch.float16, torch.float32):
        psf = torch.rand((B, H, W, 2, ks, ks), device=dev, generator=g, dtype=torch.float32)
        psf = (psf / psf.sum((-1, -2), keepdim=True)).to(dt).contiguous()
        for tone in (0, 3):
            for _ in range(2):
                E.render_local_psf(img, psf, ks, tone=tone)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            e0.record()
            for _ in range(reps):
                E.render_local_psf(img, psf, ks, tone=tone)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            px = B * H * W
            byt = px * (2 * ks * ks * psf.element_size() + 3 * 4 + 6 * 4)
            print(f"render {B}x3x{H}x{W} ks={ks} psf={str(dt).split('.')[-1]} tone={tone}: {ms:.3f} ms  {px / ms * 1e3:.3e} px/s  "
                  f"{byt / ms / 1e6:.0f} GB/s algorithmic = {byt / ms / 1e6 / hbm:.3f} of measured HBM peak {hbm:.0f} GB/s")
        del psf

main()

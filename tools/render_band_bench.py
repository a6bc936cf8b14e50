"""Scratch timing of one BAND of the render (sdirt_render_local_psf_rows as PSFNet.render calls it): rows x W pixels of nb images.
Usage: render_band_bench.py [rows nb H W ks]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sdirt_b200 import _engine as E

def main():
    a = [int(v) for v in sys.argv[1:]]
    rows, nb, H, W, ks = (a + [64, 1, 1024, 1536, 21][len(a):])[:5]
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    img = torch.rand((nb, 3, H, W), device=dev, generator=g)
    psf = torch.rand((nb, rows, W, 2, ks, ks), device=dev, generator=g)
    psf = (psf / psf.sum((-1, -2), keepdim=True)).half().contiguous()
    rl, rr = torch.empty_like(img), torch.empty_like(img)
    for tone in (0, 2):
        for _ in range(3):
            E.render_local_psf_rows(img, psf, ks, 128, rl, rr, tone=tone)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for i in range(reps):
            E.render_local_psf_rows(img, psf, ks, 64 * (i % 8), rl, rr, tone=tone)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / reps * 1e3
        byt = nb * rows * W * 2 * ks * ks * 2
        print(f"band {nb} x {rows} x {W} ks={ks} tone={tone}: {us:.1f} us per call, {byt / us / 1e3:.0f} GB/s of kernels")

main()

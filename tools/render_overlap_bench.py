"""PSFNet.render (fused MLP engine) with and without the two-stream overlap, and the band render alone at the band shape it uses.
Usage: render_overlap_bench.py [H W B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sdirt_b200 import _engine as E, lens_file
from sdirt_b200.deeplens import PSFNet

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def main():
    a = [int(v) for v in sys.argv[1:]]
    H, W, B = (a + [1024, 1536, 2][len(a):])[:3]
    dev = torch.device("cuda:0")
    lens = PSFNet(lens_file("rf50mm"), sensor_res=(H, W), kernel_size=21, device=dev)
    g = torch.Generator(device=dev).manual_seed(0)
    img = torch.rand((B, 3, H, W), device=dev, generator=g)
    low = torch.rand((B, 1, H // 64 + 2, W // 64 + 2), device=dev, generator=g)
    depth = -(torch.nn.functional.interpolate(low, size=(H, W), mode="bilinear", align_corners=False) * 9750 + 250)
    foc = torch.full((B,), -1000.0, device=dev)
    rows, nb = lens._fused_band_shape(B, H, W)
    for ov in (1, 0):
        lens.render_overlap = bool(ov)
        ms = timed(lambda: lens.render(img, depth, foc))
        print(f"PSFNet.render {B}x3x{H}x{W} overlap={ov}: {ms:.3f} ms  {B * H * W / ms * 1e3:.4e} px/s   band = {nb} x {rows} rows")
    psf = torch.rand((nb, rows, W, 2, 21, 21), device=dev, generator=g)
    psf = (psf / psf.sum((-1, -2), keepdim=True)).half().contiguous()
    rl, rr = torch.empty_like(img), torch.empty_like(img)
    t = timed(lambda: E.render_local_psf_rows(img[:nb], psf, 21, 160, rl[:nb], rr[:nb], tone=2), 20)
    print(f"band render alone (unpacked entry): {t * 1e3:.1f} us")
    if hasattr(E, "render_pack_image"):
        rec = E.render_pack_image(img[:nb].contiguous(), 21, 1)
        t = timed(lambda: E.render_local_psf_rows_packed(rec, (nb, 3, H, W), psf, 21, 160, rl[:nb], rr[:nb], tone=2), 20)
        print(f"band render alone (packed entry): {t * 1e3:.1f} us; pack of the images: {timed(lambda: E.render_pack_image(img[:nb].contiguous(), 21, 1), 10) * 1e3:.1f} us")

main()

import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
from sdirt_b200 import _engine as E, lens_file
from sdirt_b200.deeplens import PSFNet
dev = torch.device("cuda:0")
H, W = 1024, 1536
lens = PSFNet(lens_file("rf50mm"), sensor_res=(H, W), kernel_size=21, device=dev)
g = torch.Generator(device=dev).manual_seed(7)
for B in (8, 16, 8):
    img = torch.rand((B, 3, H, W), device=dev, generator=g)
    low = torch.rand((B, 1, H // 64 + 2, W // 64 + 2), device=dev, generator=g)
    depth = -(torch.nn.functional.interpolate(low, size=(H, W), mode="bilinear", align_corners=False) * 9750 + 250)
    foc = torch.full((B,), -1000.0, device=dev)
    lens.render(img[:1], depth[:1], foc[:1])
    torch.cuda.synchronize()
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(); out = lens.render(img, depth, foc); e1.record(); torch.cuda.synchronize()
        print(f"B={B} rep {rep}: {e0.elapsed_time(e1):.1f} ms (wall {1e3 * (time.perf_counter() - t0):.1f}), band {lens._fused_band_shape(B, H, W)}  {B * H * W / e0.elapsed_time(e1) / 1e3:.3e} px/s")
    del img, depth, out

"""BASELINE config 4: spatially varying DP render of 1024 x 1536 RGB-D scenes through PSFNet.render (banded: engine kernels
around the cuBLAS GEMM chain), timed per stage, next to the reference's order of operations (`render_via_pred`).
Usage: render_c4_bench.py [H W B] [band_rows band_pixels]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sdirt_b200 import _engine as E, lens_file
from sdirt_b200.deeplens import PSFNet


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    a = [int(v) for v in sys.argv[1:]]
    H, W, B = (a + [1024, 1536, 4])[:3] if len(a) < 3 else a[:3]
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    lens = PSFNet(lens_file("rf50mm"), sensor_res=(H, W), kernel_size=21, device=dev)
    if len(a) >= 5:
        lens.render_band_rows, lens.render_band_pixels = a[3], a[4]
    if len(a) >= 6:
        lens.render_overlap = bool(a[5])
    g = torch.Generator(device=dev).manual_seed(0)
    img = torch.rand((B, 3, H, W), device=dev, generator=g)
    # smooth random depth field clipped to 0.25 .. 10 m (NYUv2 range), negative millimetres
    low = torch.rand((B, 1, H // 64 + 2, W // 64 + 2), device=dev, generator=g)
    depth = -(torch.nn.functional.interpolate(low, size=(H, W), mode="bilinear", align_corners=False) * 9750 + 250)
    foc = torch.full((B,), -1000.0, device=dev)
    px = B * H * W
    flop_px = 2 * 2 * (3 * 128 + 128 * 512 + 8 * 512 * 512 + 512 * 441)          # both sides, FMA = 2
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    lens.mlp_engine = "cublas"
    ms_c = timed(lambda: lens.render(img, depth, foc))
    print(f"PSFNet.render, cuBLAS route (rows={lens.render_band_rows} band_px={lens.render_band_pixels}): {ms_c:.2f} ms  {B * H * W / ms_c * 1e3:.3e} px/s  "
          f"{B * H * W * flop_px / (ms_c * 1e-3) / 1e12:.0f} TFLOP/s")
    lens.mlp_engine = "fused"
    print("fused engine band shape (rows, images):", lens._fused_band_shape(B, H, W))
    l0 = E.launch_count()
    ms = timed(lambda: lens.render(img, depth, foc))
    tf = px * flop_px / (ms * 1e-3) / 1e12
    peak = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 0.0)) or float("nan")
    print(f"PSFNet.render banded, fused MLP kernel, {B}x3x{H}x{W} ks=21 overlap={int(lens.render_overlap)}: {ms:.2f} ms  "
          f"{px / ms * 1e3:.3e} px/s  {tf:.0f} TFLOP/s (MLP, 9.56 MFLOP/px) = {tf / peak:.3f} of measured dense 16-bit peak {peak:.0f}; "
          f"engine launches/call {(E.launch_count() - l0) // 4}; peak mem {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB")
    # per-stage split of one band batch
    (w1, b1), chain = lens._mlp_half_layers()
    rows = lens.render_band_rows
    nb = max(1, min(B, lens.render_band_pixels // (rows * W)))
    z = lens.depth2z(depth + lens.d_sensor).reshape(B, H, W).float().contiguous()
    xs, ys = torch.linspace(-1, 1, W).to(dev), torch.linspace(1, -1, H).to(dev)
    rl, rr = torch.empty_like(img), torch.empty_like(img)
    h1 = E.mlp_input_layer(xs, ys, z, 0, nb, 0, rows, w1, b1)

    def gemms():
        h = h1
        for wt, b in chain:
            h = torch._addmm_activation(b, h, wt)
        return h
    raw = gemms()
    psf = E.psf_pack(raw, 21).view(nb, rows, W, 2, 21, 21)
    bands = (B // nb) * (H // rows)
    t_in = timed(lambda: E.mlp_input_layer(xs, ys, z, 0, nb, 0, rows, w1, b1), 10)
    t_g = timed(gemms, 10)
    t_p = timed(lambda: E.psf_pack(raw, 21), 10)
    t_r = timed(lambda: E.render_local_psf_rows(img[:nb], psf, 21, 0, rl[:nb], rr[:nb], tone=3), 10)
    bpx = nb * rows * W
    print(f"  one band batch ({nb} images x {rows} rows = {bpx} px, {bands} per call): input layer {t_in * 1e3:.0f} us, GEMM chain {t_g * 1e3:.0f} us "
          f"({bpx * flop_px / (t_g * 1e-3) / 1e12:.0f} TFLOP/s), pack {t_p * 1e3:.0f} us, render {t_r * 1e3:.0f} us; sum x bands = {(t_in + t_g + t_p + t_r) * bands:.2f} ms")
    hh, per = h1, []
    for wt, b in chain:
        t = timed(lambda: torch._addmm_activation(b, hh, wt), 10)
        per.append(f"{hh.shape[1]}->{wt.shape[1]}: {t * 1e3:.0f} us ({2 * hh.shape[0] * hh.shape[1] * wt.shape[1] / (t * 1e-3) / 1e12:.0f} TF)")
        hh = torch._addmm_activation(b, hh, wt)
    print("  GEMM layers: " + ", ".join(per))
    if B * H * W <= 2 * 1024 * 1536:
        torch.cuda.reset_peak_memory_stats()
        msv = timed(lambda: lens.render_via_pred(img, depth, foc), 2)
        print(f"reference order (pred for all pixels, then one convolution): {msv:.2f} ms  {px / msv * 1e3:.3e} px/s; peak mem "
              f"{torch.cuda.max_memory_allocated() / 2**30:.2f} GiB  -> banded is {msv / ms:.2f}x")


main()

"""Generate golden vectors by RUNNING THE UNMODIFIED REFERENCE on CPU (this container only).

    python tests/golden/make_golden.py        # writes tests/golden/*.npz

The reference has no tests or fixtures of its own, so these files are what pins the
oracle (`oracle/dp_oracle.py`) and, through it, the CUDA engine.  Nothing here is imported by the
product.  The GPU box has no /root/reference: tests only read the committed .npz files.

Reproducibility.  Integer and host-side outputs (Newton loop counts, network inputs, DP-weight and trace
vectors) regenerate bit for bit.  Traced PSFs regenerate to ~1e-5 (200 k rays) ... 3e-5 (2 M rays) L1, NOT bit
for bit: the reference's own `sample_from_points` returns ray directions that differ in the last bit for ~30 %
of their components between two calls with the same seed in the same process (torch's vectorised CPU kernels
take different code paths depending on the alignment of the buffers they are handed), and everything
downstream inherits that.  Measured here with two PSFNet instances, seed 21, 100 k rays x 5 points: `o` equal,
`d` differing by 1 ulp in 440 948 of 1.5 M components, 14 rays with a different validity flag.  The committed
files are one such realisation; the parity tolerances of the task (1e-5 on hits, 1e-4 on PSFs) sit above it.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _ref_import import import_reference  # noqa: E402

dl = import_reference()
from deeplens.basics import Ray  # noqa: E402
from deeplens.monte_carlo import forward_integral, assign_points_to_pixels_small_r, assign_points_to_pixels_big_r  # noqa: E402
from deeplens.render_psf import local_psf_render_fast  # noqa: E402

torch.set_num_threads(1)          # bit-reproducible reductions / index_put_ order
LENSES = {"rf50mm": "/root/reference/lenses/rf50mm/lens_web.json",
          "rf35mm": "/root/reference/lenses/rf35mm/lens_web.json"}
DP = (0.78, 1.44, 0.3, 0.5)


def make_lens(name, res=(512, 768)):
    return dl.PSFNet(LENSES[name], sensor_res=res, kernel_size=21, device="cpu")


def newton_counts(lens, ray):
    """Trace a copy surface by surface, recording state and the global Newton loop count."""
    states, counts = [], []
    for s in lens.surfaces:
        n_calls = {"n": 0}
        orig = s.g

        def counting_g(x, y, valid=None, _o=orig, _c=n_calls):
            _c["n"] += 1
            return _o(x, y, valid)
        s.g = counting_g
        ray = s.ray_reaction(ray)
        del s.g
        counts.append(max(n_calls["n"] - 1, 0))       # loop evaluations (one extra strict evaluation)
        states.append(np.concatenate([ray.o.numpy(), ray.d.numpy(), ray.ra.numpy()[..., None]], -1))
    return ray, np.stack(states), np.asarray(counts)


def golden_setup():
    out = {}
    for name in LENSES:
        lens = make_lens(name)
        pz, pr = lens.entrance_pupil()
        epz, epr = lens.exit_pupil()
        out[f"{name}_scalars"] = np.asarray([lens.aper_idx, lens.hfov, lens.foclen, lens.fnum, pz, pr, epz, epr,
                                             lens.pixel_size, lens.d_sensor, lens.r_last], np.float64)
        etas = [[s.mat1.ior(w) / s.mat2.ior(w) for s in lens.surfaces] for w in (0.656, 0.589, 0.486)]
        out[f"{name}_eta"] = np.asarray(etas, np.float64)
        torch.manual_seed(0)
        lens.refocus(-1000 + lens.d_sensor)
        out[f"{name}_refocus"] = np.asarray([lens.d_sensor, lens.hfov], np.float64)
        torch.manual_seed(0)
        out[f"{name}_refocus_u"] = np.stack([torch.rand(2048).numpy(), torch.rand(2048).numpy()])
    np.savez_compressed(os.path.join(HERE, "setup.npz"), **out)


def golden_trace():
    """Per-surface (o, d, ra) for [spp=256, N=6] bundles, both lenses, three wavelengths for rf50mm."""
    out = {}
    for name in LENSES:
        lens = make_lens(name)
        ds = lens.d_sensor
        pts = torch.tensor([[0, 0, -2000 + ds], [0.4, 0.3, -500 + ds], [-0.7, 0.7, -1000.1 + ds],
                            [0.98, -0.98, -20000 + ds], [0.0, 0.95, -300 + ds], [-0.98, 0.0, -5000 + ds]],
                           dtype=torch.float32)
        depth = pts[:, 2]
        scale = lens.calc_scale_pinhole(depth)
        obj = pts.clone()
        obj[:, 0] = pts[:, 0] * scale * lens.sensor_size[1] / 2
        obj[:, 1] = pts[:, 1] * scale * lens.sensor_size[0] / 2
        out[f"{name}_points_norm"] = pts.numpy()
        out[f"{name}_points_obj"] = obj.numpy()
        out[f"{name}_hfov"] = np.float64(lens.hfov)
        for wv in ((0.589, 0.656, 0.486) if name == "rf50mm" else (0.589,)):
            torch.manual_seed(7)
            u = torch.rand(2, 256)
            torch.manual_seed(7)
            ray = lens.sample_from_points(o=obj, spp=256, wvln=wv)
            pz, pr = lens.entrance_pupil()
            tag = f"{name}_w{int(wv * 1000)}"
            out[f"{tag}_u"] = u.numpy()
            out[f"{tag}_pupil"] = np.asarray([pz, pr], np.float64)
            out[f"{tag}_ray0"] = np.concatenate([ray.o.numpy(), ray.d.numpy()], -1)
            ray, states, counts = newton_counts(lens, ray)
            out[f"{tag}_states"] = states
            out[f"{tag}_newton"] = counts
            ray = ray.propagate_to(lens.d_sensor)
            out[f"{tag}_sensor"] = np.concatenate([ray.o.numpy(), ray.d.numpy(), ray.ra.numpy()[..., None]], -1)
    # backward + sub-range trace (the entrance-pupil setup trace, optics.py:1335-1376)
    lens = make_lens("rf50mm")
    ap = lens.surfaces[lens.aper_idx]
    phi = torch.linspace(-0.1, 0.1, 16) / 180.0 * torch.pi
    o = torch.tensor([[1e-3, 0, ap.d.item()]]).repeat(16, 1)
    d = torch.stack((torch.sin(phi), torch.zeros_like(phi), -torch.cos(phi)), -1)
    ray = Ray(o, d, device="cpu")
    out["back_ray0"] = np.concatenate([ray.o.numpy(), ray.d.numpy()], -1)
    ray, _, _ = lens.trace(ray, lens_range=range(0, lens.aper_idx))
    out["back_final"] = np.concatenate([ray.o.numpy(), ray.d.numpy(), ray.ra.numpy()[..., None]], -1)
    np.savez_compressed(os.path.join(HERE, "trace.npz"), **out)


def golden_dp():
    x_tan = torch.linspace(-0.6, 0.6, 241)
    pts = torch.zeros(241, 2)
    ra = torch.ones(241)
    out = {"x_tan": x_tan.numpy()}
    ks, ps = 21, 0.046875
    rng = [(-ks / 2 + 0.5) * ps, (ks / 2 - 0.5) * ps]
    for tag, prm, fn in (("small", DP + ("l",), assign_points_to_pixels_small_r),
                         ("small_b", (0.6, 1.2, 0.25, 0.45, "l"), assign_points_to_pixels_small_r),
                         ("big", (0.78, 1.44, 0.3, 0.6, "l"), assign_points_to_pixels_big_r)):
        dl_, dr_ = [], []
        for i in range(241):
            L, R = fn(points=pts[i:i + 1], ks=ks, x_range=rng, y_range=rng, ra=ra[i:i + 1], x_tan=x_tan[i:i + 1],
                      param_list=prm)
            dl_.append(L.sum().item())
            dr_.append(R.sum().item())
        out[f"{tag}_params"] = np.asarray(prm[:4])
        out[f"{tag}_dl"] = np.asarray(dl_, np.float32)
        out[f"{tag}_dr"] = np.asarray(dr_, np.float32)
    # splat KAT
    p = torch.tensor([[0.3 * ps, -1.7 * ps]])
    L, R = assign_points_to_pixels_small_r(points=p, ks=ks, x_range=rng, y_range=rng, ra=torch.ones(1),
                                           x_tan=torch.tensor([0.0932]), param_list=DP + ("l",))
    out["kat_L"], out["kat_R"] = L.numpy(), R.numpy()
    np.savez_compressed(os.path.join(HERE, "dp_weights.npz"), **out)


def golden_psf():
    """psf_diff L/R on identical samples + centres; ks=21 and ks=11; small_r, big_r; RGB."""
    out = {}
    for name, spp in (("rf50mm", 200000), ("rf35mm", 200000)):
        lens = make_lens(name)
        ds = lens.d_sensor
        pts = torch.tensor([[0, 0, -2000 + ds], [0.4, 0.3, -700 + ds], [-0.7, 0.7, -1000.1 + ds],
                            [0.98, -0.98, -20000 + ds], [0.0, 0.9, -300 + ds]], dtype=torch.float32)
        out[f"{name}_points_norm"] = pts.numpy()
        out[f"{name}_hfov"] = np.float64(lens.hfov)
        pz, pr = lens.entrance_pupil()
        out[f"{name}_pupil"] = np.asarray([pz, pr], np.float64)
        for tag, prm, ks in (("l", DP + ("l",), 21), ("r", DP + ("r",), 21), ("big_l", (0.78, 1.44, 0.3, 0.6, "l"), 21),
                             ("big_r", (0.78, 1.44, 0.3, 0.6, "r"), 21), ("ks11_l", DP + ("l",), 11), ("none", None, 21)):
            torch.manual_seed(3)
            u = [torch.rand(spp).numpy() for _ in range(2)] + [torch.rand(2048).numpy() for _ in range(2)]
            torch.manual_seed(3)
            psf = lens.psf_diff(pts, ks=ks, spp=spp, param_list=prm)
            out[f"{name}_{tag}"] = psf.numpy()
            # samples are NOT stored: torch's CPU generator is portable, tests redraw them from the seed;
            # these checksums guard that assumption
            out[f"{name}_u_check"] = np.asarray([float(v.astype(np.float64).sum()) for v in u] + [spp])
        # the chief-ray centre alone, same RNG position (after the main bundle draws)
        torch.manual_seed(3)
        torch.rand(spp), torch.rand(spp)
        depth = pts[:, 2]
        scale = lens.calc_scale_pinhole(depth)
        obj = pts.clone()
        obj[:, 0] = pts[:, 0] * scale * lens.sensor_size[1] / 2
        obj[:, 1] = pts[:, 1] * scale * lens.sensor_size[0] / 2
        out[f"{name}_centre"] = lens.psf_center(obj).numpy()
        out[f"{name}_points_obj"] = obj.numpy()
        # unnormalised L from forward_integral with pointc_ref=None (RMS centre)
        torch.manual_seed(3)
        ray = lens.sample_from_points(o=obj, spp=spp)
        ray = lens.trace2sensor(ray)
        out[f"{name}_rms_raw"] = forward_integral(ray, ps=lens.pixel_size, ks=21, pointc_ref=None).numpy()
        out[f"{name}_chief_raw"] = forward_integral(ray, ps=lens.pixel_size, ks=21,
                                                    pointc_ref=torch.from_numpy(out[f"{name}_centre"])).numpy()
        if name == "rf50mm":
            torch.manual_seed(3)
            out["rf50mm_rgb"] = lens.psf_rgb(pts[:2], ks=21, spp=4000).numpy()
    np.savez_compressed(os.path.join(HERE, "psf.npz"), **out)


def golden_psf2m():
    """BASELINE config-2 sample count: 2 M rays per point, three points incl. the worst case for float32
    cancellation at the first surface (field corner at 20 m)."""
    out = {}
    spp = 2_000_000
    lens = make_lens("rf50mm")
    ds = lens.d_sensor
    pts = torch.tensor([[0, 0, -2000 + ds], [0.98, -0.98, -20000 + ds], [0.4, 0.3, -700 + ds]], dtype=torch.float32)
    out["points_norm"] = pts.numpy()
    out["hfov"] = np.float64(lens.hfov)
    pz, pr = lens.entrance_pupil()
    out["pupil"] = np.asarray([pz, pr], np.float64)
    for tag in ("l", "r"):
        torch.manual_seed(9)
        u = [torch.rand(spp).numpy() for _ in range(2)] + [torch.rand(2048).numpy() for _ in range(2)]
        torch.manual_seed(9)
        out[tag] = lens.psf_diff(pts, ks=21, spp=spp, param_list=DP + (tag,)).numpy()
        out["u_check"] = np.asarray([float(v.astype(np.float64).sum()) for v in u] + [spp])
    depth = pts[:, 2]
    scale = lens.calc_scale_pinhole(depth)
    obj = pts.clone()
    obj[:, 0] = pts[:, 0] * scale * lens.sensor_size[1] / 2
    obj[:, 1] = pts[:, 1] * scale * lens.sensor_size[0] / 2
    out["points_obj"] = obj.numpy()
    torch.manual_seed(9)
    torch.rand(spp), torch.rand(spp)
    out["centre"] = lens.psf_center(obj).numpy()
    np.savez_compressed(os.path.join(HERE, "psf2m.npz"), **out)


def golden_psf2m_sweep():
    """2 M rays per point along the depth axis (0.5 ... 20 m) at the field corner and at mid field: how the float32
    lattice of the reference's first-surface hit (ulp of |o| ~ 1e2 ... 8e3 mm) shows up in its PSFs with distance."""
    out = {}
    spp = 2_000_000
    lens = make_lens("rf50mm")
    ds = lens.d_sensor
    dist = [500, 1000, 2000, 4000, 8000, 20000]
    pts = [[0.98, -0.98, -d + ds] for d in dist] + [[-0.5, 0.45, -d + ds] for d in (1500, 3000, 6000, 12000)]
    pts = torch.tensor(pts, dtype=torch.float32)
    out["points_norm"] = pts.numpy()
    out["hfov"] = np.float64(lens.hfov)
    pz, pr = lens.entrance_pupil()
    out["pupil"] = np.asarray([pz, pr], np.float64)
    for tag in ("l", "r"):
        torch.manual_seed(21)
        u = [torch.rand(spp).numpy() for _ in range(2)]
        torch.manual_seed(21)
        out[tag] = np.concatenate([lens.psf_diff(pts[i:i + 5], ks=21, spp=spp, param_list=DP + (tag,)).numpy()
                                   if i == 0 else _psf_same_seed(lens, pts[i:i + 5], spp, tag) for i in range(0, len(pts), 5)])
        out["u_check"] = np.asarray([float(v.astype(np.float64).sum()) for v in u] + [spp])
    depth = pts[:, 2]
    scale = lens.calc_scale_pinhole(depth)
    obj = pts.clone()
    obj[:, 0] = pts[:, 0] * scale * lens.sensor_size[1] / 2
    obj[:, 1] = pts[:, 1] * scale * lens.sensor_size[0] / 2
    out["points_obj"] = obj.numpy()
    torch.manual_seed(21)
    torch.rand(spp), torch.rand(spp)
    out["centre"] = lens.psf_center(obj).numpy()
    # the bundle-global Newton loop counts of the two five-point bundles psf_diff traced (surfaces.py:543-561 `while any()`):
    # what sdirt_psf_bank replays in the `replay` parity test
    counts = []
    for i in range(0, len(pts), 5):
        torch.manual_seed(21)
        ray = lens.sample_from_points(o=obj[i:i + 5], spp=spp)
        counts.append(newton_counts(lens, ray)[2])
    out["newton_counts"] = np.asarray(counts, np.int32)
    np.savez_compressed(os.path.join(HERE, "psf2m_sweep.npz"), **out)


def golden_rf35mm2m():
    """BASELINE config 3's prescription at the config-2 sample count: 2 M rays per point through the 21 surfaces of rf35mm,
    on axis at 2 m, the field corner at 20 m (coarsest first-hit lattice) and mid field at 0.7 m; with the bundle-global
    Newton loop counts for the replay test."""
    out = {}
    spp = 2_000_000
    lens = make_lens("rf35mm")
    ds = lens.d_sensor
    pts = torch.tensor([[0, 0, -2000 + ds], [0.98, -0.98, -20000 + ds], [0.4, 0.3, -700 + ds]], dtype=torch.float32)
    out["points_norm"] = pts.numpy()
    out["hfov"] = np.float64(lens.hfov)
    pz, pr = lens.entrance_pupil()
    out["pupil"] = np.asarray([pz, pr], np.float64)
    for tag in ("l", "r"):
        torch.manual_seed(33)
        u = [torch.rand(spp).numpy() for _ in range(2)]
        torch.manual_seed(33)
        out[tag] = lens.psf_diff(pts, ks=21, spp=spp, param_list=DP + (tag,)).numpy()
        out["u_check"] = np.asarray([float(v.astype(np.float64).sum()) for v in u] + [spp])
    scale = lens.calc_scale_pinhole(pts[:, 2])
    obj = pts.clone()
    obj[:, 0] = pts[:, 0] * scale * lens.sensor_size[1] / 2
    obj[:, 1] = pts[:, 1] * scale * lens.sensor_size[0] / 2
    out["points_obj"] = obj.numpy()
    torch.manual_seed(33)
    torch.rand(spp), torch.rand(spp)
    out["centre"] = lens.psf_center(obj).numpy()
    torch.manual_seed(33)
    ray = lens.sample_from_points(o=obj, spp=spp)
    out["newton_counts"] = np.asarray(newton_counts(lens, ray)[2], np.int32)
    np.savez_compressed(os.path.join(HERE, "rf35mm2m.npz"), **out)


def golden_api():
    """Seeded outputs of the API entry points around the hot path that round 1 left untested: the two other render
    variants (render_psf.py:76-118, 157-188), psf_map (optics.py:1018-1041), get_training_data / get_test_data
    (psfnet.py:170-241), analysis_rms and calc_magnification3 (optics.py:2103-2140, 1237-1272)."""
    from deeplens.render_psf import local_psf_render, local_dp_psf_render
    out = {}
    torch.manual_seed(17)
    for ks, (b, h, w) in ((7, (2, 12, 20)), (11, (1, 10, 16))):
        img = torch.rand(b, 3, h, w)
        psf = torch.rand(b, h, w, 2, ks, ks) ** 4
        psf = psf / psf.sum((-1, -2), keepdim=True)
        rl, rr = local_psf_render(img, psf, ks)
        out[f"rv{ks}_img"], out[f"rv{ks}_psf"] = img.numpy(), psf.numpy()
        out[f"rv{ks}_rl"], out[f"rv{ks}_rr"] = rl.numpy(), rr.numpy()
        out[f"rv{ks}_dp"] = local_dp_psf_render(img, psf, ks).numpy()            # float32 arithmetic
    for name in LENSES:
        lens = make_lens(name)
        ds = lens.d_sensor
        torch.manual_seed(4)
        out[f"{name}_psf_map"] = lens.psf_map(depth=-1500.0 + ds, grid=3, ks=11, spp=20000).numpy()
        torch.manual_seed(4)
        out[f"{name}_psf_map_u"] = np.stack([torch.rand(20000).numpy(), torch.rand(20000).numpy()])
        np.random.seed(8)
        torch.manual_seed(8)
        inp, psf = lens.get_training_data(bs=8, spp=20000)
        out[f"{name}_train_inp"], out[f"{name}_train_psf"] = inp.numpy(), psf.numpy()
        torch.manual_seed(9)
        inp, psf = lens.get_test_data(bs=1024, spp=2048)
        out[f"{name}_test_inp"], out[f"{name}_test_psf"] = inp.numpy(), psf.numpy()[::16]
        torch.manual_seed(10)
        out[f"{name}_mag3"] = np.asarray([lens.calc_magnification3(-1000.0 + ds), lens.calc_magnification3(-20000.0 + ds)], np.float64)
        torch.manual_seed(11)
        out[f"{name}_rms"] = np.asarray([float(v) for v in lens.analysis_rms(depth=-1000.0 + ds)] +
                                        [float(v) for v in lens.analysis_rms(depth=-5000.0 + ds)], np.float64)
    np.savez_compressed(os.path.join(HERE, "api.npz"), **out)


def _psf_same_seed(lens, pts, spp, tag):
    torch.manual_seed(21)
    return lens.psf_diff(pts, ks=21, spp=spp, param_list=DP + (tag,)).numpy()


def golden_render():
    torch.manual_seed(11)
    lens = make_lens("rf50mm", res=(32, 48))
    out = {}
    for ks, (b, h, w) in ((7, (2, 20, 28)), (21, (1, 16, 24))):
        img = torch.rand(b, 3, h, w)
        psf = torch.rand(b, h, w, 2, ks, ks) ** 4
        psf = psf / psf.sum((-1, -2), keepdim=True)
        rl, rr = local_psf_render_fast(img, psf, ks)
        out[f"ks{ks}_img"], out[f"ks{ks}_psf"] = img.numpy(), psf.numpy().astype(np.float16)
        out[f"ks{ks}_psf32_scale"] = np.float32(1.0)
        out[f"ks{ks}_rl"], out[f"ks{ks}_rr"] = rl.numpy(), rr.numpy()
        # psf stored as fp16 (the reference casts to half first, so this loses nothing)
    x = torch.linspace(0, 1, 1001)
    out["tone_x"] = x.numpy()
    out["tone_degamma"] = lens.degamma(x.clone()).numpy()
    out["tone_gamma"] = lens.gamma(lens.degamma(x.clone())).numpy()
    lum = torch.linspace(0, 1100, 1001)
    out["tone_lum"] = lum.numpy()
    out["tone_gamma_lum"] = lens.gamma(lum.clone()).numpy()
    # full render(): seeded random-init MLP (the real checkpoint is not in the tree)
    torch.manual_seed(5)
    lens = dl.PSFNet(LENSES["rf50mm"], sensor_res=(16, 24), kernel_size=21, device="cpu")
    img = torch.rand(2, 3, 16, 24)
    depth = -(torch.rand(2, 1, 16, 24) * 9750 + 250)
    foc = torch.tensor([-1000.0, -1000.0])
    out["render_img"], out["render_depth"], out["render_foc"] = img.numpy(), depth.numpy(), foc.numpy()
    out["render_out"] = lens.render(img, depth, foc).numpy()
    x, y = torch.meshgrid(torch.linspace(-1, 1, 24), torch.linspace(1, -1, 16), indexing="xy")
    z = lens.depth2z(depth + lens.d_sensor).squeeze(1)
    o = torch.stack((x[None].repeat(2, 1, 1), y[None].repeat(2, 1, 1), z), -1).float()
    psf = lens.pred(o.clone()).detach()
    out["render_psf_probe"] = psf[:, ::5, ::7].numpy()          # a few per-pixel PSFs, fp32
    # the MLP is re-created from the seed by the tests (torch CPU RNG is portable); pin it by checksums
    sd = lens.psfnet.state_dict()
    out["mlp_checksum"] = np.asarray([[v.double().sum().item(), v.double().abs().sum().item()] for v in sd.values()])
    np.savez_compressed(os.path.join(HERE, "render.npz"), **out)


def golden_predhalf():
    """PSFNet.pred / render with the MLP in fp16, i.e. the arithmetic of the reference's CUDA run (`@autocast()` on
    MLP.forward is a no-op on CPU, so the plain CPU run above is fp32).  The reference's own `pred` and `render` lines run
    unmodified; only `lens.psfnet` is wrapped so that the seeded MLP computes in half precision on CPU."""
    torch.manual_seed(5)
    lens = dl.PSFNet(LENSES["rf50mm"], sensor_res=(16, 24), kernel_size=21, device="cpu")
    mlp32 = lens.psfnet
    mlp16 = mlp32.half()                                   # in-place conversion of the seeded weights

    class HalfMLP(torch.nn.Module):
        def forward(self, inp):
            return mlp16(inp.half())
    half_mlp = HalfMLP()
    lens.psfnet = half_mlp
    torch.manual_seed(6)
    img = torch.rand(2, 3, 16, 24)
    depth = -(torch.rand(2, 1, 16, 24) * 9750 + 250)
    foc = torch.tensor([-1000.0, -1000.0])
    out = {"img": img.numpy(), "depth": depth.numpy(), "foc": foc.numpy()}
    out["render_out"] = lens.render(img, depth, foc).float().numpy()
    x, y = torch.meshgrid(torch.linspace(-1, 1, 24), torch.linspace(1, -1, 16), indexing="xy")
    z = lens.depth2z(depth + lens.d_sensor).squeeze(1)
    o = torch.stack((x[None].repeat(2, 1, 1), y[None].repeat(2, 1, 1), z), -1).float()
    out["z"] = z.numpy()
    raw_l = half_mlp(o.clone())
    on = o.clone()
    on[..., 0] = -on[..., 0]
    raw_r = half_mlp(on)
    out["raw_l"], out["raw_r"] = raw_l.detach().numpy(), raw_r.detach().numpy()                       # [2,16,24,21,21] fp16, right unflipped
    out["psf"] = lens.pred(o.clone()).detach().numpy()                                        # [2,16,24,2,21,21] fp16
    h1 = torch.relu(torch.nn.functional.linear(o.half(), mlp16.net[0].weight, mlp16.net[0].bias))
    out["h1_l"] = h1.detach().numpy()                                                         # first activation, left rows
    np.savez_compressed(os.path.join(HERE, "predhalf.npz"), **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["setup", "trace", "dp", "psf", "psf2m", "psf2m_sweep", "rf35mm2m", "render", "predhalf", "api"]
    for w in which:
        globals()[f"golden_{w}"]()
        print("wrote", w)

"""GPU: the reference-facing Python API (sdirt_b200.deeplens) against vectors produced by the unmodified reference.
These read like calls into the reference: same class and method names, same arguments."""
import numpy as np
import pytest
import torch

from conftest import lens_path
from test_oracle_golden import l1_sumnorm

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def lenses():
    from sdirt_b200.deeplens import PSFNet
    return {n: PSFNet(lens_path(n), sensor_res=(512, 768), kernel_size=21, device=DEV) for n in ("rf50mm", "rf35mm")}


def test_lens_setup_matches_reference(golden, lenses):
    g = golden("setup")
    for name, lens in lenses.items():
        sc = g[f"{name}_scalars"]         # aper_idx, hfov, foclen, fnum, pupil z/r, exit pupil z/r, pixel, d_sensor, r_last
        assert lens.aper_idx == int(sc[0])
        assert abs(lens.hfov - sc[1]) < 2e-6
        assert abs(lens.foclen - sc[2]) < 2e-4
        pz, pr = lens.entrance_pupil()
        # torch.linalg.lstsq (float32, nearly parallel lines) answers differently on CUDA and on the CPU the goldens
        # were made on; the product calls it on the lens device exactly like the reference would
        assert abs(pz - sc[4]) < 2e-3 and abs(pr - sc[5]) / sc[5] < 2e-3
        assert lens.pixel_size == sc[8] and lens.d_sensor == sc[9] and abs(lens.r_last - sc[10]) < 1e-12


def test_refocus_matches_reference(golden):
    from sdirt_b200.deeplens import PSFNet
    g = golden("setup")
    for name in ("rf50mm", "rf35mm"):
        lens = PSFNet(lens_path(name), sensor_res=(512, 768), kernel_size=21, device=DEV)
        torch.manual_seed(0)
        lens.refocus(-1000 + lens.d_sensor)
        assert abs(lens.d_sensor - g[f"{name}_refocus"][0]) < 2e-4
        assert abs(lens.hfov - g[f"{name}_refocus"][1]) < 2e-6


def _pin_pupil(lens, pz, pr):
    lens.entrance_pupil = lambda M=32, entrance=True, shrink_pupil=False: (float(pz), float(pr) * (0.25 if shrink_pupil else 1.0))


@pytest.mark.parametrize("numerics", [None, "hybrid"])
@pytest.mark.parametrize("name", ["rf50mm", "rf35mm"])
def test_psf_api_vs_reference(golden, name, numerics):
    """lens.psf / psf_diff with the reference's seed: same RNG stream, same rays, same PSFs (L1 <= 1e-4)."""
    from sdirt_b200.deeplens import PSFNet
    g = golden("psf")
    lens = PSFNet(lens_path(name), sensor_res=(512, 768), kernel_size=21, device=DEV)
    lens.numerics = numerics
    lens.hfov = float(g[f"{name}_hfov"])
    _pin_pupil(lens, *g[f"{name}_pupil"])
    pts = torch.from_numpy(g[f"{name}_points_norm"])
    spp = int(g[f"{name}_u_check"][4])
    torch.manual_seed(3)
    psf = lens.psf(pts, ks=21, spp=spp)
    assert psf.shape == (5, 21, 21) and psf.is_cuda
    assert l1_sumnorm(psf.cpu().numpy(), g[f"{name}_l"]).max() < 1e-4
    assert abs(float(psf.max()) - 1.0) < 1e-5                              # max-normalised like psf_diff
    torch.manual_seed(3)
    psf_r = lens.psf_diff(pts, ks=21, spp=spp, param_list=(0.78, 1.44, 0.3, 0.5, "r"))
    assert l1_sumnorm(psf_r.cpu().numpy(), g[f"{name}_r"]).max() < 1e-4
    torch.manual_seed(3)
    big = lens.psf_diff(pts, ks=21, spp=spp, param_list=(0.78, 1.44, 0.3, 0.6, "l"))
    assert l1_sumnorm(big.cpu().numpy(), g[f"{name}_big_l"]).max() < 1e-4
    torch.manual_seed(3)
    one = lens.psf(pts[0], ks=21, spp=spp)                                 # 1-D input -> [ks, ks]
    assert one.shape == (21, 21)
    # the chief-ray centre at the same RNG position
    torch.manual_seed(3)
    torch.rand(spp), torch.rand(spp)
    c = lens.psf_center(lens._object_points(pts))
    np.testing.assert_allclose(c.cpu().numpy(), g[f"{name}_centre"], rtol=3e-6, atol=1e-7)


def test_psf_rgb_api(golden, lenses):
    g = golden("psf")
    lens = lenses["rf50mm"]
    lens.hfov = float(g["rf50mm_hfov"])
    _pin_pupil(lens, *g["rf50mm_pupil"])
    torch.manual_seed(3)
    rgb = lens.psf_rgb(torch.from_numpy(g["rf50mm_points_norm"][:2]), ks=21, spp=4000)
    assert rgb.shape == (2, 3, 21, 21)
    # 4000 rays: one ray is 2.5e-4 of a PSF, so this is a smoke-level bound; the 200 k-ray tests carry the tolerance
    assert l1_sumnorm(rgb.cpu().numpy(), g["rf50mm_rgb"]).max() < 2e-3
    del lens.entrance_pupil


def test_trace_api_and_forward_integral(golden, lenses):
    """sample_from_points -> trace2sensor -> forward_integral, the reference's three-call sequence."""
    from sdirt_b200.deeplens import forward_integral
    g = golden("psf")
    name = "rf50mm"
    lens = lenses[name]
    _pin_pupil(lens, *g[f"{name}_pupil"])
    obj = torch.from_numpy(g[f"{name}_points_obj"])
    spp = int(g[f"{name}_u_check"][4])
    torch.manual_seed(3)
    ray = lens.sample_from_points(o=obj, spp=spp)
    assert ray.o.shape == (spp, 5, 3) and ray.ra.shape == (spp, 5)
    ray = lens.trace2sensor(ray)
    raw = forward_integral(ray, ps=lens.pixel_size, ks=21, pointc_ref=torch.from_numpy(g[f"{name}_centre"]))
    ref = g[f"{name}_chief_raw"]
    assert (np.abs(raw.cpu().numpy() - ref).sum((1, 2)) / ref.sum((1, 2))).max() < 1e-4
    rms = forward_integral(ray, ps=lens.pixel_size, ks=21, pointc_ref=None)
    assert l1_sumnorm(rms.cpu().numpy(), g[f"{name}_rms_raw"]).max() < 1e-4
    del lens.entrance_pupil


def test_ray_reaction_single_surface(golden, lenses):
    """Aspheric.ray_reaction surface by surface equals the reference's per-surface states."""
    from sdirt_b200.deeplens import Ray
    g = golden("trace")
    lens = lenses["rf50mm"]
    r0 = g["rf50mm_w589_ray0"]
    ray = Ray(torch.from_numpy(r0[..., :3].copy()), torch.from_numpy(r0[..., 3:6].copy()), wvln=0.589, device=DEV)
    assert np.abs(ray.d.cpu().numpy() - r0[..., 3:6]).max() < 1.5e-7        # Ray() re-normalises, like the reference
    st = g["rf50mm_w589_states"]
    for i, s in enumerate(lens.surfaces):
        ray = s.ray_reaction(ray)
        assert np.array_equal(ray.ra.cpu().numpy(), st[i][..., 6])
        # Ray() re-normalised d (1 ulp), which the 2..20 m lever arm to the first surface turns into <= 3e-3 mm
        assert np.abs(ray.o.cpu().numpy() - st[i][..., :3]).max() < 3e-3
        # ... and, through the surface normal at the shifted hit (curvature up to 0.04 /mm), into <= 1e-4 of direction
        assert np.abs(ray.d.cpu().numpy() - st[i][..., 3:6]).max() < 1e-4
    ray = ray.propagate_to(lens.d_sensor)
    assert np.abs(ray.o.cpu().numpy() - g["rf50mm_w589_sensor"][..., :3]).max() < 3e-3


def test_render_api_vs_reference(golden):
    """PSFNet.render with the reference's seeded random-init MLP (the real checkpoint is not in the reference tree)."""
    from sdirt_b200.deeplens import PSFNet, local_psf_render_fast
    g = golden("render")
    torch.manual_seed(5)
    lens = PSFNet(lens_path("rf50mm"), sensor_res=(16, 24), kernel_size=21, device=DEV)
    got = np.asarray([[v.double().sum().item(), v.double().abs().sum().item()] for v in lens.psfnet.state_dict().values()])
    np.testing.assert_allclose(got, g["mlp_checksum"], rtol=1e-6)
    img = torch.from_numpy(g["render_img"]).to(DEV)
    depth = torch.from_numpy(g["render_depth"]).to(DEV)
    foc = torch.from_numpy(g["render_foc"]).to(DEV)
    out = lens.render(img, depth, foc)
    assert out.shape == (2, 6, 16, 24)
    # the MLP runs under CUDA autocast (fp16 GEMMs) here and in fp32 on the reference's CPU run: PSFs agree to
    # ~1e-3 relative, the rendered image (a 441-tap average) to a few fp16 ulps
    np.testing.assert_allclose(out.cpu().numpy(), g["render_out"], atol=4e-3)
    for ks in (7, 21):
        rl, rr = local_psf_render_fast(torch.from_numpy(g[f"ks{ks}_img"]).to(DEV), torch.from_numpy(g[f"ks{ks}_psf"]).to(DEV).float(), ks)
        np.testing.assert_allclose(rl.cpu().numpy(), g[f"ks{ks}_rl"], rtol=1.1e-3, atol=1e-6)
        np.testing.assert_allclose(rr.cpu().numpy(), g[f"ks{ks}_rr"], rtol=1.1e-3, atol=1e-6)


def test_tone_fused_matches_reference_curves(golden):
    from sdirt_b200 import _engine as E
    g = golden("render")
    x = torch.from_numpy(g["tone_x"]).to(DEV)
    img = x.reshape(1, 1, 7, 143).contiguous()
    ident = torch.zeros(1, 7, 143, 2, 1, 1, device=DEV)
    ident[...] = 1.0
    rl, _ = E.render_local_psf(img, ident, 1, tone=1)                     # degamma only, 1x1 identity PSF
    np.testing.assert_allclose(rl.reshape(-1).cpu().numpy(), g["tone_degamma"], rtol=1.1e-3, atol=1e-3)   # fp16 round trip
    rl, _ = E.render_local_psf(img, ident, 1, tone=3)
    np.testing.assert_allclose(rl.reshape(-1).cpu().numpy(), np.clip(g["tone_gamma"], 0, 1), atol=3e-3)


def test_smoke_entry():
    import __graft_entry__ as ge
    ge.smoke()


def test_psfnet_fitting_loop(tmp_path, lenses):
    """The PSF-bank workload generators and a few steps of PSFNet.train_psfnet (psfnet.py:101-241): ray-traced targets from
    the engine, MLP step in torch, no host round trip of the PSFs."""
    lens = lenses["rf50mm"]
    torch.manual_seed(0)
    np.random.seed(0)
    inp, psf = lens.get_training_data(bs=16, spp=4000)
    assert inp.shape == (16, 3) and psf.shape == (16, 21, 21) and psf.is_cuda
    assert float(psf.amax((1, 2)).min()) > 0.99 and torch.isfinite(psf).all()       # every point produced a max-normalised PSF
    assert float(inp[:, 2].min()) >= 0.0 and float(inp[:, 2].max()) <= 1.0
    tin, tpsf = lens.get_test_data(bs=1024, spp=512)
    assert tin.shape == (1024, 3) and tpsf.shape == (1024, 21, 21)
    lens.numerics = "adaptive"
    try:
        hist = lens.train_psfnet(iters=4, bs=16, spp=2000, evaluate_every=4, result_dir=str(tmp_path))
    finally:
        lens.numerics = None
    assert len(hist) == 5 and all(np.isfinite(hist))
    assert (tmp_path / "iter4_PSFNet_mlp.pkl").exists()
    # the CUDA-graph step (forward + loss + backward captured once) follows the eager step
    runs = []
    for use_graph in (False, True):
        torch.manual_seed(7)
        np.random.seed(7)
        fresh = type(lens)(lens_path("rf50mm"), sensor_res=(512, 768), kernel_size=21, device=DEV)
        fresh.numerics = "adaptive"
        runs.append(fresh.train_psfnet(iters=5, bs=16, spp=2000, evaluate_every=10 ** 9, result_dir=str(tmp_path), graph=use_graph))
    np.testing.assert_allclose(runs[0], runs[1], rtol=2e-2)
    # everything on the device: batch coordinates and pupil samples from the CUDA generator, the evaluation bank traced once
    fresh.fit_device_data = True
    torch.cuda.manual_seed(3)
    hist = fresh.train_psfnet(iters=5, bs=16, spp=2000, evaluate_every=3, result_dir=str(tmp_path))
    assert len(hist) == 6 and all(np.isfinite(hist)) and len(fresh.eval_history) == 2
    bank = fresh._test_bank
    assert bank[1].is_cuda and bank[2].shape == (1024, 21, 21) and float(bank[2].amax((1, 2)).min()) > 0.99
    a, b = fresh._training_data_device(64, 4000)
    assert a.is_cuda and b.is_cuda and float(a[:, 2].min()) >= 0 and float(a[:, 2].max()) <= 1 and float(b.amax((1, 2)).min()) > 0.99


def test_render_banded_vs_reference_half(golden):
    """PSFNet.render (banded: engine kernels around the cuBLAS GEMM chain) against the reference's own render() run with
    its MLP in fp16 (tests/golden/predhalf.npz, i.e. the arithmetic of the reference's CUDA path), against the oracle's
    restatement, and against the unbanded route through `pred`."""
    from sdirt_b200.deeplens import PSFNet
    from oracle import dp_oracle as O
    from test_oracle_golden import seeded_mlp_weights
    g = golden("predhalf")
    torch.manual_seed(5)
    lens = PSFNet(lens_path("rf50mm"), sensor_res=(16, 24), kernel_size=21, device=DEV)
    img, depth, foc = (torch.from_numpy(g[k]).to(DEV) for k in ("img", "depth", "foc"))
    out = lens.render(img, depth, foc)
    assert out.shape == (2, 6, 16, 24)
    ref = g["render_out"]
    assert np.abs(out.cpu().numpy() - ref).max() < 2e-3                  # a few fp16 ulps of a [0, 1] image
    want = O.psfnet_render_half(seeded_mlp_weights(), g["img"], g["z"], 21)
    assert np.abs(out.cpu().numpy() - want).max() < 2e-3
    # bands of 5 rows, one image at a time == one band
    lens.render_band_rows, lens.render_band_pixels = 5, 1
    try:
        out5 = lens.render(img, depth, foc)
    finally:
        del lens.render_band_rows, lens.render_band_pixels
    assert torch.equal(out5, out)
    via = lens.render_via_pred(img, depth, foc)
    assert (out - via).abs().max().item() < 2e-3


def test_render_banded_large_and_train():
    """A 64 x 96 image batch through several bands and band batches; train=True adds noise between gamma and clip."""
    from sdirt_b200.deeplens import PSFNet
    torch.manual_seed(2)
    lens = PSFNet(lens_path("rf35mm"), sensor_res=(64, 96), kernel_size=11, device=DEV)
    gen = torch.Generator(device=DEV).manual_seed(1)
    img = torch.rand((3, 3, 64, 96), device=DEV, generator=gen)
    depth = -(torch.rand((3, 1, 64, 96), device=DEV, generator=gen) * 9000 + 300)
    foc = torch.full((3,), -1000.0, device=DEV)
    lens.render_band_pixels = 2 * 16 * 96
    out = lens.render(img, depth, foc)
    via = lens.render_via_pred(img, depth, foc)
    assert out.shape == (3, 6, 64, 96) and torch.isfinite(out).all()
    assert (out - via).abs().max().item() < 2e-3
    np.random.seed(0)
    tr = lens.render(img, depth, foc, train=True)
    assert tr.shape == out.shape and float(tr.min()) >= 0.0 and float(tr.max()) <= 1.0
    assert (tr - out).abs().mean().item() > 1e-4                          # noise was added


def test_render_train_tail_and_focal_stack():
    """render(train=True): the fused gamma + noise + clip kernel against the torch ops of the reference's noise()/gamma()
    (same seeds, same draw order), and render_focal_stack against the per-image loop of 2_dfdp_net.py:166-171."""
    from sdirt_b200.deeplens import PSFNet
    torch.manual_seed(3)
    lens = PSFNet(lens_path("rf50mm"), sensor_res=(64, 96), kernel_size=11, device=DEV)
    gen = torch.Generator(device=DEV).manual_seed(9)
    img = torch.rand((3, 3, 64, 96), device=DEV, generator=gen)
    depth = -(torch.rand((3, 1, 64, 96), device=DEV, generator=gen) * 5000 + 300)
    foc = torch.full((3,), -1000.0, device=DEV)

    def seeded(fn):
        np.random.seed(4)
        torch.manual_seed(4)
        return fn()
    a = seeded(lambda: lens.render(img, depth, foc, train=True))
    b = seeded(lambda: lens.render_via_pred(img, depth, foc, train=True))
    assert a.shape == (3, 6, 64, 96) and float(a.min()) >= 0 and float(a.max()) <= 1
    assert (a - b).abs().max().item() < 2e-3
    clean = lens.render(img, depth, foc)
    assert 1e-4 < (a - clean).abs().mean().item() < 0.05                 # sigma <= 0.05 x ramp <= 1
    stack = seeded(lambda: lens.render_focal_stack(img, depth, foc, train=True))
    loop = seeded(lambda: torch.cat([lens.render(img[i:i + 1], depth[i:i + 1], foc[i:i + 1], train=True) for i in range(3)], 0))
    assert (stack - loop).abs().max().item() < 2e-3
    assert torch.equal(lens.render_focal_stack(img, depth, foc, train=False), clean)


def test_render_fused_mlp_engine(golden):
    """PSFNet.render with mlp_engine = "fused" (pred as one tcgen05 kernel per band) against the reference's fp16 golden and
    against the cuBLAS route."""
    from sdirt_b200.deeplens import PSFNet
    g = golden("predhalf")
    torch.manual_seed(5)
    lens = PSFNet(lens_path("rf50mm"), sensor_res=(16, 24), kernel_size=21, device=DEV)
    img, depth, foc = (torch.from_numpy(g[k]).to(DEV) for k in ("img", "depth", "foc"))
    lens.mlp_engine = "cublas"
    base = lens.render(img, depth, foc)
    lens.mlp_engine = "fused"                                               # the default
    out = lens.render(img, depth, foc)
    assert np.abs(out.cpu().numpy() - g["render_out"]).max() < 2e-3
    assert (out - base).abs().max().item() < 1e-3
    lens.render_band_rows, lens.render_band_pixels = 8, 1                   # several bands, one image at a time
    assert torch.equal(lens.render(img, depth, foc), out)
    from sdirt_b200 import _engine as E
    try:                                                                    # single CTAs instead of CTA pairs: same bits
        E.lib().sdirt_mlp_fused_cta_group(1)
        assert torch.equal(lens.render(img, depth, foc), out)
    finally:
        E.lib().sdirt_mlp_fused_cta_group(2)

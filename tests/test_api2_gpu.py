"""GPU: the API rows VERDICT r1 marked partial -- the other two render variants, psf_map, the fitting-data generators, the
RMS / magnification analyses, the 1_fit_psfnet.py call sequence, a PSFNet the fused kernel is not compiled for, and a lens on
a device that is not the current one -- against vectors produced by the unmodified reference (tests/golden/api.npz,
make_golden.py `api`)."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

from conftest import lens_path
from test_oracle_golden import l1_sumnorm

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _pinned_lens(golden, name, res=(512, 768), ks=21, device=DEV):
    """PSFNet with the reference's CPU values of hfov and entrance pupil (torch.linalg.lstsq answers in the last digits
    differently on CUDA, see test_api_gpu.test_lens_setup_matches_reference)."""
    from sdirt_b200.deeplens import PSFNet
    sc = golden("setup")[f"{name}_scalars"]
    lens = PSFNet(lens_path(name), sensor_res=res, kernel_size=ks, device=device)
    lens.hfov = float(sc[1])
    pz, pr = float(sc[4]), float(sc[5])
    lens.entrance_pupil = lambda M=32, entrance=True, shrink_pupil=False: (pz, pr * (0.25 if shrink_pupil else 1.0))
    return lens


def half_ulps(a, b):
    a16, b16 = a.astype(np.float16), b.astype(np.float16)
    return np.abs(a16.view(np.int16).astype(np.int32) - b16.view(np.int16).astype(np.int32))


@pytest.mark.parametrize("ks", [7, 11])
def test_render_variants_vs_reference(golden, ks):
    """local_psf_render returns the (left, right) pair in the half() arithmetic of local_psf_render_fast (render_psf.py:76-118);
    local_dp_psf_render returns cat(left, right) computed in float32 (render_psf.py:157-188)."""
    from sdirt_b200.deeplens import local_dp_psf_render, local_psf_render
    g = golden("api")
    img, psf = torch.from_numpy(g[f"rv{ks}_img"]).to(DEV), torch.from_numpy(g[f"rv{ks}_psf"]).to(DEV)
    out = local_psf_render(img, psf, ks)
    assert isinstance(out, tuple) and len(out) == 2 and out[0].shape == img.shape and out[0].dtype == img.dtype
    assert half_ulps(out[0].cpu().numpy(), g[f"rv{ks}_rl"]).max() <= 1
    assert half_ulps(out[1].cpu().numpy(), g[f"rv{ks}_rr"]).max() <= 1
    dp = local_dp_psf_render(img, psf, ks)
    assert dp.shape == (img.shape[0], 6) + img.shape[2:] and dp.dtype == torch.float32
    np.testing.assert_allclose(dp.cpu().numpy(), g[f"rv{ks}_dp"], rtol=2e-6, atol=1e-7)
    # float32 arithmetic is visibly NOT the half path: the two differ by half-precision rounding
    both = torch.cat(out, 1)
    assert 1e-5 < float((both - dp).abs().max()) < 2e-3
    # both inputs in half: the promoted dtype is half, i.e. the fast path's arithmetic
    dph = local_dp_psf_render(img.half(), psf.half(), ks)
    assert dph.dtype == torch.float16 and half_ulps(dph[:, :3].float().cpu().numpy(), g[f"rv{ks}_rl"]).max() <= 1


@pytest.mark.parametrize("name", ["rf50mm", "rf35mm"])
def test_psf_map_vs_reference(golden, name):
    """Lensgroup.psf_map (optics.py:1018-1041): a 3 x 3 grid of RGB PSFs as one [3, 33, 33] mosaic, same seed, same rays."""
    g = golden("api")
    lens = _pinned_lens(golden, name)
    torch.manual_seed(4)
    m = lens.psf_map(depth=-1500.0 + lens.d_sensor, grid=3, ks=11, spp=20000)
    want = g[f"{name}_psf_map"]
    assert tuple(m.shape) == want.shape == (3, 33, 33)
    got = m.cpu().numpy()
    tiles = lambda a: a.reshape(3, 3, 11, 3, 11).transpose(0, 1, 3, 2, 4).reshape(27, 11, 11)
    l1 = l1_sumnorm(tiles(got), tiles(want))
    print(name, "psf_map per-tile L1:", np.array2string(l1, precision=1, max_line_width=200))
    assert l1.max() < 3e-4
    assert np.abs(got.reshape(3, 3, 11, 3, 11).max((2, 4)) - 1).max() < 1e-5             # every tile max-normalised


@pytest.mark.parametrize("name", ["rf50mm", "rf35mm"])
def test_fit_data_vs_reference(golden, name):
    """get_training_data / get_test_data (psfnet.py:170-241) with the reference's seeds: the network inputs are the same
    numbers (same host RNG calls in the same order) and the traced PSFs the same PSFs."""
    g = golden("api")
    lens = _pinned_lens(golden, name)
    np.random.seed(8)
    torch.manual_seed(8)
    inp, psf = lens.get_training_data(bs=8, spp=20000)
    np.testing.assert_array_equal(inp.cpu().numpy(), g[f"{name}_train_inp"])
    l1 = l1_sumnorm(psf.cpu().numpy(), g[f"{name}_train_psf"])
    print(name, "get_training_data L1:", np.array2string(l1, precision=1))
    assert l1.max() < 3e-4
    torch.manual_seed(9)
    inp, psf = lens.get_test_data(bs=1024, spp=2048)
    np.testing.assert_array_equal(inp.cpu().numpy(), g[f"{name}_test_inp"])
    l1 = l1_sumnorm(psf.cpu().numpy()[::16], g[f"{name}_test_psf"])
    print(name, "get_test_data L1 (64 of 1024 PSFs, 2048 rays): max", l1.max(), "mean", l1.mean())
    assert l1.max() < 1.5e-3 and l1.mean() < 2e-4                  # a ray on the other side of a pixel edge is 5e-4 of a 2048-ray PSF


@pytest.mark.parametrize("name", ["rf50mm", "rf35mm"])
def test_analysis_vs_reference(golden, name):
    """calc_magnification3 (optics.py:1237-1272) and analysis_rms (optics.py:2103-2140) through the engine's trace, seeded."""
    g = golden("api")
    lens = _pinned_lens(golden, name)
    ds = lens.d_sensor
    torch.manual_seed(10)
    mag = [lens.calc_magnification3(-1000.0 + ds), lens.calc_magnification3(-20000.0 + ds)]
    np.testing.assert_allclose(mag, g[f"{name}_mag3"], rtol=2e-5)
    torch.manual_seed(11)
    rms = [float(v) for v in lens.analysis_rms(depth=-1000.0 + ds)] + [float(v) for v in lens.analysis_rms(depth=-5000.0 + ds)]
    print(name, "rms (avg, on, off) x 2 depths:", rms, "reference:", g[f"{name}_rms"])
    np.testing.assert_allclose(rms, g[f"{name}_rms"], rtol=2e-4)


def test_fit_script_sequence(tmp_path, monkeypatch, capsys):
    """The body of the reference's 1_fit_psfnet.py (:8-10, :16-38), statement by statement, with `deeplens` resolving to the
    mirror: construct, refocus, write the lens file, analyse two depths, load a checkpoint, fit, compare."""
    import sdirt_b200.deeplens as mirror
    monkeypatch.setitem(sys.modules, "deeplens", mirror)
    monkeypatch.setitem(sys.modules, "deeplens.psfnet", importlib.import_module("sdirt_b200.deeplens.psfnet"))
    monkeypatch.setitem(sys.modules, "deeplens.utils", importlib.import_module("sdirt_b200.deeplens.utils"))
    import logging
    handlers = list(logging.getLogger().handlers)
    from deeplens.psfnet import PSFNet
    from deeplens.utils import set_logger, set_seed
    result_dir = str(tmp_path)
    try:
        set_logger(result_dir)
        set_seed(0)
        ks = 21
        psfnet = PSFNet(filename=lens_path("rf50mm"), sensor_res=(512, 768), kernel_size=ks, device="cuda")
        d_sensor = psfnet.d_sensor
        infocus = -1000 + d_sensor
        psfnet.refocus(infocus)
        psfnet.write_lens_json(f"{result_dir}/lens.json")
        # golden("setup") rf50mm_refocus has the reference's least-squares sensor position for ITS 2048 samples (62.2538); here the
        # generator has already initialised the MLP, so the samples differ: agreement to the estimator's sampling noise
        assert abs(psfnet.d_sensor - 62.25384521) < 1e-2
        near_depth = -500 + d_sensor
        r_near = psfnet.analysis(save_name=f"{result_dir}/{int(near_depth)}", depth=near_depth, ks=ks)
        far_depth = -20000 + d_sensor
        r_far = psfnet.analysis(save_name=f"{result_dir}/{int(far_depth)}", depth=far_depth, ks=ks)
        assert "On-axis RMS radius" in capsys.readouterr().out
        assert all(torch.isfinite(v) and v > 0 for v in r_near + r_far)
        torch.save(psfnet.psfnet.state_dict(), f"{result_dir}/F4_PSFNet_mlp.pkl")       # (the published checkpoint is not in the tree)
        psfnet.load_net(f"{result_dir}/F4_PSFNet_mlp.pkl")
        losses = psfnet.train_psfnet(iters=3, bs=8, lr=1e-4, spp=2000, evaluate_every=2, result_dir=result_dir)
        assert len(losses) == 4 and all(np.isfinite(losses))
        assert os.path.exists(f"{result_dir}/PSFNet_mlp.pkl") and os.path.exists(f"{result_dir}/iter2_PSFNet_mlp.pkl")
        assert len(psfnet.eval_history) == 2 and all(np.isfinite(v) for row in psfnet.eval_history for v in row)
        cmp = psfnet.compare_psf()
        assert sorted(cmp) == [-20000, -500]
        for traced, pred in cmp.values():
            assert traced.shape == pred.shape == (3, 2, ks, ks) and torch.isfinite(traced).all() and torch.isfinite(pred).all()
            # on axis the right PSF is the mirror image of the left one (different samples: to sampling noise)
            assert l1_sumnorm(traced[0:1, 0].numpy(), torch.flip(traced[0:1, 1], dims=[-1]).numpy())[0] < 0.05
    finally:
        for h in list(logging.getLogger().handlers):
            if h not in handlers:
                logging.getLogger().removeHandler(h)
                h.close()


def test_render_kernel_size_without_fused_kernel():
    """a PSFNet whose kernel size the fused tensor-core kernel is not compiled for (9 here; the reference documents 35
    for f/1.8) renders through the library-GEMM route and the generic kernels -- same pixels as pred() followed by the explicit
    convolution."""
    from sdirt_b200.deeplens import PSFNet
    torch.manual_seed(2)
    lens = PSFNet(lens_path("rf50mm"), sensor_res=(32, 48), kernel_size=9, device=DEV)
    assert lens.mlp_engine == "fused" and lens._mlp_fused() is None
    g = torch.Generator(device=DEV).manual_seed(1)
    img = torch.rand((2, 3, 32, 48), device=DEV, generator=g)
    depth = -(torch.rand((2, 1, 32, 48), device=DEV, generator=g) * 9000 + 300)
    foc = torch.full((2,), -1000.0, device=DEV)
    out = lens.render(img, depth, foc)
    want = lens.render_via_pred(img, depth, foc)
    assert out.shape == (2, 6, 32, 48) and torch.isfinite(out).all()
    assert float((out - want).abs().max()) < 2e-3


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_lens_on_a_device_that_is_not_current(golden):
    """tensors on cuda:1 while cuda:0 is the current device -- every binding switches to the tensors' device for the
    duration of the call (the library launches on the current device)."""
    from sdirt_b200.deeplens import PSFNet
    assert torch.cuda.current_device() == 0
    pts = torch.tensor([[0.0, 0.0, -2000.0], [0.4, 0.3, -800.0]])
    res = []
    for dev in ("cuda:0", "cuda:1"):
        lens = PSFNet(lens_path("rf50mm"), sensor_res=(64, 96), kernel_size=11, device=dev)
        torch.manual_seed(5)
        L, R = lens.psf_dp(pts, ks=11, spp=20000)
        g = torch.Generator(device=dev).manual_seed(1)
        img = torch.rand((1, 3, 64, 96), device=dev, generator=g)
        out = lens.render(img, -(img[:, :1] * 5000 + 500), torch.full((1,), -1000.0, device=dev))
        assert L.device == torch.device(dev) and out.device == torch.device(dev) and torch.isfinite(out).all()
        res.append((L.cpu(), R.cpu()))
    assert torch.cuda.current_device() == 0
    assert l1_sumnorm(res[0][0].numpy(), res[1][0].numpy()).max() < 1e-5

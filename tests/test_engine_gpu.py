"""GPU parity: the CUDA engine (through the C ABI) against the numpy oracle and the reference goldens.

Tolerances are BASELINE.json's: hit coordinates <= 1e-5 relative, pixel assignment identical for >= 99.99 %
of rays, per-PSF L1 <= 1e-4 after sum-normalisation.  Where the engine and the oracle state the same IEEE
arithmetic (trace with a replayed or per-ray Newton schedule) the comparison is bit-for-bit.
"""
import numpy as np
import pytest
import torch

from conftest import lens_path
from oracle import dp_oracle as O
from test_oracle_golden import D_SENSOR, l1_sumnorm, make_lens, psf_golden_samples, torch_pupil

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def engine_lens(name):
    from sdirt_b200 import _engine as E
    from sdirt_b200.prescription import load_lens_json
    recs, descs, head = load_lens_json(lens_path(name))
    return E.LensHandle(recs, D_SENSOR[name])


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(DEV)


def trace_inputs(g, tag):
    name = tag.split("_")[0]
    pz, pr = g[f"{tag}_pupil"]
    px, py = torch_pupil(g[f"{tag}_u"], pr)
    return name, g[f"{name}_points_obj"], px, py, float(pz)


@pytest.mark.parametrize("tag", ["rf50mm_w589", "rf50mm_w486", "rf35mm_w589"])
@pytest.mark.parametrize("mode", ["replay", "per_ray"])
def test_trace_bitexact_vs_oracle(golden, tag, mode):
    from sdirt_b200 import _engine as E
    g = golden("trace")
    name, obj, px, py, pz = trace_inputs(g, tag)
    wv = int(tag[-3:]) / 1000
    lens = make_lens(name)
    ray = O.rays_from_points(obj, px, py, pz, wv)
    sched = "per_ray" if mode == "per_ray" else [int(v) for v in g[f"{tag}_newton"]]
    rec = []
    O.trace_to_sensor(lens, ray, record=rec, newton_iters=sched)
    r0 = g[f"{tag}_ray0"].reshape(-1, 6)
    o, d = cu(r0[:, :3]), cu(r0[:, 3:6])
    ra = torch.ones(o.shape[0], device=DEV)
    h = engine_lens(name)
    got = E.trace_rays(h, wv, o, d, ra, to_sensor=True, newton=sched, record=True).cpu().numpy()
    for i, r in enumerate(rec):
        want = np.concatenate([r.o(), r.d(), r.ra[..., None]], -1).reshape(-1, 7)
        assert np.array_equal(got[i][:, 6], want[:, 6]), f"validity differs at surface {i}"
        same = (got[i] == want).all(-1).mean()
        assert same == 1.0, f"surface {i}: only {same:.5f} of rays bit-identical, max diff {np.abs(got[i] - want).max():.3e}"
    assert np.array_equal(o.cpu().numpy(), ray.o().reshape(-1, 3))
    assert np.array_equal(ra.cpu().numpy(), ray.ra.reshape(-1))


@pytest.mark.parametrize("tag", ["rf50mm_w589", "rf35mm_w589"])
def test_trace_vs_reference_golden(golden, tag):
    """Engine (per-ray Newton, its production mode) against the reference's own per-surface states."""
    from sdirt_b200 import _engine as E
    g = golden("trace")
    name = tag.split("_")[0]
    lens = make_lens(name)
    r0 = g[f"{tag}_ray0"].reshape(-1, 6)
    o, d = cu(r0[:, :3]), cu(r0[:, 3:6])
    ra = torch.ones(o.shape[0], device=DEV)
    got = E.trace_rays(engine_lens(name), 0.589, o, d, ra, to_sensor=True, newton="per_ray", record=True).cpu().numpy()
    st = g[f"{tag}_states"]
    for i in range(st.shape[0]):
        ref = st[i].reshape(-1, 7)
        assert np.array_equal(got[i][:, 6], ref[:, 6])
        scale = np.maximum(np.linalg.norm(ref[:, :3], axis=-1), lens.surfaces[i].r)[:, None]
        assert (np.abs(got[i][:, :3] - ref[:, :3]) / scale).max() < 1e-5
        assert np.abs(got[i][:, 3:6] - ref[:, 3:6]).max() < 2e-6
    ref = g[f"{tag}_sensor"].reshape(-1, 7)
    assert np.abs(o.cpu().numpy() - ref[:, :3]).max() < 3e-5
    # pixel assignment on the sensor, 21x21 window around the reference's chief-ray centre
    rref = O.RayBundle(*(g[f"{tag}_sensor"][..., i].copy() for i in range(7)))
    centre = O.chief_ray_centre(rref)
    mine = O.RayBundle(*(np.concatenate([o.cpu().numpy(), d.cpu().numpy(), ra.cpu().numpy()[:, None]], -1)
                         .reshape(g[f"{tag}_sensor"].shape)[..., i].copy() for i in range(7)))
    idx = []
    for r in (mine, rref):
        qx, qy, w = O.crop_and_shift(r, centre, 21, lens.pixel_size)
        r0_, c0_, _, _, _, _ = O.splat_indices(qx, qy, 21, lens.pixel_size)
        idx.append((r0_, c0_, w))
    agree = (idx[0][0] == idx[1][0]) & (idx[0][1] == idx[1][1]) & (idx[0][2] == idx[1][2])
    assert agree.mean() >= 0.9999


def test_backward_subrange(golden):
    from sdirt_b200 import _engine as E
    g = golden("trace")
    lens = make_lens("rf50mm")
    r0 = g["back_ray0"]
    o, d = cu(r0[:, :3]), cu(r0[:, 3:6])
    ra = torch.ones(16, device=DEV)
    E.trace_rays(engine_lens("rf50mm"), 0.589, o, d, ra, s_begin=0, s_end=lens.aper_idx, backward=True, newton="per_ray")
    ref = g["back_final"]
    assert np.array_equal(ra.cpu().numpy(), ref[:, 6])
    np.testing.assert_allclose(o.cpu().numpy(), ref[:, :3], atol=2e-6)
    np.testing.assert_allclose(d.cpu().numpy(), ref[:, 3:6], atol=3e-7)
    ray = O.RayBundle.from_od(r0[:, :3], r0[:, 3:6], normalize=False)
    O.trace(lens, ray, range(0, lens.aper_idx), newton_iters="per_ray")
    assert np.array_equal(o.cpu().numpy(), ray.o()) and np.array_equal(d.cpu().numpy(), ray.d())


@pytest.mark.parametrize("name", ["rf50mm", "rf35mm"])
def test_psf_bank_vs_reference_golden(golden, name):
    from sdirt_b200 import _engine as E
    g = golden("psf")
    lens = make_lens(name, g[f"{name}_hfov"])
    obj = g[f"{name}_points_obj"]
    pz, pr = g[f"{name}_pupil"]
    (px, py), (cx, cy) = psf_golden_samples(g, name)
    h = engine_lens(name)
    pts, pup, cpup = cu(obj), cu(np.stack([px, py], -1)), cu(np.stack([cx, cy], -1))
    centre = E.psf_centre(h, 0.589, pts, cpup, float(pz))
    np.testing.assert_allclose(centre.cpu().numpy(), g[f"{name}_centre"], rtol=3e-6, atol=1e-8)
    gc = cu(g[f"{name}_centre"])
    L, R, cnt = E.psf_bank(h, 0.589, pts, pup, float(pz), gc, 21, lens.pixel_size, want_counts=True)
    L, R = L.cpu().numpy(), R.cpu().numpy()
    assert l1_sumnorm(L, g[f"{name}_l"]).max() < 1e-4
    assert l1_sumnorm(R, g[f"{name}_r"]).max() < 1e-4
    np.testing.assert_allclose(L, g[f"{name}_l"], atol=2e-4)
    assert l1_sumnorm(L, g[f"{name}_r"]).max() > 0.05               # an L/R swap cannot pass
    assert (cnt.cpu().numpy() > 0.5 * px.shape[0]).all()
    # end to end with the engine's own chief-ray centre
    L2, _ = E.psf_bank(h, 0.589, pts, pup, float(pz), centre, 21, lens.pixel_size)
    assert l1_sumnorm(L2.cpu().numpy(), g[f"{name}_l"]).max() < 1e-4
    # big-radius micro-lens, ks = 11, raw sums
    big = (0.78, 1.44, 0.3, 0.6)
    Lb, Rb = E.psf_bank(h, 0.589, pts, pup, float(pz), gc, 21, lens.pixel_size, dp=big)
    assert l1_sumnorm(Lb.cpu().numpy(), g[f"{name}_big_l"]).max() < 1e-4
    assert l1_sumnorm(Rb.cpu().numpy(), g[f"{name}_big_r"]).max() < 1e-4
    L11, _ = E.psf_bank(h, 0.589, pts, pup, float(pz), gc, 11, lens.pixel_size)
    assert l1_sumnorm(L11.cpu().numpy(), g[f"{name}_ks11_l"]).max() < 1e-4
    Lraw, _ = E.psf_bank(h, 0.589, pts, pup, float(pz), gc, 21, lens.pixel_size, normalise=0)
    np.testing.assert_allclose(Lraw.cpu().numpy(), g[f"{name}_chief_raw"], rtol=2e-4, atol=2e-2)


def test_psf_bank_vs_oracle_small_and_ragged():
    """Seeded inputs at oracle-sized scale, incl. ragged sample counts and a point that is fully vignetted."""
    from sdirt_b200 import _engine as E
    name = "rf50mm"
    lens = make_lens(name, 0.40959781408309937)
    h = engine_lens(name)
    rng = np.random.default_rng(5)
    ptsn = np.concatenate([rng.uniform(-1, 1, (7, 2)), rng.uniform(-6000, -400, (7, 1))], -1).astype(np.float32)
    obj = O.object_points(lens, ptsn)
    obj[6] = [4000.0, 0.0, -500.0]                                   # far outside the field: no ray survives
    for spp in (1, 255, 4097):
        u = rng.uniform(0, 1, (2, spp)).astype(np.float32)
        px, py = torch_pupil(u, 6.0193)
        centre = np.zeros((7, 2), np.float32)
        cray = O.rays_from_points(obj, *torch_pupil(rng.uniform(0, 1, (2, 512)).astype(np.float32), 6.0193 / 4), 22.5132)
        O.trace_to_sensor(lens, cray, newton_iters="per_ray")
        centre = O.chief_ray_centre(cray)
        Lo, Ro, _ = O.psf_bank(lens, obj, px, py, 22.5132, 21, centre=centre, params=O.DP_DEFAULT, normalise=False,
                               newton_iters="per_ray")
        L, R, cnt = E.psf_bank(h, 0.589, cu(obj), cu(np.stack([px, py], -1)), 22.5132, cu(centre), 21, lens.pixel_size,
                               normalise=0, want_counts=True)
        np.testing.assert_allclose(L.cpu().numpy(), Lo, rtol=1e-5, atol=2e-5)
        np.testing.assert_allclose(R.cpu().numpy(), Ro, rtol=1e-5, atol=2e-5)
        assert cnt[6].item() == 0 and float(L[6].abs().sum()) == 0.0
    # empty point list is a no-op
    L, R = E.psf_bank(h, 0.589, cu(np.zeros((0, 3))), cu(np.zeros((4, 2))), 22.5, cu(np.zeros((0, 2))), 21, lens.pixel_size)
    assert L.shape == (0, 21, 21)


def test_splat_existing_rays_vs_oracle(golden):
    from sdirt_b200 import _engine as E
    g = golden("trace")
    lens = make_lens("rf50mm")
    s = g["rf50mm_w589_sensor"]
    ray = O.RayBundle(*(s[..., i].copy() for i in range(7)))
    Lo, Ro = O.splat_points(ray, lens.pixel_size, 21, None, O.DP_DEFAULT)
    L, R = E.splat_rays(cu(s[..., :3]), cu(s[..., 3:6]), cu(s[..., 6]), None, 21, lens.pixel_size, dp=O.DP_DEFAULT[:4])
    np.testing.assert_allclose(L.cpu().numpy(), Lo, rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(R.cpu().numpy(), Ro, rtol=1e-5, atol=2e-5)


@pytest.mark.parametrize("ks", [7, 21])
@pytest.mark.parametrize("half", [False, True])
def test_render_local_psf(golden, ks, half):
    from sdirt_b200 import _engine as E
    g = golden("render")
    img = cu(g[f"ks{ks}_img"])
    psf = torch.from_numpy(g[f"ks{ks}_psf"]).to(DEV)
    psf = psf if half else psf.float()
    rl, rr = E.render_local_psf(img, psf.contiguous(), ks)
    # one fp16 ulp for the float32 summation order (the reference's own order is a torch implementation detail)
    np.testing.assert_allclose(rl.cpu().numpy(), g[f"ks{ks}_rl"], rtol=1.1e-3, atol=1e-6)
    np.testing.assert_allclose(rr.cpu().numpy(), g[f"ks{ks}_rr"], rtol=1.1e-3, atol=1e-6)
    ol, orr = O.render_local_psf(g[f"ks{ks}_img"], g[f"ks{ks}_psf"].astype(np.float32), ks)
    assert (rl.cpu().numpy() == ol).mean() > 0.95

"""CPU: pin the numpy oracle against vectors produced by the UNMODIFIED reference (tests/golden/*.npz).

Tolerances (BASELINE.json north_star): hit coordinates within 1e-5 relative, sensor-pixel assignment
identical for >= 99.99 % of rays, per-PSF L1 <= 1e-4 after sum-normalisation.  "Relative" for a hit
coordinate is taken against max(|hit vector|, surface semi-diameter): the reference's own first-surface
hit carries ~1e-4 mm of float32 cancellation noise (t ~ 2000 mm), see DESIGN.md.

The reference is not bit-reproducible outside its own build (torch's CPU sqrt goes through MKL VML and is
off by one ulp for ~0.6 % of inputs, measured), so equality is asserted statistically, not bit-wise.
"""
import numpy as np
import pytest
import torch

from conftest import lens_path
from oracle import dp_oracle as O

D_SENSOR = {"rf50mm": 62.25, "rf35mm": 80.447}       # psfnet.py:42-48


def torch_pupil(u, pupil_r):
    """The reference's own pupil arithmetic (torch CPU ops, optics.py:483-487)."""
    ut = torch.from_numpy(np.asarray(u))
    th = ut[0] * 2 * np.pi
    rr = torch.sqrt(ut[1] * float(pupil_r) ** 2)
    return (rr * torch.cos(th)).numpy(), (rr * torch.sin(th)).numpy()


def psf_golden_samples(g, name, seed=3):
    """Redraw the pupil samples of tests/golden/psf.npz from the seed (optics.py:483-484, 900)."""
    chk = g[f"{name}_u_check"]
    spp = int(chk[4])
    torch.manual_seed(seed)
    u = [torch.rand(spp).numpy(), torch.rand(spp).numpy(), torch.rand(2048).numpy(), torch.rand(2048).numpy()]
    np.testing.assert_allclose([float(v.astype(np.float64).sum()) for v in u], chk[:4], rtol=0, atol=0)
    pr = g[f"{name}_pupil"][1]
    return torch_pupil(np.stack(u[:2]), pr), torch_pupil(np.stack(u[2:]), pr * 0.25)


def make_lens(name, hfov=None, res=(512, 768)):
    lens = O.load_lens(lens_path(name), sensor_res=res, d_sensor=D_SENSOR[name])
    if hfov is not None:
        lens.hfov = float(hfov)
    return lens


def l1_sumnorm(a, b):
    a = a / a.sum((-1, -2), keepdims=True)
    b = b / b.sum((-1, -2), keepdims=True)
    return np.abs(a - b).sum((-1, -2))


# ---------------------------------------------------------------------------------------------
def test_material_eta(golden):
    g = golden("setup")
    for name in ("rf50mm", "rf35mm"):
        lens = make_lens(name)
        for wi, wv in enumerate((0.656, 0.589, 0.486)):
            eta = [O.refractive_index(s.mat1, wv) / O.refractive_index(s.mat2, wv) for s in lens.surfaces]
            np.testing.assert_allclose(eta, g[f"{name}_eta"][wi], rtol=1e-15)
        assert lens.aper_idx == int(g[f"{name}_scalars"][0])
        assert lens.pixel_size == g[f"{name}_scalars"][8]
        assert abs(lens.r_last - g[f"{name}_scalars"][10]) < 1e-12


def test_single_ray_kat():
    """SURVEY.md §8(c) known-answer ray through rf50mm (values measured on the reference)."""
    lens = make_lens("rf50mm")
    o = np.asarray([[100.0, 50.0, -1937.75]], np.float32)
    d = np.asarray([[2.0, -1.5, 22.51324462890625]], np.float32) - o
    ray = O.RayBundle.from_od(o, d)
    rec = []
    O.trace_to_sensor(lens, ray, record=rec)
    np.testing.assert_allclose(rec[0].o()[0], [3.116264, -0.913391, 0.184814], atol=2e-5)
    np.testing.assert_allclose(rec[8].o()[0], [0.428677, -1.137581, 25.350334], atol=2e-5)
    np.testing.assert_allclose(rec[11].d()[0], [-0.0927971, 0.002629, 0.9956813], atol=2e-6)
    np.testing.assert_allclose(ray.o()[0], [-2.728621, -1.294840, 62.25], atol=3e-5)
    assert abs((-ray.dx / ray.dz)[0] - 0.09319963) < 2e-6
    assert ray.ra[0] == 1


@pytest.mark.parametrize("tag", ["rf50mm_w589", "rf50mm_w656", "rf50mm_w486", "rf35mm_w589"])
def test_trace_per_surface(golden, tag):
    g = golden("trace")
    name = tag.split("_")[0]
    wv = int(tag[-3:]) / 1000
    lens = make_lens(name, g[f"{name}_hfov"])
    obj = O.object_points(lens, g[f"{name}_points_norm"])
    assert np.array_equal(obj, g[f"{name}_points_obj"])
    pz, pr = g[f"{tag}_pupil"]
    px, py = torch_pupil(g[f"{tag}_u"], pr)
    ray = O.rays_from_points(obj, px, py, pz, wv)
    r0 = g[f"{tag}_ray0"]
    assert np.array_equal(ray.o(), r0[..., :3]) and np.array_equal(ray.d(), r0[..., 3:6])
    rec = []
    counts = O.trace_to_sensor(lens, ray, record=rec)
    assert counts == list(g[f"{tag}_newton"])                       # global Newton loop counts
    st = g[f"{tag}_states"]
    for i, r in enumerate(rec):
        ref_o, ref_d, ref_ra = st[i][..., :3], st[i][..., 3:6], st[i][..., 6]
        assert np.array_equal(r.ra, ref_ra), f"validity differs at surface {i}"
        scale = np.maximum(np.linalg.norm(ref_o, axis=-1), lens.surfaces[i].r)[..., None]
        assert (np.abs(r.o() - ref_o) / scale).max() < 1e-5
        assert np.abs(r.d() - ref_d).max() < 2e-6
        assert (np.concatenate([r.o(), r.d()], -1) == st[i][..., :6]).all(-1).mean() > 0.9
    ref = g[f"{tag}_sensor"]
    assert np.abs(ray.o() - ref[..., :3]).max() < 3e-5
    assert np.array_equal(ray.ra, ref[..., 6])


def test_backward_subrange_trace(golden):
    g = golden("trace")
    lens = make_lens("rf50mm")
    r0 = g["back_ray0"]
    ray = O.RayBundle.from_od(r0[:, :3], r0[:, 3:6], normalize=False)
    O.trace(lens, ray, range(0, lens.aper_idx))
    ref = g["back_final"]
    assert np.array_equal(ray.ra, ref[:, 6])
    np.testing.assert_allclose(ray.o(), ref[:, :3], atol=2e-6)
    np.testing.assert_allclose(ray.d(), ref[:, 3:6], atol=3e-7)


def test_setup_geometry(golden):
    g = golden("setup")
    for name in ("rf50mm", "rf35mm"):
        sc = g[f"{name}_scalars"]
        lens = O.load_lens(lens_path(name), d_sensor=D_SENSOR[name])
        # PSFNet overrides d_sensor AFTER hfov was computed with the JSON value (psfnet.py:42-48)
        lens_json = O.load_lens(lens_path(name))
        pz, pr = O.pupil_paraxial(lens_json)
        # The reference solves its nearly-parallel 2x2 line intersections with float32 torch.linalg.lstsq
        # (gelsy), whose answer differs from the exact solve by ~1e-3 relative in the radius (measured; it
        # also moves by 1e-5 with the thread count).  The product calls the same torch routine on the host;
        # the oracle solves exactly, hence the loose bound here.
        assert abs(pz - sc[4]) < 2e-3 and abs(pr - sc[5]) < 1e-2
        hfov = O.calc_hfov(lens_json)
        assert abs(hfov - sc[1]) < 2e-6
        u = g[f"{name}_refocus_u"]
        d_new = O.refocus(lens, -1000 + lens.d_sensor, u[0], u[1])
        assert abs(d_new - g[f"{name}_refocus"][0]) < 2e-4


@pytest.mark.parametrize("tag", ["small", "small_b", "big"])
def test_dp_weights(golden, tag):
    g = golden("dp_weights")
    prm = tuple(float(v) for v in g[f"{tag}_params"]) + ("l",)
    d_l, d_r = O.dp_weights(g["x_tan"], prm)
    np.testing.assert_allclose(d_l, g[f"{tag}_dl"], atol=3e-7)
    np.testing.assert_allclose(d_r, g[f"{tag}_dr"], atol=3e-7)
    if tag == "small":    # SURVEY §8(c) KAT
        kat = {-0.30: 0.543722, -0.12: 0.583120, -0.05: 0.489809, 0.0: 0.411824, 0.05: 0.336129, 0.12: 0.239431,
               0.30: 0.098454}
        for xt, v in kat.items():
            dl1, dr1 = O.dp_weights(np.asarray([xt], np.float32), prm)
            dl2, dr2 = O.dp_weights(np.asarray([-xt], np.float32), prm)
            assert abs(dl1[0] - v) < 2e-6 and abs(dr2[0] - v) < 2e-6


def test_splat_kat(golden):
    g = golden("dp_weights")
    ks, ps = 21, 0.046875
    q = np.asarray([[0.3 * ps, -1.7 * ps]], np.float32)
    d_l, d_r = O.dp_weights(np.asarray([0.0932], np.float32))
    one = np.ones(1, np.float32)
    L = O._bilinear_splat(q[:, 0], q[:, 1], (one, d_l), ks, ps)
    R = O._bilinear_splat(q[:, 0], q[:, 1], (one, d_r), ks, ps)
    np.testing.assert_allclose(L, g["kat_L"], atol=1e-7)
    np.testing.assert_allclose(R, g["kat_R"], atol=1e-7)
    assert np.count_nonzero(L) == 4 and abs(L[12, 10] - 0.134616) < 1e-6


@pytest.mark.parametrize("name", ["rf50mm", "rf35mm"])
def test_psf_bank_vs_reference(golden, name):
    g = golden("psf")
    lens = make_lens(name, g[f"{name}_hfov"])
    obj = O.object_points(lens, g[f"{name}_points_norm"])
    assert np.array_equal(obj, g[f"{name}_points_obj"])
    pz, pr = g[f"{name}_pupil"]
    (px, py), (cx, cy) = psf_golden_samples(g, name)
    L, R, centre = O.psf_bank(lens, obj, px, py, pz, 21, centre_samples=(cx, cy), params=O.DP_DEFAULT)
    np.testing.assert_allclose(centre, g[f"{name}_centre"], rtol=3e-6, atol=1e-8)   # fp32 sum-order noise
    print("end-to-end L1 (own centre):", l1_sumnorm(L, g[f"{name}_l"]).max())
    L, R, _ = O.psf_bank(lens, obj, px, py, pz, 21, centre=g[f"{name}_centre"], params=O.DP_DEFAULT)
    assert l1_sumnorm(L, g[f"{name}_l"]).max() < 1e-4
    assert l1_sumnorm(R, g[f"{name}_r"]).max() < 1e-4
    assert l1_sumnorm(L, g[f"{name}_none"]).max() < 1e-4          # param_list=None == defaults, 'l'
    np.testing.assert_allclose(L, g[f"{name}_l"], atol=2e-4)         # max-normalised values
    # L and R are far apart, so an L/R mix-up cannot pass
    assert l1_sumnorm(L, g[f"{name}_r"]).max() > 0.05
    # big-radius micro-lens variant, and ks=11
    big = (0.78, 1.44, 0.3, 0.6, "l")
    Lb, Rb, _ = O.psf_bank(lens, obj, px, py, pz, 21, centre=g[f"{name}_centre"], params=big)
    assert l1_sumnorm(Lb, g[f"{name}_big_l"]).max() < 1e-4 and l1_sumnorm(Rb, g[f"{name}_big_r"]).max() < 1e-4
    L11, _, _ = O.psf_bank(lens, obj, px, py, pz, 11, centre=g[f"{name}_centre"], params=O.DP_DEFAULT)
    assert l1_sumnorm(L11, g[f"{name}_ks11_l"]).max() < 1e-4
    # raw (un-normalised) grids: chief-ray centre and RMS centre (pointc_ref=None)
    Lraw, _, _ = O.psf_bank(lens, obj, px, py, pz, 21, centre=g[f"{name}_centre"], normalise=False)
    np.testing.assert_allclose(Lraw, g[f"{name}_chief_raw"], rtol=2e-4, atol=2e-2)
    ray = O.rays_from_points(obj, px, py, pz)
    O.trace_to_sensor(lens, ray)
    Lrms, _ = O.splat_points(ray, lens.pixel_size, 21, None)
    assert l1_sumnorm(Lrms, g[f"{name}_rms_raw"]).max() < 1e-4


def test_pixel_assignment_vs_reference(golden):
    """>= 99.99 % of rays land in the same pixel as in the reference's own trace."""
    g = golden("trace")
    for tag in ("rf50mm_w589", "rf35mm_w589"):
        name = tag.split("_")[0]
        lens = make_lens(name, g[f"{name}_hfov"])
        pz, pr = g[f"{tag}_pupil"]
        px, py = torch_pupil(g[f"{tag}_u"], pr)
        ray = O.rays_from_points(g[f"{name}_points_obj"], px, py, pz)
        O.trace_to_sensor(lens, ray)
        ref = g[f"{tag}_sensor"]
        rref = O.RayBundle(*(ref[..., i].copy() for i in range(7)))
        centre = O.chief_ray_centre(rref)
        same = []
        for r in (ray, rref):
            qx, qy, w = O.crop_and_shift(r, centre, 21, lens.pixel_size)
            r0, c0, _, _, _, _ = O.splat_indices(qx, qy, 21, lens.pixel_size)
            same.append((r0, c0, w))
        agree = (same[0][0] == same[1][0]) & (same[0][1] == same[1][1]) & (same[0][2] == same[1][2])
        assert agree.mean() >= 0.9999, agree.mean()


def test_psf_rgb(golden):
    g = golden("psf")
    lens = make_lens("rf50mm", g["rf50mm_hfov"])
    obj = O.object_points(lens, g["rf50mm_points_norm"][:2])
    pz, pr = g["rf50mm_pupil"]
    torch.manual_seed(3)
    out = []
    for wv in (0.656, 0.589, 0.486):                                 # basics.py:22, optics.py:1011-1013
        u = np.stack([torch.rand(4000).numpy(), torch.rand(4000).numpy()])
        uc = np.stack([torch.rand(2048).numpy(), torch.rand(2048).numpy()])
        px, py = torch_pupil(u, pr)
        cx, cy = torch_pupil(uc, pr * 0.25)
        L, _, _ = O.psf_bank(lens, obj, px, py, pz, 21, wvln=wv, centre_samples=(cx, cy))
        out.append(L)
    rgb = np.stack(out, 1)
    assert l1_sumnorm(rgb, g["rf50mm_rgb"]).max() < 1e-4


def test_tone_curves(golden):
    g = golden("render")
    np.testing.assert_allclose(O.degamma(g["tone_x"]), g["tone_degamma"], rtol=2e-6, atol=1e-6)
    np.testing.assert_allclose(O.gamma(g["tone_lum"]), g["tone_gamma_lum"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(O.gamma(O.degamma(g["tone_x"])), g["tone_gamma"], rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize("ks", [7, 21])
def test_render_local_psf(golden, ks):
    g = golden("render")
    rl, rr = O.render_local_psf(g[f"ks{ks}_img"], g[f"ks{ks}_psf"].astype(np.float32), ks)
    # fp16 output: allow one fp16 ulp (2^-11 relative) for the float32 summation order
    np.testing.assert_allclose(rl, g[f"ks{ks}_rl"], rtol=1.1e-3, atol=1e-6)
    np.testing.assert_allclose(rr, g[f"ks{ks}_rr"], rtol=1.1e-3, atol=1e-6)
    assert (rl == g[f"ks{ks}_rl"]).mean() > 0.9


# ---- PSFNet.pred / render in the arithmetic of the reference's CUDA run (fp16 MLP), tests/golden/predhalf.npz -------
def seeded_mlp_weights(seed=5, ks=21):
    """(W, b) of the reference's random-init PSF MLP, recreated from the seed (pinned by test_mlp_matches_reference_init)."""
    from sdirt_b200.deeplens.psfnet_arch import MLP, initialize_weights
    torch.manual_seed(seed)
    net = MLP(in_features=3, out_features=ks ** 2, hidden_features=512, hidden_layers=8)
    net.apply(initialize_weights)
    lin = [m for m in net.net if isinstance(m, torch.nn.Linear)]
    return [(m.weight.detach().numpy(), m.bias.detach().numpy()) for m in lin]


def half_ulps(a, b):
    """Distance in fp16 units in the last place between two arrays of fp16-representable values."""
    a16, b16 = np.asarray(a).astype(np.float16), np.asarray(b).astype(np.float16)
    ia, ib = a16.view(np.int16).astype(np.int32), b16.view(np.int16).astype(np.int32)
    return np.abs(ia - ib)                                                  # all values here are >= 0


def test_mlp_input_layer_half(golden):
    g = golden("predhalf")
    w = seeded_mlp_weights()
    xs, ys = O._torch_linspace(-1, 1, 24), O._torch_linspace(1, -1, 16)
    rows = O.mlp_input_rows(xs, ys, g["z"], 0, 2, 0, 16)
    assert rows.shape == (2 * 2 * 16 * 24, 3)
    h1 = O.mlp_linear_relu_half(O._h(rows), *w[0])
    # left rows against torch's own fp16 Linear + ReLU on the reference's coordinate grid
    assert half_ulps(h1[0::2].reshape(2, 16, 24, -1), g["h1_l"]).max() <= 1
    assert (half_ulps(h1[0::2].reshape(2, 16, 24, -1), g["h1_l"]) == 0).mean() > 0.999
    # right rows: same y, z, mirrored x
    np.testing.assert_array_equal(rows[1::2, 0], -rows[0::2, 0])
    np.testing.assert_array_equal(rows[1::2, 1:], rows[0::2, 1:])


def test_mlp_forward_half(golden):
    """The 11-layer fp16 chain: rounding to fp16 after every layer makes the result sensitive to the fp32 accumulation
    order inside each GEMM, so equality with torch's CPU half GEMM is statistical: a few fp16 ulps, mostly none."""
    g = golden("predhalf")
    w = seeded_mlp_weights()
    xs, ys = O._torch_linspace(-1, 1, 24), O._torch_linspace(1, -1, 16)
    rows = O.mlp_input_rows(xs, ys, g["z"], 0, 1, 0, 4)                     # 96 pixels: both sides, 192 rows
    raw = O.mlp_forward_half(w, rows)
    ref_l = g["raw_l"][0, :4].reshape(-1, 441).astype(np.float32)
    ref_r = g["raw_r"][0, :4].reshape(-1, 441).astype(np.float32)
    for got, ref in ((raw[0::2], ref_l), (raw[1::2], ref_r)):
        scale = ref.max()
        assert np.abs(got - ref).max() <= 4e-3 * scale
        assert np.abs(got - ref).mean() <= 2e-4 * scale


def test_psf_pack_half(golden):
    g = golden("predhalf")
    raw = np.stack((g["raw_l"].reshape(-1, 441), g["raw_r"].reshape(-1, 441)), 1).reshape(-1, 441)
    psf = O.psf_pack_half(raw, 21).reshape(2, 16, 24, 2, 21, 21)
    ref = g["psf"].astype(np.float32)
    ok = np.isfinite(ref).all((-1, -2))                                     # 0/0 kernels are NaN in the reference
    assert ok.mean() > 0.99
    u = half_ulps(psf[ok], ref[ok])
    assert u.max() <= 1 and (u == 0).mean() > 0.999
    assert (psf[~ok] == 0).all()
    # padded rows (the GEMM's N rounded up to 448) give the same kernels
    rawp = np.concatenate((raw, np.full((raw.shape[0], 7), 7.0, np.float16)), 1)
    np.testing.assert_array_equal(O.psf_pack_half(rawp, 21).reshape(psf.shape), psf)


def test_psfnet_render_half(golden):
    g = golden("predhalf")
    w = seeded_mlp_weights()
    out = O.psfnet_render_half(w, g["img"][:1], g["z"][:1], 21)
    ref = g["render_out"][:1]
    assert out.shape == ref.shape == (1, 6, 16, 24)
    assert np.abs(out - ref).max() < 2e-3                                   # a few fp16 ulps of a [0, 1] image


_ARBITER_CACHE = {}


def arbiter_psf(g, name, point, seed, counts):
    """One point of a 2 M-ray golden, traced by the oracle with the reference's own bundle-global Newton loop counts
    (bit-identical rays) and splatted with float64 accumulation of the reference's float32 addends (oracle
    splat_points_f64).  Returns (L64, R64) max-normalised like psf_diff.  Cached per process (tens of seconds of numpy)."""
    key = (name, point, seed)
    if key not in _ARBITER_CACHE:
        lens = make_lens(name, float(g["hfov"]))
        spp = int(g["u_check"][2])
        torch.manual_seed(seed)
        u = [torch.rand(spp).numpy(), torch.rand(spp).numpy()]
        assert [float(v.astype(np.float64).sum()) for v in u] == list(g["u_check"][:2])
        pz, pr = g["pupil"]
        px, py = torch_pupil(np.stack(u), pr)
        ray = O.rays_from_points(g["points_obj"][point:point + 1], px, py, float(pz))
        O.trace_to_sensor(lens, ray, newton_iters=[int(c) for c in counts])
        L, R = O.splat_points_f64(ray, lens.pixel_size, 21, g["centre"][point:point + 1])
        _ARBITER_CACHE[key] = (L[0] / (L[0].max() + 1e-6), R[0] / (R[0].max() + 1e-6))
    return _ARBITER_CACHE[key]


def arbiter_in_focus_corner(g):
    """The rf50mm field corner exactly in focus at 2 M rays (point 1 of the depth-sweep golden)."""
    return arbiter_psf(g, "rf50mm", 1, 21, g["newton_counts"][0])


def test_in_focus_corner_reference_vs_float64_sum(golden):
    """VERDICT r1 'settle the in-focus-corner PSF': every engine mode, the bit-exact replay included, is 1.0 ... 1.3e-4 (L1)
    from the reference's PSF of the field corner exactly in focus at 2 M rays.  The arbiter (same rays, same float32
    addends, float64 accumulation) shows where that distance comes from: the reference's own float32 running sums
    (monte_carlo.py:225-235: 2 M addends of ~0.4 into four taps of ~2.5e5) are that far from the exact sum of their
    addends.  The GPU test of the same name bounds the engine's distance to the arbiter."""
    g = golden("psf2m_sweep")
    L64, R64 = arbiter_in_focus_corner(g)
    d_ref = max(l1_sumnorm(g["l"][1:2].astype(np.float64), L64[None])[0], l1_sumnorm(g["r"][1:2].astype(np.float64), R64[None])[0])
    print("reference PSF vs float64 sum of its own addends, L1:", d_ref)
    assert 5e-5 < d_ref < 3e-4
    assert (L64 > 1e-3).sum() <= 12         # the whole PSF is a handful of taps


def test_rf35mm_far_corner_reference_vs_float64_sum(golden):
    """The same dispute on BASELINE config 3's lens: the rf35mm field corner at 20 m is nearly in focus (eight taps carry the
    PSF), and the reference's sequential float32 sums put its PSF 3.4e-4 (L) / 1.2e-4 (R) from the exact sum of its own
    addends, while the oracle's float32 splat in the reference's order reproduces the reference to 2e-5."""
    g = golden("rf35mm2m")
    L64, R64 = arbiter_psf(g, "rf35mm", 1, 33, g["newton_counts"])
    d_l = l1_sumnorm(g["l"][1:2].astype(np.float64), L64[None])[0]
    d_r = l1_sumnorm(g["r"][1:2].astype(np.float64), R64[None])[0]
    print("rf35mm 20 m corner: reference PSF vs float64 sum of its own addends, L1:", d_l, d_r)
    assert 1e-4 < d_l < 6e-4 and 5e-5 < d_r < 3e-4

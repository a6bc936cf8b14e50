"""CPU: host logic of the product (no GPU, no compute calls): C-ABI exports, prescription IO, dispersion, options,
sharding arithmetic and the world_size-2 gather over gloo, loud failure without CUDA."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, lens_path


def test_build_and_exports():
    import __graft_entry__ as ge
    path = ge.build()
    lib = ctypes.CDLL(path)
    header = open(os.path.join(ROOT, "include", "sdirt_engine.h")).read()
    names = sorted(set(re.findall(r"\b(sdirt_[a-z0-9_]+)\s*\(", header)))
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), n
    lib.sdirt_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.sdirt_version()


def test_sass_is_sm100a():
    out = os.popen(f"cuobjdump -lelf {os.path.join(ROOT, 'sdirt_b200', '_lib', 'libsdirt_engine.so')} 2>/dev/null").read()
    assert "sm_100a" in out


def test_prescription_and_eta(golden):
    from sdirt_b200 import _engine as E
    from sdirt_b200.prescription import find_aperture, load_lens_json
    g = golden("setup")
    for name, n_surf, aper, n_asph in (("rf50mm", 12, 5, 2), ("rf35mm", 21, 7, 1)):
        recs, descs, head = load_lens_json(lens_path(name))
        assert len(recs) == n_surf and find_aperture(descs) == aper
        assert sum(d["kind"] == E.SURF_ASPHERE for d in descs) == n_asph
        assert descs[aper]["kind"] == E.SURF_FLAT
        h = E.LensHandle(recs, head["d_sensor"])           # host-only object: no device needed
        for wi, wv in enumerate((0.656, 0.589, 0.486)):
            np.testing.assert_allclose(h.eta(wv), g[f"{name}_eta"][wi], rtol=1e-15)
            np.testing.assert_allclose(h.eta(wv, backward=True), 1.0 / g[f"{name}_eta"][wi], rtol=1e-15)


def test_bad_arguments_are_reported():
    from sdirt_b200 import _engine as E
    with pytest.raises(RuntimeError, match="curved surface with c == 0"):
        E.LensHandle([E.make_surface(E.SURF_SPHERE, 5.0, 0.0, 0.0)], 10.0)
    with pytest.raises(RuntimeError, match="surfaces"):
        E.LensHandle([E.make_surface(E.SURF_FLAT, 5.0, float(i)) for i in range(40)], 10.0)
    with pytest.raises(ValueError):
        E.make_options(numerics="sloppy")
    o = E.make_options([3, 2, 1], "hybrid")
    assert (o.newton_mode, o.numerics, list(o.iters[:3])) == (E.NEWTON_REPLAY, E.NUMERICS_HYBRID, [3, 2, 1])


def test_no_cpu_fallback():
    """The product refuses CPU tensors / a CPU device instead of silently computing somewhere else."""
    from sdirt_b200 import _engine as E, lens_file
    from sdirt_b200.deeplens import PSFNet, local_psf_render_fast
    from sdirt_b200.prescription import load_lens_json
    recs, _, head = load_lens_json(lens_file("rf50mm"))
    h = E.LensHandle(recs, 62.25)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        E.trace_rays(h, 0.589, torch.zeros(4, 3), torch.zeros(4, 3), torch.ones(4))
    with pytest.raises(RuntimeError, match="CUDA"):
        PSFNet(lens_file("rf50mm"), sensor_res=(512, 768), kernel_size=21, device="cpu")
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        local_psf_render_fast(torch.rand(1, 3, 8, 8), torch.rand(1, 8, 8, 2, 5, 5), 5)
    import sdirt_b200
    src = "".join(open(os.path.join(ROOT, "sdirt_b200", f)).read() for f in os.listdir(os.path.join(ROOT, "sdirt_b200")) if f.endswith(".py"))
    assert "oracle" not in src.replace("oracle/dp_oracle.py", "")      # the product never imports the oracle


def test_mlp_matches_reference_init(golden):
    """Seeded construction reproduces the reference's random-init PSF MLP (checksums from the reference)."""
    from sdirt_b200.deeplens.psfnet_arch import MLP, initialize_weights
    g = golden("render")
    torch.manual_seed(5)
    net = MLP(in_features=3, out_features=21 ** 2, hidden_features=512, hidden_layers=8)
    net.apply(initialize_weights)                                       # PSFNet.init_net applies it once more
    got = np.asarray([[v.double().sum().item(), v.double().abs().sum().item()] for v in net.state_dict().values()])
    np.testing.assert_allclose(got, g["mlp_checksum"], rtol=1e-12, atol=1e-12)


def test_shard_bounds():
    from sdirt_b200.sharding import shard_bounds, shard_slice
    for n in (0, 1, 7, 4096, 131072):
        for w in (1, 2, 3, 8):
            b = shard_bounds(n, w)
            assert b[0] == 0 and b[-1] == n and len(b) == w + 1
            sizes = np.diff(b)
            assert sizes.max() - sizes.min() <= 1 and (sizes >= 0).all()
    assert shard_slice(10, 1, 4) == slice(3, 6)


def _gather_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from sdirt_b200.sharding import gather_blocks, shard_slice
    n_total = 7                                                    # ragged: 4 + 3
    full = torch.arange(n_total * 2 * 3 * 3, dtype=torch.float32).reshape(n_total, 2, 3, 3)
    sl = shard_slice(n_total)
    out = gather_blocks(full[sl].clone(), n_total)
    ret[rank] = bool(torch.equal(out, full)) and (sl == (slice(0, 4) if rank == 0 else slice(4, 7)))
    dist.destroy_process_group()


def test_gather_blocks_gloo_world2():
    world, port = 2, 29500 + os.getpid() % 2000
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_gather_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world))


def _dealt_worker(rank, world, port, ret):
    import torch.distributed as dist
    from sdirt_b200.sharding import dealt_groups, gather_dealt
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_groups, g = 8, 3
    full = torch.arange(n_groups * g * 2 * 4, dtype=torch.float32).reshape(n_groups * g, 2, 4)
    mine = dealt_groups(n_groups)
    local = torch.cat([full[k * g:(k + 1) * g] for k in mine])
    out = gather_dealt(local, n_groups)
    ret[rank] = bool(torch.equal(out, full)) and mine == list(range(rank, n_groups, world))
    dist.destroy_process_group()


def test_gather_dealt_gloo_world2():
    """The strong-scaling assembly of bench.py's `strong` leg: depth slabs dealt round-robin, one all-gather, reorder."""
    world, port = 2, 31500 + os.getpid() % 2000
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_dealt_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world))
    from sdirt_b200.sharding import dealt_groups
    assert dealt_groups(32, 3, 8) == [3, 11, 19, 27]
    with pytest.raises(ValueError):
        dealt_groups(32, 0, 3)


def test_bench_workload_shape():
    sys.path.insert(0, ROOT)
    import bench
    p = bench.bank_points(0)
    assert p.shape == (4096, 3) and abs(float(p[0, 0]) + 127 / 128) < 1e-6 and abs(float(p[0, 1]) - 127 / 128) < 1e-6
    depths = [float(bench.bank_points(s)[0, 2]) for s in range(32)]
    assert min(depths) >= -20000 - 1e-3 and max(depths) <= -200 + 1e-3 and len(set(depths)) == 32
    # N GPUs: the same slab on every rank, N distinct sub-cell-shifted field grids inside [-1, 1]; N = 1 is the plain grid
    assert torch.equal(bench.bank_points(5, 0, 1), bench.bank_points(5))
    for world in (2, 4, 8):
        assert len({bench.slab_of_step(4, r, world) for r in range(world)}) == 1 and bench.slab_of_step(4, 0, world) == bench.slab_of_step(4)
        pts = torch.cat([bench.bank_points(5, r, world) for r in range(world)])
        assert float(pts[:, :2].abs().max()) < 1.0 and len(torch.unique((pts[:, :2] * 1e5).round(), dim=0)) == world * 4096
        assert len(set(pts[:, 2].tolist())) == 1


def test_fused_kernel_sass_is_tcgen05_cta_pairs():
    """The fused PSF-MLP kernel is the tcgen05 / TMEM path it claims to be: 2-CTA UMMA, multicast commits, TMEM loads, bulk
    copies, cluster barriers (SASS mnemonics of profiles' B200 recipe); and its issue loop holds no ELECT / R2UR waterfall."""
    import subprocess
    from sdirt_b200 import build
    sass = subprocess.run(["cuobjdump", "-sass", build.build()], capture_output=True, text=True).stdout
    body, on = [], False
    for line in sass.splitlines():
        if "Function :" in line:
            on = "mlp_fused_pred_kernelILi21ELi2E" in line
        elif on:
            body.append(line)
    text = "\n".join(body)
    for mnemonic in ("UTCHMMA.2CTA", "UTCBAR.2CTA.MULTICAST", "LDTM", "UBLKCP", "UCGABAR_ARV", "FADD2", "FFMA2"):
        assert mnemonic in text, mnemonic
    parts = text.split("UTCHMMA.2CTA")
    n_mma = len(parts) - 1
    assert n_mma >= 4 and n_mma % 4 == 0                         # four K = 16 steps per stage at every (cloned) issue site
    for i in range(1, n_mma, 4):                                 # back-to-back MMAs, operands already in uniform registers
        assert not any("R2UR" in seg or "ELECT" in seg for seg in parts[i:i + 3])


def test_fused_band_shape_fills_rounds():
    """PSFNet's band chooser for the fused engine: bands whose last round of the persistent kernel is (nearly) full."""
    from sdirt_b200.deeplens.psfnet import PSFNet
    stub = PSFNet.__new__(PSFNet)
    for n, h, w in ((2, 1024, 1536), (4, 512, 768), (16, 1024, 1536), (1, 512, 768)):
        rows, nb = PSFNet._fused_band_shape(stub, n, h, w)
        assert rows % 16 == 0 and 1 <= nb <= n and rows <= h
        rounds = -(-(nb * rows * w) // 128) / 74
        assert rounds / -(-rounds // 1) > 0.95 and nb * rows * w <= max(PSFNet.render_band_pixels_fused, rows * w)
    stub.render_band_rows, stub.render_band_pixels = 8, 1           # set by hand: kept
    assert PSFNet._fused_band_shape(stub, 4, 512, 768) == (8, 1)


def test_fused_mlp_layout_host_side():
    """sdirt_mlp_fused_layout is host arithmetic: packed sizes of the flagship MLP (last layer: 21 kernel rows padded to 24
    accumulator columns = 504 -> 512), smaller windows, and the shapes it refuses."""
    import ctypes as C
    from sdirt_b200 import _engine as E
    lib = E.lib()

    def shape(n1, dims):
        sh = E.MlpShape()
        sh.n_layers, sh.n1 = len(dims), n1
        k = n1
        for l, n in enumerate(dims):
            sh.K[l], sh.N[l] = k, n
            k = n
        return sh

    w_off, b_off, bias = (C.c_int64 * 12)(), (C.c_int32 * 12)(), C.c_int64(0)
    sh = shape(128, [512] * 9 + [441])
    nbytes = lib.sdirt_mlp_fused_layout(C.byref(sh), w_off, b_off, C.byref(bias))
    assert nbytes == 2 * (512 * 128 + 8 * 512 * 512 + 512 * 512) == 4849664          # last layer packed as 512 rows
    assert bias.value == 10 * 512 and list(b_off[:10]) == [512 * i for i in range(10)]
    assert w_off[1] == 2 * 512 * 128 and w_off[9] == 2 * (512 * 128 + 8 * 512 * 512)
    sh = shape(64, [256, 128, 49])                                                     # ks = 7: 7 rows of 8 columns = 56 -> 64
    assert lib.sdirt_mlp_fused_layout(C.byref(sh), None, None, C.byref(bias)) == 2 * (256 * 64 + 128 * 256 + 64 * 128)
    assert bias.value == 256 + 128 + 64
    sh = shape(128, [512, 121])                                                        # ks = 11: 11 rows of 12 columns = 132 -> 144
    assert lib.sdirt_mlp_fused_layout(C.byref(sh), None, None, C.byref(bias)) == 2 * (512 * 128 + 144 * 512)
    for bad in (shape(128, [512, 100]), shape(128, [100, 441]), shape(96, [512, 441]), shape(128, [640, 441])):
        assert lib.sdirt_mlp_fused_layout(C.byref(bad), None, None, None) == -1
        assert lib.sdirt_last_error()


def test_render_records_size_host_side():
    """sdirt_render_records_bytes is host arithmetic: one record per padded image row and 32-pixel strip -- 3 channels x 27 words in
    two pairings, the second at a word offset of 16 (mod 32), rounded to 16 bytes (784 B at ks = 21, 592 B at 11, 560 B at 7) --,
    0 for the shapes the strip-walking render kernel does not take; the packed entries refuse those shapes without a device."""
    from sdirt_b200 import _engine as E
    lib = E.lib()
    assert lib.sdirt_render_records_bytes(2, 3, 1024, 1536, 21) == 2 * (1024 + 20) * 48 * 784 == 78575616
    assert lib.sdirt_render_records_bytes(1, 3, 512, 768, 11) == (512 + 10) * 24 * 592
    assert lib.sdirt_render_records_bytes(4, 3, 64, 32, 7) == 4 * (64 + 6) * 1 * 560
    for b, c, h, w, ks in ((1, 3, 64, 40, 21), (1, 3, 64, 64, 9), (1, 1, 64, 64, 21), (1, 4, 64, 64, 21), (0, 3, 64, 64, 21), (1, 3, 64, 65536, 21)):
        assert lib.sdirt_render_records_bytes(b, c, h, w, ks) == 0
    assert lib.sdirt_render_pack_image(None, 1, 3, 64, 40, 21, 0, None, None) < 0 and b"packed render" in lib.sdirt_last_error()
    assert lib.sdirt_render_local_psf_rows_packed(None, None, 1, 3, 64, 64, 60, 8, 21, 0, None, None, None) < 0 and b"outside the image" in lib.sdirt_last_error()
    assert lib.sdirt_render_local_psf_rows_packed(None, None, 1, 3, 64, 64, 0, 8, 21, 0, None, None, None) < 0 and b"null buffer" in lib.sdirt_last_error()


def test_every_exported_symbol_is_documented():
    """The drop-in boundary is the header: every entry point it declares appears, by its full name, in INTEGRATION.md's table of
    what it replaces in the reference."""
    import re
    hdr = open(os.path.join(ROOT, "include", "sdirt_engine.h")).read()
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    names = sorted(set(re.findall(r"\b(sdirt_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 30
    assert [n for n in names if n not in doc] == []


def test_reference_arm_runs_the_unmodified_reference():
    """bench.py's CPU legs time the UNMODIFIED reference staged under baseline/_ref (tools/stage_reference.py), not a port: every
    staged file is byte-identical to /root/reference where that exists, the runner imports the reference's own `deeplens` (never
    the mirror), and one tiny psf_diff call goes through."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import bench
    import ref_runner as R
    if not R.available():
        pytest.skip("baseline/_ref is not staged in this checkout (python tools/stage_reference.py)")
    if os.path.isdir("/root/reference/deeplens"):
        import filecmp
        for pkg in ("deeplens", "dfdp"):
            for dp, _, fn in os.walk(os.path.join(R.REF, pkg)):
                for f in fn:
                    if f.endswith(".py"):
                        staged = os.path.join(dp, f)
                        assert filecmp.cmp(staged, os.path.join("/root/reference", os.path.relpath(staged, R.REF)), shallow=False), staged
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k == "deeplens" or k.startswith("deeplens.")}
    try:
        for k in saved:
            sys.modules.pop(k, None)
        dl = R.import_reference()
        assert os.path.abspath(dl.__file__).startswith(R.REF)
        rate, times, threads = bench.reference_cpu_rays_per_s(R, 1, warmup=0, sample=(4, 512))
        assert rate > 0 and len(times) == 1 and threads >= 1
    finally:
        for k in [k for k in sys.modules if k == "deeplens" or k.startswith("deeplens.") or k == "dfdp" or k.startswith("dfdp.")]:
            sys.modules.pop(k, None)
        sys.modules.update({k: v for k, v in saved.items() if v is not None})

#!/usr/bin/env python
"""bench.py — rays/s traced + splatted into dual-pixel PSFs (BASELINE.json metric), one JSON line.

Workload (BASELINE.json configs[1]): the rf50mm F/4 PSF bank for PSFNet fitting, 64x64 field points x 32 depths,
2 M rays per point.  One "step" = one depth slab of that bank: 4096 points x 2 M rays = 8.4e9 rays through
sample -> 12-surface trace -> DP weights -> bilinear splat -> normalise, i.e. 4096 (L, R) PSF pairs.  Consecutive
steps walk the 32 depth slabs with stride 11 (near, in-focus and far depths alike).  With N GPUs the job is an N times denser
field sampling of the same bank: every rank works on its own (sub-cell shifted) 64x64 field grid of the step's slab (weak
scaling, equal work per rank, no collective on the data path); `value` is the whole-job rays/s = N x slab rays /
max-over-ranks device time.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--numerics adaptive|strict|hybrid|fast] [--impl reference]

Besides the contract keys the line carries (each a dict with its own metric / value / unit):
  conformant        the same workload in the numerics mode that meets all three north_star tolerances (strict)
  rf35mm            BASELINE configs[2]: the rf35mm bank slab at four focus distances (Lensgroup.refocus), this rank count
  strong            the FULL 131 072-point bank sharded over the N ranks + ONE all-gather of the 462 MB bank (seconds per bank)
  render_sharded    BASELINE configs[3]: PSFNet.render of 16 x 3 x 1024 x 1536 sharded by image over the N ranks
  datagen           BASELINE configs[4] shape: focal-stack generation (render + noise) for a DfDP batch, per rank
  eager_gpu_baseline  (N = 1) the UNMODIFIED reference run eagerly on the same B200 (psf_diff and PSFNet.render)
  cpu_baseline      (N = 1) the UNMODIFIED reference's torch-CPU path on this box's host cores, bounded sample

`--impl reference` times the unmodified reference (`baseline/_ref`, staged by tools/stage_reference.py; Lensgroup.psf_diff,
optics.py:934-996) on all host threads on bounded samples of the same workload and prints the same JSON shape with
"impl": "reference".  Only when baseline/_ref is absent does it fall back to the numpy port (oracle/dp_oracle.py, "kind": "port").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GRID, DEPTHS, SPP, KS = 64, 32, 2_000_000, 21
SENSOR_RES = (512, 768)
LENS = "rf50mm"          # --lens rf35mm switches the MAIN leg to BASELINE config 3's prescription (same bank shape)
# algorithmic flop/ray with the reference's minimal per-ray Newton counts (FMA = 2); the strict mode executes the
# reference's own iteration counts (3.1 / 4.9 kflop per ray)
FLOPS_PER_RAY = {"rf50mm": 2.1e3, "rf35mm": 3.3e3}
FLOPS_PER_RAY_STRICT = {"rf50mm": 3.1e3, "rf35mm": 4.9e3}
FLOP_PER_RAY = FLOPS_PER_RAY[LENS]
# The reference's own setup values (tests/test_api_gpu.py pins the engine's against them):
HFOV = {"rf50mm": 0.40959781408309937, "rf35mm": 0.5514792203903198}
PUPIL = {"rf50mm": (22.51324462890625, 6.019352912902832), "rf35mm": (14.338210105895996, 4.767455577850342)}
D_SENSOR = {"rf50mm": 62.25, "rf35mm": 80.447}
FP32_NOMINAL_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12        # 148 SMs x 128 lanes x 2 flop x 1.965 GHz boost = 74.4
TOLERANCES = {
    "strict": "all three north_star tolerances: hit coordinates bit-identical to the reference arithmetic, pixel / sub-pixel "
              "assignment identical for >= 99.99 % of rays (100 % when the reference's global Newton counts are replayed), PSF L1 <= 1e-4",
    "hybrid": "hit coordinates <= 1e-5 relative and PSF L1 <= 1e-4; pixel assignment identical for >= 99.95 % of rays (NOT the 99.99 % asked)",
    "adaptive": "hit coordinates <= 1e-5 relative and PSF L1 <= 1e-4 (2 M rays: <= 3e-5 against the reference at the depth-sweep golden points, "
                "<= 4.6e-5 against the strict mode at points right under the 2048 mm bound of its fast arithmetic); pixel assignment "
                "identical for >= 99.7 % of rays (NOT the 99.99 % asked: see `conformant` for the mode that meets it)",
    "fast": "PSF L1 <= 1e-4 up to 8 m, 1.1e-4 at the 20 m field corner; pixel assignment identical for >= 99.7 % of rays",
}


# ----------------------------------------------------------------------------------------------------
# workload
# ----------------------------------------------------------------------------------------------------
def bank_points(slab, rank=0, world=1, lens=None):
    """Normalised (x, y, depth) of depth slab `slab` (0..31): cell-centred 64x64 field grid (psfnet.py:221-225 with
    64 cells) at the slab's depth, z from the get_test_data warp of linspace(-3, 3, 32) (psfnet.py:229-232).
    With `world` GPUs the job is a `world` times denser field sampling of the same bank: rank r takes the same grid moved to
    the r-th of m x m sub-cell centres (m = ceil(sqrt(world))), so every rank traces different points of equal cost."""
    import torch
    g = GRID
    m = 1
    while m * m < world:
        m += 1
    cell = 1.0 / g                      # the grid's margin to the field edge is 1 / (2 g): shifted copies stay inside [-1, 1]
    ox, oy = (((rank % m) + 0.5) / m - 0.5) * cell, (((rank // m) + 0.5) / m - 0.5) * cell
    x, y = torch.meshgrid(torch.linspace(-1 + 1 / (2 * g), 1 - 1 / (2 * g), g) + ox,
                          torch.linspace(1 - 1 / (2 * g), -1 + 1 / (2 * g), g) + oy, indexing="xy")
    d_min, d_max, ds = -200.0, -20000.0, D_SENSOR[lens or LENS]
    foc_z = ((-1000.0 + ds) - d_min) / (d_max - d_min)
    zg = torch.linspace(-3, 3, DEPTHS)[slab % DEPTHS]
    z = (1 - foc_z) * zg / 3 + foc_z if zg > 0 else foc_z * zg / 3 + foc_z
    depth = z * (d_max - d_min) + d_min
    return torch.stack((x.reshape(-1), y.reshape(-1), torch.full((g * g,), float(depth))), -1).float()


def slab_of_step(step, rank=0, world=1):
    """Depth slab (0..31) a given step works on: a fixed stride-11 walk through the 32 depths, so that any few consecutive
    steps sample near, in-focus and far slabs alike (cost per ray varies with depth).  The same slab on every rank (each
    rank has its own field points of it, see bank_points): the ranks' steps cost the same and weak scaling measures the
    machine, not the luck of the depth draw."""
    return (step * 11) % DEPTHS


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(self.rows[0][1]) if self.rows else None, "samples": len(self.rows), "reasons": reasons}


# ----------------------------------------------------------------------------------------------------
# CPU legs.  (a) the unmodified reference through baseline/ref_runner.py; (b) the numpy port, only when (a) is absent
# ----------------------------------------------------------------------------------------------------
def _ref_runner():
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import ref_runner as R
    return R if R.available() else None


REF_SAMPLE = (64, 32768)           # points x rays of one reference step (2.1e6 rays: seconds on a host's cores)


def reference_cpu_rays_per_s(R, repeats, warmup=1, threads=None, sample=REF_SAMPLE, lens_name=None):
    """The reference's Lensgroup.psf_diff (optics.py:934-996) on torch CPU: `repeats` timed calls on `sample[0]` points of the
    workload's slabs x `sample[1]` rays.  One call traces and splats points x rays rays into the left PSFs."""
    import torch
    if threads:
        torch.set_num_threads(threads)
    name = lens_name or LENS
    lens = R.make_psfnet(name, SENSOR_RES, KS, "cpu")
    n_pts, spp = sample
    times = []
    for r in range(warmup + repeats):
        pts = bank_points(slab_of_step(r), lens=name)[:: (GRID * GRID) // n_pts][:n_pts].contiguous()
        t, _ = R.time_psf(lens, pts, KS, spp, repeats=1, seed=1000 + r)
        if r >= warmup:
            times += t
    return n_pts * spp * len(times) / sum(times), times, torch.get_num_threads()


def _cpu_chunk(args):
    """Worker of the port fallback: trace + splat `n_pts` points x one chunk of pupil samples with the numpy oracle."""
    slab, p0, n_pts, seed, spp = args
    from oracle import dp_oracle as O
    lens = O.load_lens(os.path.join(ROOT, "sdirt_b200", "lenses", LENS + ".json"), sensor_res=SENSOR_RES, d_sensor=D_SENSOR[LENS])
    lens.hfov = HFOV[LENS]
    pts = O.object_points(lens, bank_points(slab).numpy()[p0:p0 + n_pts])
    rng = np.random.default_rng(seed)
    pz, pr = PUPIL[LENS]
    px, py = O.pupil_points(rng.random(spp, dtype=np.float32), rng.random(spp, dtype=np.float32), pr)
    cx, cy = O.pupil_points(rng.random(2048, dtype=np.float32), rng.random(2048, dtype=np.float32), pr * 0.25)
    L, R, _ = O.psf_bank(lens, pts, px, py, pz, KS, centre_samples=(cx, cy), params=O.DP_DEFAULT)
    return float(L.sum() + R.sum())


def port_cpu_rays_per_s(n_pts, spp, repeats=1, slab0=0):
    """Fallback when baseline/_ref is absent: the numpy port on `n_pts` points x `spp` rays per repeat, one process per core."""
    from concurrent.futures import ProcessPoolExecutor
    cores = os.cpu_count() or 1
    per = max(1, n_pts // cores)
    times = []
    with ProcessPoolExecutor(max_workers=cores) as ex:
        list(ex.map(_cpu_chunk, [(0, 0, 1, 0, 256)] * cores))            # spin the workers up (imports)
        for r in range(repeats):
            jobs = [(slab0 + r, p0, min(per, n_pts - p0), 1000 + r, spp) for p0 in range(0, n_pts, per)]
            t0 = time.perf_counter()
            list(ex.map(_cpu_chunk, jobs))
            times.append(time.perf_counter() - t0)
    return n_pts * spp * len(times) / sum(times), times, cores


def run_reference(args):
    """`--impl reference`: K timed steps, each a bounded sample of the slab workload, on rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch
    torch.set_num_threads(os.cpu_count() or 1)        # (torchrun pins OMP_NUM_THREADS=1 for its workers: the CPU arm takes the whole host)
    R = _ref_runner()
    extra = {}
    if R is not None:
        value, times, cores = reference_cpu_rays_per_s(R, args.steps, warmup=args.warmup)
        kind = "reference"
        sample = (f"{REF_SAMPLE[0]} points x {REF_SAMPLE[1]} rays per step of the {GRID}x{GRID}x{DEPTHS} x 2M-ray bank: the unmodified "
                  f"reference's Lensgroup.psf_diff (torch CPU, {cores} threads); one call = the LEFT PSFs, the reference's right PSFs cost a "
                  f"second, mirrored trace (psfnet.py:540-544)")
        one, _, _ = reference_cpu_rays_per_s(R, 1, warmup=0, threads=1, sample=(REF_SAMPLE[0], REF_SAMPLE[1] // 8))
        extra["one_thread_rays_per_s"] = one
        n_pts, spp = REF_SAMPLE
    else:
        n_pts, spp = 64, 131072
        value, times, cores = port_cpu_rays_per_s(n_pts, spp, repeats=args.warmup + args.steps, slab0=0)
        times = times[args.warmup:]
        value = n_pts * spp * len(times) / sum(times)
        kind = "port"
        sample = (f"{n_pts} points x {spp} rays per step; baseline/_ref is absent, so this is the numpy port of the reference's "
                  f"algorithm (oracle/dp_oracle.py), one process per core")
    cfg = workload_config("cpu")
    cfg["timed_region"] = "one call of the reference's psf_diff per step (sampling, trace, chief-ray centre, forward_integral, normalisation), wall clock"
    cfg["l2"] = "n/a (CPU)"
    print(json.dumps({
        "impl": "reference", "metric": "rays/sec traced+splatted into DP L/R PSFs", "value": value, "unit": "rays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": dict({"value": value, "unit": "rays/s", "cores": cores, "kind": kind, "sample": sample}, **extra),
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "dp_psfs_per_s": value / spp / (2 if kind == "reference" else 1),
    }))


def workload_config(numerics):
    return {"workload": f"{LENS} F/4 PSF bank, {GRID}x{GRID} field x {DEPTHS} depths, {SPP} rays/point, ks={KS}, "
                        f"sensor {SENSOR_RES[0]}x{SENSOR_RES[1]}; step = one depth slab ({GRID * GRID} points), slabs visited "
                        f"with stride 11 over the {DEPTHS} depths; N GPUs = N sub-cell-shifted copies of the field grid at the "
                        f"same slab (N times denser field sampling, equal work per rank)",
            "lens": LENS, "points_per_step": GRID * GRID, "rays_per_point": SPP, "ks": KS, "numerics": numerics,
            "tolerances_met": TOLERANCES.get(numerics, "n/a"),
            "timed_region": "`value`: sdirt_psf_bank of one slab (dp_lut + psf_bank_run + psf_finalize kernels) with points, "
                            "Morton-sorted pupil samples and chief-ray centres resident; the per-slab chief-ray centres (sdirt_psf_centre, "
                            "2048 rays / point) and the one-off sort of the shared sample set (sdirt_pupil_sort) run BEFORE the timed "
                            "region (0.2 % of a step together); `e2e` includes both, every step",
            "l2": "L2 flushed (256 MiB write) between timed steps; the 16 MB shared pupil-sample set is re-read from L2 "
                  "by every point inside a step by design"}


# ----------------------------------------------------------------------------------------------------
# GPU leg
# ----------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from sdirt_b200 import _engine as E, lens_file, sharding
    from sdirt_b200.deeplens import PSFNet

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU path (use --impl reference for the CPU leg)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # The JSON line must be the only thing on stdout, and NCCL writes its version banner to file descriptor 1 when the first
    # communicator is built: point fd 1 at stderr for the whole run and keep the real stdout for the one line at the end.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        # torchrun pins OMP_NUM_THREADS=1; what is left of the host side of the end-to-end path may use this rank's share of the cores
        torch.set_num_threads(max(1, (os.cpu_count() or 1) // world))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([float(x)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def bank_leg(lens_name, numerics, steps, warmup, d_sensor=None, slabs_of=None, want_clocks=False):
        """K timed slab steps of one prescription in one numerics mode, inputs resident; returns a dict and what the e2e leg reuses."""
        lens = PSFNet(lens_file(lens_name), sensor_res=SENSOR_RES, kernel_size=KS, device=dev)
        if d_sensor is not None:
            lens.d_sensor = d_sensor
        lens.numerics = numerics
        handle = lens._engine_lens()
        pz, pr = lens.entrance_pupil()
        torch.manual_seed(1234)                                        # same pupil samples on every rank (optics.py:483-490)
        theta = torch.rand(SPP) * 2 * np.pi
        rho = torch.sqrt(torch.rand(SPP) * pr ** 2)
        pupil = torch.stack((rho * torch.cos(theta), rho * torch.sin(theta)), 1).to(dev).contiguous()
        cpupil = (pupil[:2048] * 0.25).contiguous()
        pupil = E.pupil_sort(pupil, pr)                                # one-off spatial ordering of the shared sample set
        total = warmup + steps
        which = slabs_of or (lambda s: slab_of_step(s, rank, world))
        slabs = [lens._object_points(bank_points(which(s), rank, world, lens=lens_name)).to(dev).contiguous() for s in range(total)]
        centres = [E.psf_centre(handle, 0.589, p, cpupil, pz, numerics=numerics) for p in slabs]
        n_pts = slabs[0].shape[0]

        def step(i):
            return E.psf_bank(handle, 0.589, slabs[i], pupil, pz, centres[i], KS, lens.pixel_size, numerics=numerics)

        for i in range(warmup):
            step(i)
        barrier()
        sampler = ClockSampler(local) if want_clocks else None
        if sampler:
            sampler.start()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        launches0 = E.launch_count()
        barrier()
        out = None
        for k in range(steps):
            flush.zero_()
            ev[k][0].record()
            out = step(warmup + k)
            ev[k][1].record()
        barrier()
        launches = E.launch_count() - launches0
        clocks = sampler.stop() if sampler else None
        step_ms = [a.elapsed_time(b) for a, b in ev]
        total_ms = max_over_ranks(sum(step_ms))
        assert torch.isfinite(out[0]).all() and float(out[0].max()) > 0.99
        rays_per_step = n_pts * SPP
        return {"value": world * rays_per_step * steps / (total_ms * 1e-3), "ms_per_step": total_ms / steps, "step_ms": step_ms,
                "launches": int(launches), "clocks": clocks, "rays_per_step": rays_per_step, "n_pts": n_pts,
                "per_gpu_rate": rays_per_step * steps / (sum(step_ms) * 1e-3)}, lens, (slabs, centres, pupil, pz, handle, rho, theta)

    def e2e_leg(lens, lens_name, e2e_steps):
        """The same metric end to end through the public API (PSFNet.psf_dp), host buffers in, host buffers out.
        Double-buffered like any producer / consumer loop: while the GPU works on step i the host prepares step i+1 and reads
        the PSFs of step i-1 out of pinned memory; every step's inputs cross H2D and every step's result is read on the host."""
        n_pts = GRID * GRID
        pinned_out = [torch.empty((n_pts, 2, KS, KS), dtype=torch.float32).pin_memory() for _ in range(2)]
        done = [torch.cuda.Event(), torch.cuda.Event()]
        # (call 0 warms up; the timed calls take the slabs of the first timed steps of `value`, so the two numbers are about the same work)
        host_pts = [bank_points(slab_of_step(max(0, args.warmup - 1) + s, rank, world), rank, world, lens=lens_name).pin_memory() for s in range(e2e_steps + 1)]
        host_sum = [0.0]

        def enqueue(i):
            L, R = lens.psf_dp(host_pts[i], ks=KS, spp=SPP)           # sampling + H2D + sort + centre + bank kernels
            pinned_out[i % 2][:, 0].copy_(L, non_blocking=True)
            pinned_out[i % 2][:, 1].copy_(R, non_blocking=True)
            done[i % 2].record()

        def consume(i):
            done[i % 2].synchronize()
            host_sum[0] += float(pinned_out[i % 2][:, :, KS // 2, KS // 2].sum())     # the host reads the result

        torch.manual_seed(99 + rank)
        torch.cuda.manual_seed(99)                                     # the shared sample set: the same draw on every rank
        enqueue(0)
        consume(0)
        barrier()
        t0 = time.perf_counter()
        enqueue(1)
        for i in range(2, e2e_steps + 1):
            enqueue(i)
            consume(i - 1)
        consume(e2e_steps)
        barrier()
        e2e_t = max_over_ranks(time.perf_counter() - t0)
        return world * n_pts * SPP * e2e_steps / e2e_t

    # ---- the contract workload --------------------------------------------------------------------------
    main, lens, (slabs, centres, pupil, pz, handle, rho, theta) = bank_leg(LENS, args.numerics, args.steps, args.warmup, want_clocks=True)
    value, n_pts, rays_per_step = main["value"], main["n_pts"], main["rays_per_step"]
    if args.quick:
        if rank == 0:
            os.write(real_stdout, (json.dumps({"quick": True, "numerics": args.numerics, "value": value, "ms_per_step": main["ms_per_step"],
                                               "step_ms": main["step_ms"], "lib": E._LIB_PATH}) + "\n").encode())
        if world > 1:
            dist.destroy_process_group()
        return
    e2e_steps = max(2, min(args.steps, 8))              # (the same slabs as the timed steps of `value`, when K <= 8)
    # the shared sample set of the end-to-end calls is drawn on the device (the reference draws it on the host and uploads it,
    # optics.py:483-487: that serial host work is identical on every rank and cost the 8-GPU end-to-end number 6 % in round 1);
    # the object points still come from host memory and the PSFs still go back to it, every step
    lens.sample_rng = "cuda"
    e2e_value = e2e_leg(lens, LENS, e2e_steps)
    h2d = n_pts * 3 * 4
    d2h = n_pts * 2 * KS * KS * 4

    # ---- the same workload in the mode that meets every north_star tolerance ----------------------------------------
    conformant = None
    if args.numerics != "strict" and not args.lean:
        c, clens, _ = bank_leg(LENS, "strict", max(2, min(args.steps, 4)), 1)
        clens.sample_rng = "cuda"
        c_e2e = e2e_leg(clens, LENS, 2)
        ach = FLOPS_PER_RAY_STRICT[LENS] * c["per_gpu_rate"] / 1e12
        conformant = {"metric": "rays/sec traced+splatted into DP L/R PSFs", "numerics": "strict", "value": c["value"], "unit": "rays/s",
                      "ms_per_step": c["ms_per_step"], "steps": len(c["step_ms"]), "dp_psfs_per_s": c["value"] / SPP,
                      "e2e": {"value": c_e2e, "unit": "rays/s", "steps": 2}, "tolerances_met": TOLERANCES["strict"],
                      "roofline": {"bound": "fp32", "achieved": ach, "peak": FP32_NOMINAL_TFLOPS, "unit": "TFLOP/s", "frac": ach / FP32_NOMINAL_TFLOPS,
                                   "kernel": "psf_bank_run_kernel<TraceStrictLoop>",
                                   "note": "%.0f flop/ray = the reference's own iteration counts on every surface (SURVEY 8d) incl. its IEEE "
                                           "divisions / square roots as one flop each; the kernel spends ~9 issue slots per division and ~7 "
                                           "per square root (no hardware IEEE div / sqrt), so the issue ports are 96 %% busy at this fraction "
                                           "(profiles/r02*_psf_bank_strict_ncu_summary.txt)" % FLOPS_PER_RAY_STRICT[LENS]}}
        del clens

    # ---- BASELINE configs[2]: the rf35mm bank over focus distances -------------------------------------------------------
    rf35 = None
    if LENS == "rf50mm" and not args.lean:
        flens = PSFNet(lens_file("rf35mm"), sensor_res=SENSOR_RES, kernel_size=KS, device=dev)
        sweep = []
        for foc in (-700.0, -1000.0, -2000.0, -5000.0):
            t0 = time.perf_counter()
            flens.refocus(foc)                                           # Lensgroup.refocus (optics.py:1170-1196): least-squares sensor position
            torch.cuda.synchronize()
            t_ref = time.perf_counter() - t0
            # the slab in focus and a far slab of the bank at this sensor position
            r, _, _ = bank_leg("rf35mm", args.numerics, 2, 1, d_sensor=flens.d_sensor, slabs_of=lambda s: (16, 27, 5)[s % 3])
            sweep.append({"focus_mm": foc, "d_sensor": flens.d_sensor, "refocus_ms": 1e3 * t_ref, "rays_per_s": r["value"], "ms_per_step": r["ms_per_step"]})
        tot_ms = sum(s["ms_per_step"] for s in sweep)
        rs, _, _ = bank_leg("rf35mm", "strict", 2, 1)
        rf35 = {"metric": "rays/sec traced+splatted into DP L/R PSFs, rf35mm (21 surfaces), bank slab at four focus distances", "unit": "rays/s",
                "value": world * rays_per_step * len(sweep) / (tot_ms * 1e-3), "numerics": args.numerics, "focus_sweep": sweep,
                "strict_rays_per_s": rs["value"],
                "roofline_frac_fp32_nominal": FLOPS_PER_RAY["rf35mm"] * (rays_per_step * len(sweep) / (tot_ms * 1e-3)) / 1e12 / FP32_NOMINAL_TFLOPS}
        del flens

    # ---- strong scaling: the FULL bank over the N ranks and ONE all-gather of it ------------------------------------------------
    strong = None
    if LENS == "rf50mm" and not args.lean:
        P = GRID * GRID * DEPTHS
        block = GRID * GRID
        # rank r takes the depth slabs r, r + N, r + 2N, ... (near and far slabs cost differently: contiguous blocks of depths would
        # leave the rank with the far half 15 % behind); world sizes that do not divide 32 fall back to contiguous point blocks
        strided = DEPTHS % world == 0
        if strided:
            my_slabs = sharding.dealt_groups(DEPTHS, rank, world)
            allpts = torch.cat([lens._object_points(bank_points(s, lens=LENS)) for s in my_slabs], 0).to(dev).contiguous()
        else:
            sl = sharding.shard_slice(P, rank, world)
            allpts = torch.cat([lens._object_points(bank_points(s, lens=LENS)) for s in range(DEPTHS)], 0)[sl].to(dev).contiguous()
        cp = (pupil[:2048] * 0.25).contiguous()
        local_out = torch.empty((allpts.shape[0], 2, KS, KS), dtype=torch.float32, device=dev)
        gathered = torch.empty((P, 2, KS, KS), dtype=torch.float32, device=dev) if (world > 1 and strided) else None
        bank = torch.empty((P, 2, KS, KS), dtype=torch.float32, device=dev) if (world > 1 and strided) else None

        def full_bank():
            for b0 in range(0, allpts.shape[0], block):
                p = allpts[b0:b0 + block]
                c = E.psf_centre(handle, 0.589, p, cp, pz, numerics=args.numerics)
                L, R = E.psf_bank(handle, 0.589, p, pupil, pz, c, KS, lens.pixel_size, numerics=args.numerics)
                local_out[b0:b0 + block, 0], local_out[b0:b0 + block, 1] = L, R

        def assemble():
            """The ONE collective of the path (north_star): every rank ends up with the whole 462 MB bank, in depth-major order."""
            if world == 1:
                return local_out
            if not strided:
                return sharding.gather_blocks(local_out, P)
            return sharding.gather_dealt(local_out, DEPTHS, out=bank, scratch=gathered)

        full_bank()                                                    # warm
        assemble()
        barrier()
        e0, e1, g0, g1 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
        e0.record()
        full_bank()
        e1.record()
        barrier()                                                      # (so that the gather's time is the gather's, not the wait for the slowest rank)
        g0.record()
        whole = assemble()
        g1.record()
        barrier()
        t_bank, t_gather = max_over_ranks(e0.elapsed_time(e1)), max_over_ranks(g0.elapsed_time(g1))
        nbytes = P * 2 * KS * KS * 4
        strong = {"metric": "seconds per full PSF bank (131072 points x 2 M rays), depth slabs dealt round-robin to the ranks, assembled on every rank",
                  "scaling": "strong", "value": (t_bank + (t_gather if world > 1 else 0.0)) * 1e-3, "unit": "s", "higher_is_better": False, "bank_s": t_bank * 1e-3,
                  "gather_ms": t_gather if world > 1 else None, "bank_bytes": nbytes,
                  "gather_GBps_received_per_gpu": (nbytes * (world - 1) / world / (t_gather * 1e-3) / 1e9) if world > 1 else None,
                  "gather_includes": "one NCCL all_gather_into_tensor + the device-side reorder into depth-major order" if world > 1 else None,
                  "nvlink_GBps_per_direction_measured": 770.0, "rays_per_s": P * SPP / ((t_bank + (t_gather if world > 1 else 0.0)) * 1e-3),
                  "numerics": args.numerics}
        if rank == 0:
            strong["bank_max_psf"] = float(whole.amax())
            strong["bank_slab_maxima_all_one"] = bool((whole.view(DEPTHS, -1).amax(1) > 0.999).all())
        del local_out, gathered, bank, allpts, whole

    # ---- BASELINE configs[3] / [4]: batch-16 render sharded by image; focal-stack generation -----------------------------------
    rb, rh, rw = 16, 1024, 1536
    g = torch.Generator(device=dev).manual_seed(7)
    rlens = PSFNet(lens_file(LENS), sensor_res=(rh, rw), kernel_size=KS, device=dev)
    render_sharded = datagen = None
    if not args.lean:
        sl = sharding.shard_slice(rb, rank, world)
        nloc = sl.stop - sl.start
        img = torch.rand((nloc, 3, rh, rw), device=dev, generator=g)
        low = torch.rand((nloc, 1, rh // 64 + 2, rw // 64 + 2), device=dev, generator=g)
        depth = -(torch.nn.functional.interpolate(low, size=(rh, rw), mode="bilinear", align_corners=False) * 9750 + 250)
        foc = torch.full((nloc,), -1000.0, device=dev)
        rlens.render(img, depth, foc)                                  # warm-up at the timed shape (band buffers, image records)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = rlens.render(img, depth, foc)
        e1.record()
        barrier()
        t = max_over_ranks(e0.elapsed_time(e1))
        render_sharded = {"metric": "pixels/s, PSFNet.render of 16 x 3 x 1024 x 1536 (NYUv2-shaped RGB-D batch), images sharded over the ranks, no exchange",
                          "value": rb * rh * rw / (t * 1e-3), "unit": "pixels/s", "ms": t, "images_per_rank": nloc, "scaling": "strong"}
        # The DfDP training step's data side (configs/dfdp_by_sdirt_rf50mm.yml: res 512 x 768, bs 4, n_stack 1): all-in-focus image + depth
        # -> dual-pixel training images with gamma + noise + clip (2_dfdp_net.py:161-173), then the reference's own DfDP network consumes
        # the batch (dfdp/basenet.py:18-49, forward incl. its losses).  Weak scaling: every rank generates and consumes its own batch.
        fh, fw, fb = 512, 768, 4
        flens = PSFNet(lens_file(LENS), sensor_res=(fh, fw), kernel_size=KS, device=dev)
        aif = torch.rand((fb, 3, fh, fw), device=dev, generator=g)
        low = torch.rand((fb, 1, fh // 64 + 2, fw // 64 + 2), device=dev, generator=g)
        depth_m = torch.nn.functional.interpolate(low, size=(fh, fw), mode="bilinear", align_corners=False) * 9.0 + 0.3     # metres
        ffoc = torch.full((fb,), -1000.0, device=dev)
        flens.render_focal_stack(aif, -depth_m * 1e3, ffoc, train=True)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        stack = flens.render_focal_stack(aif, -depth_m * 1e3, ffoc, train=True)
        e1.record()
        barrier()
        t = max_over_ranks(e0.elapsed_time(e1))
        datagen = {"metric": "dual-pixel training images / s, render_focal_stack(train=True) at 512 x 768 (the DfDP config's resolution), 4 scenes per rank",
                   "value": world * fb / (t * 1e-3), "unit": "images/s", "ms_per_batch": t, "scaling": "weak"}
        try:                                                            # the consumer: the UNMODIFIED reference network, when it travelled
            Rn = _ref_runner()
            if Rn is None:
                raise RuntimeError("baseline/_ref absent")
            net = Rn.make_basenet(dev)
            net.train()
            feed = {"gt_depth": depth_m.clone(), "AiF_img": aif, "stack_rgb_img": stack}
            with torch.no_grad():
                net(feed)
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                stack = flens.render_focal_stack(aif, -depth_m * 1e3, ffoc, train=True)
                feed["stack_rgb_img"], feed["gt_depth"] = stack, depth_m.clone()
                losses, _ = net(feed)
                e1.record()
                barrier()
            t2 = max_over_ranks(e0.elapsed_time(e1))
            datagen["with_dfdp_forward"] = {"ms_per_batch": t2, "images_per_s": world * fb / (t2 * 1e-3), "generation_share": t / t2,
                                            "loss_depth_est": float(losses["depth_est"]),
                                            "note": "generation + the reference's Basenet('dfdp') forward with its losses (dfdp/basenet.py), autocast as the reference decorates it"}
            del net
        except Exception as ex:
            datagen["with_dfdp_forward"] = {"unavailable": repr(ex)[:300]}
        assert torch.isfinite(stack).all() and torch.isfinite(out).all()
        del img, depth, out, aif, depth_m, stack, flens

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (psf_bank_run_kernel): FP32 pipe --------------------------------
    blocks, threads, iters = 148 * 8, 256, 1 << 15
    E.fp32_peak_probe(dev, blocks, threads, 2048)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    E.fp32_peak_probe(dev, blocks, threads, iters)
    b.record()
    torch.cuda.synchronize()
    fp32_probe = blocks * threads * iters * 16.0 / (a.elapsed_time(b) * 1e-3) / 1e12
    achieved = FLOP_PER_RAY * main["per_gpu_rate"] / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    traffic = None
    try:                                   # dram__bytes_read + write of one psf_bank_run_kernel launch of THIS workload (ncu)
        traffic = json.load(open(os.path.join(ROOT, "profiles", "bank_kernel_traffic.json")))[LENS + "_bytes_per_launch"]
    except Exception:
        pass
    roofline = {"bound": "fp32", "achieved": achieved, "peak": FP32_NOMINAL_TFLOPS, "unit": "TFLOP/s", "frac": achieved / FP32_NOMINAL_TFLOPS,
                "peak_source": "nominal 148 SMs x 128 lanes x 2 flop x 1.965 GHz (MEASURED_PEAKS.json holds no FP32 entry, hbm_gbs=%s; an "
                               "FFMA micro-benchmark run in this process reads %.1f TFLOP/s)" % (peaks.get("hbm_gbs"), fp32_probe),
                "fp32_probe_tflops": fp32_probe, "traffic": traffic, "kernel": "psf_bank_run_kernel<TraceSig<Sig_%s, ...>>" % LENS,
                "note": "the dominant kernel (99.7 %% of a step, profiles/) is FP32-issue bound, not HBM/tensor bound: it reads "
                        "12 B/point + the L2-resident 16 MB sample set and writes 7 KB/point. achieved = %.0f flop/ray (SURVEY 8d: the "
                        "reference algorithm's minimal per-ray count) x rays per launch / CUDA-event time of the step's launches (dp_lut + "
                        "bank + finalize); traffic = ncu dram bytes per launch. The count is the REFERENCE algorithm's (Newton iterations "
                        "on every surface); the kernel executes fewer flops per ray than that (closed-form sphere roots), so the fraction "
                        "is useful-work-equivalent, not pipe utilisation: ncu's own pipe view is in profiles/." % FLOP_PER_RAY}

    # ---- secondary numbers (not the contract metric): other numerics modes, and the HBM-bound render kernel --------------
    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    modes = {}
    sub = slabs[args.warmup][::16].contiguous()                      # 256 points of the first timed slab
    subc = centres[args.warmup][::16].contiguous()
    for mode in ("strict", "hybrid", "adaptive", "fast"):
        ms = timed(lambda: E.psf_bank(handle, 0.589, sub, pupil, pz, subc, KS, lens.pixel_size, numerics=mode), 2)
        modes[mode] = sub.shape[0] * SPP / (ms * 1e-3)
    rb2 = 2
    img = torch.rand((rb2, 3, rh, rw), device=dev, generator=g)
    psf = torch.rand((rb2, rh, rw, 2, KS, KS), device=dev, generator=g, dtype=torch.float16)
    ms = timed(lambda: E.render_local_psf(img, psf, KS, tone=3), 5)
    rbytes = rb2 * rh * rw * (2 * KS * KS * 2 + 3 * 4 + 6 * 4)
    hbm = peaks.get("hbm_gbs", 6650.0)
    render = {"metric": "pixels/s, spatially varying DP render (explicit fp16 per-pixel PSFs, degamma+gamma fused)",
              "value": rb2 * rh * rw / (ms * 1e-3), "unit": "pixels/s", "shape": [rb2, 3, rh, rw], "ks": KS,
              "roofline": {"bound": "hbm", "achieved": rbytes / (ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                           "frac": rbytes / (ms * 1e-3) / 1e9 / hbm, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}}
    del psf
    # BASELINE config 4 shape: PSFNet.render end to end (banded; the PSF MLP of a band is ONE tcgen05 kernel -- csrc/mlp_fused.cuh --
    # the cuBLAS route is timed next to it)
    low = torch.rand((rb2, 1, rh // 64 + 2, rw // 64 + 2), device=dev, generator=g)
    depth = -(torch.nn.functional.interpolate(low, size=(rh, rw), mode="bilinear", align_corners=False) * 9750 + 250)
    foc = torch.full((rb2,), -1000.0, device=dev)
    ms = timed(lambda: rlens.render(img, depth, foc), 3)
    rlens.mlp_engine = "cublas"
    ms_cublas = timed(lambda: rlens.render(img, depth, foc), 3)
    mlp_flop_px = 2 * 2 * (3 * 128 + 128 * 512 + 8 * 512 * 512 + 512 * KS * KS)
    tpeak = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1373.0))
    render_psfnet = {"metric": "pixels/s, PSFNet.render (coordinate grid -> PSF MLP both sides -> normalise -> degamma -> DP "
                               "gather-convolution -> gamma -> clip), banded, fused tcgen05 MLP kernel",
                     "value": rb2 * rh * rw / (ms * 1e-3), "unit": "pixels/s", "shape": [rb2, 3, rh, rw], "ks": KS,
                     "cublas_route_pixels_per_s": rb2 * rh * rw / (ms_cublas * 1e-3),
                     "roofline": {"bound": "tensor", "achieved": rb2 * rh * rw * mlp_flop_px / (ms * 1e-3) / 1e12, "peak": tpeak,
                                  "unit": "TFLOP/s", "frac": rb2 * rh * rw * mlp_flop_px / (ms * 1e-3) / 1e12 / tpeak,
                                  "note": "9.56 MFLOP/pixel of 16-bit MMA dominate (mlp_fused_pred_kernel, CTA pairs); the time is "
                                          "the whole render call, convolution included; peak = measured dense 16-bit matmul "
                                          "(MEASURED_PEAKS.json, sustained)"}}
    del img, depth, rlens

    # ---- the unmodified reference on this box: eager on the same GPU, and its CPU path on the host cores (N = 1 only) ----------------
    cpu = eager = None
    R = _ref_runner() if (world == 1 and not args.no_cpu) else None
    if R is not None:
        try:
            eager = {"kind": "reference", "device": "cuda:0 (eager torch, the reference's own code path with device='cuda')"}
            rl = R.make_psfnet(LENS, SENSOR_RES, KS, dev)
            pts = bank_points(slab_of_step(args.warmup))[::16][:256].contiguous()
            times, _ = R.time_psf(rl, pts, KS, 65536, repeats=3, seed=5, sync=torch.cuda.synchronize)
            eager["psf_diff_rays_per_s"] = 256 * 65536 / min(times[1:])
            eager["psf_diff_sample"] = "256 points x 65536 rays per call (left PSFs only), best of 2 after a warm-up call"
            rr = R.make_psfnet(LENS, (rh, rw), KS, dev)
            gi = torch.Generator(device=dev).manual_seed(3)
            img1 = torch.rand((1, 3, rh, rw), device=dev, generator=gi)
            dep1 = -(torch.rand((1, 1, rh, rw), device=dev, generator=gi) * 9000 + 300)
            times, _ = R.time_render(rr, img1, dep1, torch.full((1,), -1000.0, device=dev), repeats=3, sync=torch.cuda.synchronize)
            eager["render_pixels_per_s"] = rh * rw / min(times[1:])
            eager["render_sample"] = "PSFNet.render (psfnet.py:645-714) of 1 x 3 x 1024 x 1536, best of 2 after a warm-up call"
            del rl, rr, img1, dep1
            torch.cuda.empty_cache()
        except Exception as ex:                                         # the reference arm must never take the bench line down
            eager = dict(eager or {}, error=repr(ex)[:300])
        v, times, cores = reference_cpu_rays_per_s(R, 2, warmup=1)
        cpu = {"value": v, "unit": "rays/s", "cores": cores, "kind": "reference",
               "sample": f"{REF_SAMPLE[0]} points x {REF_SAMPLE[1]} rays per call, 2 timed calls after a warm-up: the unmodified reference's "
                         f"Lensgroup.psf_diff on torch CPU ({cores} threads); one call = the LEFT PSFs only"}
    elif world == 1 and not args.no_cpu:
        v, times, cores = port_cpu_rays_per_s(64, 131072, repeats=2)
        cpu = {"value": v, "unit": "rays/s", "cores": cores, "kind": "port",
               "sample": "64 points x 131072 rays (baseline/_ref absent: oracle/dp_oracle.py, numpy, one process per core)"}

    line = json.dumps({
        "metric": "rays/sec traced+splatted into DP L/R PSFs", "value": value, "unit": "rays/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.numerics),
        "dp_psfs_per_s": value / SPP,
        "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "api": "PSFNet.psf_dp(pinned host points) -> pinned host (L, R), double-buffered; the shared pupil-sample "
                                           "set of each call is drawn on the device (Lensgroup.sample_rng = 'cuda')"},
        "gpu_launches": main["launches"],
        "clocks": main["clocks"],
        "roofline": roofline,
        "cpu_baseline": cpu,
        "eager_gpu_baseline": eager,
        "conformant": conformant,
        "rf35mm": rf35,
        "strong": strong,
        "render_sharded": render_sharded,
        "datagen": datagen,
        "numerics_modes_rays_per_s": modes,
        "render": render,
        "render_psfnet": render_psfnet,
    })
    sys.stdout.flush()
    os.write(real_stdout, (line + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="sdirt_b200", choices=["sdirt_b200", "reference"])
    ap.add_argument("--numerics", default="adaptive", choices=["strict", "hybrid", "adaptive", "fast"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / eager_gpu_baseline legs")
    ap.add_argument("--lean", action="store_true", help="contract workload only (no conformant / rf35mm / strong / render_sharded legs)")
    ap.add_argument("--quick", action="store_true", help="kernel-only number and exit (tuning runs; not a contract line)")
    ap.add_argument("--lens", default="rf50mm", choices=sorted(HFOV), help="prescription of the main leg (rf50mm = the headline config)")
    args = ap.parse_args()
    global LENS, FLOP_PER_RAY
    LENS, FLOP_PER_RAY = args.lens, FLOPS_PER_RAY[args.lens]
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py — rays/s traced + splatted into dual-pixel PSFs (BASELINE.json metric), one JSON line.

Workload (BASELINE.json configs[1]): the rf50mm F/4 PSF bank for PSFNet fitting, 64x64 field points x 32 depths,
2 M rays per point.  One "step" = one depth slab of that bank: 4096 points x 2 M rays = 8.4e9 rays through
sample -> 12-surface trace -> DP weights -> bilinear splat -> normalise, i.e. 4096 (L, R) PSF pairs.  Consecutive
steps walk the 32 depth slabs with stride 11 (near, in-focus and far depths alike).  With N GPUs the job is an N times denser
field sampling of the same bank: every rank works on its own (sub-cell shifted) 64x64 field grid of the step's slab (weak
scaling, equal work per rank, no collective on the data path); `value` is the whole-job rays/s = N x slab rays /
max-over-ranks device time.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--numerics strict|hybrid|fast] [--impl reference]

`--impl reference` times the CPU restatement of the reference's own algorithm (oracle/dp_oracle.py, all host
cores) on bounded samples of the same workload and prints the same JSON shape with "impl": "reference".
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
LENS = "rf50mm"          # --lens rf35mm switches to BASELINE config 3's prescription (same bank shape)
# algorithmic flop/ray with the reference's minimal per-ray Newton counts (FMA = 2)
FLOPS_PER_RAY = {"rf50mm": 2.1e3, "rf35mm": 3.3e3}
FLOP_PER_RAY = FLOPS_PER_RAY[LENS]
# The reference's own setup values (tests/test_api_gpu.py pins the engine's against them):
HFOV = {"rf50mm": 0.40959781408309937, "rf35mm": 0.5514792203903198}
PUPIL = {"rf50mm": (22.51324462890625, 6.019352912902832), "rf35mm": (14.338210105895996, 4.767455577850342)}
D_SENSOR = {"rf50mm": 62.25, "rf35mm": 80.447}


# ----------------------------------------------------------------------------------------------------
# workload
# ----------------------------------------------------------------------------------------------------
def bank_points(slab, rank=0, world=1):
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
    d_min, d_max, ds = -200.0, -20000.0, D_SENSOR[LENS]
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
# CPU leg: the oracle (a restatement of the reference's torch-CPU algorithm) on all host cores
# ----------------------------------------------------------------------------------------------------
def _cpu_chunk(args):
    """Worker: trace + splat `n_pts` points x one chunk of pupil samples with the numpy oracle."""
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


def cpu_rays_per_s(n_pts, spp, repeats=1, slab0=0):
    """Time the oracle on `n_pts` points x `spp` rays per repeat, points spread over all host cores."""
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
    return [n_pts * spp / t for t in times], times, cores


def run_reference(args):
    """`--impl reference`: K timed steps, each a bounded sample (64 points x 32768 rays) of the slab workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_pts, spp = 64, 131072
    rates, times, cores = cpu_rays_per_s(n_pts, spp, repeats=args.warmup + args.steps, slab0=0)
    rates, times = rates[args.warmup:], times[args.warmup:]
    value = n_pts * spp * len(times) / sum(times)
    sample = f"{n_pts} points x {spp} rays per step of the {GRID}x{GRID}x{DEPTHS} x 2M-ray bank"
    print(json.dumps({
        "impl": "reference", "metric": "rays/sec traced+splatted into DP L/R PSFs", "value": value, "unit": "rays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config("cpu"),
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "dp_psfs_per_s": value / spp,
    }))


def workload_config(numerics):
    return {"workload": f"{LENS} F/4 PSF bank, {GRID}x{GRID} field x {DEPTHS} depths, {SPP} rays/point, ks={KS}, "
                        f"sensor {SENSOR_RES[0]}x{SENSOR_RES[1]}; step = one depth slab ({GRID * GRID} points), slabs visited "
                        f"with stride 11 over the {DEPTHS} depths; N GPUs = N sub-cell-shifted copies of the field grid at the "
                        f"same slab (N times denser field sampling, equal work per rank)",
            "lens": LENS, "points_per_step": GRID * GRID, "rays_per_point": SPP, "ks": KS, "numerics": numerics,
            "l2": "L2 flushed (256 MiB write) between timed steps; the 16 MB shared pupil-sample set is re-read from L2 "
                  "by every point inside a step by design"}


# ----------------------------------------------------------------------------------------------------
# GPU leg
# ----------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from sdirt_b200 import _engine as E, lens_file
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
        # torchrun pins OMP_NUM_THREADS=1; the host side of the end-to-end path (the reference's CPU RNG + polar transform of
        # the shared sample set, optics.py:483-487) may use this rank's share of the cores
        torch.set_num_threads(max(1, (os.cpu_count() or 1) // world))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    lens = PSFNet(lens_file(LENS), sensor_res=SENSOR_RES, kernel_size=KS, device=dev)
    lens.numerics = args.numerics
    handle = lens._engine_lens()
    pz, pr = lens.entrance_pupil()

    # ---- device-resident inputs for the kernel-only number ------------------------------------------
    torch.manual_seed(1234)                                        # same pupil samples on every rank (optics.py:483-490)
    theta = torch.rand(SPP) * 2 * np.pi
    rho = torch.sqrt(torch.rand(SPP) * pr ** 2)
    pupil = torch.stack((rho * torch.cos(theta), rho * torch.sin(theta)), 1).to(dev).contiguous()
    cpupil = (pupil[:2048] * 0.25).contiguous()
    pupil = E.pupil_sort(pupil, pr)                                # one-off spatial ordering of the shared sample set
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    total_steps = args.warmup + args.steps
    slabs = [lens._object_points(bank_points(slab_of_step(s, rank, world), rank, world)).to(dev).contiguous() for s in range(total_steps)]
    centres = [E.psf_centre(handle, 0.589, p, cpupil, pz, numerics=args.numerics) for p in slabs]
    n_pts = slabs[0].shape[0]

    def step(i):
        return E.psf_bank(handle, 0.589, slabs[i], pupil, pz, centres[i], KS, lens.pixel_size, numerics=args.numerics)

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = E.launch_count()
    barrier()
    for k in range(args.steps):
        flush.zero_()
        ev[k][0].record()
        out = step(args.warmup + k)
        ev[k][1].record()
    barrier()
    launches = E.launch_count() - launches0
    clocks = sampler.stop()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    rays_per_step = n_pts * SPP
    value = world * rays_per_step * args.steps / (total_ms * 1e-3)
    assert torch.isfinite(out[0]).all() and float(out[0].max()) > 0.99
    if args.quick:
        if rank == 0:
            os.write(real_stdout, (json.dumps({"quick": True, "numerics": args.numerics, "value": value, "ms_per_step": total_ms / args.steps,
                                               "step_ms": step_ms, "lib": E._LIB_PATH}) + "\n").encode())
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- end to end through the public API, host buffers in, host buffers out -------------------------
    # Double-buffered like any producer/consumer loop: while the GPU works on step i the host draws the sample set of step
    # i+1 (the reference's CPU RNG) and reads the PSFs of step i-1 out of pinned memory; every step's inputs cross H2D and
    # every step's result is read on the host inside the timed region.
    pinned_out = [torch.empty((n_pts, 2, KS, KS), dtype=torch.float32).pin_memory() for _ in range(2)]
    done = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_steps = max(2, min(args.steps, 4))
    host_pts = [bank_points(slab_of_step(s, rank, world), rank, world) for s in range(e2e_steps + 1)]
    host_sum = [0.0]

    def enqueue(i):
        torch.manual_seed(99 + i)
        L, R = lens.psf_dp(host_pts[i], ks=KS, spp=SPP)           # CPU sampling + H2D + sort + centre + bank kernels
        pinned_out[i % 2][:, 0].copy_(L, non_blocking=True)
        pinned_out[i % 2][:, 1].copy_(R, non_blocking=True)
        done[i % 2].record()

    def consume(i):
        done[i % 2].synchronize()
        host_sum[0] += float(pinned_out[i % 2][:, :, KS // 2, KS // 2].sum())     # the host reads the result

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
    e2e_t = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = world * rays_per_step * e2e_steps / float(e2e_t.item())
    h2d = SPP * 2 * 4 + 2048 * 2 * 4 + n_pts * 3 * 4
    d2h = n_pts * 2 * KS * KS * 4

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (psf_bank_kernel): FP32 pipe --------------------------------
    blocks, threads, iters = 148 * 8, 256, 1 << 15
    E.fp32_peak_probe(dev, blocks, threads, 2048)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    E.fp32_peak_probe(dev, blocks, threads, iters)
    b.record()
    torch.cuda.synchronize()
    fp32_peak = blocks * threads * iters * 16.0 / (a.elapsed_time(b) * 1e-3) / 1e12
    per_gpu_rate = rays_per_step * args.steps / (sum(step_ms) * 1e-3)
    achieved = FLOP_PER_RAY * per_gpu_rate / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    traffic = None
    try:                                   # dram__bytes_read + write of one psf_bank_run_kernel launch of THIS workload (ncu)
        if LENS == "rf50mm":
            traffic = json.load(open(os.path.join(ROOT, "profiles", "bank_kernel_traffic.json")))["bytes_per_launch"]
    except Exception:
        pass
    roofline = {"bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
                "traffic": traffic, "kernel": "psf_bank_run_kernel<TraceSig<Sig_%s, ...>>" % LENS,
                "note": "the dominant kernel (99.7 %% of a step, profiles/) is FP32-issue bound, not HBM/tensor bound: it reads "
                        "12 B/point + the L2-resident 16 MB sample set and writes 7 KB/point. peak = FFMA micro-benchmark "
                        "measured in this run (MEASURED_PEAKS.json has no fp32 entry; its hbm_gbs=%s). achieved = %.0f "
                        "flop/ray (SURVEY 8d: the reference algorithm's minimal per-ray count) x rays per launch / CUDA-event "
                        "time of the step's launches (dp_lut + bank + finalize); traffic = ncu dram bytes per launch. The count is "
                        "the REFERENCE algorithm's (Newton iterations on every surface); the kernel reaches a fraction near 1 "
                        "because it executes fewer flops per ray than that (closed-form sphere roots) and issues its FMA-pipe "
                        "arithmetic for two rays at once (FFMA2), not because it exceeds the pipe: ncu's own pipe utilisation is in profiles/."
                        % (peaks.get("hbm_gbs"), FLOP_PER_RAY)}

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
    torch.manual_seed(1234)
    raw = torch.stack((rho * torch.cos(theta), rho * torch.sin(theta)), 1).to(dev).contiguous()
    for mode in ("strict", "hybrid", "adaptive", "fast"):
        pup = pupil
        ms = timed(lambda: E.psf_bank(handle, 0.589, sub, pup, pz, subc, KS, lens.pixel_size, numerics=mode), 2)
        modes[mode] = sub.shape[0] * SPP / (ms * 1e-3)
    rb, rh, rw = 2, 1024, 1536
    g = torch.Generator(device=dev).manual_seed(0)
    img = torch.rand((rb, 3, rh, rw), device=dev, generator=g)
    psf = torch.rand((rb, rh, rw, 2, KS, KS), device=dev, generator=g, dtype=torch.float16)
    ms = timed(lambda: E.render_local_psf(img, psf, KS, tone=3), 5)
    rbytes = rb * rh * rw * (2 * KS * KS * 2 + 3 * 4 + 6 * 4)
    hbm = peaks.get("hbm_gbs", 6650.0)
    render = {"metric": "pixels/s, spatially varying DP render (explicit fp16 per-pixel PSFs, degamma+gamma fused)",
              "value": rb * rh * rw / (ms * 1e-3), "unit": "pixels/s", "shape": [rb, 3, rh, rw], "ks": KS,
              "roofline": {"bound": "hbm", "achieved": rbytes / (ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                           "frac": rbytes / (ms * 1e-3) / 1e9 / hbm, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}}
    del psf, img
    # BASELINE config 4 shape: PSFNet.render end to end (banded; the PSF MLP of a band is ONE tcgen05 kernel -- csrc/mlp_fused.cuh --
    # the cuBLAS route is timed next to it)
    rlens = PSFNet(lens_file(LENS), sensor_res=(rh, rw), kernel_size=KS, device=dev)
    img = torch.rand((rb, 3, rh, rw), device=dev, generator=g)
    low = torch.rand((rb, 1, rh // 64 + 2, rw // 64 + 2), device=dev, generator=g)
    depth = -(torch.nn.functional.interpolate(low, size=(rh, rw), mode="bilinear", align_corners=False) * 9750 + 250)
    foc = torch.full((rb,), -1000.0, device=dev)
    ms = timed(lambda: rlens.render(img, depth, foc), 3)
    rlens.mlp_engine = "cublas"
    ms_cublas = timed(lambda: rlens.render(img, depth, foc), 3)
    mlp_flop_px = 2 * 2 * (3 * 128 + 128 * 512 + 8 * 512 * 512 + 512 * KS * KS)
    tpeak = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1373.0))
    render_psfnet = {"metric": "pixels/s, PSFNet.render (coordinate grid -> PSF MLP both sides -> normalise -> degamma -> DP "
                               "gather-convolution -> gamma -> clip), banded, fused tcgen05 MLP kernel",
                     "value": rb * rh * rw / (ms * 1e-3), "unit": "pixels/s", "shape": [rb, 3, rh, rw], "ks": KS,
                     "cublas_route_pixels_per_s": rb * rh * rw / (ms_cublas * 1e-3),
                     "roofline": {"bound": "tensor", "achieved": rb * rh * rw * mlp_flop_px / (ms * 1e-3) / 1e12, "peak": tpeak,
                                  "unit": "TFLOP/s", "frac": rb * rh * rw * mlp_flop_px / (ms * 1e-3) / 1e12 / tpeak,
                                  "note": "9.56 MFLOP/pixel of 16-bit MMA dominate (mlp_fused_pred_kernel, CTA pairs); the time is "
                                          "the whole render call, convolution included; peak = measured dense 16-bit matmul "
                                          "(MEASURED_PEAKS.json, sustained)"}}
    del img, depth, rlens

    # ---- CPU baseline on this box's host cores (bounded sample) --------------------------------------
    cpu = None
    if not args.no_cpu:
        rates, times, cores = cpu_rays_per_s(64, 131072, repeats=2)
        cpu = {"value": rates[-1], "unit": "rays/s", "cores": cores, "kind": "port",
               "sample": "64 points x 131072 rays of depth slab 1 (oracle/dp_oracle.py, numpy, one process per core)"}

    line = json.dumps({
        "metric": "rays/sec traced+splatted into DP L/R PSFs", "value": value, "unit": "rays/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.numerics),
        "dp_psfs_per_s": value / SPP,
        "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "api": "PSFNet.psf_dp(host points) -> pinned host (L, R), double-buffered"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
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
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--quick", action="store_true", help="kernel-only number and exit (tuning runs; not a contract line)")
    ap.add_argument("--lens", default="rf50mm", choices=sorted(HFOV), help="prescription (rf50mm = the headline config)")
    args = ap.parse_args()
    global LENS, FLOP_PER_RAY
    LENS, FLOP_PER_RAY = args.lens, FLOPS_PER_RAY[args.lens]
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()

"""Host-side mirror of the reference's `Lensgroup` (deeplens/optics.py) for the dual-pixel hot path.

Same public names, arguments and return shapes as the reference methods used by `1_fit_psfnet.py`,
`2_dfdp_net.py` and `PSFNet`; every ray-level computation is executed by libsdirt_engine on the GPU:

    sample_from_points  optics.py:460-494      -> sdirt_sample_rays (compat) / fused in sdirt_psf_bank
    trace / trace2sensor optics.py:601-717     -> sdirt_trace_rays
    psf_center          optics.py:889-914      -> sdirt_psf_centre
    psf / psf_diff / psf_rgb / psf_map :916-1041 -> sdirt_psf_centre + sdirt_psf_bank (rays never materialised)
    refocus / calc_fov / entrance_pupil :1170-1396 -> the same tiny setup traces, through sdirt_trace_rays

Random pupil samples are drawn from torch's CPU generator in the reference's order (optics.py:483-484, 900), so a
seeded call sees exactly the reference's rays.  Plotting / lens-design helpers are out of scope.
"""
import json

import numpy as np
import torch

from .. import _engine as E
from .basics import DEFAULT_WAVE, DEPTH, DEVICE, EPSILON, GEO_SPP, WAVE_RGB, DeepObj, Material, Ray
from .monte_carlo import forward_integral
from .surfaces import Aspheric


class Lensgroup(DeepObj):
    def __init__(self, filename=None, sensor_res=(1024, 1024), use_roc=False, post_computation=True, device=DEVICE):
        self.device = torch.device(device) if not isinstance(device, torch.device) else device
        self.numerics = None            # None -> engine default ('strict'); 'hybrid' / 'fast' see include/sdirt_engine.h
        self.newton = "per_ray"
        self._handle, self._handle_key = None, None
        if filename is not None:
            self.lens_name = filename
            self.load_file(filename, use_roc, sensor_res)
        else:
            self.sensor_res = sensor_res
            self.surfaces = []
            self.materials = []

    # ==============================================================================================
    # IO and bookkeeping (optics.py:118-201, 2145-2198)
    # ==============================================================================================
    def load_file(self, filename, use_roc, sensor_res):
        if filename[-5:] == ".json":
            self.read_lens_json(filename)
        else:
            raise Exception("File format not supported.")       # the reference's .txt reader does not exist
        self.find_aperture()
        self.prepare_sensor(sensor_res)
        self.diff_surf_range = self.find_diff_surf()
        self.post_computation()

    def read_lens_json(self, filename="./test.json"):
        self.surfaces, self.materials = [], []
        with open(filename, "r") as f:
            data = json.load(f)
        for sd in data["surfaces"]:
            if sd["type"] == "Aspheric":
                s = Aspheric(r=sd["r"], d=sd["d"], c=sd["c"], k=sd["k"], ai=sd["ai"], mat1=sd["mat1"], mat2=sd["mat2"],
                             device=self.device)
            elif sd["type"] in ("Stop", "Spheric"):
                s = Aspheric(r=sd["r"], d=sd["d"], c=sd["c"], mat1=sd["mat1"], mat2=sd["mat2"], device=self.device)
            else:
                raise Exception("Surface type not implemented.")
            self.surfaces.append(s)
            self.materials.append(Material(sd["mat1"]))
        self.materials.append(Material(sd["mat2"]))
        self.r_last = data["r_last"]
        self.d_sensor = data["d_sensor"]

    def write_lens_json(self, filename="./test.json"):
        data = {"foclen": self.foclen, "fnum": self.fnum, "r_last": self.r_last, "d_sensor": self.d_sensor,
                "sensor_size": self.sensor_size, "surfaces": []}
        for i, s in enumerate(self.surfaces):
            sd = s.surf_dict()
            nxt = self.surfaces[i + 1].d.item() if i < len(self.surfaces) - 1 else self.d_sensor
            sd["d_next"] = nxt - s.d.item()
            data["surfaces"].append(sd)
        with open(filename, "w") as f:
            json.dump(data, f, indent=4)

    def load_external(self, surfaces, materials, r_last, d_sensor):
        self.surfaces, self.materials, self.r_last, self.d_sensor = surfaces, materials, r_last, d_sensor

    def prepare_sensor(self, sensor_res=[512, 512], sensor_size=[24., 36.], sensor_directly=True):
        sensor_res = [sensor_res, sensor_res] if isinstance(sensor_res, int) else sensor_res
        self.sensor_res = sensor_res
        H, W = sensor_res
        if sensor_size is None:
            self.sensor_size = [2 * self.r_last * H / np.sqrt(H ** 2 + W ** 2), 2 * self.r_last * W / np.sqrt(H ** 2 + W ** 2)]
        else:
            self.sensor_size = sensor_size
            self.r_last = np.sqrt(sensor_size[0] ** 2 + sensor_size[1] ** 2) / 2
        assert self.sensor_size[0] / self.sensor_size[1] == H / W, "Pixel is not square."
        self.pixel_size = self.sensor_size[0] / sensor_res[0]

    def post_computation(self):
        self.find_aperture()
        self.hfov = self.calc_fov()
        self.foclen = self.calc_efl()
        avg_pupilz, avg_pupilx = self.entrance_pupil()
        self.fnum = self.foclen / avg_pupilx / 2

    def find_aperture(self):
        self.aper_idx = None
        for i in range(len(self.surfaces) - 1):
            if self.surfaces[i].mat1.A < 1.0003 and self.surfaces[i].mat2.A < 1.0003:
                self.aper_idx = i
                return

    def find_diff_surf(self):
        if self.aper_idx is None:
            return range(len(self.surfaces))
        return list(range(0, self.aper_idx)) + list(range(self.aper_idx + 1, len(self.surfaces)))

    # ---- engine handle: rebuilt whenever the prescription or the sensor position changed ------------
    def _engine_lens(self):
        key = (tuple(s._state_key() for s in self.surfaces), float(self.d_sensor))
        if self._handle is None or self._handle_key != key:
            self._handle = E.LensHandle([s.engine_record() for s in self.surfaces], float(self.d_sensor))
            self._handle_key = key
        return self._handle

    def _require_cuda(self):
        if self.device.type != "cuda":
            raise RuntimeError("sdirt_b200: Lensgroup needs a CUDA device; the engine has no CPU path")

    # ==============================================================================================
    # Sampling (optics.py:460-494, 816-858)
    # ==============================================================================================
    def _pupil_samples(self, spp, shrink_pupil=False, spatial_order=False):
        """Shared pupil disc samples, CPU generator, theta first then rho^2 (optics.py:482-487).  `spatial_order`
        returns the same set Morton-sorted on the device (a PSF is a sum over the set, so its order is free)."""
        pupilz, pupilr = self.entrance_pupil(shrink_pupil=shrink_pupil)
        if getattr(self, "sample_rng", "cpu") == "cuda":
            # engine extension: the same two uniform draws from the DEVICE generator (torch.cuda.manual_seed), no host RNG and no
            # 8 B / sample upload per call -- for throughput runs; seeded comparisons with the reference need the default
            theta = torch.rand(spp, device=self.device) * 2 * np.pi
            r = torch.sqrt(torch.rand(spp, device=self.device) * pupilr ** 2)
            xy = torch.stack((r * torch.cos(theta), r * torch.sin(theta)), 1).contiguous()
        else:
            theta = torch.rand(spp) * 2 * np.pi
            r = torch.sqrt(torch.rand(spp) * pupilr ** 2)
            xy = torch.stack((r * torch.cos(theta), r * torch.sin(theta)), 1)
            xy = xy.to(self.device, non_blocking=True).contiguous()
        if spatial_order:
            xy = E.pupil_sort(xy, pupilr)
        return xy, pupilz

    @torch.no_grad()
    def sample_from_points(self, o=[[0, 0, -10000]], spp=256, wvln=DEFAULT_WAVE, shrink_pupil=False, normalized=False):
        """Forward rays from N object points towards spp shared pupil samples; Ray of shape [spp, N, 3]."""
        self._require_cuda()
        if not torch.is_tensor(o):
            o = torch.tensor(o)
        pts = o.to(self.device, torch.float32).reshape(-1, 3).contiguous()
        xy, pupilz = self._pupil_samples(spp, shrink_pupil)
        ro, rd = E.sample_rays(pts, xy, pupilz)
        return Ray._from_engine(ro, rd, wvln)

    def point_source_grid(self, depth, grid=9, normalized=True, quater=False, center=False):
        if grid == 1:
            x, y = torch.tensor([[0.]]), torch.tensor([[0.]])
            assert not quater, "Quater should be False when grid is 1."
        elif center:
            hb = 1 / 2 / (grid - 1)
            x, y = torch.meshgrid(torch.linspace(-1 + hb, 1 - hb, grid), torch.linspace(1 - hb, -1 + hb, grid), indexing="xy")
        else:
            x, y = torch.meshgrid(torch.linspace(-0.98, 0.98, grid), torch.linspace(0.98, -0.98, grid), indexing="xy")
        z = torch.full((grid, grid), depth)
        point_source = torch.stack([x, y, z], dim=-1)
        if quater:
            bound_i = grid // 2 if grid % 2 == 0 else grid // 2 + 1
            point_source = point_source[0:bound_i, 0:bound_i, :]
        if not normalized:
            scale = self.calc_scale_pinhole(depth)
            point_source[..., 0] *= scale * self.sensor_size[0] / 2
            point_source[..., 1] *= scale * self.sensor_size[1] / 2
        return point_source

    @torch.no_grad()
    def sample_pupil(self, res=(512, 512), spp=16, num_angle=8, pupilr=None, pupilz=None):
        """Points [spp, H, W, 3] on the entrance-pupil disc: uniform, or stratified into `num_angle` sectors x spp / num_angle
        rings when spp is small (optics.py:542-594).  The draws come from the CPU generator in the reference's order (the
        reference draws on its own device), so a seeded run reproduces the reference's CPU run."""
        H, W = res
        if pupilr is None or pupilz is None:
            pupilz, pupilr = self.entrance_pupil()
        if spp % num_angle != 0 or spp >= 10000:
            theta = torch.rand((spp, H, W)) * 2 * np.pi
            r = torch.sqrt(torch.rand((spp, H, W)) * pupilr ** 2)
            x, y = r * torch.cos(theta), r * torch.sin(theta)
        else:
            xs, ys = [], []
            for i in range(num_angle):
                for j in range(spp // num_angle):
                    theta = torch.rand((1, H, W)) * 2 * np.pi / num_angle + i * 2 * np.pi / num_angle
                    r2 = torch.rand((1, H, W)) * pupilr ** 2 / spp * num_angle + j * pupilr ** 2 / spp * num_angle
                    r = torch.sqrt(r2)
                    xs.append(r * torch.cos(theta))
                    ys.append(r * torch.sin(theta))
            x, y = torch.cat(xs, dim=0), torch.cat(ys, dim=0)
        return torch.stack((x, y, torch.full_like(x, pupilz)), -1).to(self.device)

    @torch.no_grad()
    def sample_point_source(self, R=None, depth=-10.0, M=11, spp=16, fov=10.0, forward=True, pupil=True, wvln=DEFAULT_WAVE,
                            importance_sampling=False):
        """Forward rays [spp, M, M, 3] from an M x M grid on the plane z = depth towards per-point pupil samples
        (optics.py:403-456); used by the spot / RMS / magnification analyses."""
        self._require_cuda()
        if R is None:
            R = self.surfaces[0].r
        Rw = R * self.sensor_res[1] / self.sensor_res[0]
        x, y = torch.meshgrid(torch.linspace(-1, 1, M), torch.linspace(1, -1, M), indexing="xy")
        if importance_sampling:
            x, y = torch.sqrt(x.abs()) * x.sign(), torch.sqrt(y.abs()) * y.sign()
        x, y = x * Rw, y * R
        o = torch.stack((x, y, torch.full_like(x, depth)), -1).to(self.device).unsqueeze(0).repeat(spp, 1, 1, 1)
        if not pupil:
            raise Exception("Cone sampling specified by fov has been abandoned. Use pupil sampling instead.")
        d = self.sample_pupil(res=(M, M), spp=spp) - o
        return Ray(o, d, wvln, device=self.device)              # (Ray normalises d with the reference's rounding)

    # ==============================================================================================
    # Ray tracing (optics.py:601-717)
    # ==============================================================================================
    def trace(self, ray, lens_range=None, record=False):
        """Trace `ray` in place through `lens_range` (default: all surfaces).  Direction follows the sign of the
        first ray's d_z.  Returns (ray, valid, oss) like the reference."""
        self._require_cuda()
        is_forward = bool(ray.d.reshape(-1, 3)[0, 2] > 0)
        idx = list(range(len(self.surfaces))) if lens_range is None else [int(i) for i in lens_range]
        rec, o_start = None, None
        if record:
            o_start = ray.o.reshape(-1, 3).cpu().numpy()
        if len(idx) == 0:
            pass
        elif idx == list(range(idx[0], idx[-1] + 1)):
            rec = E.trace_rays(self._engine_lens(), ray.wvln, ray.o, ray.d, ray.ra, idx[0], idx[-1] + 1,
                               backward=not is_forward, newton=self.newton, record=record, numerics=self.numerics)
        else:                       # non-contiguous ranges: surface by surface, in the reference's visiting order
            for i in (idx if is_forward else idx[::-1]):
                self.surfaces[i].ray_reaction(ray, numerics=self.numerics)
        valid = ray.ra == 1
        oss = None
        if record:
            oss = [[p] for p in o_start]
            if rec is not None:
                rec = rec.cpu().numpy()
                for k in range(rec.shape[0]):
                    for j in np.nonzero(rec[k][:, 6] == 1)[0]:
                        oss[j].append(rec[k][j, :3])
        return ray, valid, oss

    def trace2obj(self, ray, depth=DEPTH):
        ray, _, _ = self.trace(ray)
        return ray.propagate_to(depth)

    def trace2sensor(self, ray, record=False, ignore_invalid=False):
        """All surfaces then Ray.propagate_to(d_sensor) (optics.py:638-664); one fused launch."""
        if record:
            ray_out, valid, oss = self.trace(ray, record=True)
            ray_out = ray_out.propagate_to(self.d_sensor)
            p = ray_out.o.reshape(-1, 3)
            vm = (ray_out.ra == 1).reshape(-1).cpu().numpy()
            for v, os_, pp in zip(vm, oss, p.cpu().numpy()):
                if v:
                    os_.append(pp)
            return (p[torch.from_numpy(vm).to(p.device)] if ignore_invalid else p), oss
        self._require_cuda()
        is_forward = bool(ray.d.reshape(-1, 3)[0, 2] > 0)
        E.trace_rays(self._engine_lens(), ray.wvln, ray.o, ray.d, ray.ra, 0, len(self.surfaces), backward=not is_forward,
                     to_sensor=True, newton=self.newton, numerics=self.numerics)
        return ray

    # ==============================================================================================
    # PSF (optics.py:889-1041)
    # ==============================================================================================
    @torch.no_grad()
    def psf_center(self, point, method="chief_ray"):
        """Reference PSF centre [N, 2] on the sensor plane (flipped) for object points [N, 3] in mm."""
        if method == "chief_ray":
            self._require_cuda()
            pts = point.to(self.device, torch.float32).reshape(-1, 3).contiguous()
            xy, pupilz = self._pupil_samples(GEO_SPP, shrink_pupil=True)
            return E.psf_centre(self._engine_lens(), DEFAULT_WAVE, pts, xy, pupilz, newton=self.newton, numerics=self.numerics)
        if method == "pinhole":
            scale = self.calc_scale_pinhole(point[..., 2])
            return -point[..., :2] / scale
        raise Exception("Unsupported method.")

    def psf(self, points, ks=31, wvln=DEFAULT_WAVE, spp=GEO_SPP, center=True):
        """Left sub-pixel PSFs [N, ks, ks] (or [ks, ks] for a single point), max-normalised."""
        return self.psf_diff(points=points, wvln=wvln, ks=ks, spp=spp, center=center)

    def _object_points(self, points):
        """Normalised (x, y, depth) -> object-space mm, the reference's arithmetic (optics.py:955-960)."""
        depth = points[:, 2]
        scale = self.calc_scale_pinhole(depth)
        point_obj = points.clone()
        point_obj[..., 0] = points[..., 0] * scale * self.sensor_size[1] / 2
        point_obj[..., 1] = points[..., 1] * scale * self.sensor_size[0] / 2
        return point_obj

    def psf_dp(self, points, ks=31, wvln=DEFAULT_WAVE, spp=GEO_SPP, center=True, param_list=None, normalise=1):
        """Both sub-pixel PSFs of the SAME trace: (L, R), each [N, ks, ks].  (Engine extension: the reference
        obtains R from a second, mirrored call with different samples, psfnet.py:540-544.)"""
        self._require_cuda()
        if not torch.is_tensor(points):
            points = torch.tensor(points)
        points = points.float() if points.is_cuda else points.float().cpu()      # (device points stay there: no host round trip)
        if points.dim() == 1:
            points = points.unsqueeze(0)
        point_obj = self._object_points(points)
        pts = point_obj.to(self.device).contiguous()
        # (the run-length splat of the specialised kernels wants neighbouring samples in a row; the generic strict kernel that
        # replays given Newton loop counts splats with shared-memory atomics and is faster on the unsorted set)
        xy, pupilz = self._pupil_samples(spp, spatial_order=self.numerics in ("fast", "hybrid", "adaptive")
                                         or self.newton == "per_ray")                                   # main bundle first ...
        if center:
            centre = self.psf_center(point_obj)                          # ... then the chief-ray bundle (RNG order)
        else:
            ideal = points.clone()[:, :2]
            ideal[:, 0] *= self.sensor_size[1] / 2
            ideal[:, 1] *= self.sensor_size[0] / 2
            centre = ideal.to(self.device).contiguous()
        dp = None if param_list is None else tuple(float(v) for v in param_list[:4])
        return E.psf_bank(self._engine_lens(), wvln, pts, xy, pupilz, centre, ks, self.pixel_size, dp=dp, newton=self.newton,
                          normalise=normalise, numerics=self.numerics)

    def psf_diff(self, points, wvln=DEFAULT_WAVE, ks=31, spp=GEO_SPP, center=True, param_list=None):
        """optics.py:934-996.  Returns the grid the reference returns: L, or R when param_list[4] != 'l'."""
        single = (not torch.is_tensor(points) and np.ndim(points) == 1) or (torch.is_tensor(points) and points.dim() == 1)
        L, R = self.psf_dp(points, ks=ks, wvln=wvln, spp=spp, center=center, param_list=param_list, normalise=1)
        psf = L if (param_list is None or param_list[4] == "l") else R
        return psf.squeeze(0) if single else psf

    def psf_rgb(self, points, ks=31, spp=GEO_SPP, center=True, param_list=None):
        psfs = [self.psf_diff(points=points, wvln=w, ks=ks, spp=spp, center=center, param_list=param_list) for w in WAVE_RGB]
        return torch.stack(psfs, dim=-3)

    def psf_map(self, depth=DEPTH, grid=7, ks=51, spp=GEO_SPP, center=True):
        """[3, grid*ks, grid*ks] mosaic of RGB PSFs (optics.py:1018-1041; make_grid with nrow=grid, no padding)."""
        points = self.point_source_grid(depth=depth, grid=grid).reshape(-1, 3)
        psf = self.psf_rgb(points=points, ks=ks, center=center, spp=spp).reshape(grid, grid, 3, ks, ks)
        return psf.permute(2, 0, 3, 1, 4).reshape(3, grid * ks, grid * ks)

    # ==============================================================================================
    # Geometry that needs (tiny) traces (optics.py:1112-1117, 1170-1233, 1302-1396, 1471-1514)
    # ==============================================================================================
    def calc_efl(self):
        return self.r_last / np.tan(self.hfov)

    @torch.no_grad()
    def refocus(self, depth=DEPTH):
        """Move the sensor to the least-squares focus of an on-axis point at `depth` (optics.py:1170-1196)."""
        o = self.surfaces[0].surface_sample(GEO_SPP)
        d = o - torch.tensor([0, 0, depth], dtype=torch.float32).to(self.device)
        ray = Ray(o, d, wvln=DEFAULT_WAVE, device=self.device)
        ray, _, _ = self.trace(ray)
        t = (ray.d[..., 0] * ray.o[..., 0] + ray.d[..., 1] * ray.o[..., 1]) / (ray.d[..., 0] ** 2 + ray.d[..., 1] ** 2)
        t = t * ray.ra
        focus_d = (ray.o[..., 2] - ray.d[..., 2] * t).cpu().numpy()
        focus_d = focus_d[ray.ra.cpu() > 0]
        focus_d = focus_d[~np.isnan(focus_d) & (focus_d > 0)]
        d_sensor_new = float(np.mean(focus_d))
        assert d_sensor_new > 0, "sensor position is negative."
        self.d_sensor = d_sensor_new
        self.post_computation()

    @torch.no_grad()
    def calc_fov(self):
        """Half diagonal field of view from 100 backward rays off the sensor corner (optics.py:1203-1233)."""
        M = 100
        pupilz, pupilx = self.exit_pupil(shrink_pupil=True)
        o1 = torch.tensor([self.r_last, 0, self.d_sensor]).repeat(M, 1).to(torch.float32)
        x2 = torch.linspace(-pupilx, pupilx, M)
        o2 = torch.stack((x2, torch.full_like(x2, 0), torch.full_like(x2, pupilz)), axis=-1)
        ray = Ray(o1, o2 - o1, device=self.device)
        ray, _, _ = self.trace(ray)
        tan_fov = ray.d[..., 0] / ray.d[..., 2]
        fov = torch.atan(torch.sum(tan_fov * ray.ra) / torch.sum(ray.ra))
        if torch.isnan(fov):
            print("computed fov is NaN, use 0.5 rad instead.")
            return 0.5
        return fov.item()

    @torch.no_grad()
    def calc_scale_pinhole(self, depth):
        return -depth * np.tan(self.hfov) / self.r_last

    @torch.no_grad()
    def calc_magnification3(self, depth):
        """Magnification from the traced image of a 21 x 21 object grid, 512 rays per point (optics.py:1237-1272)."""
        M, spp = 21, 512
        ray = self.sample_point_source(M=M, spp=spp, depth=depth, R=-depth * np.tan(self.hfov) * 0.5, pupil=True)
        o1 = torch.flip(ray.o.detach()[..., :2], [1, 2])
        ray, _, _ = self.trace(ray)
        o2 = ray.project_to(self.d_sensor)
        x1 = o1[0, :, :, 0]
        x2 = torch.sum(o2[..., 0] * ray.ra, axis=0) / torch.sum(ray.ra, axis=0).add(EPSILON)
        tmp = (x1 / x2)[:M // 2, :M // 2]
        mag = 1 / torch.mean(tmp[~tmp.isnan()]).item()
        if mag == 0:
            return 1 / (-depth * np.tan(self.hfov) / self.r_last)
        return mag

    @torch.no_grad()
    def calc_scale_ray(self, depth):
        """Object-to-sensor scale by ray tracing (optics.py:1310-1322)."""
        if isinstance(depth, torch.Tensor) and len(depth.shape) == 1:
            return torch.tensor([1 / self.calc_magnification3(d) for d in depth])
        return 1 / self.calc_magnification3(depth)

    @torch.no_grad()
    def analysis_rms(self, depth=DEPTH, ref=True):
        """(average, on-axis, off-axis) RMS spot radius in mm over the three design wavelengths, referred to the green
        spot centre (optics.py:2103-2140), on a 31 x 31 field grid with GEO_SPP rays per point."""
        H = 31
        scale = self.calc_scale_ray(depth)
        R = self.sensor_size[0] / 2 * scale
        if ref:
            ray = self.sample_point_source(M=H, spp=GEO_SPP, depth=depth, R=R, pupil=True, wvln=DEFAULT_WAVE)
            ray, _, _ = self.trace(ray)
            p_green = ray.project_to(self.d_sensor)
            p_center_ref = (p_green * ray.ra.unsqueeze(-1)).sum(0) / ray.ra.sum(0).add(0.0001).unsqueeze(-1)
        rms, rms_on_axis, rms_off_axis = [], [], []
        for wvln in WAVE_RGB:
            ray = self.sample_point_source(M=H, spp=GEO_SPP, depth=depth, R=R, pupil=True, wvln=wvln)
            ray, _, _ = self.trace(ray)
            o2 = ray.project_to(self.d_sensor)
            o2_center = (o2 * ray.ra.unsqueeze(-1)).sum(0) / ray.ra.sum(0).add(0.0001).unsqueeze(-1)
            o2_norm = (o2 - (p_center_ref if ref else o2_center)) * ray.ra.unsqueeze(-1)
            rms.append(torch.sqrt(torch.sum(o2_norm ** 2 * ray.ra.unsqueeze(-1)) / torch.sum(ray.ra)))
            rms_on_axis.append(torch.sqrt(torch.sum(o2_norm[:, H // 2 + 1, H // 2 + 1, :] ** 2 * ray.ra[:, H // 2 + 1, H // 2 + 1].unsqueeze(-1))
                                          / torch.sum(ray.ra[:, H // 2, H // 2])))
            rms_off_axis.append(torch.sqrt(torch.sum(o2_norm[:, 0, 0, :] ** 2 * ray.ra[:, 0, 0].unsqueeze(-1)) / torch.sum(ray.ra[:, 0, 0])))
        return sum(rms) / len(rms), sum(rms_on_axis) / len(rms_on_axis), sum(rms_off_axis) / len(rms_off_axis)

    @torch.no_grad()
    def analysis(self, save_name="./test", ks=None, render=False, multi_plot=False, plot_invalid=True, zmx_format=False, depth=DEPTH,
                 render_unwarp=False, lens_title=None):
        """The numeric part of Lensgroup.analysis (optics.py:1663-1683): the RMS spot radii, printed as the reference prints
        them and returned.  Its drawings (layout with ray paths, PSF map PNG, rendered resolution chart) are matplotlib / OpenCV
        output outside the hot path (out of scope) and are not produced."""
        rms_avg, rms_on, rms_off = self.analysis_rms(depth=depth)
        print(f"On-axis RMS radius: {round(rms_on.item() * 1000, 3)}um, Off-axis RMS radius: {round(rms_off.item() * 1000, 3)}um, "
              f"Avg RMS spot size (radius): {round(rms_avg.item() * 1000, 3)}um.")
        return rms_avg, rms_on, rms_off

    @torch.no_grad()
    def exit_pupil(self, shrink_pupil=False):
        return self.entrance_pupil(entrance=False, shrink_pupil=shrink_pupil)

    @torch.no_grad()
    def calc_entrance_pupil_paraxial(self, entrance=True):
        """Pupil (z, radius) from 16 paraxial rays off the stop edge (optics.py:1335-1376).  The reference repeats this
        trace inside every sample_from_points call; it depends on the prescription only, so the result is kept until a
        surface changes (same value, one trace instead of two per psf() call)."""
        key = (bool(entrance), self.aper_idx, tuple(s._state_key() for s in self.surfaces), str(self.device))
        cache = self.__dict__.setdefault("_pupil_cache", {})
        if key not in cache:
            if len(cache) > 16:
                cache.clear()
            cache[key] = self._calc_entrance_pupil_paraxial(entrance)
        return cache[key]

    def _calc_entrance_pupil_paraxial(self, entrance=True):
        aper_z = self.surfaces[self.aper_idx].d.item()
        aper_r = self.surfaces[self.aper_idx].r
        delta_r = 1e-3
        ray_o = torch.tensor([[delta_r, 0, aper_z]]).repeat(16, 1)
        phi = torch.linspace(-0.1, 0.1, 16) / 180.0 * torch.pi
        sgn = -1.0 if entrance else 1.0
        d = torch.stack((torch.sin(phi), torch.zeros_like(phi), sgn * torch.cos(phi)), axis=-1)
        ray = Ray(ray_o, d, device=self.device)
        rng = range(0, self.aper_idx) if entrance else range(self.aper_idx + 1, len(self.surfaces))
        ray, _, _ = self.trace(ray, lens_range=rng)
        keep = ray.ra != 0
        ray_o = torch.stack([ray.o[keep][:, 0], ray.o[keep][:, 2]], dim=-1)
        ray_d = torch.stack([ray.d[keep][:, 0], ray.d[keep][:, 2]], dim=-1)
        pts = self.compute_intersection_points_2d(ray_o, ray_d)
        if len(pts) == 0:
            print("No intersection points found, use the first surface as pupil.")
            return self.surfaces[0].d.item(), self.surfaces[0].r
        avg_pupilr = torch.abs((torch.mean(pts[:, 0])) / delta_r * aper_r).item()
        avg_pupilz = torch.mean(pts[:, 1]).item()
        return avg_pupilz, avg_pupilr

    @torch.no_grad()
    def entrance_pupil(self, M=32, entrance=True, shrink_pupil=False):
        if self.aper_idx is None:
            s = self.surfaces[0] if entrance else self.surfaces[-1]
            return s.d.item(), s.r
        avg_pupilz, avg_pupilr = self.calc_entrance_pupil_paraxial(entrance=entrance)
        if shrink_pupil:
            avg_pupilr = avg_pupilr * 0.25
        return avg_pupilz, avg_pupilr

    @staticmethod
    def compute_intersection_points_2d(origins, directions):
        """Pairwise intersections of 2-D lines; the same torch.linalg.lstsq call as the reference, on the same
        device (optics.py:1471-1514) — its float32 answer is part of the reference's pupil definition."""
        N = origins.shape[0]
        idx_i, idx_j = torch.combinations(torch.arange(N), r=2).unbind(1)
        Oi, Oj, Di, Dj = origins[idx_i], origins[idx_j], directions[idx_i], directions[idx_j]
        b = Oj - Oi
        A = torch.stack([Di, -Dj], dim=-1)
        x = torch.linalg.lstsq(A, b.unsqueeze(-1))[0].squeeze(-1)
        P_i = Oi + x[:, 0].unsqueeze(-1) * Di
        P_j = Oj + x[:, 1].unsqueeze(-1) * Dj
        return (P_i + P_j) / 2

    def set_aperture(self, fnum=None, foclen=None, aper_r=None):
        """Change the stop radius (optics.py:1527-1545); the engine handle is rebuilt on the next call."""
        if aper_r is None:
            foclen = self.calc_efl() if foclen is None else foclen
            aper_r = foclen / fnum / 2
        self.surfaces[self.aper_idx].r = float(aper_r)
        self.fnum = self.foclen / aper_r / 2

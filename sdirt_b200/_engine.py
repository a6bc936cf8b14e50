"""ctypes binding of libsdirt_engine.so (include/sdirt_engine.h).  No CPU fallback: every compute entry point
requires CUDA tensors and raises if the library or the device is missing."""
import ctypes as C
import os

import torch

# SDIRT_ENGINE_LIB: another build of the same library (kernel-tuning A/B runs on the GPU box: tools/build_variants.py)
_LIB_PATH = os.environ.get("SDIRT_ENGINE_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "_lib", "libsdirt_engine.so")
MAX_SURFACES, MAX_AI, MAX_POINTS_PER_CALL = 32, 8, 65535
SURF_FLAT, SURF_SPHERE, SURF_ASPHERE = 0, 1, 2


class Surface(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_ai", C.c_int32), ("square", C.c_int32), ("reserved", C.c_int32),
                ("r", C.c_double), ("d", C.c_float), ("c", C.c_float), ("k", C.c_float),
                ("ai", C.c_float * MAX_AI),
                ("n1_A", C.c_double), ("n1_B", C.c_double), ("n2_A", C.c_double), ("n2_B", C.c_double)]


class Options(C.Structure):
    _fields_ = [("newton_mode", C.c_int32), ("numerics", C.c_int32), ("iters", C.c_int32 * MAX_SURFACES)]


NEWTON_REPLAY, NEWTON_PER_RAY = 0, 1
NUMERICS_STRICT, NUMERICS_FAST, NUMERICS_HYBRID, NUMERICS_ADAPTIVE = 0, 1, 2, 3
DEFAULT_NUMERICS = NUMERICS_STRICT


MLP_MAX_LAYERS = 12


class MlpShape(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("n1", C.c_int32), ("K", C.c_int32 * MLP_MAX_LAYERS), ("N", C.c_int32 * MLP_MAX_LAYERS)]


class DPParams(C.Structure):
    _fields_ = [("h", C.c_float), ("f", C.c_float), ("w", C.c_float), ("r", C.c_float)]


_lib = None


def lib():
    """Load the engine; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise RuntimeError(f"{_LIB_PATH} is missing: run `python -m sdirt_b200.build` (there is no CPU fallback)")
    L = C.CDLL(_LIB_PATH)
    vp, i64, dbl, cint = C.c_void_p, C.c_int64, C.c_double, C.c_int
    L.sdirt_last_error.restype = C.c_char_p
    L.sdirt_version.restype = C.c_char_p
    L.sdirt_launch_count.restype = C.c_uint64
    L.sdirt_device_sm_count.restype = cint
    L.sdirt_lens_create.argtypes = [C.POINTER(Surface), cint, dbl, C.POINTER(vp)]
    L.sdirt_lens_set_sensor.argtypes = [vp, dbl]
    L.sdirt_lens_set_surface.argtypes = [vp, cint, C.POINTER(Surface)]
    L.sdirt_lens_num_surfaces.argtypes = [vp]
    L.sdirt_lens_eta.argtypes = [vp, dbl, cint, C.POINTER(dbl)]
    L.sdirt_lens_destroy.argtypes = [vp]
    L.sdirt_lens_destroy.restype = None
    L.sdirt_trace_rays.argtypes = [vp, dbl, vp, vp, vp, i64, cint, cint, cint, cint, C.POINTER(Options), vp, vp]
    L.sdirt_sample_rays.argtypes = [vp, i64, vp, i64, dbl, vp, vp, vp]
    L.sdirt_normalize_rays.argtypes = [vp, i64, vp]
    L.sdirt_propagate_rays.argtypes = [vp, vp, i64, dbl, vp]
    L.sdirt_psf_centre.argtypes = [vp, dbl, vp, i64, vp, i64, dbl, C.POINTER(Options), vp, vp]
    L.sdirt_psf_bank_workspace.argtypes = [i64, i64, cint]
    L.sdirt_psf_bank_workspace.restype = i64
    L.sdirt_psf_bank.argtypes = [vp, dbl, vp, i64, vp, i64, dbl, vp, cint, dbl, C.POINTER(DPParams), C.POINTER(Options),
                                 cint, vp, vp, vp, vp, i64, vp]
    L.sdirt_pupil_sort_workspace.argtypes = [i64]
    L.sdirt_pupil_sort_workspace.restype = i64
    L.sdirt_pupil_sort.argtypes = [vp, i64, dbl, vp, vp, i64, vp]
    L.sdirt_splat_rays.argtypes = [vp, vp, vp, i64, i64, vp, cint, dbl, C.POINTER(DPParams), vp, vp, vp, i64, vp]
    L.sdirt_render_local_psf.argtypes = [vp, vp, cint, cint, cint, cint, cint, cint, cint, vp, vp, vp]
    L.sdirt_render_local_psf_rows.argtypes = [vp, vp, cint, cint, cint, cint, cint, cint, cint, cint, cint, vp, vp, vp]
    L.sdirt_render_local_psf_f32.argtypes = [vp, vp, cint, cint, cint, cint, cint, vp, vp, vp]
    if hasattr(L, "sdirt_render_records_bytes") or not os.environ.get("SDIRT_ENGINE_LIB"):     # (an older build loaded for an A/B run may lack them)
        L.sdirt_render_records_bytes.argtypes = [cint, cint, cint, cint, cint]
        L.sdirt_render_records_bytes.restype = i64
        L.sdirt_render_pack_image.argtypes = [vp, cint, cint, cint, cint, cint, cint, vp, vp]
        L.sdirt_render_local_psf_rows_packed.argtypes = [vp, vp, cint, cint, cint, cint, cint, cint, cint, cint, vp, vp, vp]
    L.sdirt_mlp_input_layer.argtypes = [vp, vp, vp, cint, cint, cint, cint, cint, cint, cint, vp, vp, cint, vp, vp]
    L.sdirt_psf_pack.argtypes = [vp, i64, cint, cint, vp, vp]
    L.sdirt_mlp_fused_layout.argtypes = [C.POINTER(MlpShape), C.POINTER(i64), C.POINTER(C.c_int32), C.POINTER(i64)]
    L.sdirt_mlp_fused_layout.restype = i64
    L.sdirt_mlp_fused_pack_layer.argtypes = [C.POINTER(MlpShape), cint, vp, vp, vp, vp, vp]
    L.sdirt_mlp_fused_cta_group.argtypes = [cint]
    L.sdirt_mlp_fused_cta_group.restype = cint
    L.sdirt_mlp_fused_pred.argtypes = [C.POINTER(MlpShape), vp, vp, vp, vp, vp, vp, vp, cint, cint, cint, cint, cint, cint, cint, cint, vp, vp]
    L.sdirt_tone_degamma.argtypes = [vp, vp, i64, vp]
    L.sdirt_gamma_noise_clip.argtypes = [vp, vp, vp, vp, cint, cint, cint, cint, vp]
    L.sdirt_fp32_peak_probe.argtypes = [vp, cint, cint, cint, vp]
    L.sdirt_debug_trace_strict2.argtypes = [vp, dbl, vp, vp, i64, dbl, vp, vp]
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise RuntimeError("sdirt_engine: " + lib().sdirt_last_error().decode())


def version():
    return lib().sdirt_version().decode()


def launch_count():
    return int(lib().sdirt_launch_count())


def _dev(t, name, dtype=torch.float32):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"sdirt_engine: `{name}` must be a CUDA tensor (the engine has no CPU path)")
    if t.dtype != dtype or not t.is_contiguous():
        raise RuntimeError(f"sdirt_engine: `{name}` must be contiguous {dtype}")
    return C.c_void_p(t.data_ptr())


def _stream(t):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _on_device(fn):
    """Run a binding with the CUDA device of its first tensor argument current: the library launches on the CURRENT device (it
    never calls cudaSetDevice) with the stream handle taken from the tensors' device, so a call with tensors on cuda:1 while
    cuda:0 is current would otherwise launch on the wrong device (`PSFNet(..., device='cuda:1')` in a single process)."""
    import functools

    def first_cuda(a, depth=0):
        if isinstance(a, torch.Tensor):
            return a.device if a.is_cuda else None
        if isinstance(a, torch.device):
            return a if (a.type == "cuda" and a.index is not None) else None
        if isinstance(a, (list, tuple)) and depth < 2:
            for x in a:
                d = first_cuda(x, depth + 1)
                if d is not None:
                    return d
        return None

    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        for a in list(args) + list(kwargs.values()):
            d = first_cuda(a)
            if d is not None:
                if d.index == torch.cuda.current_device():
                    break
                with torch.cuda.device(d):
                    return fn(*args, **kwargs)
        return fn(*args, **kwargs)
    return wrapped


def make_surface(kind, r, d, c=0.0, k=0.0, ai=None, n1=(1.0, 0.0), n2=(1.0, 0.0), square=False):
    s = Surface()
    s.kind, s.square, s.r, s.d, s.c, s.k = kind, int(bool(square)), float(r), float(d), float(c), float(k)
    ai = [] if ai is None else list(ai)
    if len(ai) > MAX_AI:
        raise ValueError(f"at most {MAX_AI} even-asphere coefficients")
    s.n_ai = len(ai)
    for i, a in enumerate(ai):
        s.ai[i] = float(a)
    s.n1_A, s.n1_B, s.n2_A, s.n2_B = float(n1[0]), float(n1[1]), float(n2[0]), float(n2[1])
    return s


def make_options(newton=None, numerics=None):
    """newton: None / 'per_ray' -> per-ray loop; a sequence of ints -> replay those loop counts per lens surface.
    numerics: 'strict' (bit-exact restatement of the reference arithmetic), 'fast', or None for the default."""
    n = Options()
    if newton is None or (isinstance(newton, str) and newton == "per_ray"):
        n.newton_mode = NEWTON_PER_RAY
    else:
        n.newton_mode = NEWTON_REPLAY
        for i, v in enumerate(newton):
            n.iters[i] = int(v)
    if numerics is None:
        n.numerics = DEFAULT_NUMERICS
    elif numerics in ("strict", NUMERICS_STRICT):
        n.numerics = NUMERICS_STRICT
    elif numerics in ("fast", NUMERICS_FAST):
        n.numerics = NUMERICS_FAST
    elif numerics in ("hybrid", NUMERICS_HYBRID):
        n.numerics = NUMERICS_HYBRID
    elif numerics in ("adaptive", NUMERICS_ADAPTIVE):
        n.numerics = NUMERICS_ADAPTIVE
    else:
        raise ValueError(f"numerics must be 'strict', 'hybrid', 'adaptive' or 'fast', got {numerics!r}")
    return n


def make_dp(dp):
    if dp is None:
        return None
    p = DPParams()
    p.h, p.f, p.w, p.r = (float(v) for v in dp[:4])
    return p


class LensHandle:
    """Owns one sdirt_lens*; mirrors the geometric state of a Lensgroup."""

    def __init__(self, surfaces, d_sensor):
        arr = (Surface * len(surfaces))(*surfaces)
        h = C.c_void_p()
        _check(lib().sdirt_lens_create(arr, len(surfaces), float(d_sensor), C.byref(h)))
        self._h, self.n = h, len(surfaces)

    def set_sensor(self, d_sensor):
        _check(lib().sdirt_lens_set_sensor(self._h, float(d_sensor)))

    def set_surface(self, i, surf):
        _check(lib().sdirt_lens_set_surface(self._h, int(i), C.byref(surf)))

    def eta(self, wvln, backward=False):
        out = (C.c_double * self.n)()
        _check(lib().sdirt_lens_eta(self._h, float(wvln), int(backward), out))
        return list(out)

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.sdirt_lens_destroy(self._h)
            self._h = None


@_on_device
def trace_rays(lens, wvln, o, d, ra, s_begin=0, s_end=None, backward=False, to_sensor=False, newton=None, record=False, numerics=None):
    """In-place trace of AoS rays o[n,3], d[n,3], ra[n]; returns the per-surface record if requested."""
    n = ra.numel()
    s_end = lens.n if s_end is None else s_end
    rec = torch.empty((max(s_end - s_begin, 0), n, 7), device=o.device, dtype=torch.float32) if record else None
    nt = make_options(newton, numerics)
    _check(lib().sdirt_trace_rays(lens._h, float(wvln), _dev(o, "o"), _dev(d, "d"), _dev(ra, "ra"), n, s_begin, s_end,
                                  int(backward), int(to_sensor), C.byref(nt),
                                  _dev(rec, "record") if record else None, _stream(o)))
    return rec


@_on_device
def sample_rays(points, pupil_xy, pupil_z):
    """[spp, N, 3] origins and unit directions of sample_from_points (sample-major, as the reference's Ray)."""
    n, m = points.shape[0], pupil_xy.shape[0]
    o = torch.empty((m, n, 3), device=points.device, dtype=torch.float32)
    d = torch.empty((m, n, 3), device=points.device, dtype=torch.float32)
    _check(lib().sdirt_sample_rays(_dev(points, "points"), n, _dev(pupil_xy, "pupil_xy"), m, float(pupil_z),
                                   _dev(o, "o"), _dev(d, "d"), _stream(points)))
    return o, d


@_on_device
def normalize_rays(d):
    _check(lib().sdirt_normalize_rays(_dev(d, "d"), d.numel() // 3, _stream(d)))


@_on_device
def propagate_rays(o, d, z):
    _check(lib().sdirt_propagate_rays(_dev(o, "o"), _dev(d, "d"), o.numel() // 3, float(z), _stream(o)))


@_on_device
def psf_centre(lens, wvln, points, pupil_xy, pupil_z, newton=None, numerics=None):
    n = points.shape[0]
    out = torch.empty((n, 2), device=points.device, dtype=torch.float32)
    nt = make_options(newton, numerics)
    for a in range(0, n, MAX_POINTS_PER_CALL * 1024):
        pts = points[a:a + MAX_POINTS_PER_CALL * 1024]
        _check(lib().sdirt_psf_centre(lens._h, float(wvln), _dev(pts, "points"), pts.shape[0], _dev(pupil_xy, "pupil_xy"),
                                      pupil_xy.shape[0], float(pupil_z), C.byref(nt),
                                      C.c_void_p(out[a:].data_ptr()), _stream(points)))
    return out


_ws_cache = {}


def _workspace(device, nbytes):
    # one scratch buffer per (device, stream): calls issued on different streams must not share partial tiles, and a buffer that
    # is replaced by a larger one is returned to the caching allocator, which keeps it off other streams while still in flight
    key = (device.type, device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(int(nbytes), device=device, dtype=torch.uint8)
        _ws_cache[key] = ws
    return ws


@_on_device
def pupil_sort(pupil_xy, radius):
    """Morton-ordered copy of the shared pupil samples [m, 2] (the order of a PSF's ray sum is free; the fused kernel's
    run-length splat is fastest when consecutive samples are neighbours in the pupil)."""
    m = pupil_xy.shape[0]
    out = torch.empty_like(pupil_xy)
    if m == 0:
        return out
    nbytes = lib().sdirt_pupil_sort_workspace(m)
    ws = torch.empty(int(nbytes), device=pupil_xy.device, dtype=torch.uint8)
    _check(lib().sdirt_pupil_sort(_dev(pupil_xy, "pupil_xy"), m, float(radius), _dev(out, "sorted"), C.c_void_p(ws.data_ptr()),
                                  ws.numel(), _stream(pupil_xy)))
    return out


@_on_device
def psf_bank(lens, wvln, points, pupil_xy, pupil_z, centre, ks, pixel_size, dp=None, newton=None, normalise=1,
             want_counts=False, numerics=None):
    """Fused trace + DP splat.  Returns (L [N,ks,ks], R [N,ks,ks][, valid_count [N]])."""
    n, m = points.shape[0], pupil_xy.shape[0]
    dev = points.device
    out_l = torch.empty((n, ks, ks), device=dev, dtype=torch.float32)
    out_r = torch.empty((n, ks, ks), device=dev, dtype=torch.float32)
    cnt = torch.empty((n,), device=dev, dtype=torch.int64) if want_counts else None
    nt, dpp = make_options(newton, numerics), make_dp(dp)
    for a in range(0, n, MAX_POINTS_PER_CALL):
        b = min(a + MAX_POINTS_PER_CALL, n)
        nbytes = lib().sdirt_psf_bank_workspace(b - a, m, ks)
        ws = _workspace(dev, nbytes)
        _check(lib().sdirt_psf_bank(lens._h, float(wvln), _dev(points[a:b], "points"), b - a, _dev(pupil_xy, "pupil_xy"), m,
                                    float(pupil_z), _dev(centre[a:b], "centre"), int(ks), float(pixel_size),
                                    C.byref(dpp) if dpp is not None else None, C.byref(nt), int(normalise),
                                    C.c_void_p(out_l[a:].data_ptr()), C.c_void_p(out_r[a:].data_ptr()),
                                    C.c_void_p(cnt[a:].data_ptr()) if want_counts else None,
                                    C.c_void_p(ws.data_ptr()), ws.numel(), _stream(points)))
    return (out_l, out_r, cnt) if want_counts else (out_l, out_r)


@_on_device
def splat_rays(o, d, ra, centre, ks, pixel_size, dp=None):
    """forward_integral on an existing [spp,N] Ray: returns raw (L, R) [N,ks,ks]."""
    m, n = ra.shape
    dev = o.device
    out_l = torch.empty((n, ks, ks), device=dev, dtype=torch.float32)
    out_r = torch.empty((n, ks, ks), device=dev, dtype=torch.float32)
    if n > MAX_POINTS_PER_CALL:
        raise RuntimeError("sdirt_engine: split forward_integral calls above 65535 points")
    nbytes = lib().sdirt_psf_bank_workspace(n, m, ks) + 32 * n + 128
    ws = _workspace(dev, nbytes)
    dpp = make_dp(dp)
    _check(lib().sdirt_splat_rays(_dev(o, "o"), _dev(d, "d"), _dev(ra, "ra"), m, n,
                                  _dev(centre, "centre") if centre is not None else None, int(ks), float(pixel_size),
                                  C.byref(dpp) if dpp is not None else None, _dev(out_l, "out_l"), _dev(out_r, "out_r"),
                                  C.c_void_p(ws.data_ptr()), ws.numel(), _stream(o)))
    return out_l, out_r


@_on_device
def tone_degamma(img, out=None):
    """PSFNet.degamma of a float32 image (any shape), elementwise; `out` may be `img`."""
    if out is None:
        out = torch.empty_like(img)
    _check(lib().sdirt_tone_degamma(_dev(img, "img"), _dev(out, "out"), img.numel(), _stream(img)))
    return out


@_on_device
def render_local_psf(img, psf, ks, tone=0):
    """img [B,C,H,W] float32, psf [B,H,W,2,ks,ks] float32/float16 -> (rl, rr) float32.
    tone bits: 1 = degamma the input, 2 = gamma + clip the output."""
    b, c, h, w = img.shape
    # degamma once per pixel instead of once per tile halo inside the kernel (same bits) -- unless the call takes the strip-walking
    # kernel, whose image-record pack applies it once per pixel itself
    if tone & 1 and not (psf.dtype == torch.float16 and lib().sdirt_render_records_bytes(b, c, h, w, int(ks)) > 0):
        img, tone = tone_degamma(img), tone & ~1
    if psf.dtype not in (torch.float32, torch.float16):
        raise RuntimeError("sdirt_engine: psf must be float32 or float16")
    if psf.numel() != b * h * w * 2 * ks * ks:
        raise RuntimeError("sdirt_engine: psf has the wrong number of elements for [B,H,W,2,ks,ks]")
    rl, rr = torch.empty_like(img), torch.empty_like(img)
    _check(lib().sdirt_render_local_psf(_dev(img, "img"), _dev(psf, "psf", psf.dtype), int(psf.dtype == torch.float16),
                                        b, c, h, w, int(ks), int(tone), _dev(rl, "out_l"), _dev(rr, "out_r"),
                                        _stream(img)))
    return rl, rr


@_on_device
def render_local_psf_f32(img, psf, ks):
    """img [B,C,H,W] float32, psf [B,H,W,2,ks,ks] float32 -> (rl, rr) float32, float32 arithmetic throughout
    (local_dp_psf_render, render_psf.py:157-188)."""
    b, c, h, w = img.shape
    if psf.dtype != torch.float32 or psf.numel() != b * h * w * 2 * ks * ks:
        raise RuntimeError("sdirt_engine: psf must be float32 [B,H,W,2,ks,ks]")
    rl, rr = torch.empty_like(img), torch.empty_like(img)
    _check(lib().sdirt_render_local_psf_f32(_dev(img, "img"), _dev(psf, "psf"), b, c, h, w, int(ks), _dev(rl, "out_l"),
                                            _dev(rr, "out_r"), _stream(img)))
    return rl, rr


@_on_device
def render_local_psf_rows(img, psf_rows, ks, row0, out_l, out_r, tone=0):
    """Rows [row0, row0 + n_rows) of every image: img [B,C,H,W] float32, psf_rows [B,n_rows,W,2,ks,ks] float32/float16,
    written into the whole-image outputs out_l / out_r [B,C,H,W]."""
    b, c, h, w = img.shape
    n_rows = psf_rows.shape[1]
    if psf_rows.dtype not in (torch.float32, torch.float16):
        raise RuntimeError("sdirt_engine: psf must be float32 or float16")
    if psf_rows.dim() != 6 or psf_rows.numel() != b * n_rows * w * 2 * ks * ks:
        raise RuntimeError("sdirt_engine: psf_rows must be [B,n_rows,W,2,ks,ks]")
    if out_l.shape != img.shape or out_r.shape != img.shape:
        raise RuntimeError("sdirt_engine: outputs must have the image's shape")
    _check(lib().sdirt_render_local_psf_rows(_dev(img, "img"), _dev(psf_rows, "psf_rows", psf_rows.dtype),
                                             int(psf_rows.dtype == torch.float16), b, c, h, w, int(row0), int(n_rows), int(ks),
                                             int(tone), _dev(out_l, "out_l"), _dev(out_r, "out_r"), _stream(img)))
    return out_l, out_r


@_on_device
def render_pack_image(img, ks, tone=0):
    """The image as the strip-walking render kernel streams it (records: padded, degamma'd if tone & 1, fp16), packed once for all
    the bands of a PSFNet.render call; None when that kernel does not take the shape (the caller then uses render_local_psf_rows)."""
    b, c, h, w = img.shape
    n = lib().sdirt_render_records_bytes(b, c, h, w, int(ks))
    if n <= 0:
        return None
    rec = torch.empty((n,), device=img.device, dtype=torch.uint8)
    _check(lib().sdirt_render_pack_image(_dev(img, "img"), b, c, h, w, int(ks), int(tone), _dev(rec, "records", torch.uint8), _stream(img)))
    return rec


@_on_device
def render_local_psf_rows_packed(rec, shape, psf_rows, ks, row0, out_l, out_r, tone=0):
    """render_local_psf_rows from the records of render_pack_image (`shape` = the image's [B,C,H,W]); psf_rows float16."""
    b, c, h, w = shape
    n_rows = psf_rows.shape[1]
    if psf_rows.dtype != torch.float16 or psf_rows.dim() != 6 or psf_rows.numel() != b * n_rows * w * 2 * ks * ks:
        raise RuntimeError("sdirt_engine: psf_rows must be float16 [B,n_rows,W,2,ks,ks]")
    if tuple(out_l.shape) != tuple(shape) or tuple(out_r.shape) != tuple(shape):
        raise RuntimeError("sdirt_engine: outputs must have the image's shape")
    _check(lib().sdirt_render_local_psf_rows_packed(_dev(rec, "records", torch.uint8), _dev(psf_rows, "psf_rows", torch.float16), b, c, h, w,
                                                    int(row0), int(n_rows), int(ks), int(tone), _dev(out_l, "out_l"), _dev(out_r, "out_r"),
                                                    _stream(out_l)))
    return out_l, out_r


@_on_device
def mlp_input_layer(xs, ys, z, b0, nb, row0, n_rows, w1, b1, out=None):
    """First activation of the PSF MLP for the pixels of images [b0, b0+nb), rows [row0, row0+n_rows): out [2P, n1] float16,
    row 2p = left (x, y, z), row 2p+1 = right (-x, y, z), p = ((b-b0)*n_rows + (y-row0))*W + x.  xs [W], ys [H], z [B,H,W]
    float32; w1 [n1,3], b1 [n1] float16."""
    bsz, h, w = z.shape
    n1 = w1.shape[0]
    rows = 2 * nb * n_rows * w
    if out is None:
        out = torch.empty((rows, n1), device=z.device, dtype=torch.float16)
    elif out.shape != (rows, n1):
        raise RuntimeError("sdirt_engine: `out` must be [2 * nb * n_rows * W, n1]")
    if xs.numel() != w or ys.numel() != h or w1.shape != (n1, 3) or b1.numel() != n1:
        raise RuntimeError("sdirt_engine: mlp_input_layer shapes do not match")
    _check(lib().sdirt_mlp_input_layer(_dev(xs, "xs"), _dev(ys, "ys"), _dev(z, "z"), bsz, h, w, int(b0), int(nb), int(row0),
                                       int(n_rows), _dev(w1, "w1", torch.float16), _dev(b1, "b1", torch.float16), n1,
                                       _dev(out, "out", torch.float16), _stream(z)))
    return out


@_on_device
def psf_pack(raw, ks, out=None):
    """raw [2P, ld] float16 (ld >= ks*ks; row 2p left, 2p+1 right, unflipped) -> [P,2,ks,ks] float16 normalised kernels
    (PSFNet.pred: flip the right side, stack, divide by sum + 1e-9)."""
    rows, ld = raw.shape
    if rows % 2:
        raise RuntimeError("sdirt_engine: psf_pack needs an even number of rows (left / right pairs)")
    if out is None:
        out = torch.empty((rows // 2, 2, ks, ks), device=raw.device, dtype=torch.float16)
    elif out.numel() != rows * ks * ks:
        raise RuntimeError("sdirt_engine: `out` must hold [P,2,ks,ks]")
    _check(lib().sdirt_psf_pack(_dev(raw, "raw", torch.float16), rows // 2, int(ld), int(ks), _dev(out, "out", torch.float16),
                                _stream(raw)))
    return out


class FusedMlp:
    """The PSF MLP handed over to the fused tensor-core kernel (sdirt_mlp_fused_pred): first Linear (w1 [n1,3], b1 [n1]) kept
    as is, every later Linear packed once into tcgen05 operand tiles.  `linears`: list of (weight [N,K], bias [N]) CUDA
    tensors in layer order, the first one being the 3 -> n1 layer."""

    @_on_device
    def __init__(self, linears):
        (w1, b1), rest = linears[0], linears[1:]
        if w1.shape[1] != 3 or not rest:
            raise RuntimeError("sdirt_engine: FusedMlp expects Linear(3 -> n1) followed by at least one more Linear")
        dev = w1.device
        sh = MlpShape()
        sh.n_layers, sh.n1 = len(rest), w1.shape[0]
        if sh.n_layers > MLP_MAX_LAYERS:
            raise RuntimeError(f"sdirt_engine: at most {MLP_MAX_LAYERS} layers after the first")
        for l, (w, _) in enumerate(rest):
            sh.N[l], sh.K[l] = w.shape
        bias_floats = C.c_int64(0)
        nbytes = lib().sdirt_mlp_fused_layout(C.byref(sh), None, None, C.byref(bias_floats))
        if nbytes < 0:
            raise RuntimeError("sdirt_engine: " + lib().sdirt_last_error().decode())
        self.shape = sh
        self.w1 = w1.detach().half().contiguous()
        self.b1 = b1.detach().half().contiguous()
        self.packed_w = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        self.packed_b = torch.empty(bias_floats.value, dtype=torch.float32, device=dev)
        for l, (w, b) in enumerate(rest):
            w16, b16 = w.detach().half().contiguous(), b.detach().half().contiguous()
            _check(lib().sdirt_mlp_fused_pack_layer(C.byref(sh), l, _dev(w16, "w", torch.float16), _dev(b16, "b", torch.float16),
                                                    C.c_void_p(self.packed_w.data_ptr()), _dev(self.packed_b, "bias"), _stream(w16)))
        self.n_out = int(sh.N[sh.n_layers - 1])

    @_on_device
    def pred(self, xs, ys, z, b0, nb, row0, n_rows, ks, out=None):
        """Normalised L/R kernels [P,2,ks,ks] float16 of the pixels of images [b0, b0+nb), rows [row0, row0+n_rows)."""
        bsz, h, w = z.shape
        px = nb * n_rows * w
        if out is None:
            out = torch.empty((px, 2, ks, ks), device=z.device, dtype=torch.float16)
        elif out.numel() != px * 2 * ks * ks:
            raise RuntimeError("sdirt_engine: `out` must hold [P,2,ks,ks]")
        if xs.numel() != w or ys.numel() != h:
            raise RuntimeError("sdirt_engine: xs / ys do not match z")
        _check(lib().sdirt_mlp_fused_pred(C.byref(self.shape), C.c_void_p(self.packed_w.data_ptr()), _dev(self.packed_b, "bias"),
                                          _dev(self.w1, "w1", torch.float16), _dev(self.b1, "b1", torch.float16), _dev(xs, "xs"),
                                          _dev(ys, "ys"), _dev(z, "z"), bsz, h, w, int(b0), int(nb), int(row0), int(n_rows), int(ks),
                                          _dev(out, "out", torch.float16), _stream(z)))
        return out


@_on_device
def gamma_noise_clip(x, randn, noise_range, weight):
    """In place on x [N,2C,H,W] float32: clip(gamma(x) + (randn * noise_range[n]) * ramp, 0, 1); ramp = weight[n, col] for the
    left channels, weight[n, W-1-col] for the right ones (PSFNet.gamma / noise / clip of render(train=True))."""
    n, c2, h, w = x.shape
    if randn.shape != x.shape or noise_range.numel() != n or weight.shape != (n, w):
        raise RuntimeError("sdirt_engine: gamma_noise_clip shapes do not match")
    _check(lib().sdirt_gamma_noise_clip(_dev(x, "x"), _dev(randn, "randn"), _dev(noise_range, "noise_range"),
                                        _dev(weight, "weight"), n, c2, h, w, _stream(x)))
    return x


@_on_device
def fp32_peak_probe(device, blocks, threads, iters):
    out = torch.empty(blocks * threads, device=device, dtype=torch.float32)
    _check(lib().sdirt_fp32_peak_probe(_dev(out, "out"), blocks, threads, iters, _stream(out)))
    return out


@_on_device
def debug_trace_strict2(lens, wvln, point, pupil_xy, pupil_z):
    """Testing aid: sensor-plane states [m, 7] = (o, d, alive) of the m rays from one object point, traced by the packed strict
    tracer of the specialised parity kernel (csrc/strict_path.cuh)."""
    m = pupil_xy.shape[0]
    out = torch.zeros((m, 7), device=point.device, dtype=torch.float32)
    _check(lib().sdirt_debug_trace_strict2(lens._h, float(wvln), _dev(point, "point"), _dev(pupil_xy, "pupil_xy"), m, float(pupil_z),
                                           _dev(out, "out"), _stream(point)))
    return out

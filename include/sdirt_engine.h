/*
 * sdirt_engine.h — C ABI of the B200-native dual-pixel ray-tracing engine (libsdirt_engine.so).
 *
 * The reference (LinYark/Sdirt) has no FFI seam: its hot path is eager PyTorch behind Python methods.
 * Each entry point below replaces the body of one such method; the host-side mirror of the reference
 * API (sdirt_b200/deeplens/*.py) binds them with ctypes (see INTEGRATION.md for the stub a reference
 * maintainer would add).  Citations are paths relative to the reference repository.
 *
 *   sdirt_lens_create / _set_*      <- Lensgroup.read_lens_json            deeplens/optics.py:2173-2198
 *                                      Material.ior (Cauchy "n/V")         deeplens/basics.py:316-340
 *   sdirt_trace_rays                <- Lensgroup.trace/_forward_tracing    deeplens/optics.py:601-717
 *                                      Aspheric.ray_reaction               deeplens/surfaces.py:391-520
 *                                      Aspheric._newtons_method/_refract   deeplens/surfaces.py:523-679
 *                                      Ray.propagate_to                    deeplens/basics.py:256-264
 *   sdirt_sample_rays               <- Lensgroup.sample_from_points        deeplens/optics.py:460-494
 *   sdirt_normalize_rays            <- Ray.__init__ (F.normalize)          deeplens/basics.py:245
 *   sdirt_propagate_rays            <- Ray.propagate_to                    deeplens/basics.py:256-264
 *   sdirt_psf_centre                <- Lensgroup.psf_center('chief_ray')   deeplens/optics.py:889-904
 *   sdirt_psf_bank                  <- Lensgroup.psf_diff                  deeplens/optics.py:934-996
 *                                      sample_from_points                  deeplens/optics.py:476-494
 *                                      forward_integral                    deeplens/monte_carlo.py:9-68
 *                                      assign_points_to_pixels_small_r/big_r  deeplens/monte_carlo.py:135-372
 *   sdirt_splat_rays                <- forward_integral on an existing Ray deeplens/monte_carlo.py:9-68
 *   sdirt_render_local_psf(_rows)   <- local_psf_render_fast               deeplens/render_psf.py:120-155
 *                                      PSFNet.degamma / gamma / clip       deeplens/psfnet.py:589-620,706-713
 *   sdirt_mlp_input_layer           <- coordinate grid + first Linear+ReLU deeplens/psfnet.py:681-694; psfnet_arch.py:40-41
 *   sdirt_psf_pack                  <- PSFNet.pred: flip, stack, normalise deeplens/psfnet.py:326-333
 *   sdirt_mlp_fused_pred            <- PSFNet.pred (grid, MLP x2, flip, stack, normalise) deeplens/psfnet.py:317-336, 681-705; psfnet_arch.py:32-56
 *   sdirt_tone_degamma              <- PSFNet.degamma deeplens/psfnet.py:589-603
 *   sdirt_gamma_noise_clip          <- PSFNet.gamma / noise / clip (train) deeplens/psfnet.py:605-620, 629-642, 708-713
 *
 * Conventions: every pointer marked "dev" is device memory owned by the caller (a torch CUDA tensor's
 * data_ptr); the library allocates nothing on the device.  All work is enqueued on `stream`
 * (a cudaStream_t passed as void*, NULL = legacy default stream) and returns without synchronising.
 * Return value: 0 on success, negative SDIRT_E_* otherwise; sdirt_last_error() gives the message of the
 * calling thread's last failure.  Units are millimetres, wavelengths micrometres; float32 arithmetic in
 * the reference's operation order (IEEE add/mul/div/sqrt, no FMA contraction) unless a flag says otherwise.
 */
#ifndef SDIRT_ENGINE_H
#define SDIRT_ENGINE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDIRT_MAX_SURFACES 32
#define SDIRT_MAX_AI 8
#define SDIRT_MAX_KS 63

enum {
    SDIRT_OK = 0,
    SDIRT_E_ARG = -1,      /* bad argument (message says which) */
    SDIRT_E_CUDA = -2,     /* CUDA runtime error */
    SDIRT_E_NODEVICE = -3  /* no CUDA device: there is NO CPU fallback */
};

/* Surface kinds follow the three branches of Aspheric.ray_reaction (surfaces.py:409, 456, 491). */
enum { SDIRT_SURF_FLAT = 0, SDIRT_SURF_SPHERE = 1, SDIRT_SURF_ASPHERE = 2 };

/* One sequential surface as read from the lens JSON (optics.py:2173-2198). */
typedef struct sdirt_surface {
    int32_t kind;                 /* SDIRT_SURF_*; FLAT covers the aperture stop */
    int32_t n_ai;                 /* number of even-asphere coefficients a2,a4,... (0..8) */
    int32_t square;               /* flat only: square aperture of half-side r (surfaces.py:416-419) */
    int32_t reserved;
    double r;                     /* semi-diameter (python float in the reference) */
    float d;                      /* vertex z (float32 tensor in the reference) */
    float c;                      /* curvature */
    float k;                      /* conic constant */
    float ai[SDIRT_MAX_AI];
    double n1_A, n1_B;            /* Cauchy n = A + B/lambda_nm^2 of the object-side medium (basics.py:336-338) */
    double n2_A, n2_B;            /* image-side medium */
} sdirt_surface;

typedef struct sdirt_lens sdirt_lens;   /* opaque */

/* Trace options.
 *
 * newton_mode (surfaces.py:543-561, the loose Newton loop):
 *   SDIRT_NEWTON_PER_RAY  each ray leaves the loop as soon as ITS residual is <= 50e-6 mm;
 *   SDIRT_NEWTON_REPLAY   replay fixed loop counts, iters[i] loop evaluations at lens surface i (this
 *                         reproduces the reference's bundle-global `while any()` count when the caller
 *                         knows it).  Both are followed by the reference's one extra strict evaluation.
 * numerics:
 *   SDIRT_NUMERICS_STRICT float32 in the reference's exact operation order, IEEE add/mul/div/sqrt, no FMA
 *                         contraction: bit-identical to oracle/dp_oracle.py (the verification mode);
 *   SDIRT_NUMERICS_FAST   same geometry, B200-shaped arithmetic: closed-form ray/sphere intersection from
 *                         the vertex plane, FMA, Newton (from the vertex plane, step-converged) only on
 *                         conic/aspheric surfaces.  Agrees with STRICT to float32 rounding noise; see
 *                         DESIGN.md for the measured parity.  newton_mode / iters are ignored;
 *   SDIRT_NUMERICS_HYBRID FAST, except that the first visited surface is traced with the STRICT arithmetic.
 *                         For distant objects the reference's first hit carries up to ~1e-3 mm of float32
 *                         cancellation noise (t ~ 2e3..2e4 mm); reproducing it bit-for-bit is what keeps the
 *                         sensor-pixel assignment identical to the reference's for >= 99.95 % of rays;
 *   SDIRT_NUMERICS_ADAPTIVE HYBRID only where that lattice is coarse: the reference computes the first hit as
 *                         o + d*t in float32, so its lateral position sits on a lattice of ulp(|o_x|), ulp(|o_y|);
 *                         above 2048 mm that is >= 2.4e-4 mm and shows in a 2 M-ray PSF (measured PSF L1 of FAST against
 *                         the reference: <= 2e-5 up to 2048 mm off axis, 5e-5 at 2800 mm, 1.1e-4 at 7000 mm).  Points
 *                         (rays) beyond that bound take the STRICT first surface, the others FAST; the choice is uniform
 *                         per object point. */
enum { SDIRT_NEWTON_REPLAY = 0, SDIRT_NEWTON_PER_RAY = 1 };
enum { SDIRT_NUMERICS_STRICT = 0, SDIRT_NUMERICS_FAST = 1, SDIRT_NUMERICS_HYBRID = 2, SDIRT_NUMERICS_ADAPTIVE = 3 };
/* ADAPTIVE: object points (rays) whose lateral coordinate max(|x|, |y|) exceeds this many mm get the STRICT first surface */
#define SDIRT_ADAPTIVE_LATTICE_MM 2048.0f
typedef struct sdirt_options {
    int32_t newton_mode;
    int32_t numerics;
    int32_t iters[SDIRT_MAX_SURFACES];
} sdirt_options;

/* Dual-pixel sub-aperture model (monte_carlo.py:157-164), pixel units. */
typedef struct sdirt_dp_params {
    float h, f, w, r;
} sdirt_dp_params;

const char *sdirt_last_error(void);
const char *sdirt_version(void);
/* Number of kernels this library has launched in this process (bench.py reports it as gpu_launches). */
uint64_t sdirt_launch_count(void);
/* Device properties the host needs for sizing: SM count of the current device, or <0. */
int sdirt_device_sm_count(void);

/* ---- lens handle -------------------------------------------------------------------------------- */
int sdirt_lens_create(const sdirt_surface *surfaces, int n_surfaces, double d_sensor, sdirt_lens **out);
int sdirt_lens_set_sensor(sdirt_lens *lens, double d_sensor);          /* refocus(), psfnet.py:42-48 */
int sdirt_lens_set_surface(sdirt_lens *lens, int index, const sdirt_surface *s);  /* set_aperture() etc. */
int sdirt_lens_num_surfaces(const sdirt_lens *lens);
/* eta = n_in/n_out at `wvln_um` for each surface, float64 as Material.ior computes it. */
int sdirt_lens_eta(const sdirt_lens *lens, double wvln_um, int backward, double *eta_out);
void sdirt_lens_destroy(sdirt_lens *lens);

/* ---- generic trace (Lensgroup.trace / Aspheric.ray_reaction) -------------------------------------
 * In-place on AoS rays: o[n,3], d[n,3] (unit), ra[n] (0/1 float).  Surfaces [s_begin, s_end) are visited
 * in ascending order, or descending when backward != 0 (optics.py:666-717).  to_sensor != 0 adds
 * Ray.propagate_to(d_sensor).  record (optional, dev) receives [s_end-s_begin, n, 7] = (o, d, ra) after
 * every visited surface. */
int sdirt_trace_rays(const sdirt_lens *lens, double wvln_um,
                     float *o_dev, float *d_dev, float *ra_dev, int64_t n,
                     int s_begin, int s_end, int backward, int to_sensor,
                     const sdirt_options *opts, float *record_dev, void *stream);

/* ---- ray bundle of sample_from_points (optics.py:476-494, basics.py:233-245) ----------------------
 * o_out, d_out: [m, N, 3] sample-major AoS, d = normalize(pupil_j - point_i).  Compatibility path for callers
 * that want a Ray object; sdirt_psf_bank never materialises rays. */
int sdirt_sample_rays(const float *points_dev, int64_t n_points, const float *pupil_xy_dev, int64_t n_samples,
                      double pupil_z, float *o_out_dev, float *d_out_dev, void *stream);

/* ---- Ray.__init__ normalisation (basics.py:245): d /= max(|d|, 1e-12), in place on AoS directions - */
int sdirt_normalize_rays(float *d_dev, int64_t n, void *stream);

/* ---- Ray.propagate_to(z) (basics.py:256-264): o += d * (z - o_z) / d_z, in place on AoS rays ------- */
int sdirt_propagate_rays(float *o_dev, const float *d_dev, int64_t n, double z, void *stream);

/* ---- chief-ray PSF centre (Lensgroup.psf_center) --------------------------------------------------
 * points[N,3] object-space mm; pupil_xy[m,2] samples on the (shrunken) entrance pupil at z = pupil_z,
 * shared by all points.  centre_out[N,2] = -(ra-weighted centroid of sensor hits). */
int sdirt_psf_centre(const sdirt_lens *lens, double wvln_um,
                     const float *points_dev, int64_t n_points,
                     const float *pupil_xy_dev, int64_t n_samples, double pupil_z,
                     const sdirt_options *opts, float *centre_out_dev, void *stream);

/* ---- fused sample -> trace -> DP weights -> splat -> normalise (Lensgroup.psf_diff) ---------------
 * For every point i and pupil sample j a ray from points[i] towards (pupil_xy[j], pupil_z) is traced to
 * the sensor, recentred on centre[i], cropped to the ks x ks window of pixel size `pixel_size`, weighted
 * by the left / right sub-pixel areas and bilinearly splatted into out_l[i], out_r[i] (each [ks,ks]).
 *   normalise: 0 = raw sums, 1 = divide by (max + 1e-6) as psf_diff does, 2 = divide by the sum.
 *   workspace_dev: at least sdirt_psf_bank_workspace(...) bytes, contents undefined on entry and exit.
 *   valid_count_dev (optional): [N] int64, number of rays of each point that reached the window. */
int64_t sdirt_psf_bank_workspace(int64_t n_points, int64_t n_samples, int ks);
int sdirt_psf_bank(const sdirt_lens *lens, double wvln_um,
                   const float *points_dev, int64_t n_points,
                   const float *pupil_xy_dev, int64_t n_samples, double pupil_z,
                   const float *centre_dev, int ks, double pixel_size,
                   const sdirt_dp_params *dp, const sdirt_options *opts, int normalise,
                   float *out_l_dev, float *out_r_dev, int64_t *valid_count_dev,
                   void *workspace_dev, int64_t workspace_bytes, void *stream);

/* ---- spatial ordering of the shared pupil samples (setup of the fused kernel's run-length splat) -----
 * The PSF of a point is a SUM over the pupil samples of sample_from_points (optics.py:483-490), so their order is
 * free.  sdirt_psf_bank gives each thread a contiguous run of samples and accumulates the bilinear taps of
 * consecutive rays in registers while they fall on the same sensor pixel; samples sorted along a Morton curve
 * over the pupil disc (|x|,|y| <= radius) make those runs compact patches of the pupil and hence of the PSF.
 * sorted_out[n,2] receives the permuted samples (must not alias pupil_xy).  Unsorted input to sdirt_psf_bank is
 * still correct, only slower. */
int64_t sdirt_pupil_sort_workspace(int64_t n_samples);
int sdirt_pupil_sort(const float *pupil_xy_dev, int64_t n_samples, double radius, float *sorted_out_dev,
                     void *workspace_dev, int64_t workspace_bytes, void *stream);

/* ---- splat of already-traced rays (forward_integral) ----------------------------------------------
 * o[spp,N,3], d[spp,N,3], ra[spp,N] as the reference's Ray holds them (sample-major).  centre_dev may
 * be NULL: the ra-weighted centroid of each point's hits is used (monte_carlo.py:28-31).  Raw sums. */
int sdirt_splat_rays(const float *o_dev, const float *d_dev, const float *ra_dev,
                     int64_t n_samples, int64_t n_points, const float *centre_dev,
                     int ks, double pixel_size, const sdirt_dp_params *dp,
                     float *out_l_dev, float *out_r_dev, void *workspace_dev, int64_t workspace_bytes,
                     void *stream);

/* ---- spatially varying DP render (local_psf_render_fast) ------------------------------------------
 * img[B,C,H,W] float32, psf[B,H,W,2,ks,ks] (psf_is_half: 0 = float32, 1 = float16), outputs [B,C,H,W]
 * float32.  Products and the final sum are rounded to float16 exactly as the reference's half() path.
 * tone (bit mask): 1 = degamma the input (psfnet.py:589-603, :706); 2 = gamma + clip(0,1) on the output
 * (psfnet.py:605-620, :708-713).  PSFNet.render(train=False) is tone = 3. */
int sdirt_render_local_psf(const float *img_dev, const void *psf_dev, int psf_is_half,
                           int B, int C, int H, int W, int ks, int tone,
                           float *out_l_dev, float *out_r_dev, void *stream);

/* local_dp_psf_render (deeplens/render_psf.py:157-188): the one variant of the three the reference runs in the INPUT's dtype --
 * float32 image, float32 kernels, float32 products and sums, nothing rounded to half.  img[B,C,H,W], psf[B,H,W,2,ks,ks],
 * outputs [B,C,H,W], all float32. */
int sdirt_render_local_psf_f32(const float *img_dev, const float *psf_dev, int B, int C, int H, int W, int ks,
                               float *out_l_dev, float *out_r_dev, void *stream);

/* The same for a window of rows [row0, row0 + n_rows) of every image: psf_rows_dev is [B, n_rows, W, 2, ks, ks] (the
 * kernels of those rows only), img and the outputs are the whole [B,C,H,W] tensors (the window's halo rows are read
 * from img, replicate padding applies at the IMAGE border only).  PSFNet.render walks an image in such bands so that
 * the per-pixel PSF tensor (psfnet.py:705; 2.8 GB per 1024 x 1536 image in fp16) is only ever a band. */
int sdirt_render_local_psf_rows(const float *img_dev, const void *psf_rows_dev, int psf_is_half,
                                int B, int C, int H, int W, int row0, int n_rows, int ks, int tone,
                                float *out_l_dev, float *out_r_dev, void *stream);

/* The banded render with the image prepared ONCE per PSFNet.render call instead of once per band (psfnet.py:706-708: the
 * reference's degamma / local_psf_render_fast see the whole image once).  The strip-walking render kernel reads the image as
 * "records" (replicate-padded rows, degamma applied, rounded to fp16, column-mirrored per 32-pixel strip):
 *   sdirt_render_records_bytes: size of that buffer for a [B,C,H,W] image, or 0 when the kernel does not take the shape
 *     (it takes C = 3, W a multiple of 32, ks 7 / 11 / 21: other callers use sdirt_render_local_psf_rows);
 *   sdirt_render_pack_image: img[B,C,H,W] float32 -> records (tone bit 1 = degamma first);
 *   sdirt_render_local_psf_rows_packed: sdirt_render_local_psf_rows from those records and float16 kernels
 *     psf_rows_half_dev[B, n_rows, W, 2, ks, ks] (tone bit 2 = gamma + clip on the output).  Both buffers 16-byte aligned. */
int64_t sdirt_render_records_bytes(int B, int C, int H, int W, int ks);
int sdirt_render_pack_image(const float *img_dev, int B, int C, int H, int W, int ks, int tone, void *records_dev, void *stream);
int sdirt_render_local_psf_rows_packed(const void *records_dev, const void *psf_rows_half_dev, int B, int C, int H, int W,
                                       int row0, int n_rows, int ks, int tone, float *out_l_dev, float *out_r_dev, void *stream);

/* ---- the two ends of PSFNet.pred inside PSFNet.render (psfnet.py:317-336, 681-705; psfnet_arch.py:40-41) ----------
 * sdirt_mlp_input_layer: for the pixels of images [b0, b0 + nb), rows [row0, row0 + n_rows) build the MLP's first
 * activation.  Pixel p = ((b - b0) * n_rows + (y - row0)) * W + x owns output rows 2p (left: input (xs[x], ys[y],
 * z[b,y,x])) and 2p + 1 (right: (-xs[x], ys[y], z[b,y,x]), psfnet.py:328).  Row = relu(W1 . input + b1) with CUDA
 * autocast's arithmetic: inputs / weights / bias in fp16, fp32 accumulation, one rounding to fp16.
 *   xs[W], ys[H]: the linspace(-1,1,W) / linspace(1,-1,H) axes (psfnet.py:684-688); z[B,H,W]: depth2z(depth);
 *   w1_half[n1,3], b1_half[n1]: the first Linear; out_half[2 * nb * n_rows * W, n1] fp16.
 * sdirt_psf_pack: raw_half[2 * n_pixels, ld] fp16 = the last Linear + ReLU of those rows (ld >= ks*ks: the GEMM may
 * pad its N) -> psf_half[n_pixels, 2, ks, ks] fp16 = stack(left, flip(right, -1)) / (sum(-1).sum(-1) + 1e-9) with
 * torch's fp16 rounding points (psfnet.py:329-333).  All-zero kernels stay zero (the reference writes NaN there). */
int sdirt_mlp_input_layer(const float *xs_dev, const float *ys_dev, const float *z_dev, int B, int H, int W,
                          int b0, int nb, int row0, int n_rows, const void *w1_half_dev, const void *b1_half_dev,
                          int n1, void *out_half_dev, void *stream);
int sdirt_psf_pack(const void *raw_half_dev, int64_t n_pixels, int ld, int ks, void *psf_half_dev, void *stream);

/* ---- PSFNet.pred for a window of pixels as ONE tensor-core kernel (csrc/mlp_fused.cuh) -------------------------------
 * Replaces, for the pixels of images [b0, b0 + nb), rows [row0, row0 + n_rows): the coordinate grid, both MLP
 * evaluations under CUDA autocast (psfnet_arch.py:32-56), flip / stack / normalise (psfnet.py:317-336) -- i.e.
 * sdirt_mlp_input_layer + the cuBLAS GEMM chain + sdirt_psf_pack -- and writes psf_half_dev[n_pixels, 2, ks, ks].
 * The MLP is described by its shape: a first Linear(3 -> n1) + ReLU, then n_layers Linear(K[l] -> N[l]) + ReLU whose
 * weights are handed over once through sdirt_mlp_fused_pack_layer (tcgen05 operand tiles, SWIZZLE_128B) into caller-owned
 * buffers of sdirt_mlp_fused_layout's sizes.  Hidden widths: multiples of 64 up to 512; N[last] = ks*ks; ks in {7, 11, 21};
 * the window must hold a multiple of 4 pixels. */
#define SDIRT_MLP_MAX_LAYERS 12
typedef struct sdirt_mlp_shape {
    int32_t n_layers;                     /* tensor-core layers, i.e. every Linear after the first */
    int32_t n1;                           /* width of the first Linear (64 or 128) */
    int32_t K[SDIRT_MLP_MAX_LAYERS];      /* input width of each layer (K[0] = n1, K[l] = N[l-1]) */
    int32_t N[SDIRT_MLP_MAX_LAYERS];      /* output width of each layer (true, unpadded) */
} sdirt_mlp_shape;
/* Returns the bytes of the packed-weight buffer (or -1); fills per-layer offsets and the float count of the bias buffer. */
int64_t sdirt_mlp_fused_layout(const sdirt_mlp_shape *shape, int64_t *w_off_out, int32_t *b_off_out, int64_t *bias_floats_out);
int sdirt_mlp_fused_pack_layer(const sdirt_mlp_shape *shape, int layer, const void *w_half_dev /*[N,K]*/,
                               const void *b_half_dev /*[N]*/, void *packed_w_dev, float *packed_bias_dev, void *stream);
int sdirt_mlp_fused_pred(const sdirt_mlp_shape *shape, const void *packed_w_dev, const float *packed_bias_dev,
                         const void *w1_half_dev /*[n1,3]*/, const void *b1_half_dev /*[n1]*/,
                         const float *xs_dev, const float *ys_dev, const float *z_dev, int B, int H, int W,
                         int b0, int nb, int row0, int n_rows, int ks, void *psf_half_dev, void *stream);

/* Measurement helper: a device buffer [grid][8] of int64 that the next sdirt_mlp_fused_pred launches fill with cycle counters
 * (MMA loop, its waits for weights / for the epilogue; epilogue loop, its waits); NULL switches it off. */
void sdirt_mlp_fused_debug(long long *counters_dev);
/* Kernel variant of sdirt_mlp_fused_pred: 2 (default) = CTA pairs (thread-block cluster of 2, tcgen05 cta_group::2, each CTA
 * stages half of every weight tile), 1 = single CTAs.  Same results either way; returns the previous setting (other values of
 * `ncta` only query). */
int sdirt_mlp_fused_cta_group(int ncta);

/* gamma -> sensor noise -> clip(0,1), the tail of PSFNet.render(train=True) (psfnet.py:605-620, 629-642, 708-713), in
 * place on x_dev[N, 2C, H, W] (the convolved linear image, left channels first).  randn_dev: standard-normal draws of
 * that shape; noise_range_dev[N]; weight_dev[N, W]: the linspace(range1, range2, W) ramp, read mirrored for the right
 * channels.  x <- clip(gamma(x) + (randn * noise_range) * weight, 0, 1). */
/* PSFNet.degamma (psfnet.py:589-603) elementwise, out_dev may equal in_dev: the linear image that sdirt_render_local_psf(_rows)
 * then takes with tone bit 1 clear (same bits as the fused degamma; evaluated once per pixel instead of once per tile halo). */
int sdirt_tone_degamma(const float *in_dev, float *out_dev, int64_t n, void *stream);

int sdirt_gamma_noise_clip(float *x_dev, const float *randn_dev, const float *noise_range_dev, const float *weight_dev,
                           int N, int C2, int H, int W, void *stream);

/* ---- measurement helper: dependent-free FP32 FMA loop, returns nothing; timed by the caller -------- */
/* Testing aid: runs the packed (two rays per thread) strict first-surface step and the one-ray one on the m rays from one object
 * point and counts, per field (o.x o.y o.z d.x d.y d.z, alive flag, rays), those whose bits differ: mismatch_dev[8], example_dev[16]. */
int sdirt_debug_strict_pair(const sdirt_lens *lens, double wvln_um, const float *point_dev, const float *pupil_xy_dev, int64_t n_samples,
                            double pupil_z, int *mismatch_dev, float *example_dev, void *stream);

/* Testing aid: the m rays from one object point towards pupil_xy[m,2], traced by the packed (two rays per thread) strict tracer of
 * the specialised parity kernel (csrc/strict_path.cuh) and propagated to the sensor: out_dev[m,7] = (o, d, alive).  Compared bit
 * for bit with sdirt_trace_rays (numerics STRICT, per-ray Newton) by the tests. */
int sdirt_debug_trace_strict2(const sdirt_lens *lens, double wvln_um, const float *point_dev, const float *pupil_xy_dev, int64_t n_samples,
                              double pupil_z, float *out_dev, void *stream);

int sdirt_fp32_peak_probe(float *out_dev, int blocks, int threads, int iters, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SDIRT_ENGINE_H */

// strict_path.cuh — the reference's own float32 arithmetic on EVERY surface, at the speed of the specialised kernels.
//
// Included by engine.cu after fast_path.cuh.  `psf_bank_kernel<STRICT>` (engine.cu) restates the reference operation by
// operation but pays for a run-time surface loop, one ray per thread, six acos per ray and eight shared-memory atomics per
// ray.  Here the same arithmetic -- separately rounded products and sums, IEEE division and square root, the per-ray Newton
// loop of surfaces.py:543-578 followed by the one extra strict evaluation, the validity rules of surfaces.py:464 / 584, the
// vector Snell step of surfaces.py:633-679 -- runs
//   * with the lens structure as a template parameter (lens_sigs.inc): no kind is tested at run time and every prescription
//     value is a constant-bank operand;
//   * for TWO rays per thread in packed fp32 (FMUL2 / FADD2 / FFMA2 issue once for both halves; `mul2s` keeps ptxas from
//     contracting a product into the sum that uses it), the two Newton loops in lock step;
//   * into the run-length register splat of the fast kernels, with the pixel index computed by the reference's own
//     expressions (monte_carlo.py:209-216) and the sub-pixel weights d_l / d_r read from the table.
// Only operations whose result is EXACTLY the reference's are restated differently:
//   x = m ? nx : 0, y = m ? ny : 0, x*x + y*y          ==  m ? (nx*nx + ny*ny) : 0
//   ((1 + k) r2) c2 with k = 0 (spheres)                ==  r2 * c2
//   dg * (2 * (a t + b)) - dz                           ==  fma(2, dg * (a t + b), -dz)     (scaling by 2 is exact)
//   normalize((+-2x, +-2y, +-(2z - 2(d+R))))            ==  +-(x, y, z - (d+R)) / sqrt(x^2 + y^2 + (z-(d+R))^2)   (powers of two)
//   sqrt(x^2 + y^2) <= r  (stop)                        ==  x^2 + y^2 <= max{v : fl(sqrt(v)) <= r}   (sqrt is monotonic)
// so the per-surface ray states are bit-identical to `trace_rays_kernel<STRICT>` with the per-ray Newton schedule
// (tests/test_engine_gpu.py::test_strict_pair_tracer_is_bit_identical), which is bit-identical to oracle/dp_oracle.py.
#pragma once

// Even-asphere terms for N <= 6 coefficients (surfaces.py:787-830, the explicit-powers branches), on each half: rho^4, rho^6 are
// float32 products, higher powers the float64 product rounded once (poly_terms<N> in engine.cu).
template <int N, bool WANT_G>
__device__ __forceinline__ void poly_terms2(const SurfDev &s, f2 r2, f2 &g, f2 &dg) {
    static_assert(N >= 1 && N <= 6, "packed strict polynomial: 1..6 coefficients");
    f2 p[N + 1];
    p[1] = r2;
    if (N >= 2) p[2] = mul2s(r2, r2);
    if (N >= 3) p[3] = mul2s(p[2], r2);
    if (N >= 4) {
        const double x0 = (double)r2.x, x1 = (double)r2.y;
        double q0 = x0 * x0 * x0, q1 = x1 * x1 * x1;
#pragma unroll
        for (int i = 4; i <= N; ++i) { q0 = q0 * x0; q1 = q1 * x1; p[i] = make_float2((float)q0, (float)q1); }
    }
    if (WANT_G) {
#pragma unroll
        for (int i = 1; i <= N; ++i) g = add2(g, mul2s(bc2(s.ai[i - 1]), p[i]));
    }
    dg = add2(dg, bc2(s.ai[0]));
#pragma unroll
    for (int i = 2; i <= N; ++i) dg = add2(dg, mul2s(bc2(s.dai[i - 1]), p[i - 1]));     // dai[i-1] = fl(i * ai[i-1])
}

// IEEE quotient a / b on each half from a reciprocal SEED r0 ~ 1/b good to a few ulp: one Newton step on the reciprocal, the
// quotient, its exact remainder, the correction -- div_rn2 with the seed supplied by the caller instead of MUFU.RCP (the
// refined reciprocal is within half an ulp of 1/b either way, which is all the correction step asks for).
__device__ __forceinline__ f2 div_seed2(f2 a, f2 b, f2 r0) {
    const f2 r = fma2(r0, fma2(neg2(b), r0, bc2(1.0f)), r0);
    const f2 q = mul2(a, r);
    return fma2(r, fma2(neg2(b), q, a), q);
}
#ifndef SDIRT_STRICT_SEED_SHARE
#define SDIRT_STRICT_SEED_SHARE 1     // 1: 1/sf seeded by the square root's own MUFU.RSQ, 1/(1+sf)^2 by the square of 1/(1+sf): 3 MUFU per evaluation, not 5
#endif
#ifndef SDIRT_STRICT_SHORT_DIV
#define SDIRT_STRICT_SHORT_DIV 0      // experiment only: quotients without the reciprocal's Newton step (misrounds ~4e-7 of them)
#endif
__device__ __forceinline__ f2 sdiv2(f2 a, f2 b) {
#if SDIRT_STRICT_SHORT_DIV
    const f2 r = rcp2(b);
    const f2 q = mul2(a, r);
    return fma2(r, fma2(neg2(b), q, a), q);
#else
    return div_rn2(a, b);
#endif
}

// sag G(rho^2) and slope G'(rho^2) (sag_and_slope in engine.cu) on each half; the square root is shared.
template <int KIND, int NAI, bool WANT_G>
__device__ __forceinline__ void sag_slope_strict2(const SurfDev &s, f2 r2, f2 &g, f2 &dg) {
    const f2 kr2c2 = KIND == SIG_SPHERE ? mul2s(r2, bc2(s.c2)) : mul2s(mul2s(bc2(s.onek), r2), bc2(s.c2));
#if SDIRT_STRICT_SEED_SHARE && !SDIRT_STRICT_SHORT_DIV
    const f2 x = add2(bc2(1.0f), neg2(kr2c2));
    const f2 y = rsq2(x);                                                    // sqrt_rn2, keeping its seed
    const f2 sq = mul2(x, y);
    const f2 sf = fma2(fma2(neg2(sq), sq, x), mul2(y, bc2(0.5f)), sq);
    const f2 one_sf = add2(bc2(1.0f), sf);
    const f2 r1 = rcp2(one_sf);
    const f2 r1n = fma2(r1, fma2(neg2(one_sf), r1, bc2(1.0f)), r1);          // 1 / (1 + sf), refined once (div_rn2's reciprocal)
    if (WANT_G) { const f2 num = mul2s(r2, bc2(s.c)); const f2 q = mul2(num, r1n); g = fma2(r1n, fma2(neg2(one_sf), q, num), q); }
    const f2 inner = div_seed2(mul2s(kr2c2, bc2(0.5f)), sf, y);
    dg = div_seed2(mul2s(add2(one_sf, inner), bc2(s.c)), mul2s(one_sf, one_sf), mul2(r1n, r1n));
#else
    const f2 sf = sqrt_rn2(add2(bc2(1.0f), neg2(kr2c2)));
    const f2 one_sf = add2(bc2(1.0f), sf);
    if (WANT_G) g = sdiv2(mul2s(r2, bc2(s.c)), one_sf);
    dg = sdiv2(mul2s(add2(one_sf, sdiv2(mul2s(kr2c2, bc2(0.5f)), sf)), bc2(s.c)), mul2s(one_sf, one_sf));
#endif
    if constexpr (KIND == SIG_ASPHERE && NAI > 0) poly_terms2<NAI, WANT_G>(s, r2, g, dg);
}

template <int KIND>
__device__ __forceinline__ bool newton_mask(const SurfDev &s, float r2u, bool strict) {
    if (KIND == SIG_SPHERE) return r2u < (strict ? s.thr_strict : s.bound);       // k = 0 > -1: both masks are upper bounds
    return strict ? strict_mask(s, r2u) : loose_mask(s, r2u);
}

// Newton intersection (surfaces.py:523-586) for a pair of rays with the per-ray schedule: a half leaves the loose loop when ITS
// residual is within 50e-6 mm (or after 10 evaluations), the loop runs until both have left, then one strict evaluation for both.
// FIRST (the first surface of the lens): the object may be metres away, t sits on a coarse float32 lattice and the map
// t -> t_new ends in a fixed point or a 2-cycle long before the residual test is met; both are detected and the rest of the
// loop is skipped with its known outcome (newton_strict in engine.cu does the same).
template <int KIND, int NAI, bool FIRST>
__device__ __forceinline__ void newton_strict2(const SurfDev &s, const Ray2 &r, f2 &t_out, f2 &ft_last) {
    const f2 t0 = div_rn2(add2(bc2(s.d), neg2(r.oz)), r.dz);
    const f2 a = add2(mul2s(r.dx, r.dx), mul2s(r.dy, r.dy));
    const f2 b = add2(mul2s(r.dx, r.ox), mul2s(r.dy, r.oy));
    f2 t = t0, ft = bc2(MAXT_F), tb = t0;
    bool run0 = r.a0, run1 = r.a1;                     // still inside the loose loop
    bool strict = false;
    int it = 0;
    for (;;) {
        if (!strict) {
            run0 = run0 && fabsf(ft.x) > NEWTON_LOOSE;
            run1 = run1 && fabsf(ft.y) > NEWTON_LOOSE;
            if (!((run0 || run1) && it < NEWTON_MAXIT)) {
                strict = true;
                t = add2(t0, add2(t, neg2(t0)));                                   // surfaces.py:563-567
            }
        }
        const f2 nx = add2(r.ox, mul2s(r.dx, t)), ny = add2(r.oy, mul2s(r.dy, t)), nz = add2(r.oz, mul2s(r.dz, t));
        const f2 r2u = add2(mul2s(nx, nx), mul2s(ny, ny));
        const f2 r2 = make_float2(newton_mask<KIND>(s, r2u.x, strict) ? r2u.x : 0.0f, newton_mask<KIND>(s, r2u.y, strict) ? r2u.y : 0.0f);
        f2 g, dg;
        sag_slope_strict2<KIND, NAI, true>(s, r2, g, dg);
        const f2 ftn = add2(add2(g, bc2(s.d)), neg2(nz));
        const f2 dfdt = fma2(bc2(2.0f), mul2s(dg, add2(mul2s(a, t), b)), neg2(r.dz));
        f2 step = sdiv2(ftn, add2(dfdt, bc2(EPS_F)));
        step = make_float2(fminf(fmaxf(step.x, -NEWTON_STEP), NEWTON_STEP), fminf(fmaxf(step.y, -NEWTON_STEP), NEWTON_STEP));
        const f2 tn = add2(t, neg2(step));
        if (strict) { ft_last = ftn; t_out = tn; return; }
        ++it;
        if (FIRST) {
            // period 1 (t_new == t): every further evaluation repeats this one.  period 2 (t_new == the t before this
            // evaluation's input) with the residual still above the tolerance: t alternates until the cap, the parity of
            // the evaluations left picks the survivor.
            const bool p1x = tn.x == t.x, p1y = tn.y == t.y;
            const bool p2x = !p1x && it >= 2 && tn.x == tb.x && fabsf(ftn.x) > NEWTON_LOOSE;
            const bool p2y = !p1y && it >= 2 && tn.y == tb.y && fabsf(ftn.y) > NEWTON_LOOSE;
            const bool odd = ((NEWTON_MAXIT - it) & 1) != 0;
            const bool keepx = p2x && odd, keepy = p2y && odd;
            if (run0) { ft.x = ftn.x; tb.x = t.x; if (!keepx) t.x = tn.x; if (p1x || p2x) run0 = false; }
            if (run1) { ft.y = ftn.y; tb.y = t.y; if (!keepy) t.y = tn.y; if (p1y || p2y) run1 = false; }
        } else {
            if (run0) { ft.x = ftn.x; t.x = tn.x; }
            if (run1) { ft.y = ftn.y; t.y = tn.y; }
        }
    }
}

// Snell (surfaces.py:633-679, forward direction) for a pair, given q = the unit normal up to the sign `sigma` the reference
// gives it (n = sigma q): cos_i = sigma (d.q), and in d' = sr n + eta (d - cos_i n) the sign survives only on the sr term.
__device__ __forceinline__ void refract_strict2(const SurfDev &s, Ray2 &r, f2 qx, f2 qy, f2 qz, float sigma, bool &v0, bool &v1) {
    const f2 cq = add2(add2(mul2s(r.dx, qx), mul2s(r.dy, qy)), mul2s(r.dz, qz));
    const f2 c2 = mul2s(cq, cq);
    const f2 e = mul2s(bc2(s.eta2), add2(bc2(1.0f), neg2(c2)));
    v0 = v0 && (c2.x > 0.1f) && (e.x < 1.0f);
    v1 = v1 && (c2.y > 0.1f) && (e.y < 1.0f);
    const f2 sr = mul2(bc2(sigma), sqrt_rn2(add2(bc2(1.0f), neg2(e))));
    r.dx = add2(mul2s(sr, qx), mul2s(bc2(s.eta), add2(r.dx, neg2(mul2s(cq, qx)))));
    r.dy = add2(mul2s(sr, qy), mul2s(bc2(s.eta), add2(r.dy, neg2(mul2s(cq, qy)))));
    r.dz = add2(mul2s(sr, qz), mul2s(bc2(s.eta), add2(r.dz, neg2(mul2s(cq, qz)))));
}

// Aspheric.ray_reaction (surfaces.py:391-520) for surface J of structure SIG and a pair of rays; r.oz is ABSOLUTE here (the
// reference's coordinates), unlike the fast tracer's vertex-relative z.  A dead half carries garbage that nothing reads.
template <class SIG, int J>
__device__ __forceinline__ void strict_step2(const LensDev &L, Ray2 &r) {
    const SurfDev &s = L.s[J];
    constexpr int K = SIG::kind(J);
    if (K == SIG_STOP || K == SIG_FLATREFR) {
        const f2 t = div_rn2(add2(bc2(s.d), neg2(r.oz)), r.dz);
        r.ox = add2(r.ox, mul2s(t, r.dx)); r.oy = add2(r.oy, mul2s(t, r.dy)); r.oz = add2(r.oz, mul2s(t, r.dz));
        const f2 r2u = add2(mul2s(r.ox, r.ox), mul2s(r.oy, r.oy));
        bool v0 = r.a0 && (r2u.x <= s.r2_sqrt_le), v1 = r.a1 && (r2u.y <= s.r2_sqrt_le);     // sqrt(x^2 + y^2) <= r, surfaces.py:421
        if (K == SIG_FLATREFR) refract_strict2(s, r, bc2(0.0f), bc2(0.0f), bc2(1.0f), 1.0f, v0, v1);   // n = -normalize((0,0,-1)) = (0,0,1)
        r.a0 = v0;
        r.a1 = v1;
        return;
    }
    f2 t, ft_last;
    newton_strict2<K, SIG::nai(J), J == 0>(s, r, t, ft_last);
    r.ox = add2(r.ox, mul2s(t, r.dx)); r.oy = add2(r.oy, mul2s(t, r.dy)); r.oz = add2(r.oz, mul2s(t, r.dz));
    const f2 r2u = add2(mul2s(r.ox, r.ox), mul2s(r.oy, r.oy));
    bool v0, v1;
    if (K == SIG_SPHERE) {
        v0 = r.a0 && (r2u.x <= s.r2) && (t.x >= 0.0f);                                       // surfaces.py:464
        v1 = r.a1 && (r2u.y <= s.r2) && (t.y >= 0.0f);
        // gradient (+-2x, +-2y, +-(2z - 2(d+R))) = +-2 (x, y, w): the normalised vector is +-(x, y, w) / |(x, y, w)| bit for bit
        const f2 w = add2(r.oz, bc2(-s.dR));
        const f2 nrm = sqrt_rn2(fma2(w, w, fma2(r.oy, r.oy, mul2s(r.ox, r.ox))));             // norm3
        refract_strict2(s, r, div_rn2(r.ox, nrm), div_rn2(r.oy, nrm), div_rn2(w, nrm), (s.flags & F_CPOS) ? -1.0f : 1.0f, v0, v1);
    } else {
        v0 = r.a0 && strict_mask(s, r2u.x) && (fabsf(ft_last.x) < NEWTON_TIGHT) && (t.x > 0.0f);   // surfaces.py:584
        v1 = r.a1 && strict_mask(s, r2u.y) && (fabsf(ft_last.y) < NEWTON_TIGHT) && (t.y > 0.0f);
        f2 g, dg;
        sag_slope_strict2<K, SIG::nai(J), false>(s, r2u, g, dg);                              // (x, y masked by ra > 0: alive here)
        const f2 dg2 = mul2s(dg, bc2(2.0f));
        const f2 gx = mul2s(dg2, r.ox), gy = mul2s(dg2, r.oy);
        f2 nrm = sqrt_rn2(add2(fma2(gy, gy, mul2s(gx, gx)), bc2(1.0f)));                     // norm3(gx, gy, -1)
        refract_strict2(s, r, div_rn2(gx, nrm), div_rn2(gy, nrm), div_rn2(bc2(-1.0f), nrm), -1.0f, v0, v1);
    }
    r.a0 = v0;
    r.a1 = v1;
}

template <class SIG, int J>
__device__ __forceinline__ void trace_strict2(const LensDev &L, Ray2 &r) {
    if constexpr (J < SIG::N) {
        strict_step2<SIG, J>(L, r);
        if (!(r.a0 || r.a1)) return;
        trace_strict2<SIG, J + 1>(L, r);
    }
}

// can the packed strict tracer take this (resolved) lens?  per-ray Newton schedule, forward, structure SIG, <= 6 polynomial terms
template <class SIG>
static bool strict_sig_ok(const LensDev &L) {
    if (!sig_matches<SIG>(L)) return false;
    for (int j = 0; j < SIG::N; ++j) {
        if (SIG::kind(j) == SIG_ASPHERE && SIG::nai(j) > 6) return false;
        if (L.s[j].fixed_iters >= 0) return false;
    }
    return true;
}
template <class SIG>
__host__ __device__ constexpr bool strict_sig_compiles() {
    for (int j = 0; j < SIG::N; ++j)
        if (SIG::kind(j) == SIG_ASPHERE && SIG::nai(j) > 6) return false;
    return true;
}

#ifndef SDIRT_STRICT_MIN_CTAS
#define SDIRT_STRICT_MIN_CTAS 3       // 256-thread CTAs per SM: 80 registers per thread for the pair's state + the Newton loop
#endif
template <class SIG>
struct TraceStrictSig {
    static constexpr bool PAIR = true;
    static constexpr bool STRICT_SPLAT = true;
    static constexpr int MIN_CTAS = SDIRT_STRICT_MIN_CTAS;
    static __device__ __forceinline__ bool trace(const LensDev &L, RayReg &r, bool) {      // one-ray loop (debug switch): the generic strict trace
        trace_lens<STRICT, false>(L, r, nullptr, 0, 0, false);
        return r.alive;
    }
    static __device__ __forceinline__ float sensor_distance(const LensDev &L, const RayReg &r) { return L.d_sensor - r.oz; }
    static __device__ __forceinline__ Ray2 trace2(const LensDev &L, float px, float py, float pz, float2 s0, float2 s1, float pupil_z, bool) {
        Ray2 r = ray2_from_point(px, py, pz, s0, s1, pupil_z);
        if constexpr (strict_sig_compiles<SIG>()) trace_strict2<SIG, 0>(L, r);
        else r.a0 = r.a1 = false;
        return r;
    }
};

// ---- testing aid: the packed strict tracer's sensor-plane ray states, to be compared bit for bit with sdirt_trace_rays ----
template <class SIG>
__global__ void __launch_bounds__(128)
debug_trace_strict2_kernel(const __grid_constant__ LensDev L, const float *__restrict__ point, const float2 *__restrict__ pupil, int64_t m,
                           float pupil_z, float *__restrict__ out /*[m,7]*/) {
    const int64_t j = 2 * ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
    if (j >= m) return;
    const bool two = j + 1 < m;
    Ray2 r = TraceStrictSig<SIG>::trace2(L, point[0], point[1], point[2], pupil[j], two ? pupil[j + 1] : pupil[j], pupil_z, false);
    // Ray.propagate_to(d_sensor), basics.py:262-263
    const f2 t = div_rn2(add2(bc2(L.d_sensor), neg2(r.oz)), r.dz);
    const f2 sx = add2(r.ox, mul2s(r.dx, t)), sy = add2(r.oy, mul2s(r.dy, t)), sz = add2(r.oz, mul2s(r.dz, t));
    float *p = out + j * 7;
    p[0] = sx.x; p[1] = sy.x; p[2] = sz.x; p[3] = r.dx.x; p[4] = r.dy.x; p[5] = r.dz.x; p[6] = r.a0 ? 1.f : 0.f;
    if (two) { p += 7; p[0] = sx.y; p[1] = sy.y; p[2] = sz.y; p[3] = r.dx.y; p[4] = r.dy.y; p[5] = r.dz.y; p[6] = r.a1 ? 1.f : 0.f; }
}

// The parity mode of sdirt_psf_bank (numerics STRICT, per-ray Newton schedule) on the specialised kernel when the lens
// structure was compiled; returns 1 if no structure matched (the caller then takes the generic kernel).
static int launch_bank_strict(const LensDev &L, const SplatDev &P, dim3 grid, cudaStream_t st,
                              const float *points, const float2 *pupil, int64_t m, float pupil_z, const float *centre,
                              float4 *lut, int64_t chunk, int run, float *partial, int *hits) {
#define SDIRT_SIG_TRY(NAME)                                                                                                     \
    if (strict_sig_compiles<NAME>() && strict_sig_ok<NAME>(L)) {                                                                \
        dp_lut_kernel<<<DP_LUT_N / 256, 256, 0, st>>>(P, lut);                                                                  \
        if (int rc = check_launch("dp_lut_kernel")) return rc;                                                                  \
        return launch_bank_run<TraceStrictSig<NAME>>(L, P, grid, st, points, pupil, m, pupil_z, centre, lut, chunk, run, partial, hits); \
    }
    SDIRT_SIG_LIST(SDIRT_SIG_TRY)
#undef SDIRT_SIG_TRY
    return 1;
}

static int launch_debug_trace_strict2(const LensDev &L, cudaStream_t st, const float *point, const float2 *pupil, int64_t m, float pupil_z, float *out) {
    const unsigned blocks = (unsigned)(((m + 1) / 2 + 127) / 128);
#define SDIRT_SIG_TRY(NAME)                                                                                                     \
    if (strict_sig_compiles<NAME>() && strict_sig_ok<NAME>(L)) {                                                                \
        debug_trace_strict2_kernel<NAME><<<blocks, 128, 0, st>>>(L, point, pupil, m, pupil_z, out);                             \
        return check_launch("debug_trace_strict2_kernel");                                                                     \
    }
    SDIRT_SIG_LIST(SDIRT_SIG_TRY)
#undef SDIRT_SIG_TRY
    return fail(SDIRT_E_ARG, "sdirt_debug_trace_strict2: no compiled lens structure matches this lens");
}

// strict_path.cuh — the reference's own float32 arithmetic on EVERY surface, at the speed of the specialised kernels.
//
// Included by engine.cu after fast_path.cuh.  `psf_bank_kernel<STRICT>` (engine.cu) restates the reference operation by
// operation but pays for a run-time surface loop, one ray per thread, six acos per ray and eight shared-memory atomics per
// ray.  Here the same arithmetic -- separately rounded products and sums, IEEE division and square root, the per-ray Newton
// loop of surfaces.py:543-578 followed by the one extra strict evaluation, the validity rules of surfaces.py:464 / 584, the
// vector Snell step of surfaces.py:633-679 -- runs
//   * as ONE copy of the sphere / flat / asphere code in a run-time loop over the surfaces, every prescription value a
//     uniform load (any lens, not only the compiled structures);
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
    // a2 = 0 (the usual case: the vertex curvature is carried by c): g + 0 * rho^2 and dg + 0 are g and dg themselves
    const bool a2 = !(s.flags & F_A2ZERO);
    if (WANT_G) {
        if (a2) g = add2(g, mul2s(bc2(s.ai[0]), p[1]));
#pragma unroll
        for (int i = 2; i <= N; ++i) g = add2(g, mul2s(bc2(s.ai[i - 1]), p[i]));
    }
    if (a2) dg = add2(dg, bc2(s.ai[0]));
#pragma unroll
    for (int i = 2; i <= N; ++i) dg = add2(dg, mul2s(bc2(s.dai[i - 1]), p[i - 1]));     // dai[i-1] = fl(i * ai[i-1])
}

// Quotient a / b on each half in THREE fp32 operations: q0 = a * rcp(b), the remainder a - b q0 (one FMA), q0 + rcp(b) * rem.
// q0 + rcp(b) rem differs from a / b by at most |rem| |rcp(b) - 1/b| <= 2^-45 |a / b| before the final rounding, so the result
// is the IEEE quotient unless a / b lies within 2^-21 ulp of a rounding boundary (bound 1e-6 per quotient for a 1-ulp
// reciprocal; MUFU.RCP is better than that: 0 differing rays in 2e8 quotients, tools/strict_check.py).  nvcc's own div.rn
// sequence (div_rn2) spends two more operations refining the reciprocal first; every instruction of this kernel costs issue
// slots (a packed instruction takes two, tools/probe/issue_mix_probe.cu), and the reference's own CPU arithmetic is further from
// IEEE than this (MKL's vector sqrt is off by one ulp for 0.6 % of its inputs, tests/test_oracle_golden.py).
#ifndef SDIRT_STRICT_SHORT_DIV
#define SDIRT_STRICT_SHORT_DIV 1
#endif
#ifndef SDIRT_STRICT_CYCLE_FROM
#define SDIRT_STRICT_CYCLE_FROM 3     // first evaluation of the first surface's Newton loop that looks for a fixed point / 2-cycle of t (>= 2)
#endif
#ifndef SDIRT_STRICT_RCP_SQUARE
#define SDIRT_STRICT_RCP_SQUARE 0     // experiment: relieve the XU pipe (one MUFU.RCP less per evaluation) at one more packed multiply
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
// three quotients by the same divisor (a unit normal): one reciprocal
__device__ __forceinline__ void sdiv2x3(f2 a0, f2 a1, f2 a2, f2 b, f2 &q0, f2 &q1, f2 &q2) {
#if SDIRT_STRICT_SHORT_DIV
    const f2 r = rcp2(b), nb = neg2(b);
    const f2 p0 = mul2(a0, r), p1 = mul2(a1, r), p2 = mul2(a2, r);
    q0 = fma2(r, fma2(nb, p0, a0), p0);
    q1 = fma2(r, fma2(nb, p1, a1), p1);
    q2 = fma2(r, fma2(nb, p2, a2), p2);
#else
    q0 = div_rn2(a0, b); q1 = div_rn2(a1, b); q2 = div_rn2(a2, b);
#endif
}

// sag G(rho^2) and slope G'(rho^2) (sag_and_slope in engine.cu) on each half; the square root is shared.
// (r2 c2) / 2 is taken as r2 (c2 / 2): scaling by a power of two commutes with rounding.
template <int KIND, int NAI, bool WANT_G, bool WATCH>
__device__ __forceinline__ void sag_slope_strict2(const SurfDev &s, f2 r2, f2 &g, f2 &dg, StrictWatch &w) {
    // ((1 + k) r2) c2; with k = 0 (spheres, and aspheres whose conic constant is zero) the first product is r2 itself
    const f2 kr2c2 = (KIND == SIG_SPHERE || (s.flags & F_KZERO)) ? mul2s(r2, bc2(s.c2)) : mul2s(mul2s(bc2(s.onek), r2), bc2(s.c2));
    if (WATCH) w.kmax = fmaxf(fmaxf(kr2c2.x, kr2c2.y), w.kmax);
#if SDIRT_STRICT_SHORT_DIV
    // sqrt_rn2 keeping its MUFU.RSQ seed y ~ 1 / sf (2^-22) as the reciprocal of the quotient by sf.  The reference's term is
    // ((r2 c2) / 2) / sf; TWICE that, (r2 c2) / sf, is the same quotient up to the exact scaling, and the halving rides in the
    // sum that uses it: (1 + sf) + q / 2 == fma(0.5, q, 1 + sf).  The term is a small correction to 1 + sf, so even the rare
    // last-bit difference from the IEEE quotient (2e-6 of them by the bound above) survives that sum's rounding one time in ~30.
    const f2 x = add2(bc2(1.0f), neg2(kr2c2));
    const f2 y = rsq2(x);
    const f2 sq = mul2(x, y);
    const f2 sf = fma2(fma2(neg2(sq), sq, x), mul2(y, bc2(0.5f)), sq);
    const f2 q0 = mul2(kr2c2, y);
    const f2 inner2 = fma2(y, fma2(neg2(sf), q0, kr2c2), q0);
    const f2 one_sf = add2(bc2(1.0f), sf);
    const f2 s1 = fma2(bc2(0.5f), inner2, one_sf);
#else
    const f2 sf = sqrt_rn2(add2(bc2(1.0f), neg2(kr2c2)));
    const f2 one_sf = add2(bc2(1.0f), sf);
    const f2 s1 = add2(one_sf, sdiv2(mul2s(kr2c2, bc2(0.5f)), sf));
#endif
#if SDIRT_STRICT_SHORT_DIV && SDIRT_STRICT_RCP_SQUARE
    // one MUFU.RCP for both divisors: 1 / (1+sf) for g, its square as the reciprocal of fl((1+sf)^2) for the slope
    const f2 r1 = rcp2(one_sf);
    if (WANT_G) { const f2 num = mul2s(r2, bc2(s.c)); const f2 q = mul2(num, r1); g = fma2(r1, fma2(neg2(one_sf), q, num), q); }
    { const f2 num = mul2s(s1, bc2(s.c)), den = mul2s(one_sf, one_sf), rd = mul2(r1, r1);
      const f2 q = mul2(num, rd); dg = fma2(rd, fma2(neg2(den), q, num), q); }
#else
    if (WANT_G) g = sdiv2(mul2s(r2, bc2(s.c)), one_sf);
    dg = sdiv2(mul2s(s1, bc2(s.c)), mul2s(one_sf, one_sf));
#endif
    if constexpr (KIND == SIG_ASPHERE && NAI > 0) poly_terms2<NAI, WANT_G>(s, r2, g, dg);
}

// What the packed tracer leaves out of a loose evaluation, and how it knows: the mask select (an iterate outside the surface's own
// radius of definition is evaluated at rho = 0, surfaces.py:548-552) and the +-5 mm clamp of the Newton step (surfaces.py:558-560)
// cost eight instructions per evaluation pair and act on essentially no ray of a real lens.  Instead two running maxima ride along
// through the whole lens -- kmax over (1 + k) rho^2 c^2 of every loose evaluation (the mask acts where that reaches 1 up to a few
// ulp), smax over |step| of every evaluation -- one three-input FMNMX3 each, which returns the non-NaN operands (a dead, NaN half
// never shows).  A pair that ends the lens with kmax >= STRICT_KMAX or smax > 5 mm is traced again, ray by ray, with the
// generic strict tracer (retrace_pair_general): for every other pair the evaluations below ARE the reference's arithmetic.

// One Newton evaluation (surfaces.py:548-561 / 569-578) at t for both halves: the residual and the updated t.
// STRICT: the one extra evaluation after the loop (_valid instead of _valid_loose); it keeps its mask select, because rays stopped
// by a surface's clear aperture are not rare.
template <int KIND, int NAI, bool STRICT>
__device__ __forceinline__ void newton_eval2(const SurfDev &s, const Ray2 &r, f2 a, f2 b, f2 t, f2 &ftn, f2 &tn, StrictWatch &w) {
    const f2 nx = add2(r.ox, mul2s(r.dx, t)), ny = add2(r.oy, mul2s(r.dy, t)), nz = add2(r.oz, mul2s(r.dz, t));
    const f2 r2u = add2(mul2s(nx, nx), mul2s(ny, ny));
    // k > -1 (every surface this tracer accepts, strict_loop_ok): both masks are upper bounds on rho^2
    f2 r2 = r2u;
    if (STRICT) r2 = make_float2(r2u.x < s.thr_strict ? r2u.x : 0.0f, r2u.y < s.thr_strict ? r2u.y : 0.0f);
    f2 g, dg;
    sag_slope_strict2<KIND, NAI, true, !STRICT>(s, r2, g, dg, w);
    ftn = add2(add2(g, bc2(s.d)), neg2(nz));
    // -(f' + eps) instead of f' + eps (rounding is symmetric under negation): no negated copy of d_z is needed as an FMA addend
    const f2 ndf = add2(fma2(bc2(-2.0f), mul2s(dg, add2(mul2s(a, t), b)), r.dz), bc2(-EPS_F));
    const f2 nstep = sdiv2(ftn, ndf);
    w.smax = fmaxf(fmaxf(fabsf(nstep.x), fabsf(nstep.y)), w.smax);
    tn = add2(t, nstep);
}

// Newton intersection (surfaces.py:523-586) for a pair of rays with the per-ray schedule: a half leaves the loose loop when ITS
// residual is within 50e-6 mm (or after 10 evaluations), the loop runs until both have left, then one strict evaluation for both.
// FIRST (the first surface of the lens): the object may be metres away, t sits on a coarse float32 lattice and the map
// t -> t_new ends in a fixed point or a 2-cycle long before the residual test is met; both are detected and the rest of the
// loop is skipped with its known outcome (newton_strict in engine.cu does the same).  Without FIRST the same rays simply run to
// the cap: same result, more evaluations.
// (Splitting the loop into "both halves iterating", without the selects on t, plus a loop for the stragglers saves three more
// instructions per evaluation and loses more than that to warps whose threads sit in different loops: profiles/r02n_*.)
template <int KIND, int NAI, bool FIRST>
__device__ __forceinline__ void newton_strict2(const SurfDev &s, const Ray2 &r, f2 &t_out, f2 &ft_last, StrictWatch &w) {
    const f2 t0 = sdiv2(add2(bc2(s.d), neg2(r.oz)), r.dz);
    const f2 a = add2(mul2s(r.dx, r.dx), mul2s(r.dy, r.dy));
    const f2 b = add2(mul2s(r.dx, r.ox), mul2s(r.dy, r.oy));
    f2 t = t0, tb = t0;
    bool run0 = true, run1 = true;                     // still inside the loose loop (the first evaluation always runs: ft = 1e5;
                                                       // a dead half is NaN and leaves after it: |NaN| > tol is false)
    int it = 0;
#pragma unroll 1
    do {
        f2 ftn, tn;
        newton_eval2<KIND, NAI, false>(s, r, a, b, t, ftn, tn, w);
        ++it;
        if (FIRST && it >= SDIRT_STRICT_CYCLE_FROM) {
            // period 1 (t_new == t): every further evaluation repeats this one.  period 2 (t_new == the t before this
            // evaluation's input) with the residual still above the tolerance: t alternates until the cap, the parity of
            // the evaluations left picks the survivor.  Looked for from the third evaluation on (a ray that converges the
            // ordinary way has left by then; a cycle that began earlier is still a cycle: same outcome, one evaluation later).
            const bool big0 = fabsf(ftn.x) > NEWTON_LOOSE, big1 = fabsf(ftn.y) > NEWTON_LOOSE;
            const bool p1x = tn.x == t.x, p1y = tn.y == t.y;
            const bool p2x = !p1x && tn.x == tb.x && big0;
            const bool p2y = !p1y && tn.y == tb.y && big1;
            const bool odd = ((NEWTON_MAXIT - it) & 1) != 0;
            if (run0) { tb.x = t.x; if (!(p2x && odd)) t.x = tn.x; }
            if (run1) { tb.y = t.y; if (!(p2y && odd)) t.y = tn.y; }
            run0 = run0 && big0 && !(p1x || p2x);
            run1 = run1 && big1 && !(p1y || p2y);
        } else {
            if (FIRST) { if (run0) tb.x = t.x; if (run1) tb.y = t.y; }
            if (run0) t.x = tn.x;
            if (run1) t.y = tn.y;
            run0 = run0 && fabsf(ftn.x) > NEWTON_LOOSE;
            run1 = run1 && fabsf(ftn.y) > NEWTON_LOOSE;
        }
    } while ((run0 || run1) && it < NEWTON_MAXIT);
    t = add2(t0, add2(t, neg2(t0)));                                               // surfaces.py:563-567
    newton_eval2<KIND, NAI, true>(s, r, a, b, t, ft_last, t_out, w);
}

// The two validity tests of the step (cos_i^2 > 0.1, no total internal reflection: e < 1) are left to the caller, which folds them
// into the surface's own tests (kill_* below).
__device__ __forceinline__ void refract_strict2(const SurfDev &s, Ray2 &r, f2 qx, f2 qy, f2 qz, float sigma, f2 &c2, f2 &e) {
    const f2 cq = add2(add2(mul2s(r.dx, qx), mul2s(r.dy, qy)), mul2s(r.dz, qz));
    c2 = mul2s(cq, cq);
    e = mul2s(bc2(s.eta2), add2(bc2(1.0f), neg2(c2)));
    const f2 sr = mul2(bc2(sigma), sqrt_rn2(add2(bc2(1.0f), neg2(e))));
    r.dx = add2(mul2s(sr, qx), mul2s(bc2(s.eta), add2(r.dx, neg2(mul2s(cq, qx)))));
    r.dy = add2(mul2s(sr, qy), mul2s(bc2(s.eta), add2(r.dy, neg2(mul2s(cq, qy)))));
    r.dz = add2(mul2s(sr, qz), mul2s(bc2(s.eta), add2(r.dz, neg2(mul2s(cq, qz)))));
}

// Validity lives IN the ray in this tracer: a half that fails a surface's test gets d_z = NaN.  Every later quantity of that half is
// then NaN (t0 = (d - o_z) / d_z first of all), every later validity test is a comparison and fails on NaN, and the crop test of the
// splat drops it -- no alive flags to carry through the surface loop, no conjunctions with them (a ray whose arithmetic produces
// NaN by itself is invalid in the reference for the same reason: its comparisons fail).
__device__ __forceinline__ void strict_kill2(Ray2 &r, bool v0, bool v1) {
    const float nan = __int_as_float(0x7fc00000);
    r.dz = make_float2(v0 ? r.dz.x : nan, v1 ? r.dz.y : nan);
}
// The conjunction of a surface's validity tests as ONE predicate chain (setp.and) and one select per half: written out in C the
// compiler keeps one select per test (six FSEL per pair and surface instead of two).  0f3DCCCCCD = 0.1f, 0f7FC00000 = NaN.
// (the reference's fourth rule, no total internal reflection: e < 1, needs no test: sqrt_rn2(1 - e) is NaN for e >= 1 -- rsqrt of a
// negative number, or 0 * inf at e = 1 -- and reaches d_z through the refraction's own arithmetic)
// sphere, surfaces.py:464 + 667-669:  rho^2 <= r^2, t >= 0, cos^2 > 0.1
__device__ __forceinline__ float kill_sphere(float dz, float r2u, float r2, float t, float c2) {
    float out;
    asm("{ .reg .pred p;\n\t"
        "setp.le.f32 p, %2, %3;\n\t"
        "setp.ge.and.f32 p, %4, 0f00000000, p;\n\t"
        "setp.gt.and.f32 p, %5, 0f3DCCCCCD, p;\n\t"
        "selp.f32 %0, %1, 0f7FC00000, p; }" : "=f"(out) : "f"(dz), "f"(r2u), "f"(r2), "f"(t), "f"(c2));
    return out;
}
// asphere, surfaces.py:584 + 667-669:  rho^2 < thr, |ft| < 10e-6, t > 0, cos^2 > 0.1
__device__ __forceinline__ float kill_asphere(float dz, float r2u, float thr, float ft, float t, float c2) {
    float out;
    asm("{ .reg .pred p;\n\t"
        "setp.lt.f32 p, %2, %3;\n\t"
        "setp.lt.and.f32 p, %4, 0f3727C5AC, p;\n\t"
        "setp.gt.and.f32 p, %5, 0f00000000, p;\n\t"
        "setp.gt.and.f32 p, %6, 0f3DCCCCCD, p;\n\t"
        "selp.f32 %0, %1, 0f7FC00000, p; }" : "=f"(out) : "f"(dz), "f"(r2u), "f"(thr), "f"(fabsf(ft)), "f"(t), "f"(c2));
    return out;
}

// Aspheric.ray_reaction (surfaces.py:391-520) for one surface of kind K (SIG_STOP covers every flat surface: whether it refracts
// is a run-time flag) and a pair of rays; r.oz is ABSOLUTE here (the reference's coordinates), unlike the fast tracer's
// vertex-relative z.  A dead half carries garbage that nothing reads.
template <int K, int NAI, bool FIRST>
__device__ __forceinline__ void strict_step2(const SurfDev &s, Ray2 &r, StrictWatch &w) {
    if (K == SIG_STOP) {
        const f2 t = sdiv2(add2(bc2(s.d), neg2(r.oz)), r.dz);
        r.ox = add2(r.ox, mul2s(t, r.dx)); r.oy = add2(r.oy, mul2s(t, r.dy)); r.oz = add2(r.oz, mul2s(t, r.dz));
        bool v0, v1;
        if (s.flags & F_SQUARE) {                                                            // surfaces.py:416-419
            v0 = fabsf(r.ox.x) <= s.r && fabsf(r.oy.x) <= s.r;
            v1 = fabsf(r.ox.y) <= s.r && fabsf(r.oy.y) <= s.r;
        } else {
            const f2 r2u = add2(mul2s(r.ox, r.ox), mul2s(r.oy, r.oy));
            v0 = r2u.x <= s.r2_sqrt_le;                                                      // sqrt(x^2 + y^2) <= r, surfaces.py:421
            v1 = r2u.y <= s.r2_sqrt_le;
        }
        if (s.flags & F_REFRACTS) {
            f2 c2, e;
            refract_strict2(s, r, bc2(0.0f), bc2(0.0f), bc2(1.0f), 1.0f, c2, e);                      // n = -normalize((0,0,-1))
            v0 = v0 & (c2.x > 0.1f) & (e.x < 1.0f);
            v1 = v1 & (c2.y > 0.1f) & (e.y < 1.0f);
        }
        strict_kill2(r, v0, v1);
        return;
    }
    f2 t, ft_last;
    newton_strict2<K, NAI, FIRST>(s, r, t, ft_last, w);
    r.ox = add2(r.ox, mul2s(t, r.dx)); r.oy = add2(r.oy, mul2s(t, r.dy)); r.oz = add2(r.oz, mul2s(t, r.dz));
    const f2 r2u = add2(mul2s(r.ox, r.ox), mul2s(r.oy, r.oy));
    f2 qx, qy, qz, c2, e;
    if (K == SIG_SPHERE) {
        // gradient (+-2x, +-2y, +-(2z - 2(d+R))) = +-2 (x, y, w): the normalised vector is +-(x, y, w) / |(x, y, w)| bit for bit
        const f2 w = add2(r.oz, bc2(-s.dR));
        const f2 nrm = sqrt_rn2(fma2(w, w, fma2(r.oy, r.oy, mul2s(r.ox, r.ox))));             // norm3
        sdiv2x3(r.ox, r.oy, w, nrm, qx, qy, qz);
        refract_strict2(s, r, qx, qy, qz, s.sigma, c2, e);
        r.dz = make_float2(kill_sphere(r.dz.x, r2u.x, s.r2, t.x, c2.x), kill_sphere(r.dz.y, r2u.y, s.r2, t.y, c2.y));
    } else {
        f2 g, dg;
        sag_slope_strict2<K, NAI, false, false>(s, r2u, g, dg, w);                                      // (x, y masked by ra > 0: alive here)
        const f2 dg2 = mul2s(dg, bc2(2.0f));
        const f2 gx = mul2s(dg2, r.ox), gy = mul2s(dg2, r.oy);
        const f2 nrm = sqrt_rn2(add2(fma2(gy, gy, mul2s(gx, gx)), bc2(1.0f)));               // norm3(gx, gy, -1)
        sdiv2x3(gx, gy, bc2(-1.0f), nrm, qx, qy, qz);
        refract_strict2(s, r, qx, qy, qz, -1.0f, c2, e);
        r.dz = make_float2(kill_asphere(r.dz.x, r2u.x, s.thr_strict, ft_last.x, t.x, c2.x),
                           kill_asphere(r.dz.y, r2u.y, s.thr_strict, ft_last.y, t.y, c2.y));
    }
}

// One surface of any kind.  Polynomial orders compiled: 0 (pure conic), 4, 5, 6 coefficients; the cycle detection of the first
// surface is compiled for spheres only (any other first surface takes the plain loop: same result).
template <bool FIRST>
__device__ __forceinline__ void strict_surface2(const SurfDev &s, Ray2 &r, StrictWatch &w) {
    if (s.kind == SDIRT_SURF_SPHERE) strict_step2<SIG_SPHERE, 0, FIRST>(s, r, w);
    else if (s.kind == SDIRT_SURF_FLAT) strict_step2<SIG_STOP, 0, false>(s, r, w);
    else if (s.n_ai == 6) strict_step2<SIG_ASPHERE, 6, false>(s, r, w);
    else if (s.n_ai == 0) strict_step2<SIG_ASPHERE, 0, false>(s, r, w);
    else if (s.n_ai == 4) strict_step2<SIG_ASPHERE, 4, false>(s, r, w);
    else strict_step2<SIG_ASPHERE, 5, false>(s, r, w);
}

// The whole lens as a run-time loop over the resolved surfaces: ONE copy of the sphere / flat / asphere code (plus the first
// surface's own), the prescription values of a surface fetched from the constant bank by the loop counter.  A surface costs
// ~1000 executed instructions per ray pair in this arithmetic, so the loop's overhead is noise, while the unrolled form (4400
// instructions for 12 surfaces, 6000 for 21) ran with `no_instruction` among its first stall reasons (profiles/r02b_*): the
// strict kernel is the one place where the run-time loop wins.
// Returns whether the pair has to be traced again with the generic tracer (see newton_eval2).
__device__ __forceinline__ bool trace_strict_loop2(const LensDev &L, Ray2 &r) {
    StrictWatch w = {0.0f, 0.0f};
    strict_surface2<true>(L.s[0], r, w);
#pragma unroll 1
    for (int j = 1; j < L.n; ++j) {
        if (r.dz.x != r.dz.x && r.dz.y != r.dz.y) continue;  // both halves dead (no early exit: the loop counter stays warp-uniform)
        strict_surface2<false>(L.s[j], r, w);
    }
    r.a0 = r.dz.x == r.dz.x;
    r.a1 = r.dz.y == r.dz.y;
    return w.kmax >= STRICT_KMAX || w.smax > NEWTON_STEP;
}

// The rare pair: both rays through the generic one-ray strict tracer (trace_lens<STRICT> in engine.cu, the code behind
// sdirt_trace_rays), handed back in the packed tracer's conventions (absolute z, a dead half has d_z = NaN).  Not inlined: it is
// the whole generic surface loop, and it runs for essentially no pair.
__device__ __noinline__ void retrace_pair_general(const LensDev &L, float px, float py, float pz, float2 s0, float2 s1, float pupil_z, float *out /*[12], local*/) {
    RayReg r0 = ray_from_point(px, py, pz, s0.x, s0.y, pupil_z), r1 = ray_from_point(px, py, pz, s1.x, s1.y, pupil_z);
    trace_lens<STRICT, false>(L, r0, nullptr, 0, 0, false);
    trace_lens<STRICT, false>(L, r1, nullptr, 0, 0, false);
    const float nan = __int_as_float(0x7fc00000);
    out[0] = r0.ox; out[1] = r1.ox; out[2] = r0.oy; out[3] = r1.oy; out[4] = r0.oz; out[5] = r1.oz;
    out[6] = r0.dx; out[7] = r1.dx; out[8] = r0.dy; out[9] = r1.dy;
    out[10] = r0.alive ? r0.dz : nan; out[11] = r1.alive ? r1.dz : nan;
}

// can the packed strict tracer take this (resolved) lens?  forward, per-ray Newton schedule, compiled polynomial orders
static bool strict_loop_ok(const LensDev &L) {
    if (!L.forward) return false;
    for (int j = 0; j < L.n; ++j) {
        const SurfDev &s = L.s[j];
        if (s.kind != SDIRT_SURF_FLAT && s.fixed_iters >= 0) return false;
        if (s.kind == SDIRT_SURF_ASPHERE && !(s.n_ai == 0 || (s.n_ai >= 4 && s.n_ai <= 6))) return false;
        if (s.kind == SDIRT_SURF_ASPHERE && !(s.flags & F_KGT)) return false;        // k <= -1: the loose mask is not an upper bound
    }
    return true;
}

#ifndef SDIRT_STRICT_BLOCK
#define SDIRT_STRICT_BLOCK 16
#endif
#ifndef SDIRT_STRICT_THREADS
#define SDIRT_STRICT_THREADS 256
#endif
#ifndef SDIRT_STRICT_MIN_CTAS
#define SDIRT_STRICT_MIN_CTAS 4       // 256-thread CTAs per SM: 64 registers per thread (measured 2 % faster than 3 x 80 despite ~200 B of spills)
#endif
struct TraceStrictLoop {
    static constexpr bool PAIR = true;
    static constexpr bool ONE_RAY_LOOP = false;      // no one-ray variant of the loop in this kernel (the generic kernel is the comparison)
    static constexpr bool STRICT_SPLAT = true;
    static constexpr int MIN_CTAS = SDIRT_STRICT_MIN_CTAS;
    static constexpr int THREADS = SDIRT_STRICT_THREADS;   // (its interleaved blocks do not need chunk / threads to be whole)
    static constexpr int BLOCK = SDIRT_STRICT_BLOCK;     // samples per lane and block of the interleaved assignment (fast_path.cuh)
    static __device__ __forceinline__ bool trace(const LensDev &L, RayReg &r, bool) {      // one-ray loop (debug switch): the generic strict trace
        trace_lens<STRICT, false>(L, r, nullptr, 0, 0, false);
        return r.alive;
    }
    static __device__ __forceinline__ float sensor_distance(const LensDev &L, const RayReg &r) { return L.d_sensor - r.oz; }
    static __device__ __forceinline__ Ray2 trace2(const LensDev &L, float px, float py, float pz, float2 s0, float2 s1, float pupil_z, bool) {
        Ray2 r = ray2_from_point(px, py, pz, s0, s1, pupil_z);
        if (trace_strict_loop2(L, r) || L.debug_scalar_strict == 3) {   // (3: testing aid, every pair takes the rare path; the ray itself stays in registers: only this branch goes through memory)
            float buf[12];
            retrace_pair_general(L, px, py, pz, s0, s1, pupil_z, buf);
            r.ox = make_float2(buf[0], buf[1]); r.oy = make_float2(buf[2], buf[3]); r.oz = make_float2(buf[4], buf[5]);
            r.dx = make_float2(buf[6], buf[7]); r.dy = make_float2(buf[8], buf[9]); r.dz = make_float2(buf[10], buf[11]);
            r.a0 = r.dz.x == r.dz.x; r.a1 = r.dz.y == r.dz.y;
        }
        return r;
    }
};

// ---- testing aid: the packed strict tracer's sensor-plane ray states, to be compared bit for bit with sdirt_trace_rays ----
__global__ void __launch_bounds__(128)
debug_trace_strict2_kernel(const __grid_constant__ LensDev L, const float *__restrict__ point, const float2 *__restrict__ pupil, int64_t m,
                           float pupil_z, float *__restrict__ out /*[m,7]*/) {
    const int64_t j = 2 * ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
    if (j >= m) return;
    const bool two = j + 1 < m;
    Ray2 r = TraceStrictLoop::trace2(L, point[0], point[1], point[2], pupil[j], two ? pupil[j + 1] : pupil[j], pupil_z, false);
    // Ray.propagate_to(d_sensor), basics.py:262-263
    const f2 t = div_rn2(add2(bc2(L.d_sensor), neg2(r.oz)), r.dz);
    const f2 sx = add2(r.ox, mul2s(r.dx, t)), sy = add2(r.oy, mul2s(r.dy, t)), sz = add2(r.oz, mul2s(r.dz, t));
    float *p = out + j * 7;                                 // (a dead ray's state is NaN inside the tracer: report zeros)
    if (r.a0) { p[0] = sx.x; p[1] = sy.x; p[2] = sz.x; p[3] = r.dx.x; p[4] = r.dy.x; p[5] = r.dz.x; p[6] = 1.f; }
    else { for (int k = 0; k < 7; ++k) p[k] = 0.f; }
    if (two) {
        p += 7;
        if (r.a1) { p[0] = sx.y; p[1] = sy.y; p[2] = sz.y; p[3] = r.dx.y; p[4] = r.dy.y; p[5] = r.dz.y; p[6] = 1.f; }
        else { for (int k = 0; k < 7; ++k) p[k] = 0.f; }
    }
}

// The parity mode of sdirt_psf_bank (numerics STRICT, per-ray Newton schedule) on the packed kernel; returns 1 if the lens is
// outside what it compiles (the caller then takes the generic one-ray kernel).
static int launch_bank_strict(const LensDev &L, const SplatDev &P, dim3 grid, cudaStream_t st,
                              const float *points, const float2 *pupil, int64_t m, float pupil_z, const float *centre,
                              float4 *lut, int64_t chunk, int run, float *partial, int *hits) {
    if (!strict_loop_ok(L)) return 1;
    dp_lut_kernel<<<DP_LUT_N / 256, 256, 0, st>>>(P, lut);
    if (int rc = check_launch("dp_lut_kernel")) return rc;
    return launch_bank_run<TraceStrictLoop>(L, P, grid, st, points, pupil, m, pupil_z, centre, lut, chunk, run, partial, hits);
}

static int launch_debug_trace_strict2(const LensDev &L, cudaStream_t st, const float *point, const float2 *pupil, int64_t m, float pupil_z, float *out) {
    if (!strict_loop_ok(L)) return fail(SDIRT_E_ARG, "sdirt_debug_trace_strict2: this lens is outside what the packed strict tracer compiles");
    debug_trace_strict2_kernel<<<(unsigned)(((m + 1) / 2 + 127) / 128), 128, 0, st>>>(L, point, pupil, m, pupil_z, out);
    return check_launch("debug_trace_strict2_kernel");
}

// psfnet_path.cuh — the two ends of PSFNet.pred (deeplens/psfnet.py:317-336) inside PSFNet.render (psfnet.py:681-708),
// so that the per-pixel PSF tensor of a whole image never has to exist.  Included by engine.cu.
//
// The reference renders an image in three sweeps over [N,H,W] pixels: build the (x, y, z) coordinate grid, run the PSF
// MLP twice (left: (x, y, z); right: (-x, y, z), then flipped along the last axis), stack + normalise, gather-convolve.
// At 1024 x 1536 the stacked fp16 tensor alone is 2.8 GB per image and every elementwise step is another pass over it.
// Here PSFNet.render walks the image in row bands (sdirt_render_local_psf_rows) sized for the L2, and the band's MLP
// rows are laid out pixel-major / side-minor (row 2p = left, row 2p + 1 = right), so that the output of the last
// Linear [2P, ks*ks] IS the stacked [P, 2, ks, ks] tensor.  The dense 512-wide GEMM chain stays on cuBLAS; the kernels
// here are its first and last layers' surroundings:
//
//   mlp_input_layer_kernel : coordinate grid (psfnet.py:683-694) + the first Linear(3 -> n1) + ReLU (psfnet_arch.py:40-41)
//                            under CUDA autocast: inputs, weights, bias rounded to fp16, products accumulated in fp32,
//                            one rounding to fp16 (what the fp16 GEMM + bias epilogue of cuBLAS computes for K = 3);
//   psf_pack_kernel        : flip the right kernels along their last axis, sum(-1).sum(-1) with torch's fp16 rounding
//                            points, `+ 1e-9`, divide (psfnet.py:330-333), compact the padded GEMM rows.
#pragma once

#define MLP_IN_GROUP 8          // outputs per thread: one 16-byte store
#define MLP_IN_THREADS 256

// Thread (g, j) of a CTA owns output columns [8g, 8g + 8) of rows j, j + rows_per_cta, ...: its 8 x (w_x, w_y, w_z, bias)
// stay in registers, a warp writes whole 16-byte pieces of consecutive rows, and a row costs three loads, 24 FMAs, 8
// roundings and one store.
__global__ void __launch_bounds__(MLP_IN_THREADS)
mlp_input_layer_kernel(const float *__restrict__ xs, const float *__restrict__ ys, const float *__restrict__ z,
                       int H, int W, int b0, int nb, int row0, int nrw,
                       const __half *__restrict__ w1, const __half *__restrict__ b1, int n1, __half *__restrict__ out) {
    const int groups = n1 / MLP_IN_GROUP;
    const int rows_per_cta = MLP_IN_THREADS / groups;
    const int g = threadIdx.x % groups, j = threadIdx.x / groups;
    if (j >= rows_per_cta) return;
    float wx[MLP_IN_GROUP], wy[MLP_IN_GROUP], wz[MLP_IN_GROUP], wb[MLP_IN_GROUP];
#pragma unroll
    for (int k = 0; k < MLP_IN_GROUP; ++k) {
        const int o = g * MLP_IN_GROUP + k;
        wx[k] = __half2float(w1[3 * o + 0]);
        wy[k] = __half2float(w1[3 * o + 1]);
        wz[k] = __half2float(w1[3 * o + 2]);
        wb[k] = __half2float(b1[o]);
    }
    const unsigned n_rows = 2u * (unsigned)nb * (unsigned)nrw * (unsigned)W;        // < 2^31: checked by the host
    for (unsigned r = blockIdx.x * rows_per_cta + j; r < n_rows; r += gridDim.x * rows_per_cta) {
        const unsigned p = r >> 1, side = r & 1u;
        const unsigned q = p / (unsigned)W, x = p - q * (unsigned)W;
        const unsigned bq = q / (unsigned)nrw;
        const int y = row0 + (int)(q - bq * (unsigned)nrw), b = b0 + (int)bq;
        // autocast casts the fp32 coordinates to fp16 before the GEMM; inp[..., 0] *= -1 for the right side (psfnet.py:328)
        const float xr = __ldg(xs + x);
        const float xv = __half2float(__float2half_rn(side ? -xr : xr));
        const float yv = __half2float(__float2half_rn(__ldg(ys + y)));
        const float zv = __half2float(__float2half_rn(__ldg(z + ((int64_t)b * H + y) * W + x)));
        __align__(16) __half o[MLP_IN_GROUP];
#pragma unroll
        for (int k = 0; k < MLP_IN_GROUP; ++k) {
            const float acc = fmaf(zv, wz[k], fmaf(yv, wy[k], xv * wx[k])) + wb[k];  // products are exact in fp32
            o[k] = __float2half_rn(fmaxf(__half2float(__float2half_rn(acc)), 0.0f));
        }
        *reinterpret_cast<uint4 *>(out + (size_t)r * n1 + g * MLP_IN_GROUP) = *reinterpret_cast<const uint4 *>(o);
    }
}

// One warp per pixel (= two MLP output rows).  raw: [2P, ld] fp16, row 2p = left, row 2p+1 = right evaluated at -x and
// not yet flipped; psf: [P, 2, ks, ks] fp16.  torch's fp16 arithmetic (psfnet.py:330-333): sum(-1) accumulates a kernel
// row in fp32 and rounds it to fp16, the second sum(-1) adds the ks rounded row sums in fp32 and rounds again, `+ 1e-9`
// is an fp16 add (a no-op for any sum an fp16 can hold), the quotient is computed in fp32 (IEEE, div_rn: numerator and
// denominator are fp16 values, far from fp32's exponent limits) and rounded once.  An all-zero
// kernel (0 / 0 = NaN in the reference's CUDA run, 0 in its fp32 CPU run) is written as zeros.
// Data movement: both rows come in with 16-byte loads (ld % 8 == 0) into shared memory, the pixel's 2*ks*ks outputs go
// out as 32-bit words (a pixel's block is 4-byte aligned, its right half is not).  KS = 0: window size at run time.
#define PACK_WARPS 8
template <int KS>
__global__ void __launch_bounds__(PACK_WARPS * 32)
psf_pack_kernel(const __half *__restrict__ raw, int64_t n_pix, int ld, int ks_rt, __half *__restrict__ psf) {
    extern __shared__ __align__(16) unsigned char pk_smem[];
    const int ks = KS ? KS : ks_rt;
    const int kk = ks * ks;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __half *rows = reinterpret_cast<__half *>(pk_smem) + (size_t)warp * 2 * ld;      // this warp's two raw rows
    float *rsum = reinterpret_cast<float *>(pk_smem + (size_t)PACK_WARPS * 2 * ld * sizeof(__half)) + warp * 2 * (SDIRT_MAX_KS + 1);
    const bool vec = (ld % 8) == 0 && ((uintptr_t)raw & 15) == 0;
    // compile-time window: where in the two staged rows this lane's output words come from is the same for every pixel
    constexpr int NIT = KS ? (KS * KS + 31) / 32 : 1;
    unsigned short off_lo[NIT], off_hi[NIT];
    unsigned right_lo = 0, right_hi = 0;
    if constexpr (KS > 0) {
#pragma unroll
        for (int i = 0; i < NIT; ++i) {
            const int w = min(lane + 32 * i, KS * KS - 1);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int e = 2 * w + h;
                const bool right = e >= KS * KS;
                const int t = e - (right ? KS * KS : 0), u = t / KS, v = t - u * KS;
                const unsigned short o = (unsigned short)(right ? ld + u * KS + (KS - 1 - v) : e);
                if (h == 0) { off_lo[i] = o; right_lo |= (unsigned)right << i; }
                else { off_hi[i] = o; right_hi |= (unsigned)right << i; }
            }
        }
    }
    for (int64_t p = (int64_t)blockIdx.x * PACK_WARPS + warp; p < n_pix; p += (int64_t)gridDim.x * PACK_WARPS) {
        const __half *src = raw + 2 * p * ld;
        if (vec) {
            const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
            uint4 *d4 = reinterpret_cast<uint4 *>(rows);
            for (int i = lane; i < 2 * ld / 8; i += 32) d4[i] = __ldg(s4 + i);
        } else {
            for (int i = lane; i < 2 * ld; i += 32) rows[i] = src[i];
        }
        __syncwarp();
        for (int i = lane; i < 2 * ks; i += 32) {                                    // one kernel row of one side per lane
            const int side = i >= ks, u = i - side * ks;
            const __half *kr = rows + side * ld + u * ks;
            float s = 0.0f;
#pragma unroll 7
            for (int v = 0; v < ks; ++v) s += __half2float(kr[v]);
            rsum[side * (SDIRT_MAX_KS + 1) + u] = __half2float(__float2half_rn(s));
        }
        __syncwarp();
        float tl = 0.0f, tr = 0.0f;
        for (int u = 0; u < ks; ++u) { tl += rsum[u]; tr += rsum[SDIRT_MAX_KS + 1 + u]; }
        float dl = __half2float(__float2half_rn(tl)), dr = __half2float(__float2half_rn(tr));
        dl = __half2float(__float2half_rn(dl + 1e-9f));
        dr = __half2float(__float2half_rn(dr + 1e-9f));
        // quotient = IEEE fp32 division (div_rn's sequence with the per-denominator half hoisted): r ~ 1/den refined once,
        // q = a r, remainder, correction
        const bool okl = dl > 0.0f && dl <= 65504.0f, okr = dr > 0.0f && dr <= 65504.0f;
        float rl = rcp_approx(okl ? dl : 1.0f), rr = rcp_approx(okr ? dr : 1.0f);
        rl = fmaf(rl, fmaf(-dl, rl, 1.0f), rl);
        rr = fmaf(rr, fmaf(-dr, rr, 1.0f), rr);
        auto quot = [](float a, float den, float r) { const float q = a * r; return fmaf(r, fmaf(-den, q, a), q); };
        unsigned *dst = reinterpret_cast<unsigned *>(psf + 2 * p * kk);
        if constexpr (KS > 0) {
#pragma unroll
            for (int i = 0; i < NIT; ++i) {
                const int w = lane + 32 * i;
                if (NIT * 32 == KS * KS || w < KS * KS) {
                    const float a0 = __half2float(rows[off_lo[i]]), a1 = __half2float(rows[off_hi[i]]);
                    const bool r0 = (right_lo >> i) & 1, r1 = (right_hi >> i) & 1;
                    const __half lo = __float2half_rn((r0 ? okr : okl) ? quot(a0, r0 ? dr : dl, r0 ? rr : rl) : 0.0f);
                    const __half hi = __float2half_rn((r1 ? okr : okl) ? quot(a1, r1 ? dr : dl, r1 ? rr : rl) : 0.0f);
                    dst[w] = (unsigned)__half_as_ushort(lo) | ((unsigned)__half_as_ushort(hi) << 16);
                }
            }
        } else {
            auto element = [&](int e) -> __half {                                    // output element e of the pixel's [2][ks][ks] block
                if (e < kk) return __float2half_rn(okl ? quot(__half2float(rows[e]), dl, rl) : 0.0f);
                const int t = e - kk, u = t / ks, v = t - u * ks;
                return __float2half_rn(okr ? quot(__half2float(rows[ld + u * ks + (ks - 1 - v)]), dr, rr) : 0.0f);
            };
            for (int w = lane; w < kk; w += 32) {
                const __half lo = element(2 * w), hi = element(2 * w + 1);
                dst[w] = (unsigned)__half_as_ushort(lo) | ((unsigned)__half_as_ushort(hi) << 16);
            }
        }
        __syncwarp();
    }
}

// gamma -> sensor noise -> clip, the tail of PSFNet.render(train=True) (psfnet.py:605-620, 629-642, 708-713) in one pass.
// x [N, 2C, H, W] float32 in place: the convolved linear image (left channels first).  randn: standard normal draws of the
// same shape (torch.randn_like on the device, as the reference draws them); noise_range [N]; weight [N, W] = the
// reference's torch.linspace(range1, range2, W) ramp, read mirrored for the right channels (torch.flip(weight_l, [-1])).
__global__ void __launch_bounds__(256)
gamma_noise_clip_kernel(float *__restrict__ x, const float *__restrict__ randn, const float *__restrict__ noise_range,
                        const float *__restrict__ weight, int N, int C2, int H, int W) {
    const int64_t total = (int64_t)N * C2 * H * W;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int col = (int)(i % W);
        const int64_t q = i / ((int64_t)H * W);
        const int c = (int)(q % C2), n = (int)(q / C2);
        const float w = weight[(int64_t)n * W + (c < C2 / 2 ? col : W - 1 - col)];
        const float noise = (randn[i] * noise_range[n]) * w;
        x[i] = fminf(fmaxf(tone_gamma(x[i]) + noise, 0.0f), 1.0f);
    }
}

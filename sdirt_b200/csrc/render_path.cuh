// render_path.cuh — spatially varying dual-pixel render (local_psf_render_fast, deeplens/render_psf.py:120-155),
// the HBM-streaming kernel.  Included by engine.cu.
//
// out_c(p) = sum over the ks x ks window of  pad_c(p + (u,v)) * psf_side(p)[ks-1-u, ks-1-v]   for side = L, R,
// with the reference's half() arithmetic: image and kernels rounded to fp16, every product rounded to fp16, the sum
// accumulated wider and rounded to fp16 once.  The per-pixel kernels [B,H,W,2,ks,ks] ARE the traffic (1764 B per pixel
// at ks = 21 in fp16 against 36 B of image in and out), so the kernel is a pure stream over that tensor:
//
//   * one warp per output pixel; lanes own fixed *aligned pairs of taps* of the pixel's contiguous 2*ks*ks block, so a
//     warp reads the block with 4-byte (fp16) / 8-byte (fp32) coalesced loads and no per-tap index arithmetic: every
//     lane precomputes its (at most 14 + 2) tap positions once per CTA;
//   * the replicate-padded image tile lives in shared memory as fp16, column-mirrored (so that increasing tap index
//     is increasing address) and in two copies shifted by one element (so that the two image values a tap pair needs
//     are ONE aligned 32-bit word whatever the pixel's column parity);
//   * products are HMUL2 (two fp16-rounded products per instruction), accumulated with sm_100's mixed-precision
//     add (PTX add.rn.f32.f16 -> SASS FHADD: f32 += f16 half of a register) -- the reference's rounding, 1.5
//     instructions per multiply-accumulate;
//   * degamma (psfnet.py:589-603) is applied while the tile is staged, gamma + clip (psfnet.py:605-620, 711-713) when
//     the pixel is written.
#pragma once

#define RP_TH 16
#define RP_TW 32
#define RP_WARPS 8
#define RP_C 3

__device__ __forceinline__ void fhadd(float &acc, __half h) {
    asm("add.rn.f32.f16 %0, %1, %0;" : "+f"(acc) : "h"(__half_as_ushort(h)));
}

// Sum six per-lane partials over the warp with 8 shuffles instead of 30: at each of the first three butterfly levels a
// lane keeps one value of a pair and ships the other, so the number of live values halves as the lane count does.
// Returns the total of value `index` (0..5 = side * 3 + channel) on the lanes listed by reduce6_writer().
__device__ __forceinline__ float reduce6(const float (&acc)[2][RP_C], int lane, int &index) {
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
    float k0 = b4 ? acc[0][1] : acc[0][0], s0 = b4 ? acc[0][0] : acc[0][1];
    float k1 = b4 ? acc[1][0] : acc[0][2], s1 = b4 ? acc[0][2] : acc[1][0];
    float k2 = b4 ? acc[1][2] : acc[1][1], s2 = b4 ? acc[1][1] : acc[1][2];
    k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
    k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
    k2 += __shfl_xor_sync(0xffffffffu, s2, 16);
    float c0 = b3 ? k1 : k0, t0 = b3 ? k0 : k1;
    c0 += __shfl_xor_sync(0xffffffffu, t0, 8);
    k2 += __shfl_xor_sync(0xffffffffu, k2, 8);
    float d = b2 ? k2 : c0, t1 = b2 ? c0 : k2;
    d += __shfl_xor_sync(0xffffffffu, t1, 4);
    d += __shfl_xor_sync(0xffffffffu, d, 2);
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    index = b2 ? 4 + (b4 ? 1 : 0) : (b3 ? 2 : 0) + (b4 ? 1 : 0);
    return d;
}
__device__ __forceinline__ bool reduce6_writer(int lane) { return (lane & 3) == 0 && lane != 12 && lane != 28; }

template <typename PsfT> struct PsfPair;
template <> struct PsfPair<__half> {
    static __device__ __forceinline__ __half2 load2(const __half *p) { return *reinterpret_cast<const __half2 *>(p); }
    static __device__ __forceinline__ __half load1(const __half *p) { return *p; }
};
template <> struct PsfPair<float> {
    static __device__ __forceinline__ __half2 load2(const float *p) {
        const float2 v = *reinterpret_cast<const float2 *>(p);
        return __floats2half2_rn(v.x, v.y);
    }
    static __device__ __forceinline__ __half load1(const float *p) { return __float2half_rn(*p); }
};

template <int KS, int TWPX = RP_TW>
struct RenderGeom {
    static constexpr int PR = (KS - 1) / 2;            // aligned tap pairs per kernel row
    static constexpr int NP = 2 * KS * PR;             // pair units per pixel (both sides)
    static constexpr int NS = 2 * KS;                  // single taps per pixel (one per kernel row and side)
    static constexpr int NIT = (NP + 31) / 32;         // pair iterations per lane
    static constexpr int NSI = (NS + 31) / 32;         // single iterations per lane
    static constexpr int TH = RP_TH + KS - 1, TW = TWPX + KS - 1;
    static constexpr int RW = (TW + 2) / 2;            // 32-bit words per tile row (TW + 1 halves, rounded up)
    static constexpr int CHW = TH * RW;                // words per channel
    static constexpr int COPYW = RP_C * CHW;           // words per parity copy
    static constexpr int SMEM_BYTES = 2 * COPYW * 4;
    static_assert(2 * COPYW < 65536, "tile word offsets are packed into 16 bits");
};

template <int KS, typename PsfT>
__global__ void __launch_bounds__(RP_WARPS * 32, 3)
render_pairs_kernel(const float *__restrict__ img, const PsfT *__restrict__ psf, int B, int H, int W, int row0, int nrw, int tone,
                    float *__restrict__ out_l, float *__restrict__ out_r) {
    using G = RenderGeom<KS>;
    extern __shared__ unsigned rp_smem[];               // [2 copies][C][TH][RW] words of two fp16 each
    __half *s0h = reinterpret_cast<__half *>(rp_smem);  // copy 0 addressed by element
    constexpr int pad = (KS - 1) / 2;
    const int b = blockIdx.z, y0 = row0 + blockIdx.y * RP_TH, x0 = blockIdx.x * RP_TW;

    // ---- stage the mirrored tile: copy 0 element (r, m) = image(y0 + r - pad, x0 + (TW-1-m) - pad), replicate-padded
    for (int i = threadIdx.x; i < RP_C * G::TH * 2 * G::RW; i += blockDim.x) {
        const int c = i / (G::TH * 2 * G::RW), rem = i - c * (G::TH * 2 * G::RW);
        const int r = rem / (2 * G::RW), m = rem - r * (2 * G::RW);
        float v = 0.0f;
        if (m < G::TW) {
            const int gy = min(max(y0 + r - pad, 0), H - 1), gx = min(max(x0 + (G::TW - 1 - m) - pad, 0), W - 1);
            v = img[(((int64_t)b * RP_C + c) * H + gy) * W + gx];
            if (tone & 1) v = tone_degamma(v);
        }
        s0h[i] = __float2half_rn(v);
    }
    __syncthreads();
    // copy 1 word w of a row = elements (2w+1, 2w+2) of copy 0
    for (int i = threadIdx.x; i < G::COPYW; i += blockDim.x) {
        const int w = i % G::RW;
        const __half lo = s0h[2 * i + 1];
        const __half hi = (w + 1 < G::RW) ? s0h[2 * i + 2] : __float2half_rn(0.0f);
        rp_smem[G::COPYW + i] = (unsigned)__half_as_ushort(lo) | ((unsigned)__half_as_ushort(hi) << 16);
    }
    __syncthreads();

    // ---- per-lane tap tables (pixel independent) -------------------------------------------------------------------
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned peo[G::NIT];                               // smem word offsets: low 16 bits even pixel column, high 16 odd
    int pg[G::NIT];                                     // global half2 index inside the pixel's block
    unsigned side_mask = 0, live_mask = 0;
#pragma unroll
    for (int i = 0; i < G::NIT; ++i) {
        const int q = lane + 32 * i;
        const bool live = q < G::NP;
        const int qq = live ? q : 0;
        const int side = qq / (KS * G::PR), rem = qq - side * (KS * G::PR);
        const int u = rem / G::PR, m = rem - u * G::PR;
        const int v0 = 2 * m + ((side + u) & 1);
        const int K = RP_TW - 1 + v0, A = (KS - 1 - u) * G::RW;
        peo[i] = (unsigned)(A + (K >> 1) + (K & 1) * G::COPYW) | ((unsigned)(A + ((K - 1) >> 1) + ((K - 1) & 1) * G::COPYW) << 16);
        pg[i] = (side * KS * KS + u * KS + v0) >> 1;
        side_mask |= (unsigned)side << i;
        live_mask |= (unsigned)live << i;
    }
    int sh[G::NSI], sg[G::NSI];                         // smem half offset (copy 0), global element index
    unsigned sside = 0, slive = 0;
#pragma unroll
    for (int i = 0; i < G::NSI; ++i) {
        const int s = lane + 32 * i;
        const bool live = s < G::NS;
        const int ss = live ? s : 0;
        const int side = ss / KS, u = ss - side * KS;
        const int v = ((side + u) & 1) ? 0 : KS - 1;
        sh[i] = (KS - 1 - u) * 2 * G::RW + RP_TW - 1 + v;
        sg[i] = side * KS * KS + u * KS + v;
        sside |= (unsigned)side << i;
        slive |= (unsigned)live << i;
    }

    // ---- pixels: warp w owns tile rows w, w + 8 -----------------------------------------------------------------------
    for (int ly = warp; ly < RP_TH; ly += RP_WARPS) {
        const int y = y0 + ly;
        if (y >= row0 + nrw) break;
        for (int lx = 0; lx < RP_TW; ++lx) {
            const int x = x0 + lx;
            if (x >= W) break;
            const PsfT *kp = psf + (((int64_t)b * nrw + (y - row0)) * W + x) * (2 * KS * KS);
            const int base = ly * G::RW - (lx >> 1);
            const bool odd = lx & 1;
            float acc[2][RP_C];
#pragma unroll
            for (int s = 0; s < 2; ++s)
#pragma unroll
                for (int c = 0; c < RP_C; ++c) acc[s][c] = 0.0f;
            __half2 kv[G::NIT];
#pragma unroll
            for (int i = 0; i < G::NIT; ++i)
                kv[i] = ((live_mask >> i) & 1) ? PsfPair<PsfT>::load2(kp + 2 * pg[i]) : __float2half2_rn(0.0f);
            __half ks1[G::NSI];
#pragma unroll
            for (int i = 0; i < G::NSI; ++i)
                ks1[i] = ((slive >> i) & 1) ? PsfPair<PsfT>::load1(kp + sg[i]) : __float2half_rn(0.0f);
#pragma unroll
            for (int i = 0; i < G::NIT; ++i) {
                const int w = base + (int)(odd ? (peo[i] >> 16) : (peo[i] & 0xffffu));
                constexpr int half_units = KS * G::PR;
                const bool all0 = 32 * i + 31 < half_units, all1 = 32 * i >= half_units;   // compile time after unrolling
                const bool side = all1 ? true : (all0 ? false : (bool)((side_mask >> i) & 1));
#pragma unroll
                for (int c = 0; c < RP_C; ++c) {
                    const unsigned word = rp_smem[w + c * G::CHW];
                    const __half2 p = __hmul2(kv[i], *reinterpret_cast<const __half2 *>(&word));
                    if (!side) { fhadd(acc[0][c], __low2half(p)); fhadd(acc[0][c], __high2half(p)); }
                    else { fhadd(acc[1][c], __low2half(p)); fhadd(acc[1][c], __high2half(p)); }
                }
            }
#pragma unroll
            for (int i = 0; i < G::NSI; ++i) {
                const int hidx = ly * 2 * G::RW + sh[i] - lx;
                const bool side = (sside >> i) & 1;
#pragma unroll
                for (int c = 0; c < RP_C; ++c) {
                    const __half p = __hmul(ks1[i], s0h[hidx + c * 2 * G::CHW]);
                    if (side) fhadd(acc[1][c], p); else fhadd(acc[0][c], p);
                }
            }
            int oi;
            float v = reduce6(acc, lane, oi);
            if (reduce6_writer(lane)) {
                const int s = oi / RP_C, c = oi - s * RP_C;
                v = __half2float(__float2half_rn(v));
                if (tone & 2) v = fminf(fmaxf(tone_gamma(v), 0.0f), 1.0f);
                (s ? out_r : out_l)[(((int64_t)b * RP_C + c) * H + y) * W + x] = v;
            }
        }
    }
}

template <int KS>
static int launch_render_pairs(const float *img, const void *psf, int psf_is_half, int B, int H, int W, int row0, int nrw, int tone,
                               float *out_l, float *out_r, cudaStream_t st) {
    using G = RenderGeom<KS>;
    dim3 grid((W + RP_TW - 1) / RP_TW, (nrw + RP_TH - 1) / RP_TH, B);
    if (psf_is_half) {
        CUDA_TRY(cudaFuncSetAttribute(render_pairs_kernel<KS, __half>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES));
        render_pairs_kernel<KS, __half><<<grid, RP_WARPS * 32, G::SMEM_BYTES, st>>>(img, (const __half *)psf, B, H, W, row0, nrw, tone, out_l, out_r);
    } else {
        CUDA_TRY(cudaFuncSetAttribute(render_pairs_kernel<KS, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES));
        render_pairs_kernel<KS, float><<<grid, RP_WARPS * 32, G::SMEM_BYTES, st>>>(img, (const float *)psf, B, H, W, row0, nrw, tone, out_l, out_r);
    }
    return check_launch("render_pairs_kernel");
}

// ------------------------------------------------------------------------------------------------
// The same arithmetic with the kernels streamed by the TMA unit.
//
// The direct kernel above keeps one pixel's 1764 B in flight per warp (registers), ~42 KB per SM: about what
// Little's law asks for at HBM3e latency, i.e. no slack -> 0.33 of the HBM roofline (r01f).  Here the per-pixel kernel
// blocks of SEG consecutive pixels of a tile row -- one contiguous, 16-byte aligned run of the [B,H,W,2,ks,ks] tensor
// -- are fetched by ONE cp.async.bulk (UBLKCP) into a 3-stage shared-memory ring guarded by mbarriers; two stages
// (113 KB) are in flight while the 16 warps of the CTA consume the third.  Needs W % SEG == 0 (the host picks).
// ------------------------------------------------------------------------------------------------
#define RS_WARPS 16
#define RS_STAGES 3

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int KS, typename PsfT, int SEG>
struct StreamGeom {
    using G = RenderGeom<KS, SEG>;
    static constexpr int PIX_BYTES = 2 * KS * KS * (int)sizeof(PsfT);
    static constexpr int STAGE_BYTES = SEG * PIX_BYTES;
    static constexpr int TILE_OFF = RS_STAGES * STAGE_BYTES;                 // image tile after the ring
    static constexpr int BAR_OFF = (TILE_OFF + G::SMEM_BYTES + 15) / 16 * 16;
    static constexpr int SMEM_BYTES = BAR_OFF + 2 * RS_STAGES * 8;
    static_assert(STAGE_BYTES % 16 == 0, "bulk copies move multiples of 16 bytes");
};

template <int KS, typename PsfT, int SEG>
__global__ void __launch_bounds__(RS_WARPS * 32, 1)
render_stream_kernel(const float *__restrict__ img, const PsfT *__restrict__ psf, int B, int H, int W, int row0, int nrw, int tone,
                     float *__restrict__ out_l, float *__restrict__ out_r) {
    using SG = StreamGeom<KS, PsfT, SEG>;
    using G = typename SG::G;
    extern __shared__ __align__(128) unsigned char rs_raw[];
    unsigned *tile = reinterpret_cast<unsigned *>(rs_raw + SG::TILE_OFF);
    __half *s0h = reinterpret_cast<__half *>(tile);
    unsigned long long *full = reinterpret_cast<unsigned long long *>(rs_raw + SG::BAR_OFF), *empty = full + RS_STAGES;
    constexpr int pad = (KS - 1) / 2;
    const int b = blockIdx.z, y0 = row0 + blockIdx.y * RP_TH, x0 = blockIdx.x * SEG;
    const int nrows = min(RP_TH, row0 + nrw - y0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    if (threadIdx.x == 0) {
        for (int s = 0; s < RS_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, RS_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int row) {
        const int st = row % RS_STAGES;
        mbar_expect_tx(full + st, SG::STAGE_BYTES);
        bulk_g2s(rs_raw + st * SG::STAGE_BYTES, psf + (((int64_t)b * nrw + (y0 - row0 + row)) * W + x0) * (2 * KS * KS), SG::STAGE_BYTES, full + st);
    };
    if (threadIdx.x == 0)
        for (int row = 0; row < min(RS_STAGES, nrows); ++row) issue(row);

    // ---- image tile (mirrored, two parity copies), while the first rows are in flight ------------------------------------
    for (int i = threadIdx.x; i < RP_C * G::TH * 2 * G::RW; i += blockDim.x) {
        const int c = i / (G::TH * 2 * G::RW), rem = i - c * (G::TH * 2 * G::RW);
        const int r = rem / (2 * G::RW), m = rem - r * (2 * G::RW);
        float v = 0.0f;
        if (m < G::TW) {
            const int gy = min(max(y0 + r - pad, 0), H - 1), gx = min(max(x0 + (G::TW - 1 - m) - pad, 0), W - 1);
            v = img[(((int64_t)b * RP_C + c) * H + gy) * W + gx];
            if (tone & 1) v = tone_degamma(v);
        }
        s0h[i] = __float2half_rn(v);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < G::COPYW; i += blockDim.x) {
        const int w = i % G::RW;
        const __half lo = s0h[2 * i + 1];
        const __half hi = (w + 1 < G::RW) ? s0h[2 * i + 2] : __float2half_rn(0.0f);
        tile[G::COPYW + i] = (unsigned)__half_as_ushort(lo) | ((unsigned)__half_as_ushort(hi) << 16);
    }
    __syncthreads();

    // ---- per-lane tap tables: byte offsets, for THIS warp's pixel-column parity (lx = warp, warp + 16, ...: fixed parity) ----
    static_assert(RS_WARPS % 2 == 0, "a warp's pixel columns must share their parity");
    const int par = warp & 1;
    int ia[G::NIT], ga[G::NIT];                          // image word / kernel pair byte offsets
    unsigned side_mask = 0;
#pragma unroll
    for (int i = 0; i < G::NIT; ++i) {
        const int q = min(lane + 32 * i, G::NP - 1);    // lanes past the last unit repeat it with a zero kernel
        const int side = q / (KS * G::PR), rem = q - side * (KS * G::PR);
        const int u = rem / G::PR, m = rem - u * G::PR;
        const int v0 = 2 * m + ((side + u) & 1);
        const int K = SEG - 1 + v0 - par, A = (KS - 1 - u) * G::RW;
        ia[i] = 4 * (A + (K >> 1) + (K & 1) * G::COPYW);
        ga[i] = (side * KS * KS + u * KS + v0) * (int)sizeof(PsfT);
        side_mask |= (unsigned)side << i;
    }
    int sh[G::NSI], sg[G::NSI];
    unsigned sside = 0;
#pragma unroll
    for (int i = 0; i < G::NSI; ++i) {
        const int ss = min(lane + 32 * i, G::NS - 1);
        const int side = ss / KS, u = ss - side * KS;
        const int v = ((side + u) & 1) ? 0 : KS - 1;
        sh[i] = 2 * ((KS - 1 - u) * 2 * G::RW + SEG - 1 + v);
        sg[i] = (side * KS * KS + u * KS + v) * (int)sizeof(PsfT);
        sside |= (unsigned)side << i;
    }
    const unsigned char *tile_b = reinterpret_cast<const unsigned char *>(tile);

    // ---- rows of the tile, one ring stage each; a warp takes pixels warp, warp + 16, ... of the row ----------------------
    for (int ly = 0; ly < nrows; ++ly) {
        const int st = ly % RS_STAGES;
        const unsigned phase = (unsigned)(ly / RS_STAGES) & 1u;
        mbar_wait(full + st, phase);
        const PsfT *stage = reinterpret_cast<const PsfT *>(rs_raw + st * SG::STAGE_BYTES);
        const int y = y0 + ly;
        for (int lx = warp; lx < SEG; lx += RS_WARPS) {
            const unsigned char *kb = rs_raw + st * SG::STAGE_BYTES + lx * SG::PIX_BYTES;
            const unsigned char *pb = tile_b + 4 * (ly * G::RW - (lx >> 1));
            float acc[2][RP_C];
#pragma unroll
            for (int s = 0; s < 2; ++s)
#pragma unroll
                for (int c = 0; c < RP_C; ++c) acc[s][c] = 0.0f;
#pragma unroll
            for (int i = 0; i < G::NIT; ++i) {
                constexpr int half_units = KS * G::PR;
                const bool all_live = 32 * i + 31 < G::NP;                                   // compile time after unrolling
                const bool all0 = 32 * i + 31 < half_units, all1 = 32 * i >= half_units;
                __half2 kv = PsfPair<PsfT>::load2(reinterpret_cast<const PsfT *>(kb + ga[i]));
                if (!all_live && lane + 32 * i >= G::NP) kv = __float2half2_rn(0.0f);
                const unsigned char *wp = pb + ia[i];
                const bool side = all1 ? true : (all0 ? false : (bool)((side_mask >> i) & 1));
#pragma unroll
                for (int c = 0; c < RP_C; ++c) {
                    const unsigned word = *reinterpret_cast<const unsigned *>(wp + c * (4 * G::CHW));
                    const __half2 p = __hmul2(kv, *reinterpret_cast<const __half2 *>(&word));
                    if (!side) { fhadd(acc[0][c], __low2half(p)); fhadd(acc[0][c], __high2half(p)); }
                    else { fhadd(acc[1][c], __low2half(p)); fhadd(acc[1][c], __high2half(p)); }
                }
            }
#pragma unroll
            for (int i = 0; i < G::NSI; ++i) {
                const bool all_live = 32 * i + 31 < G::NS;
                __half k1 = PsfPair<PsfT>::load1(reinterpret_cast<const PsfT *>(kb + sg[i]));
                if (!all_live && lane + 32 * i >= G::NS) k1 = __float2half_rn(0.0f);
                const unsigned char *hp = tile_b + 2 * (ly * 2 * G::RW - lx) + sh[i];
                const bool side = (sside >> i) & 1;
#pragma unroll
                for (int c = 0; c < RP_C; ++c) {
                    const __half p = __hmul(k1, *reinterpret_cast<const __half *>(hp + c * (4 * G::CHW)));
                    if (side) fhadd(acc[1][c], p); else fhadd(acc[0][c], p);
                }
            }
            int oi;
            float v = reduce6(acc, lane, oi);
            if (reduce6_writer(lane)) {
                const int s = oi / RP_C, c = oi - s * RP_C;
                v = __half2float(__float2half_rn(v));
                if (tone & 2) v = fminf(fmaxf(tone_gamma(v), 0.0f), 1.0f);
                (s ? out_r : out_l)[(((int64_t)b * RP_C + c) * H + y) * W + (x0 + lx)] = v;
            }
        }
        // this warp is done with the stage; warp 0 refills it with row ly + STAGES once every warp has let go
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + st);
        if (threadIdx.x == 0 && ly + RS_STAGES < nrows) {
            mbar_wait(empty + st, phase);
            issue(ly + RS_STAGES);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Lane-per-pixel streamed kernel (fp16 kernels, the tensor PSFNet.pred produces under CUDA autocast).
//
// The kernels above give every pixel to a warp (lanes over taps): each lane needs its own table of tap positions,
// the six partial sums are warp-reduced per pixel, and the 10-pairs-then-a-single rhythm of a kernel row makes most
// shared-memory loads 2-way bank conflicted (measured: 0.44 of the HBM roofline, shared-memory pipe bound).  Here the
// 32 lanes of a warp ARE the 32 pixels of the staged row segment and a warp owns a few whole kernel rows of BOTH sides:
//   * the per-pixel kernel blocks are KS*KS words apart in the stage (an odd stride): the 32 lanes of a load hit 32
//     different banks; the image words of 32 neighbouring pixels are 16 consecutive words in each of two parity
//     copies placed half a bank-cycle apart: conflict free as well;
//   * all lanes walk the same taps, so every address is lane base + immediate: no tables, no integer arithmetic;
//   * kernel row u of the left and of the right kernel multiply the SAME image row, so its words are loaded once for
//     the two (round 1 / early round 2 gave a warp rows of one side only: 8 shared-memory loads per pair of taps of
//     the two sides; now 5, and 12 % fewer instructions per row in all -- the issue slots, not the shared-memory pipe,
//     were what the earlier kernel ran out of at the HBM roofline's pace).  The two rows start at
//     elements of different parity in the pixel's block (KS*KS is odd), so one of them arrives shifted by one half
//     against the image words: its kernel words are re-paired with one PRMT each (the image words are shared by three
//     channels, the kernel words are not: permuting the kernel side costs a third of permuting the image side);
//   * every lane keeps its own pixel's sums: no shuffles; the KS/TPW warps add their partial sums through shared
//     memory once per row.
// Per pair of taps of BOTH sides: 2 LDS (kernel pairs) + 3 LDS (image, 3 channels) + 1 PRMT + 6 HMUL2 + 12 FHADD.
// ------------------------------------------------------------------------------------------------
#ifndef SDIRT_RENDER_SPLIT_ACC
#define SDIRT_RENDER_SPLIT_ACC 1     // the two taps of a pair sum into separate fp32 accumulators (12 independent chains per warp instead of 6)
#endif
#ifndef SDIRT_RENDER_REDUCERS_SAME_SMSP
#define SDIRT_RENDER_REDUCERS_SAME_SMSP 1
#endif
#define RL_PBUF 4     // partial-sum buffers between the compute warps and the reducer warps
#define RL_RING 32    // image rows resident (a power of two >= KS + RS_STAGES)

template <int KS>
struct LaneGeom {
    static constexpr int SEG = 32;
    static constexpr int TPW = (KS % 3 == 0) ? 3 : 1;          // kernel rows (of both sides) per warp
    static constexpr int NW = KS / TPW;                         // compute warps
#if SDIRT_RENDER_REDUCERS_SAME_SMSP
    // + one reducer per side + the streamer.  Warp w issues from scheduler w % 4: with 7 (11) compute warps scheduler 3 has one
    // compute warp fewer than the others, so BOTH reducers go there (warp ids = 3 mod 4: their ~200 instructions per row -- sums,
    // rounding, tone curve, stores -- would otherwise delay the two compute warps of scheduler 0, and the slowest compute warp
    // paces the CTA); the warps in between have no role and leave at once.
    static constexpr int RED0 = NW + ((3 - NW % 4) + 4) % 4;     // the first warp id past the compute warps that is 3 mod 4
    static constexpr int RED1 = RED0 + 4, STREAM = RED0 + 1;
    static constexpr int NWARPS = RED1 + 1;
#else
    static constexpr int RED0 = NW, RED1 = NW + 1, STREAM = NW + 2;
    static constexpr int NWARPS = NW + 3;                       // + one reducer per side + the streamer
#endif
    static_assert(RED0 >= NW && STREAM > RED0 && STREAM != RED1, "warp roles");
    static constexpr int TW = SEG + KS - 1;
    static constexpr int RW = (TW + 2) / 2;                     // words per image row and channel
    // an image-row record: what one 32-pixel strip needs of one padded image row, as the compute warps read it:
    // copy 0 = [channel][RW words of elements (2w, 2w+1)], copy 1 = the same row shifted by one element (2w+1, 2w+2),
    // placed at a word offset == 16 (mod 32) so that the even and the odd lanes of a load fall into different banks
    static constexpr int COPY_RAW = RP_C * RW;
    static constexpr int COPY1 = COPY_RAW + ((16 - COPY_RAW % 32) + 32) % 32;
    static constexpr int REC_WORDS = (COPY1 + COPY_RAW + 3) / 4 * 4;
    static constexpr int REC_BYTES = REC_WORDS * 4;
    static constexpr int PIX_BYTES = 2 * KS * KS * 2;
    static constexpr int STAGE_BYTES = SEG * PIX_BYTES;
    static constexpr int TILE_OFF = RS_STAGES * STAGE_BYTES;    // the ring of image-row records after the ring of kernel stages
    static constexpr int PART_OFF = (TILE_OFF + RL_RING * REC_BYTES + 15) / 16 * 16;
    static constexpr int PART_BYTES = RL_PBUF * NW * 2 * RP_C * 32 * 4;   // RL_PBUF buffers of [NW][side][C][32] floats
    static constexpr int BAR_OFF = PART_OFF + PART_BYTES;
    static constexpr int SMEM_BYTES = BAR_OFF + (2 * RS_STAGES + 2 * RL_PBUF) * 8;
    static_assert(STAGE_BYTES % 16 == 0 && TILE_OFF % 16 == 0, "bulk copies move multiples of 16 bytes");
    static_assert(KS % TPW == 0 && NWARPS <= 32, "warp roles");
    static_assert((RL_RING & (RL_RING - 1)) == 0 && RL_RING >= KS + RS_STAGES, "an image row is replaced once the output rows that read it are done");
};

// The image as the strips read it, built once per launch: record (b, yb, tx) = padded image row row0 - pad + yb (replicate
// padding, PSFNet.degamma applied, rounded to fp16) over the columns of strip tx, column-mirrored (so that increasing tap index
// is increasing address), in the two pairings of LaneGeom.  ~25 B per pixel against the 1764 B of its kernels.
// One CTA per padded image row: the row is tone-mapped, rounded, replicate-padded and MIRRORED once into shared memory
// (element j = padded column W + KS-2 - j); strip tx's TW elements then start at j = W - 32 (tx + 1), an aligned word: copy 0 of
// its record is a run of those words, copy 1 the same run shifted by one element (a funnel shift of two neighbouring words).
template <int KS>
__global__ void __launch_bounds__(256)
render_pack_image_kernel(const float *__restrict__ img, int B, int H, int W, int row0, int nrw, int tone, unsigned *__restrict__ rec) {
    using G = LaneGeom<KS>;
    static_assert(G::RW <= 32, "a warp writes one channel row of a record per pass");
    extern __shared__ unsigned rpk_row[];                // [RP_C][pitch] words of two fp16 each
    constexpr int pad = (KS - 1) / 2;
    const int tx_n = W / G::SEG, nyb = nrw + KS - 1;
    const int wp = W + KS - 1, pitch = wp / 2 + 2;       // (W is a multiple of 32 and KS is odd: wp is even) + one word read past the end by copy 1
    const int yb = blockIdx.x % nyb, b = blockIdx.x / nyb;
    const int gy = min(max(row0 - pad + yb, 0), H - 1);
#pragma unroll
    for (int ch = 0; ch < RP_C; ++ch) {
        const float *rowp = img + (((int64_t)b * RP_C + ch) * H + gy) * W;
#pragma unroll 4
        for (int jw = threadIdx.x; jw < pitch; jw += blockDim.x) {
            // elements j = 2 jw, 2 jw + 1 = image columns wp - 1 - j - pad, replicate-padded; zero past the row
            const int x1 = wp - 1 - pad - 2 * jw;
            float v0 = __ldg(rowp + min(max(x1, 0), W - 1)), v1 = __ldg(rowp + min(max(x1 - 1, 0), W - 1));
            if (tone & 1) { v0 = tone_degamma(v0); v1 = tone_degamma(v1); }
            const unsigned word = (unsigned)__half_as_ushort(__float2half_rn(v0)) | ((unsigned)__half_as_ushort(__float2half_rn(v1)) << 16);
            rpk_row[ch * pitch + jw] = 2 * jw < wp ? word : 0u;
        }
    }
    __syncthreads();
    // a warp per record, a lane per word of a channel row (the words between and after the two copies are never read)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int tx = warp; tx < tx_n; tx += nwarp) {
        unsigned *out = rec + ((int64_t)blockIdx.x * tx_n + tx) * G::REC_WORDS;
        const unsigned *src = rpk_row + (W - G::SEG * (tx + 1)) / 2 + lane;
        if (lane < G::RW) {
#pragma unroll
            for (int ch = 0; ch < RP_C; ++ch) {
                const unsigned w0 = src[ch * pitch], w1 = src[ch * pitch + 1];
                // elements m = 2 lane (+1) of copy 0, 2 lane + 1 (+1) of copy 1: nothing past the strip's TW elements
                unsigned c0 = w0, c1 = __funnelshift_r(w0, w1, 16);
                if (2 * lane + 1 >= G::TW) c0 = 2 * lane < G::TW ? (c0 & 0xffffu) : 0u;
                if (2 * lane + 2 >= G::TW) c1 = 2 * lane + 1 < G::TW ? (c1 & 0xffffu) : 0u;
                out[ch * G::RW + lane] = c0;
                out[G::COPY1 + ch * G::RW + lane] = c1;
            }
        }
    }
}

// Kernel row u of both sides for this lane's pixel.  In the pixel's block one of the two rows starts at an even element
// (aligned pairs at v = 0, 2, ..., single tap v = KS-1: `ka`) and the other at an odd one (single tap v = 0, aligned pairs at
// v = 1, 3, ...: `kb`).  The image words are fetched once, in the pairing of the even row; the odd row's words
// (k[2j+1], k[2j+2]) are re-paired into (k[2j], k[2j+1]) -- one PRMT each -- and its last tap is the high half of its last word.
//   ka / kb: this lane's kernel block + byte offset of the row's first element (even / odd side)
//   ib: this lane's image word address for the pair at v = 0 (copy already chosen by lane parity)
//   sb: address of the image element (copy 0) under tap v = KS-1
template <int KS, int CHB>
__device__ __forceinline__ void lane_row_pair(const unsigned char *ka, const unsigned char *kb, const unsigned char *ib, const unsigned char *sb,
                                              float (&acc_a)[RP_C], float (&acc_b)[RP_C], float (&hi_a)[RP_C], float (&hi_b)[RP_C]) {
    constexpr int PR = (KS - 1) / 2;
    unsigned prev = *reinterpret_cast<const unsigned short *>(kb);                 // k[0] of the odd row, low half
#pragma unroll
    for (int k = 0; k < PR; ++k) {
        const unsigned wa = *reinterpret_cast<const unsigned *>(ka + 4 * k);
        const unsigned wb = *reinterpret_cast<const unsigned *>(kb + 2 + 4 * k);
        const unsigned wr = (k == 0) ? __byte_perm(prev, wb, 0x5410) : __byte_perm(prev, wb, 0x5432);
        prev = wb;
        const __half2 kva = *reinterpret_cast<const __half2 *>(&wa), kvb = *reinterpret_cast<const __half2 *>(&wr);
#pragma unroll
        for (int c = 0; c < RP_C; ++c) {
            const __half2 im = *reinterpret_cast<const __half2 *>(ib + 4 * k + c * CHB);
            const __half2 pa = __hmul2(kva, im), pb = __hmul2(kvb, im);
            fhadd(acc_a[c], __low2half(pa));
            fhadd(hi_a[c], __high2half(pa));
            fhadd(acc_b[c], __low2half(pb));
            fhadd(hi_b[c], __high2half(pb));
        }
    }
    const __half ka1 = *reinterpret_cast<const __half *>(ka + 2 * (KS - 1));
    const __half kb1 = __high2half(*reinterpret_cast<const __half2 *>(&prev));
#pragma unroll
    for (int c = 0; c < RP_C; ++c) {
        const __half im = *reinterpret_cast<const __half *>(sb + c * CHB);
        fhadd(acc_a[c], __hmul(ka1, im));
        fhadd(acc_b[c], __hmul(kb1, im));
    }
}

// Warp roles: warps 0 .. NW-1 compute (each a few kernel rows of both sides, lanes = the 32 pixels of the row segment);
// warps NW, NW+1 reduce one side each (add its NW partial sums, round, tone-map, write the pixels); lane 0 of warp NW+2 streams:
// it issues the bulk copies.  Everything between the roles is an mbarrier: no CTA-wide barrier after the first.
//
// Persistent, strip-walking: the launch's 32-pixel row segments, ordered strip by strip (strip = a 32-pixel column of one image,
// top to bottom), are cut into gridDim.x equal contiguous chunks, one per CTA (148 chunks of ~664 rows for two 1024 x 1536
// images: balanced to one row).  Walking DOWN a strip, consecutive output rows share all but one of their KS image rows, so
// the image lives in a ring of RL_RING row records: the bulk copy of output row i's kernels (56 KB) is followed by one of image
// row i + KS-1 (0.8 KB, from render_pack_image_kernel's records) on the same mbarrier, into the slot of the image row that output
// row i - (RL_RING - KS + 1) was the last to read.  No thread of the CTA touches the image on its way in -- the earlier version's
// 16-row tiles had their 36 image rows loaded, tone-mapped and converted by the whole CTA between two tiles with every
// compute warp waiting, about a sixth of the kernel's time (r02H profile).  Where a chunk crosses into the next strip (at most
// twice per CTA for these images) the streamer lets the rows of the old strip finish before it sends the new strip's first
// KS image rows (they would overwrite rows still being read): one copy latency, nothing else.
template <int KS>
__global__ void __launch_bounds__(LaneGeom<KS>::NWARPS * 32, 1)
render_lanes_kernel(const unsigned *__restrict__ rec, int rec_nyb, int rec_y0, const __half *__restrict__ psf, int B, int H, int W, int row0, int nrw,
                    int tone, float *__restrict__ out_l, float *__restrict__ out_r) {
    using G = LaneGeom<KS>;
    extern __shared__ __align__(128) unsigned char rl_raw[];
    float *part = reinterpret_cast<float *>(rl_raw + G::PART_OFF);
    unsigned long long *full = reinterpret_cast<unsigned long long *>(rl_raw + G::BAR_OFF), *empty = full + RS_STAGES;
    unsigned long long *pfull = empty + RS_STAGES, *pempty = pfull + RL_PBUF;
    constexpr int SEG = G::SEG, CHB = 4 * G::RW;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tx_n = W / SEG;
    // (rec holds rec_nyb record rows per image; record row rec_y0 is padded image row row0 - pad)
    // this CTA's chunk of the strip-major row order: rows j0 .. j0 + n - 1; row j = row j % nrw of strip j / nrw
    const int64_t total = (int64_t)B * tx_n * nrw;
    const int64_t j0 = total * blockIdx.x / gridDim.x;
    const int n = (int)(total * (blockIdx.x + 1) / gridDim.x - j0);
    const int strip0 = (int)(j0 / nrw), yl0 = (int)(j0 - (int64_t)strip0 * nrw);

    if (threadIdx.x == 0) {
        for (int s = 0; s < RS_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, G::NW); }
        for (int s = 0; s < RL_PBUF; ++s) { mbar_init(pfull + s, G::NW); mbar_init(pempty + s, 2); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp >= G::NW && warp != G::RED0 && warp != G::RED1 && warp != G::STREAM) return;       // no role
    if (warp == G::STREAM) {
        // ---- streamer ---------------------------------------------------------------------------------------------------------
        if (lane != 0) return;
        // next row to send: the p_g-th of the chunk = row p_yl of strip p_strip; its run (the chunk's rows in that strip) began
        // with chunk row p_run0 at strip row p_ys
        int p_strip = strip0, p_yl = yl0, p_g = 0, p_run0 = 0, p_ys = yl0;
        auto issue_next = [&]() {
            const int b = p_strip / tx_n, tx = p_strip - b * tx_n, st = p_g % RS_STAGES;
            const int i = p_g - p_run0;                                   // row of the run
            const int nimg = i == 0 ? KS : 1, r0 = i == 0 ? 0 : i + KS - 1;   // image rows of the run riding with it
            mbar_expect_tx(full + st, G::STAGE_BYTES + nimg * G::REC_BYTES);
            bulk_g2s(rl_raw + st * G::STAGE_BYTES, psf + (((int64_t)b * nrw + p_yl) * W + tx * SEG) * (2 * KS * KS), G::STAGE_BYTES, full + st);
            for (int k = 0; k < nimg; ++k) {
                const int r = r0 + k;                                     // image row row0 - pad + p_ys + r = record row rec_y0 + p_ys + r
                bulk_g2s(rl_raw + G::TILE_OFF + (r & (RL_RING - 1)) * G::REC_BYTES,
                         rec + (((int64_t)b * rec_nyb + rec_y0 + p_ys + r) * tx_n + tx) * G::REC_WORDS, G::REC_BYTES, full + st);
            }
            ++p_g;
            if (++p_yl == nrw) { p_yl = 0; p_ys = 0; ++p_strip; p_run0 = p_g; }
        };
        while (p_g < n && p_g < RS_STAGES && p_run0 == 0) issue_next();
        for (int g = 0; g < n; ++g) {
            mbar_wait(empty + g % RS_STAGES, (unsigned)(g / RS_STAGES) & 1u);   // every compute warp has left row g (and every row before it)
            // a stage is free once its previous row is done; a new run's image rows wait for ALL rows of the run before
            while (p_g < n && p_g <= g + RS_STAGES && p_run0 <= g + 1) issue_next();
        }
        return;
    }

    // compute warps: this warp's kernel rows and this lane's pixel (constant over the chunk)
    // pixel lx = lane; tap (u, v) multiplies mirrored element m = SEG-1-lane + v of image row i + KS-1-u of the run
    const int u0 = warp * G::TPW;
    const unsigned char *tb = rl_raw + G::TILE_OFF;
    // the pair at v = 0: m0 = 31 - lane: odd lanes -> copy 0 word (31-lane)/2, even lanes -> copy 1 word (30-lane)/2
    const int img0 = (lane & 1) ? 4 * ((SEG - 1 - lane) >> 1) : 4 * (G::COPY1 + ((SEG - 2 - lane) >> 1));
    const int sgl0 = 2 * (SEG - 1 - lane + KS - 1);                                    // the single tap v = KS-1 (copy 0, by element)

    int strip = strip0, yl = yl0;
    for (int g0 = 0; g0 < n;) {
        // ---- a run: rows yl .. yl + rows - 1 of one strip = rows g0 .. g0 + rows - 1 of the chunk -----------------------------
        const int rows = min(nrw - yl, n - g0);
        if (warp >= G::NW) {
            // ---- reducer warps (one per side) -----------------------------------------------------------------------------------
            const int s = warp == G::RED1;
            const int b = strip / tx_n, x0 = (strip - b * tx_n) * SEG;
            float *outp = (s ? out_r : out_l) + ((int64_t)b * RP_C * H + row0 + yl) * W + x0 + lane;
            for (int i = 0; i < rows; ++i) {
                const int g = g0 + i, pb = g % RL_PBUF;
                mbar_wait(pfull + pb, (unsigned)(g / RL_PBUF) & 1u);
                const float *pr = part + pb * (G::NW * 2 * RP_C * 32);
                float v[RP_C];
#pragma unroll
                for (int c = 0; c < RP_C; ++c) {
                    const float *ps = pr + s * (RP_C * 32) + c * 32 + lane;
                    v[c] = 0.0f;
#pragma unroll
                    for (int k = 0; k < G::NW; ++k) v[c] += ps[k * (2 * RP_C * 32)];
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(pempty + pb);           // the partial sums are in registers: the buffer goes back before the tone curve
#pragma unroll
                for (int c = 0; c < RP_C; ++c) {
                    float o = __half2float(__float2half_rn(v[c]));
                    if (tone & 2) o = fminf(fmaxf(tone_gamma(o), 0.0f), 1.0f);
                    outp[((int64_t)c * H + i) * W] = o;
                }
            }
        } else {
            // ---- compute warps ---------------------------------------------------------------------------------------------
            for (int i = 0; i < rows; ++i) {
                const int g = g0 + i, st = g % RS_STAGES, pb = g % RL_PBUF;
                mbar_wait(full + st, (unsigned)(g / RS_STAGES) & 1u);     // row g's kernels and image row i + KS-1 of the run (the earlier ones came before)
                const unsigned char *kblock = rl_raw + st * G::STAGE_BYTES + lane * G::PIX_BYTES;
                float acc[2][RP_C];
#pragma unroll
                for (int c = 0; c < RP_C; ++c) acc[0][c] = acc[1][c] = 0.0f;
#if SDIRT_RENDER_SPLIT_ACC
                float hi[2][RP_C];                          // the second tap of every pair sums apart: twice the independent chains
#pragma unroll
                for (int c = 0; c < RP_C; ++c) hi[0][c] = hi[1][c] = 0.0f;
#else
                float (&hi)[2][RP_C] = acc;
#endif
#pragma unroll
                for (int tk = 0; tk < G::TPW; ++tk) {
                    const int u = u0 + tk;
                    const unsigned char *kl = kblock + 2 * u * KS, *kr = kl + 2 * KS * KS;      // row u of the left / right kernel
                    const unsigned char *rowb = tb + ((i + KS - 1 - u) & (RL_RING - 1)) * G::REC_BYTES;
                    // element u*KS of the block is even for even u (KS is odd): then the left row is the aligned one
                    if (u & 1) lane_row_pair<KS, CHB>(kr, kl, rowb + img0, rowb + sgl0, acc[1], acc[0], hi[1], hi[0]);
                    else lane_row_pair<KS, CHB>(kl, kr, rowb + img0, rowb + sgl0, acc[0], acc[1], hi[0], hi[1]);
                }
#if SDIRT_RENDER_SPLIT_ACC
#pragma unroll
                for (int c = 0; c < RP_C; ++c) { acc[0][c] += hi[0][c]; acc[1][c] += hi[1][c]; }
#endif
                if (g >= RL_PBUF) mbar_wait(pempty + pb, (unsigned)(g / RL_PBUF - 1) & 1u);   // the reducers are done with this buffer
                float *pw = part + (pb * G::NW + warp) * (2 * RP_C * 32);
#pragma unroll
                for (int s = 0; s < 2; ++s)
#pragma unroll
                    for (int c = 0; c < RP_C; ++c) pw[(s * RP_C + c) * 32 + lane] = acc[s][c];
                __syncwarp();
                if (lane == 0) { mbar_arrive(empty + st); mbar_arrive(pfull + pb); }
            }
        }
        g0 += rows;
        yl = 0;
        ++strip;
    }
}

// the stream-ordered pool keeps what it is given back (the default is to return it to the driver at the next synchronisation)
static int render_pool_keep(cudaStream_t) {
    static bool done[64] = {};
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !done[dev]) {
        cudaMemPool_t pool;
        CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, dev));
        unsigned long long keep = ~0ull;
        CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        done[dev] = true;
    }
    return 0;
}

// can the strip-walking kernel take this shape?  (fp16 kernels, 16-byte aligned, are the caller's to check)
template <int KS>
static bool render_lanes_ok(int W) { return W % 32 == 0 && W <= 32768 && LaneGeom<KS>::SMEM_BYTES <= 227 * 1024; }
template <int KS>
static int64_t render_records_bytes(int B, int n_rows, int W) {
    return (int64_t)B * (n_rows + KS - 1) * (W / LaneGeom<KS>::SEG) * LaneGeom<KS>::REC_BYTES;
}

// records of padded image rows row0 - pad .. row0 + nrw - 1 + pad of every image
template <int KS>
static int launch_render_pack(const float *img, int B, int H, int W, int row0, int nrw, int tone, unsigned *rec, cudaStream_t st) {
    const size_t pack_smem = (size_t)RP_C * ((W + KS - 1) / 2 + 2) * sizeof(unsigned);
    if (pack_smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(render_pack_image_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pack_smem));
    render_pack_image_kernel<KS><<<(unsigned)(B * (nrw + KS - 1)), 256, pack_smem, st>>>(img, B, H, W, row0, nrw, tone, rec);
    return check_launch("render_pack_image_kernel");
}

// rows [row0, row0 + nrw) from records: rec holds rec_nyb record rows per image, record row rec_y0 = padded image row row0 - pad
template <int KS>
static int launch_render_strips(const unsigned *rec, int rec_nyb, int rec_y0, const __half *psf, int B, int H, int W, int row0, int nrw, int tone,
                                float *out_l, float *out_r, cudaStream_t st) {
    using G = LaneGeom<KS>;
    const int64_t total = (int64_t)B * (W / G::SEG) * nrw;                       // 32-pixel row segments
    if (total >= ((int64_t)1 << 31)) return fail(SDIRT_E_ARG, "render: too many rows");
    if (total == 0) return 0;
    const int sms = std::max(sdirt_device_sm_count(), 1);
    // one chunk per SM (a chunk starts by fetching KS image rows: not fewer than 8 output rows per chunk)
    dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>(sms, total / 8)));
    CUDA_TRY(cudaFuncSetAttribute(render_lanes_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES));
    render_lanes_kernel<KS><<<grid, G::NWARPS * 32, G::SMEM_BYTES, st>>>(rec, rec_nyb, rec_y0, psf, B, H, W, row0, nrw, tone, out_l, out_r);
    return check_launch("render_lanes_kernel");
}

template <int KS>
static int launch_render_lanes(const float *img, const __half *psf, int B, int H, int W, int row0, int nrw, int tone,
                               float *out_l, float *out_r, cudaStream_t st) {
    if (int rc = render_pool_keep(st)) return rc;
    unsigned *rec = nullptr;
    CUDA_TRY(cudaMallocAsync((void **)&rec, (size_t)render_records_bytes<KS>(B, nrw, W), st));
    int rc = launch_render_pack<KS>(img, B, H, W, row0, nrw, tone, rec, st);
    if (!rc) rc = launch_render_strips<KS>(rec, nrw + KS - 1, 0, psf, B, H, W, row0, nrw, tone, out_l, out_r, st);
    cudaFreeAsync(rec, st);
    return rc;
}

// W % SEG == 0 and 16-byte aligned rows are what the bulk copies need; anything else runs the direct kernel.
template <int KS>
static int launch_render(const float *img, const void *psf, int psf_is_half, int B, int H, int W, int row0, int nrw, int tone,
                         float *out_l, float *out_r, cudaStream_t st) {
    if (psf_is_half && ((uintptr_t)psf & 15) == 0 && render_lanes_ok<KS>(W))
        return launch_render_lanes<KS>(img, (const __half *)psf, B, H, W, row0, nrw, tone, out_l, out_r, st);
    if (psf_is_half && W % 32 == 0 && ((uintptr_t)psf & 15) == 0) {
        using SG = StreamGeom<KS, __half, 32>;
        if (SG::SMEM_BYTES <= 227 * 1024) {
            dim3 grid(W / 32, (nrw + RP_TH - 1) / RP_TH, B);
            CUDA_TRY(cudaFuncSetAttribute(render_stream_kernel<KS, __half, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, SG::SMEM_BYTES));
            render_stream_kernel<KS, __half, 32><<<grid, RS_WARPS * 32, SG::SMEM_BYTES, st>>>(img, (const __half *)psf, B, H, W, row0, nrw, tone, out_l, out_r);
            return check_launch("render_stream_kernel");
        }
    }
    if (!psf_is_half && W % 16 == 0 && ((uintptr_t)psf & 15) == 0) {
        using SG = StreamGeom<KS, float, 16>;
        if (SG::SMEM_BYTES <= 227 * 1024) {
            dim3 grid(W / 16, (nrw + RP_TH - 1) / RP_TH, B);
            CUDA_TRY(cudaFuncSetAttribute(render_stream_kernel<KS, float, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SG::SMEM_BYTES));
            render_stream_kernel<KS, float, 16><<<grid, RS_WARPS * 32, SG::SMEM_BYTES, st>>>(img, (const float *)psf, B, H, W, row0, nrw, tone, out_l, out_r);
            return check_launch("render_stream_kernel");
        }
    }
    return launch_render_pairs<KS>(img, psf, psf_is_half, B, H, W, row0, nrw, tone, out_l, out_r, st);
}


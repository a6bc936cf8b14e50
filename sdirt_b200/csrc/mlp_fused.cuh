// mlp_fused.cuh — PSFNet.pred for a band of pixels as ONE kernel on the 5th-generation tensor cores (tcgen05 / TMEM).
// Included by engine.cu.
//
// What it replaces (deeplens/psfnet.py:317-336, 681-705; deeplens/psfnet_arch.py:32-56): coordinate grid -> MLP(x, y, z) and
// MLP(-x, y, z) under CUDA autocast (fp16 operands, fp32 accumulation, one fp16 rounding + ReLU per layer) -> flip the right
// kernels -> stack -> divide by sum(-1).sum(-1) + 1e-9.  The cuBLAS route (PSFNet.render with mlp_engine = "cublas") runs
// this as 13 launches per band and moves every activation [rows, 512] through L2 twice per layer.  Here a CTA owns a tile
// of 128 MLP rows (64 pixels x 2 sides) from the coordinates to the normalised kernels:
//
//   * the activation tile [128, K <= 512] fp16 lives in shared memory (128 KB) in the canonical K-major SWIZZLE_128B
//     layout of a tcgen05 A operand and never leaves the SM;
//   * the weights arrive PRE-SWIZZLED (sdirt_mlp_fused_pack_weights lays every [<= 256 x 64] tile out as its shared-memory
//     image), so a pipeline stage is one contiguous cp.async.bulk (UBLKCP) into a 3-stage mbarrier ring: no tensor maps;
//   * one elected thread issues tcgen05.mma (M = 128, N = 256 / 192, K = 16) into a [128 lanes x 512 columns] fp32
//     accumulator that fills the SM's TMEM; tcgen05.commit releases ring stages and hands the accumulator over;
//   * four epilogue warps (thread = row = TMEM lane) read it back with tcgen05.ld, add the bias, ReLU, round to fp16 and
//     write the next layer's A operand straight into the swizzled tile; after the last layer they compute torch's two-stage
//     fp16 sums from TMEM, normalise, flip the right rows and stage the packed [64, 2, ks, ks] block, which leaves with one
//     bulk store;
//   * the first Linear (K = 3) is computed by the same four warps on the CUDA cores directly into the A tile.
//
// Roles: warps 0-3 epilogue / first layer, warp 4 weight producer, warp 5 MMA issuer + TMEM owner.  Persistent: CTAs loop
// over tiles.  Numerics: identical rounding points to the cuBLAS route (fp32 accumulate, bias added in fp32, one rounding);
// only the fp32 summation order inside a dot product differs, as it does between any two GEMM implementations.
#pragma once

namespace mlpf {
constexpr int TM = 128;                          // rows per tile = TMEM lanes
constexpr int A_KB_BYTES = TM * 128;             // one 64-wide k-block of the activation tile
constexpr int A_BYTES = 8 * A_KB_BYTES;          // K <= 512
constexpr int STAGE_BYTES = 256 * 128;           // one weight tile: <= 256 output rows x 64 k
constexpr int STAGES = 3;
constexpr int MAX_LAYERS = 12;
constexpr int MAX_N1 = 128;
constexpr int OFF_W = A_BYTES;
constexpr int OFF_W1 = OFF_W + STAGES * STAGE_BYTES;
constexpr int OFF_BAR = OFF_W1 + MAX_N1 * 16;
constexpr int SMEM_BYTES = OFF_BAR + 128;
constexpr int THREADS = 192;
static_assert(SMEM_BYTES <= 227 * 1024, "fused MLP tile does not fit in shared memory");

struct Net {
    int n_layers;                 // tensor-core layers (everything after the first Linear)
    int n1;                       // width of the first Linear = K of layer 0 (multiple of 64, <= MAX_N1)
    int K[MAX_LAYERS], N[MAX_LAYERS];        // K multiple of 64 <= 512; N multiple of 16 <= 512 (padded)
    long long w_off[MAX_LAYERS];  // byte offset of the layer's pre-swizzled tiles
    int b_off[MAX_LAYERS];        // float offset of the layer's bias (padded to N)
};

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(unsigned long long *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(unsigned d_tmem, unsigned long long a_desc, unsigned long long b_desc, unsigned idesc, unsigned acc) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
// K-major, SWIZZLE_128B operand: 8-row atoms of 1024 B (stride byte offset), rows of 128 B, descriptor version 1 (sm_100)
__device__ __forceinline__ unsigned long long smem_desc(unsigned addr) {
    return (unsigned long long)((addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16: D = f32, A = B = f16, both K-major, M = 128
__device__ __forceinline__ unsigned instr_desc(int n) { return (1u << 4) | ((unsigned)(n >> 3) << 17) | ((unsigned)(TM >> 4) << 24); }

// 32 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(unsigned taddr, float (&v)[32]) {
    unsigned r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        "tcgen05.wait::ld.sync.aligned;\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// byte offset of the 16-byte chunk holding columns [col, col + 8) of row `row` in the swizzled activation tile
__device__ __forceinline__ unsigned a_chunk_off(int row, int col) {
    const int kb = col >> 6, c = (col & 63) >> 3;
    return (unsigned)(kb * A_KB_BYTES + (row >> 3) * 1024 + (row & 7) * 128 + ((c ^ (row & 7)) << 4));
}

__device__ __forceinline__ unsigned pack_relu_h2(float a, float b) {
    const __half2 h = __floats2half2_rn(fmaxf(a, 0.0f), fmaxf(b, 0.0f));
    return *reinterpret_cast<const unsigned *>(&h);
}
__device__ __forceinline__ float relu_h(float a) { return __half2float(__float2half_rn(fmaxf(a, 0.0f))); }

template <int KS>
__global__ void __launch_bounds__(THREADS, 1)
mlp_fused_pred_kernel(const __grid_constant__ Net net, const unsigned char *__restrict__ wsw, const float *__restrict__ bias,
                      const __half *__restrict__ w1, const __half *__restrict__ b1,
                      const float *__restrict__ xs, const float *__restrict__ ys, const float *__restrict__ z,
                      int H, int W, int b0, int nb, int row0, int nrw, __half *__restrict__ psf) {
    extern __shared__ __align__(1024) unsigned char fm_smem[];
    constexpr int KK = KS * KS;
    unsigned char *sA = fm_smem;
    unsigned char *sW = fm_smem + OFF_W;
    float4 *sW1 = reinterpret_cast<float4 *>(fm_smem + OFF_W1);
    unsigned long long *full = reinterpret_cast<unsigned long long *>(fm_smem + OFF_BAR), *empty = full + STAGES;
    unsigned long long *a_ready = empty + STAGES, *acc_ready = a_ready + 1;
    unsigned *tmem_slot = reinterpret_cast<unsigned *>(acc_ready + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int L = net.n_layers;
    const unsigned n_rows = 2u * (unsigned)nb * (unsigned)nrw * (unsigned)W;
    const unsigned n_tiles = (n_rows + TM - 1) / TM;

    if ((smem_u32(fm_smem) & 1023u) != 0) __trap();         // the swizzled operand atoms need a 1024-byte aligned base
    for (int i = threadIdx.x; i < net.n1; i += blockDim.x)
        sW1[i] = make_float4(__half2float(w1[3 * i]), __half2float(w1[3 * i + 1]), __half2float(w1[3 * i + 2]), __half2float(b1[i]));
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(a_ready, TM);
        mbar_init(acc_ready, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {                                          // this warp owns the TMEM allocation (all 512 columns)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem = *tmem_slot;

    if (warp == 4) {
        // ---- weight producer: one contiguous bulk copy per (layer, half, k-block), in the order the MMA warp consumes them ----
        if (lane == 0) {
            int s = 0;
            unsigned ph = 0;
            for (unsigned tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int l = 0; l < L; ++l) {
                    const int nkb = net.K[l] >> 6, N = net.N[l];
                    const unsigned char *src = wsw + net.w_off[l];
                    for (int h = 0; h * 256 < N; ++h) {
                        const unsigned bytes = (unsigned)min(256, N - h * 256) * 128u;
                        for (int kb = 0; kb < nkb; ++kb) {
                            mbar_wait(empty + s, ph ^ 1u);
                            mbar_expect_tx(full + s, bytes);
                            bulk_g2s(sW + s * STAGE_BYTES, src, bytes, full + s);
                            src += bytes;
                            if (++s == STAGES) { s = 0; ph ^= 1u; }
                        }
                    }
                }
            }
        }
    } else if (warp == 5) {
        // ---- MMA issuer ----------------------------------------------------------------------------------------------------
        if (lane == 0) {
            int s = 0;
            unsigned ph = 0, a_ph = 0;
            const unsigned a_base = smem_u32(sA), w_base = smem_u32(sW);
            for (unsigned tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int l = 0; l < L; ++l) {
                    const int nkb = net.K[l] >> 6, N = net.N[l];
                    mbar_wait(a_ready, a_ph);                 // the A tile of this layer is written and TMEM is drained
                    a_ph ^= 1u;
                    tc_fence_after();
                    for (int h = 0; h * 256 < N; ++h) {
                        const unsigned idesc = instr_desc(min(256, N - h * 256));
                        const unsigned d_tmem = tmem + (unsigned)(h * 256);
                        for (int kb = 0; kb < nkb; ++kb) {
                            mbar_wait(full + s, ph);
                            tc_fence_after();
#pragma unroll
                            for (int k = 0; k < 4; ++k)       // UMMA K = 16 fp16 = 32 bytes inside the 128-byte swizzle row
                                tc_mma_f16(d_tmem, smem_desc(a_base + kb * A_KB_BYTES + k * 32), smem_desc(w_base + s * STAGE_BYTES + k * 32),
                                           idesc, (unsigned)((kb | k) != 0));
                            tc_commit(empty + s);             // the stage is free once these MMAs have read it
                            if (++s == STAGES) { s = 0; ph ^= 1u; }
                        }
                    }
                    tc_commit(acc_ready);                     // every MMA of the layer has completed
                }
            }
        }
    } else {
        // ---- epilogue warps: thread t = row t of the tile = TMEM lane t --------------------------------------------------------
        const int t = threadIdx.x;
        const unsigned lane_addr = tmem + ((unsigned)(warp * 32) << 16);
        unsigned acc_ph = 0;
        for (unsigned tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            // first Linear (K = 3) + ReLU for this row, straight into the swizzled A tile (mlp_input_layer_kernel's arithmetic)
            {
                const unsigned r = min(tile * TM + (unsigned)t, n_rows - 1);
                const unsigned p = r >> 1, side = r & 1u;
                const unsigned q = p / (unsigned)W, x = p - q * (unsigned)W;
                const unsigned bq = q / (unsigned)nrw;
                const int y = row0 + (int)(q - bq * (unsigned)nrw), b = b0 + (int)bq;
                const float xr = __ldg(xs + x);
                const float xv = __half2float(__float2half_rn(side ? -xr : xr));
                const float yv = __half2float(__float2half_rn(__ldg(ys + y)));
                const float zv = __half2float(__float2half_rn(__ldg(z + ((int64_t)b * H + y) * W + x)));
                for (int c = 0; c < net.n1; c += 8) {
                    unsigned o[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float4 wa = sW1[c + 2 * k], wb = sW1[c + 2 * k + 1];
                        const float va = fmaf(zv, wa.z, fmaf(yv, wa.y, xv * wa.x)) + wa.w;
                        const float vb = fmaf(zv, wb.z, fmaf(yv, wb.y, xv * wb.x)) + wb.w;
                        o[k] = pack_relu_h2(va, vb);
                    }
                    *reinterpret_cast<uint4 *>(sA + a_chunk_off(t, c)) = make_uint4(o[0], o[1], o[2], o[3]);
                }
            }
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(a_ready);
            for (int l = 0; l + 1 < L; ++l) {
                // hidden layer: accumulator -> + bias -> ReLU -> fp16 -> the next layer's A operand
                const int N = net.N[l];
                const float4 *bp = reinterpret_cast<const float4 *>(bias + net.b_off[l]);
                mbar_wait(acc_ready, acc_ph);
                acc_ph ^= 1u;
                tc_fence_after();
                for (int j = 0; j < N; j += 32) {
                    float v[32];
                    tmem_ld32(lane_addr + (unsigned)j, v);
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const float4 ba = __ldg(bp + (j >> 2) + 2 * g), bb = __ldg(bp + (j >> 2) + 2 * g + 1);
                        const uint4 o = make_uint4(pack_relu_h2(v[8 * g] + ba.x, v[8 * g + 1] + ba.y), pack_relu_h2(v[8 * g + 2] + ba.z, v[8 * g + 3] + ba.w),
                                                   pack_relu_h2(v[8 * g + 4] + bb.x, v[8 * g + 5] + bb.y), pack_relu_h2(v[8 * g + 6] + bb.z, v[8 * g + 7] + bb.w));
                        *reinterpret_cast<uint4 *>(sA + a_chunk_off(t, j + 8 * g)) = o;
                    }
                }
                fence_proxy_async();
                tc_fence_before();
                mbar_arrive(a_ready);
            }
            // last layer: torch's fp16 sums (sum(-1) rounds every kernel row, the second sum(-1) rounds the total), then the
            // normalised kernels, the right rows flipped along their last axis, staged as the packed [64, 2, KS, KS] block
            {
                const float *bl = bias + net.b_off[L - 1];
                mbar_wait(acc_ready, acc_ph);
                acc_ph ^= 1u;
                tc_fence_after();
                float tot = 0.0f, rowsum = 0.0f;
#pragma unroll
                for (int j = 0; j < KK; j += 32) {
                    float v[32];
                    tmem_ld32(lane_addr + (unsigned)j, v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int col = j + i;
                        if (col < KK) {
                            rowsum += relu_h(v[i] + __ldg(bl + col));
                            if ((col + 1) % KS == 0) { tot += __half2float(__float2half_rn(rowsum)); rowsum = 0.0f; }
                        }
                    }
                }
                float den = __half2float(__float2half_rn(tot));
                den = __half2float(__float2half_rn(den + 1e-9f));
                const bool ok = den > 0.0f && den <= 65504.0f;
                float rc = rcp_approx(ok ? den : 1.0f);
                rc = fmaf(rc, fmaf(-den, rc, 1.0f), rc);
                __half *dst = reinterpret_cast<__half *>(sA) + (size_t)t * KK;
                const bool flip = t & 1;                      // tiles start at an even row: odd rows are right kernels
#pragma unroll
                for (int j = 0; j < KK; j += 32) {
                    float v[32];
                    tmem_ld32(lane_addr + (unsigned)j, v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int col = j + i;
                        if (col < KK) {
                            const float a = relu_h(v[i] + __ldg(bl + col));
                            const float qv = a * rc;
                            const float quo = fmaf(rc, fmaf(-den, qv, a), qv);        // IEEE a / den (div_rn's sequence)
                            const int u = col / KS, vv = col - u * KS;
                            dst[flip ? u * KS + (KS - 1 - vv) : col] = __float2half_rn(ok ? quo : 0.0f);
                        }
                    }
                }
                tc_fence_before();
                fence_proxy_async();
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (t == 0) {
                    const unsigned rows_here = min((unsigned)TM, n_rows - tile * TM);
                    const unsigned bytes = rows_here * (unsigned)(KK * 2);
                    __half *g = psf + (size_t)tile * TM * KK;
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(smem_u32(sA)), "r"(bytes) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

// Weights of one layer [n_true, K] fp16 row-major -> the layer's pre-swizzled tiles: for each half of <= 256 output rows and
// each 64-wide k-block, rows of 128 bytes whose 16-byte chunk c sits at position c ^ (row & 7) (Swizzle<3,4,3>, the
// shared-memory image of a K-major SWIZZLE_128B operand).  Rows >= n_true (padding) are zero.  One thread per chunk.
__global__ void __launch_bounds__(256)
swizzle_weights_kernel(const __half *__restrict__ w, int n_true, int n_pad, int K, unsigned char *__restrict__ out) {
    const int nkb = K >> 6;
    const int64_t chunks = (int64_t)n_pad * nkb * 8;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < chunks; i += (int64_t)gridDim.x * blockDim.x) {
        const int pch = (int)(i & 7);                         // chunk position inside the 128-byte row
        int64_t rest = i >> 3;                                // (half, kb, r) with r fastest
        const int rows0 = min(256, n_pad), rows1 = n_pad - rows0;
        const int64_t half0 = (int64_t)rows0 * nkb;
        int h, kb, r;
        if (rest < half0) { h = 0; kb = (int)(rest / rows0); r = (int)(rest - (int64_t)kb * rows0); }
        else { rest -= half0; h = 1; kb = (int)(rest / rows1); r = (int)(rest - (int64_t)kb * rows1); }
        const int n = h * 256 + r, c = pch ^ (r & 7);
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (n < n_true) v = *reinterpret_cast<const uint4 *>(w + (int64_t)n * K + kb * 64 + c * 8);
        *reinterpret_cast<uint4 *>(out + i * 16) = v;
    }
}

__global__ void pad_bias_kernel(const __half *__restrict__ b, int n_true, int n_pad, float *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_pad) out[i] = i < n_true ? __half2float(b[i]) : 0.0f;
}
}  // namespace mlpf

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
constexpr int MAX_LAYERS = 12;
constexpr int MAX_N1 = 128;
constexpr int EPI_THREADS = 256, PROD_WARP = 8, MMA_WARP = 9;
constexpr int THREADS = 320;
// NCTA = 1: one CTA per tile, a weight tile [<= 256 rows x 64 k] per ring stage, biases read from global memory.
// NCTA = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2, UMMA M = 256) works on two tiles at once; each CTA stages only ITS
// half of every weight tile (the tensor cores of both SMs read both halves), which halves the L2 -> SM weight traffic -- the
// bound of the 1-CTA kernel: 512 KB per tile-layer per SM against ~42 B/clk/SM of L2 throughput is 12.3 k cycles for 8.2 k
// cycles of MMA -- and frees shared memory for a fourth stage and for all the biases.
template <int NCTA>
struct Cfg {
    static constexpr int STAGE_BYTES = 256 * 128 / NCTA;
    static constexpr int STAGES = NCTA == 1 ? 3 : 4;
    static constexpr int BIAS_FLOATS = NCTA == 1 ? 0 : 11 * 512;     // capacity of the resident bias table
    static constexpr int OFF_W = A_BYTES;
    static constexpr int OFF_BIAS = OFF_W + STAGES * STAGE_BYTES;
    static constexpr int OFF_W1 = OFF_BIAS + BIAS_FLOATS * 4;
    static constexpr int OFF_BAR = OFF_W1 + MAX_N1 * 16;
    static constexpr int SMEM_BYTES = OFF_BAR + 256;                  // barriers, TMEM slot
    static_assert(SMEM_BYTES <= 227 * 1024, "fused MLP tile does not fit in shared memory");
};

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
// tcgen05.commit: the mbarrier is signalled once every MMA issued so far by this thread has completed; NCTA = 2: the barrier
// at the same offset in BOTH CTAs of the pair
template <int NCTA>
__device__ __forceinline__ void tc_commit(unsigned long long *bar) {
    if constexpr (NCTA == 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    else
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(bar)), "h"((unsigned short)3) : "memory");
}
template <int NCTA>
__device__ __forceinline__ void tc_mma_f16(unsigned d_tmem, unsigned long long a_desc, unsigned long long b_desc, unsigned idesc, unsigned acc) {
    if constexpr (NCTA == 1)
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
            ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}"
            ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
// CTA pair plumbing
__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `cta` of the cluster.  Default (CTA-scope) semantics on
// purpose: the cluster-scope release / acquire forms compile to MEMBAR.ALL.GPU and CCTL.IVALL around every hand-over, and
// what is handed over here is read by the tensor cores through the async proxy (the writers have fenced it already).
__device__ __forceinline__ void mbar_arrive_remote(unsigned long long *bar, unsigned cta) {
    asm volatile("{\n.reg .b32 ra;\nmapa.shared::cluster.u32 ra, %0, %1;\n"
                 "mbarrier.arrive.shared::cluster.b64 _, [ra];\n}" ::"r"(smem_u32(bar)), "r"(cta) : "memory");
}
// K-major, SWIZZLE_128B operand: 8-row atoms of 1024 B (stride byte offset), rows of 128 B, descriptor version 1 (sm_100)
__device__ __forceinline__ unsigned long long smem_desc(unsigned addr) {
    return (unsigned long long)((addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16: D = f32, A = B = f16, both K-major; m = 128 (one CTA) or 256 (CTA pair: 128 rows from each)
__device__ __forceinline__ unsigned instr_desc(int n, int m) { return (1u << 4) | ((unsigned)(n >> 3) << 17) | ((unsigned)(m >> 4) << 24); }

// 32 consecutive fp32 columns of this thread's TMEM lane: the load is asynchronous, the registers may be read only after
// tmem_wait32 (which names them as operands so that the compiler keeps every use behind the wait).
#define TM_R(i) "=r"(r[i])
#define TM_RW(i) "+r"(r[i])
__device__ __forceinline__ void tmem_ld32_issue(unsigned taddr, unsigned (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : TM_R(0), TM_R(1), TM_R(2), TM_R(3), TM_R(4), TM_R(5), TM_R(6), TM_R(7), TM_R(8), TM_R(9), TM_R(10), TM_R(11), TM_R(12),
          TM_R(13), TM_R(14), TM_R(15), TM_R(16), TM_R(17), TM_R(18), TM_R(19), TM_R(20), TM_R(21), TM_R(22), TM_R(23), TM_R(24),
          TM_R(25), TM_R(26), TM_R(27), TM_R(28), TM_R(29), TM_R(30), TM_R(31)
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait32(unsigned (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;\n"
        : TM_RW(0), TM_RW(1), TM_RW(2), TM_RW(3), TM_RW(4), TM_RW(5), TM_RW(6), TM_RW(7), TM_RW(8), TM_RW(9), TM_RW(10), TM_RW(11),
          TM_RW(12), TM_RW(13), TM_RW(14), TM_RW(15), TM_RW(16), TM_RW(17), TM_RW(18), TM_RW(19), TM_RW(20), TM_RW(21), TM_RW(22),
          TM_RW(23), TM_RW(24), TM_RW(25), TM_RW(26), TM_RW(27), TM_RW(28), TM_RW(29), TM_RW(30), TM_RW(31)
        :: "memory");
}
__device__ __forceinline__ void tmem_ld16_issue(unsigned taddr, unsigned (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : TM_R(0), TM_R(1), TM_R(2), TM_R(3), TM_R(4), TM_R(5), TM_R(6), TM_R(7), TM_R(8), TM_R(9), TM_R(10), TM_R(11), TM_R(12),
          TM_R(13), TM_R(14), TM_R(15)
        : "r"(taddr) : "memory");
}
// The same load, ordered after every use of `dep` that precedes it and before every use that follows it in program order:
// naming `dep` as in/out operands keeps the compiler from sinking the issue below the arithmetic on `dep`, which would
// serialise "load, wait, compute" again.
__device__ __forceinline__ void tmem_ld16_issue_after(unsigned taddr, unsigned (&r)[16], unsigned (&dep)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%32];\n"
        : TM_R(0), TM_R(1), TM_R(2), TM_R(3), TM_R(4), TM_R(5), TM_R(6), TM_R(7), TM_R(8), TM_R(9), TM_R(10), TM_R(11), TM_R(12),
          TM_R(13), TM_R(14), TM_R(15),
          "+r"(dep[0]), "+r"(dep[1]), "+r"(dep[2]), "+r"(dep[3]), "+r"(dep[4]), "+r"(dep[5]), "+r"(dep[6]), "+r"(dep[7]),
          "+r"(dep[8]), "+r"(dep[9]), "+r"(dep[10]), "+r"(dep[11]), "+r"(dep[12]), "+r"(dep[13]), "+r"(dep[14]), "+r"(dep[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait16(unsigned (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;\n"
        : TM_RW(0), TM_RW(1), TM_RW(2), TM_RW(3), TM_RW(4), TM_RW(5), TM_RW(6), TM_RW(7), TM_RW(8), TM_RW(9), TM_RW(10), TM_RW(11),
          TM_RW(12), TM_RW(13), TM_RW(14), TM_RW(15)
        :: "memory");
}
#undef TM_R
#undef TM_RW

// 16-byte read-only load that stays where it is written (a plain __ldg may be hoisted across the asm statements around it; in
// the fully unrolled last layer that piles hundreds of registers up)
__device__ __forceinline__ float4 ldg_f4_here(const float4 *p) {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

__device__ __forceinline__ float4 lds_f4_here(const float4 *p) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
    return v;
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

// 32 accumulator columns + bias -> ReLU -> fp16, packed two per register: one packed fp32 add (FADD2), one conversion (F2FP)
// and one packed maximum (HMNMX2) per pair of columns; relu(round(x)) = round(relu(x)) since rounding is monotonic
__device__ __forceinline__ unsigned add_relu_h2(unsigned a, unsigned b, float ba, float bb) {
    const float2 v = __fadd2_rn(make_float2(__uint_as_float(a), __uint_as_float(b)), make_float2(ba, bb));
    const __half2 h = __hmax2(__floats2half2_rn(v.x, v.y), __float2half2_rn(0.0f));
    return *reinterpret_cast<const unsigned *>(&h);
}
__device__ __forceinline__ void bias_relu_pack(const unsigned (&r)[32], const float4 (&bv)[8], unsigned (&o)[16]) {
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        o[2 * g] = add_relu_h2(r[4 * g], r[4 * g + 1], bv[g].x, bv[g].y);
        o[2 * g + 1] = add_relu_h2(r[4 * g + 2], r[4 * g + 3], bv[g].z, bv[g].w);
    }
}
__device__ __forceinline__ void store_a_chunks(unsigned char *sA, int row, int col, const unsigned (&o)[16]) {
#pragma unroll
    for (int g = 0; g < 4; ++g)
        *reinterpret_cast<uint4 *>(sA + a_chunk_off(row, col + 8 * g)) = make_uint4(o[4 * g], o[4 * g + 1], o[4 * g + 2], o[4 * g + 3]);
}

// Last layer, one 16-column piece [col0, col0 + 16) of this thread's half row of the accumulator: relu(acc + bias) rounded to
// fp16.  PASS 0 adds it to the kernel-row sums, PASS 1 (the accumulator is read a second time: TMEM loads are cheap, a
// hundred live registers are not) divides and stages the result at its final position; RIGHT: the warp's rows are right
// kernels (rows 64..127 of a tile), whose columns are flipped inside every kernel row.  The position of the piece is a
// run-time value on purpose: unrolling the row over compile-time columns (every address an immediate) made 12 k instructions
// of straight-line code per kernel, four times the instruction cache, and the tail ran at the speed of instruction fetch.
struct TailState {
    float part, rs, rs_x, den, rc;      // total of finished kernel rows, running row, the straddling row's piece; denominator
    __half *dst;                        // this row of the staged block
    const float *bl;                    // last layer's bias (SB: in shared memory)
    unsigned lane_addr;
    int c_lo, c_end;                    // this thread's live columns
};
template <int KS>
__device__ __forceinline__ int wrap_ks(int w) {              // w mod KS for w < KS + 16
#pragma unroll
    for (int k = 0; k < (16 + KS - 1) / KS; ++k) w = w >= KS ? w - KS : w;
    return w;
}
template <int KS, int PASS, int RIGHT, bool SB>
__device__ __forceinline__ void tail_piece(const int col0, const int w0 /* col0 mod KS */, const unsigned (&cur)[16], TailState &st) {
    constexpr int KK = KS * KS, KC = (KK + 31) / 32, HC = (KC + 1) / 2, SPLIT = 32 * HC;
    constexpr int U_X = (SPLIT < KK) ? SPLIT / KS : -1;
    constexpr bool STRADDLE = U_X >= 0 && (SPLIT % KS) != 0;
    float4 bv[4];                                             // this piece's biases (the padded bias row covers every piece)
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const float4 *bp = reinterpret_cast<const float4 *>(st.bl + col0) + g;
        bv[g] = SB ? lds_f4_here(bp) : ldg_f4_here(bp);
    }
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
        const int c0 = col0 + i;                              // even: col0 is a multiple of 16
        const bool live0 = c0 < st.c_end, live1 = c0 + 1 < st.c_end;
        const float b0 = (i & 2) ? bv[i >> 2].z : bv[i >> 2].x, b1 = (i & 2) ? bv[i >> 2].w : bv[i >> 2].y;
        // relu(round_fp16(acc + bias)) for two columns at once
        const float2 s2 = __fadd2_rn(make_float2(__uint_as_float(cur[i]), __uint_as_float(cur[i + 1])), make_float2(b0, b1));
        const __half2 h = __hmax2(__floats2half2_rn(s2.x, s2.y), __float2half2_rn(0.0f));
        const float v0 = live0 ? __low2float(h) : 0.0f, v1 = live1 ? __high2float(h) : 0.0f;
        const int wa = wrap_ks<KS>(w0 + i), wb = wrap_ks<KS>(w0 + i + 1);       // positions inside their kernel rows
        if (PASS == 0) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int c = c0 + k;
                st.rs += k ? v1 : v0;
                // end of a kernel row (or of this thread's piece of the straddling row); warp-uniform
                if ((k ? live1 : live0) && ((k ? wb : wa) == KS - 1 || c + 1 == st.c_end)) {
                    if (STRADDLE && c >= U_X * KS && c < (U_X + 1) * KS) st.rs_x = st.rs;
                    else st.part += __half2float(__float2half_rn(st.rs));
                    st.rs = 0.0f;
                }
            }
        } else {
            // IEEE v / den from the refined reciprocal (div_rn's sequence); den = rc = 0 for an all-zero kernel -> 0
            const float2 rc2 = make_float2(st.rc, st.rc), v2 = make_float2(v0, v1);
            const float2 q = __fmul2_rn(v2, rc2);
            const float2 r = __ffma2_rn(rc2, __ffma2_rn(make_float2(-st.den, -st.den), q, v2), q);
            if (!RIGHT) {                                     // left rows start 4-byte aligned: one 32-bit store per pair
                if (live1) *reinterpret_cast<__half2 *>(st.dst + c0) = __floats2half2_rn(r.x, r.y);
                else if (live0) st.dst[c0] = __float2half_rn(r.x);
            } else {                                          // column w of a kernel row goes to KS - 1 - w
                if (live0) st.dst[c0 + (KS - 1) - 2 * wa] = __float2half_rn(r.x);
                if (live1) st.dst[c0 + 1 + (KS - 1) - 2 * wb] = __float2half_rn(r.y);
            }
        }
    }
}
// One pass over this thread's half row: 2 * HC pieces, the next one in flight while the current one is used (the two
// register arrays alternate, so an iteration handles two pieces); the loop is NOT unrolled (see tail_piece).
template <int KS, int PASS, int RIGHT, bool SB>
__device__ __forceinline__ void tail_pass(TailState &st) {
    constexpr int KK = KS * KS, KC = (KK + 31) / 32, HC = (KC + 1) / 2;
    unsigned ra[16], rb[16];
    int w = st.c_lo % KS;
    tmem_ld16_issue(st.lane_addr + (unsigned)st.c_lo, ra);
#pragma unroll 1
    for (int pc = 0; pc < HC; ++pc) {
        const int col0 = st.c_lo + 32 * pc;
        tmem_wait16(ra);
        tmem_ld16_issue_after(st.lane_addr + (unsigned)(col0 + 16), rb, ra);
        tail_piece<KS, PASS, RIGHT, SB>(col0, w, ra, st);
        w = wrap_ks<KS>(w + 16);
        tmem_wait16(rb);
        if (pc + 1 < HC) tmem_ld16_issue_after(st.lane_addr + (unsigned)(col0 + 32), ra, rb);
        tail_piece<KS, PASS, RIGHT, SB>(col0 + 16, w, rb, st);
        w = wrap_ks<KS>(w + 16);
    }
}

// Warp roles: warps 0-7 epilogue (thread = row (warp & 3) * 32 + lane = TMEM lane; the two warps of a lane quarter split the
// columns), warp 8 weight producer, warp 9 MMA issuer + TMEM owner (NCTA = 2: the leader CTA's warp 9 issues for the pair, the
// other CTA's warp 9 relays its CTA's "operand ready" barriers to the leader).
//
// Hand-over of the activation tile between layers (a_lo / a_hi): the epilogue drains accumulator half 0 (output columns
// 0..255 = k-blocks 0..3 of the next layer) into registers while the MMAs of half 1 still read the A tile; when the layer is
// complete it stores them and signals a_lo, and the next layer's MMAs over k-blocks 0..3 run while the epilogue converts half
// 1 (k-blocks 4..7, signalled by a_hi).
template <int KS, int NCTA>
__global__ void __launch_bounds__(THREADS, 1)
mlp_fused_pred_kernel(const __grid_constant__ Net net, const unsigned char *__restrict__ wsw, const float *__restrict__ bias, int bias_floats,
                      const __half *__restrict__ w1, const __half *__restrict__ b1,
                      const float *__restrict__ xs, const float *__restrict__ ys, const float *__restrict__ z,
                      int H, int W, int b0, int nb, int row0, int nrw, __half *__restrict__ psf, long long *__restrict__ dbg) {
    extern __shared__ __align__(1024) unsigned char fm_smem[];
    using CF = Cfg<NCTA>;
    constexpr int STAGES = CF::STAGES, STAGE_BYTES = CF::STAGE_BYTES;
    constexpr bool SB = NCTA == 2;                           // biases resident in shared memory
    constexpr int KK = KS * KS;
    constexpr int KC = (KK + 31) / 32;                       // 32-column chunks of the last layer
    unsigned char *sA = fm_smem;
    unsigned char *sW = fm_smem + CF::OFF_W;
    float *sBias = reinterpret_cast<float *>(fm_smem + CF::OFF_BIAS);
    float4 *sW1 = reinterpret_cast<float4 *>(fm_smem + CF::OFF_W1);
    unsigned long long *full = reinterpret_cast<unsigned long long *>(fm_smem + CF::OFF_BAR), *empty = full + STAGES;
    unsigned long long *a_lo = empty + STAGES, *a_hi = a_lo + 1, *acc_ready = a_hi + 1;     // acc_ready[2]: one per accumulator half
    unsigned *tmem_slot = reinterpret_cast<unsigned *>(acc_ready + 2);
    float *s_part = reinterpret_cast<float *>(sA + TM * KK * 2);                     // [4][TM] partial sums of the last layer, behind the packed block
    static_assert(TM * KK * 2 + 4 * TM * 4 <= A_BYTES, "the packed block and its partial sums must fit in the activation tile");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int L = net.n_layers;
    const unsigned n_rows = 2u * (unsigned)nb * (unsigned)nrw * (unsigned)W;
    const unsigned n_tiles = (n_rows + TM - 1) / TM;
    const unsigned rank = NCTA == 2 ? cluster_ctarank() : 0u;
    // a CTA (pair) takes tile group g = its index, g + number of CTAs (pairs), ...; the CTA of rank r works on tile NCTA * g + r
    // (a pair's second tile may lie past the end: it is computed on clamped coordinates and not stored)
    const unsigned n_groups = (n_tiles + NCTA - 1) / NCTA, g0 = blockIdx.x / NCTA, g_step = gridDim.x / NCTA;

    if ((smem_u32(fm_smem) & 1023u) != 0) __trap();         // the swizzled operand atoms need a 1024-byte aligned base
    for (int i = threadIdx.x; i < net.n1; i += blockDim.x)
        sW1[i] = make_float4(__half2float(w1[3 * i]), __half2float(w1[3 * i + 1]), __half2float(w1[3 * i + 2]), __half2float(b1[i]));
    if (SB) for (int i = threadIdx.x; i < bias_floats; i += blockDim.x) sBias[i] = bias[i];
    if (threadIdx.x == 0) {
        const unsigned peer = (NCTA == 2 && rank == 0) ? 1u : 0u;      // the leader's barriers also count the other CTA's relay
        for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1 + peer); mbar_init(empty + s, 1); }
        mbar_init(a_lo, EPI_THREADS + peer);
        mbar_init(a_hi, EPI_THREADS + peer);
        mbar_init(acc_ready, 1);
        mbar_init(acc_ready + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == MMA_WARP) {                                   // this warp owns the TMEM allocation (all 512 columns)
        if constexpr (NCTA == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    if constexpr (NCTA == 2) { __syncthreads(); cluster_sync_all(); } else __syncthreads();
    tc_fence_after();
    const unsigned tmem = *tmem_slot;

    if (warp == PROD_WARP) {
        // ---- weight producer: one contiguous bulk copy per (layer, half, k-block), in the order the MMA warp consumes them; a
        // CTA of a pair copies rows [rank * rows / 2, (rank + 1) * rows / 2) of the tile ----
        if (lane == 0) {
            int s = 0;
            unsigned ph = 0;
            for (unsigned g = g0; g < n_groups; g += g_step) {
                for (int l = 0; l < L; ++l) {
                    const int nkb = net.K[l] >> 6, N = net.N[l];
                    const unsigned char *src = wsw + net.w_off[l];
                    for (int h = 0; h * 256 < N; ++h) {
                        const unsigned bytes = (unsigned)min(256, N - h * 256) * 128u, mine = bytes / NCTA;
                        for (int kb = 0; kb < nkb; ++kb) {
                            mbar_wait(empty + s, ph ^ 1u);
                            mbar_expect_tx(full + s, mine);
                            bulk_g2s(sW + s * STAGE_BYTES, src + rank * mine, mine, full + s);
                            src += bytes;
                            if (++s == STAGES) { s = 0; ph ^= 1u; }
                        }
                    }
                }
            }
        }
    } else if (warp == MMA_WARP) {
        if (lane == 0 && rank == 0) {
            // ---- MMA issuer (NCTA = 2: for both CTAs of the pair) ---------------------------------------------------------------
            int s = 0;
            unsigned ph = 0, a_ph = 0;
            const unsigned a_base = smem_u32(sA), w_base = smem_u32(sW);
            long long t_start = clock64(), t_full = 0, t_aready = 0, t0;
            for (unsigned g = g0; g < n_groups; g += g_step) {
                for (int l = 0; l < L; ++l) {
                    const int nkb = net.K[l] >> 6, N = net.N[l], kb_lo = min(4, nkb);
                    const bool tl = dbg && blockIdx.x == 0 && g == g0 + g_step;    // timeline of this CTA's second tile group
                    long long *tlp = dbg + 148 * 8 + l * 16;
                    t0 = clock64();
                    // k-blocks [0, kb_lo) of this layer's A tile are written and accumulator half 0 is drained
                    mbar_wait(a_lo, a_ph);
                    if (kb_lo == nkb) mbar_wait(a_hi, a_ph);
                    t_aready += clock64() - t0;
                    if (tl) { tlp[0] = t0; tlp[1] = clock64(); }
                    tc_fence_after();
                    for (int h = 0; h < 2; ++h) {
                        if (h * 256 < N) {
                            const unsigned idesc = instr_desc(min(256, N - h * 256), TM * NCTA);
                            const unsigned d_tmem = tmem + (unsigned)(h * 256);
                            for (int kb = 0; kb < nkb; ++kb) {
                                if (h == 0 && kb == kb_lo) {  // the rest of the A tile, and accumulator half 1 drained
                                    t0 = clock64();
                                    mbar_wait(a_hi, a_ph);
                                    t_aready += clock64() - t0;
                                    if (tl) { tlp[2] = t0; tlp[3] = clock64(); }
                                    tc_fence_after();
                                }
                                t0 = clock64();
                                mbar_wait(full + s, ph);
                                t_full += clock64() - t0;
                                tc_fence_after();
#pragma unroll
                                for (int k = 0; k < 4; ++k)   // UMMA K = 16 fp16 = 32 bytes inside the 128-byte swizzle row
                                    tc_mma_f16<NCTA>(d_tmem, smem_desc(a_base + kb * A_KB_BYTES + k * 32), smem_desc(w_base + s * STAGE_BYTES + k * 32),
                                                     idesc, (unsigned)((kb | k) != 0));
                                tc_commit<NCTA>(empty + s);   // the stage is free once these MMAs have read it
                                if (++s == STAGES) { s = 0; ph ^= 1u; }
                            }
                        }
                        tc_commit<NCTA>(acc_ready + h);       // this half of the accumulator is complete (h = 1: the whole layer)
                        if (tl) tlp[4 + h] = clock64();
                    }
                    a_ph ^= 1u;
                }
            }
            if (dbg) { dbg[blockIdx.x * 8 + 0] = clock64() - t_start; dbg[blockIdx.x * 8 + 1] = t_full; dbg[blockIdx.x * 8 + 2] = t_aready; }
        } else if (NCTA == 2 && lane == 0) {
            // ---- relay (second CTA of a pair): pass this CTA's "A tile written" and "weights landed" on to the leader's barriers,
            // in the order the leader waits for them ----
            int s = 0;
            unsigned ph = 0, a_ph = 0;
            for (unsigned g = g0; g < n_groups; g += g_step) {
                for (int l = 0; l < L; ++l) {
                    const int nkb = net.K[l] >> 6, N = net.N[l], kb_lo = min(4, nkb);
                    mbar_wait(a_lo, a_ph);
                    mbar_arrive_remote(a_lo, 0);
                    if (kb_lo == nkb) { mbar_wait(a_hi, a_ph); mbar_arrive_remote(a_hi, 0); }
                    for (int h = 0; h * 256 < N; ++h) {
                        for (int kb = 0; kb < nkb; ++kb) {
                            if (h == 0 && kb == kb_lo) { mbar_wait(a_hi, a_ph); mbar_arrive_remote(a_hi, 0); }
                            mbar_wait(full + s, ph);
                            mbar_arrive_remote(full + s, 0);
                            if (++s == STAGES) { s = 0; ph ^= 1u; }
                        }
                    }
                    a_ph ^= 1u;
                }
            }
        }
    } else {
        // ---- epilogue warps --------------------------------------------------------------------------------------------------
        const int t = (warp & 3) * 32 + lane;                 // row of the tile = TMEM lane
        const int grp = warp >> 2;                            // which of the two warps of this lane quarter
        const unsigned lane_addr = tmem + ((unsigned)((warp & 3) * 32) << 16);
        unsigned acc_ph = 0;
        long long e_start = clock64(), e_wait = 0, e_last = 0, e_p0 = 0, e_st = 0, e0, e1;
        for (unsigned g = g0; g < n_groups; g += g_step) {
            const unsigned tile = NCTA * g + rank;
            // first Linear (K = 3) + ReLU for this row, straight into the swizzled A tile (mlp_input_layer_kernel's arithmetic);
            // the two warps of a row take alternate groups of 8 columns
            {
                // tile row t: pixel (t & 63) of the tile's 64, left kernel for t < 64, right kernel for t >= 64 (a warp is
                // all-left or all-right: the flip of the right kernels is then a compile-time address pattern)
                const unsigned side = (unsigned)t >> 6;
                const unsigned p = min(tile * (TM / 2) + ((unsigned)t & 63u), (n_rows >> 1) - 1);
                const unsigned q = p / (unsigned)W, x = p - q * (unsigned)W;
                const unsigned bq = q / (unsigned)nrw;
                const int y = row0 + (int)(q - bq * (unsigned)nrw), b = b0 + (int)bq;
                const float xr = __ldg(xs + x);
                const float xv = __half2float(__float2half_rn(side ? -xr : xr));
                const float yv = __half2float(__float2half_rn(__ldg(ys + y)));
                const float zv = __half2float(__float2half_rn(__ldg(z + ((int64_t)b * H + y) * W + x)));
                for (int c = 8 * grp; c < net.n1; c += 16) {
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
            mbar_arrive(a_lo);
            mbar_arrive(a_hi);
            for (int l = 0; l + 1 < L; ++l) {
                // hidden layer: accumulator -> + bias -> ReLU -> fp16 -> the next layer's A operand.  A warp takes every second
                // 32-column chunk of a half.
                const int N = net.N[l];
                const float4 *bp = reinterpret_cast<const float4 *>((SB ? sBias : bias) + net.b_off[l]);
                const int nc0 = min(256, N) >> 5, nc1 = max(N - 256, 0) >> 5;       // chunks in each half
                unsigned held[4][16];
                const bool tl = dbg && blockIdx.x == 0 && threadIdx.x == 0 && g == g0 + g_step;
                long long *tlp = dbg + 148 * 8 + l * 16 + 8;
                e0 = clock64();
                mbar_wait(acc_ready, acc_ph);
                e_wait += clock64() - e0;
                if (tl) { tlp[0] = e0; tlp[1] = clock64(); }
                tc_fence_after();
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int c = 2 * i + grp;
                    if (c < nc0) {
                        unsigned r[32];
                        float4 bv[8];
                        tmem_ld32_issue(lane_addr + (unsigned)(32 * c), r);
#pragma unroll
                        for (int g4 = 0; g4 < 8; ++g4) bv[g4] = SB ? bp[8 * c + g4] : __ldg(bp + 8 * c + g4);
                        tmem_wait32(r);
                        bias_relu_pack(r, bv, held[i]);
                    }
                }
                e0 = clock64();
                mbar_wait(acc_ready + 1, acc_ph);
                e_wait += clock64() - e0;
                if (tl) { tlp[2] = e0; tlp[3] = clock64(); }
                acc_ph ^= 1u;
                tc_fence_after();
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int c = 2 * i + grp;
                    if (c < nc0) store_a_chunks(sA, t, 32 * c, held[i]);
                }
                fence_proxy_async();
                tc_fence_before();
                mbar_arrive(a_lo);
                if (tl) tlp[4] = clock64();
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int c = 2 * i + grp;
                    if (c < nc1) {
                        unsigned r[32], o[16];
                        float4 bv[8];
                        tmem_ld32_issue(lane_addr + (unsigned)(256 + 32 * c), r);
#pragma unroll
                        for (int g4 = 0; g4 < 8; ++g4) bv[g4] = SB ? bp[64 + 8 * c + g4] : __ldg(bp + 64 + 8 * c + g4);
                        tmem_wait32(r);
                        bias_relu_pack(r, bv, o);
                        store_a_chunks(sA, t, 256 + 32 * c, o);
                    }
                }
                fence_proxy_async();
                tc_fence_before();
                mbar_arrive(a_hi);
                if (tl) tlp[5] = clock64();
            }
            // last layer: torch's fp16 sums (sum(-1) rounds every kernel row, the second sum(-1) rounds the total), then the
            // normalised kernels, the right rows flipped along their last axis, staged as the packed [64, 2, KS, KS] block.
            // The two warps of a row split the columns in the middle ([0, SPLIT) and [SPLIT, 32 * KC)); a thread keeps its
            // half of the row in registers (fp16 pairs) from the accumulator to the staged block.  Kernel-row sums are taken in
            // column order; the one kernel row that straddles SPLIT is merged from two partial sums, and the totals of the two
            // halves are added through shared memory (fp32 sums of a few fp16 values: exact, whatever the order).
            {
                constexpr int HC = (KC + 1) / 2;              // 32-column chunks per thread
                constexpr int SPLIT = 32 * HC;                // first column of the second warp
                constexpr int U_X = (SPLIT < KK) ? SPLIT / KS : -1;                 // kernel row that straddles SPLIT ...
                constexpr bool STRADDLE = U_X >= 0 && (SPLIT % KS) != 0;            // ... unless SPLIT is a row boundary
                const float *bl = (SB ? sBias : bias) + net.b_off[L - 1];
                e0 = clock64();
                mbar_wait(acc_ready, acc_ph);
                mbar_wait(acc_ready + 1, acc_ph);
                acc_ph ^= 1u;
                e_wait += clock64() - e0;
                e0 = clock64();
                tc_fence_after();
                const int c_lo = grp ? SPLIT : 0;             // this thread's columns: [c_lo, c_lo + 32 * HC), live below c_end
                const int c_end = grp ? KK : (SPLIT < KK ? SPLIT : KK);
                const int right = t >> 6;                     // warp-uniform: rows 64..127 are right kernels
                __half *dst = reinterpret_cast<__half *>(sA) + (size_t)(2 * (t & 63) + right) * KK;   // staged block [64 px][2][KK]
                TailState st{0.0f, 0.0f, 0.0f, 0.0f, 0.0f, dst, bl, lane_addr, c_lo, c_end};
                tail_pass<KS, 0, 0, SB>(st);
                e_p0 += clock64() - e0;
                s_part[grp * TM + t] = st.part;
                s_part[(2 + grp) * TM + t] = st.rs_x;
                asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
                float tot = s_part[t] + s_part[TM + t];
                if (STRADDLE) tot += __half2float(__float2half_rn(s_part[2 * TM + t] + s_part[3 * TM + t]));
                st.den = __half2float(__float2half_rn(tot));
                st.den = __half2float(__float2half_rn(st.den + 1e-9f));
                if (st.den > 0.0f && st.den <= 65504.0f) {
                    st.rc = rcp_approx(st.den);
                    st.rc = fmaf(st.rc, fmaf(-st.den, st.rc, 1.0f), st.rc);
                } else {
                    st.den = 0.0f;                            // all-zero kernel: every quotient becomes 0
                    st.rc = 0.0f;
                }
                if (right) tail_pass<KS, 1, 1, SB>(st); else tail_pass<KS, 1, 0, SB>(st);
                e1 = clock64();
                tc_fence_before();
                fence_proxy_async();
                asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
                if (threadIdx.x == 0 && tile < n_tiles) {
                    const unsigned rows_here = min((unsigned)TM, n_rows - tile * TM);
                    const unsigned bytes = rows_here * (unsigned)(KK * 2);
                    __half *gp = psf + (size_t)tile * TM * KK;
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gp), "r"(smem_u32(sA)), "r"(bytes) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                }
                asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
                e_last += clock64() - e0;
                e_st += clock64() - e1;
            }
        }
        if (dbg && threadIdx.x == 0) { dbg[blockIdx.x * 8 + 3] = clock64() - e_start; dbg[blockIdx.x * 8 + 4] = e_wait; dbg[blockIdx.x * 8 + 5] = e_last; dbg[blockIdx.x * 8 + 6] = e_p0; dbg[blockIdx.x * 8 + 7] = e_st; }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (NCTA == 2) cluster_sync_all();             // the other CTA may still be signalling this one's barriers
    if (warp == MMA_WARP) {
        tc_fence_after();
        if constexpr (NCTA == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
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

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
//   * the weights arrive PRE-SWIZZLED (sdirt_mlp_fused_pack_layer lays every [<= 256 x 64] tile out as its shared-memory
//     image), so a pipeline stage is one contiguous cp.async.bulk (UBLKCP) into an mbarrier ring: no tensor maps;
//   * CTA pairs (cluster of 2): tcgen05.mma.cta_group::2 (UMMA M = 256: 128 rows from each CTA, N = 256, K = 16) into a
//     [128 lanes x 512 columns] fp32 accumulator per CTA that fills the SM's TMEM; each CTA stages half of every weight
//     tile; tcgen05.commit (multicast to both CTAs) releases ring stages and hands the accumulator over;
//   * eight epilogue warps (thread = row = TMEM lane, two warps per lane quarter) read it back with tcgen05.ld, add the bias,
//     ReLU, round to fp16 and write the next layer's A operand straight into the swizzled tile; after the last layer they
//     compute torch's two-stage fp16 sums from TMEM, normalise, flip the right rows and stage the packed [64, 2, ks, ks]
//     block, which leaves with two bulk stores;
//   * the first Linear (K = 3) is computed by the same warps on the CUDA cores, one tile ahead.
//
// Roles: warps 0-7 epilogue / first layer, warp 8 weight producer, warp 9 MMA issuer + TMEM owner.  Persistent: CTA pairs loop
// over groups of two tiles.  Numerics: identical rounding points to the cuBLAS route (fp32 accumulate, bias added in fp32, one
// rounding); only the fp32 summation order inside a dot product differs, as it does between any two GEMM implementations.
// DESIGN.md 3.4 has the measurements behind each choice.
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
// The MMA warp runs its loop with all 32 lanes (warp-uniform control flow and operands, so descriptors and addresses live in
// uniform registers); `issue` is non-zero in the one elected lane that actually issues.  (Inside an `if (lane == 0)` region
// the compiler wraps every UTCHMMA in an ELECT / 7 x R2UR.BROADCAST / BRA.U.ANY loop -- ~25 instructions per MMA, more than
// the issue slots a warp that shares its scheduler with two epilogue warps gets in the 128 cycles an MMA takes.)
__device__ __forceinline__ unsigned elect_one() {
    unsigned p;
    asm volatile("{\n.reg .pred pe;\nelect.sync _|pe, 0xffffffff;\nselp.u32 %0, 1, 0, pe;\n}" : "=r"(p));
    return p;
}
// tcgen05.commit: the mbarrier is signalled once every MMA issued so far by this thread has completed; NCTA = 2: the barrier
// at the same offset in BOTH CTAs of the pair
template <int NCTA>
__device__ __forceinline__ void tc_commit(unsigned issue, unsigned long long *bar) {
    if constexpr (NCTA == 1)
        asm volatile("{\n.reg .pred q;\nsetp.ne.b32 q, %1, 0;\n"
                     "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}" ::"r"(smem_u32(bar)), "r"(issue) : "memory");
    else
        asm volatile("{\n.reg .pred q;\nsetp.ne.b32 q, %2, 0;\n"
                     "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n}"
                     ::"r"(smem_u32(bar)), "h"((unsigned short)3), "r"(issue) : "memory");
}
template <int NCTA>
__device__ __forceinline__ void tc_mma_f16(unsigned issue, unsigned d_tmem, unsigned long long a_desc, unsigned long long b_desc, unsigned idesc, unsigned acc) {
    if constexpr (NCTA == 1)
        asm volatile(
            "{\n.reg .pred p, q;\nsetp.ne.b32 p, %4, 0;\nsetp.ne.b32 q, %5, 0;\n"
            "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
            ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc), "r"(issue) : "memory");
    else
        asm volatile(
            "{\n.reg .pred p, q;\nsetp.ne.b32 p, %4, 0;\nsetp.ne.b32 q, %5, 0;\n"
            "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}"
            ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc), "r"(issue) : "memory");
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
// ROWP (8, 12 or 24) consecutive fp32 columns of this thread's TMEM lane: one kernel row of the last layer (see row_pad).
// `dep`: registers of the previous row; naming them as in/out operands keeps the compiler from sinking the issue below the
// arithmetic on them, which would serialise "load, wait, compute" again.
#define TM_O8(b) TM_R(b), TM_R(b + 1), TM_R(b + 2), TM_R(b + 3), TM_R(b + 4), TM_R(b + 5), TM_R(b + 6), TM_R(b + 7)
#define TM_D8(b) "+r"(dep[b]), "+r"(dep[b + 1]), "+r"(dep[b + 2]), "+r"(dep[b + 3]), "+r"(dep[b + 4]), "+r"(dep[b + 5]), "+r"(dep[b + 6]), "+r"(dep[b + 7])
#define TM_W8(b) TM_RW(b), TM_RW(b + 1), TM_RW(b + 2), TM_RW(b + 3), TM_RW(b + 4), TM_RW(b + 5), TM_RW(b + 6), TM_RW(b + 7)
template <int ROWP>
__device__ __forceinline__ void tmem_ld_row_issue(unsigned taddr, unsigned (&r)[ROWP], unsigned (&dep)[ROWP]) {
    static_assert(ROWP == 8 || ROWP == 12 || ROWP == 24, "row loads are written for 8, 12 and 24 columns");
    // loads of 8 (4) columns, each starting at a multiple of its width
    if constexpr (ROWP == 24) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n" : TM_O8(0) : "r"(taddr) : "memory");
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n" : TM_O8(8) : "r"(taddr + 8u) : "memory");
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%32];\n"       // %8..%31: dep
                     : TM_O8(16), TM_D8(0), TM_D8(8), TM_D8(16) : "r"(taddr + 16u) : "memory");
    } else if constexpr (ROWP == 12) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n" : TM_R(0), TM_R(1), TM_R(2), TM_R(3) : "r"(taddr) : "memory");
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n" : TM_R(4), TM_R(5), TM_R(6), TM_R(7) : "r"(taddr + 4u) : "memory");
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%16];\n"                        // %4..%15: dep
                     : TM_R(8), TM_R(9), TM_R(10), TM_R(11), TM_D8(0), "+r"(dep[8]), "+r"(dep[9]), "+r"(dep[10]), "+r"(dep[11]) : "r"(taddr + 8u) : "memory");
    } else {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%16];\n" : TM_O8(0), TM_D8(0) : "r"(taddr) : "memory");
    }
}
template <int ROWP>
__device__ __forceinline__ void tmem_wait_row(unsigned (&r)[ROWP]) {
    if constexpr (ROWP == 24) asm volatile("tcgen05.wait::ld.sync.aligned;\n" : TM_W8(0), TM_W8(8), TM_W8(16) :: "memory");
    else if constexpr (ROWP == 12) asm volatile("tcgen05.wait::ld.sync.aligned;\n" : TM_W8(0), TM_RW(8), TM_RW(9), TM_RW(10), TM_RW(11) :: "memory");
    else asm volatile("tcgen05.wait::ld.sync.aligned;\n" : TM_W8(0) :: "memory");
}
#undef TM_O8
#undef TM_D8
#undef TM_W8
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

// The last layer is packed with every kernel row padded to row_pad(KS) columns (sdirt_mlp_fused_pack_layer: output
// u * KS + w sits in accumulator column u * ROWP + w, the padding has zero weights), so the tail below loops over kernel
// rows at RUN time while every column inside a row -- where a row sum ends, which staged element a value goes to, flipped
// or not -- is a COMPILE-time position.  (A tail unrolled over all 441 compile-time columns was 12 k instructions of
// straight-line code per kernel, four times the instruction cache, and ran at the speed of instruction fetch; one indexed at
// run time paid ~12 integer instructions and two branches per element.)
__host__ __device__ constexpr int row_pad(int ks) { return (ks + 3) / 4 * 4; }

struct TailState {
    float part, den, rc;                // sum of this thread's rounded kernel-row sums; the kernel's denominator, its reciprocal
    __half *dst;                        // this row of the staged block
    const float *bl;                    // last layer's bias, padded like the accumulator (SB: in shared memory)
    unsigned lane_addr;
};
// One kernel row u of tile row t: relu(acc + bias) rounded to fp16, two columns per FADD2 / F2FP / HMNMX2.
// PASS 0: add the row's sum, rounded to fp16 (torch's sum(-1) of a half tensor), to st.part.
// PASS 1 (the accumulator is read a second time: TMEM loads are cheap, two hundred live registers are not): the IEEE quotient
// by the denominator, staged at its final position -- RIGHT: rows 64..127 of a tile are right kernels, whose columns are
// flipped inside every kernel row -- with 32-bit stores where the staged row allows (ALIGNED: it starts on a word).
template <int KS, int PASS, int RIGHT, int ALIGNED, bool SB>
__device__ __forceinline__ void tail_row(const int u, const unsigned (&cur)[row_pad(KS)], TailState &st) {
    constexpr int ROWP = row_pad(KS), NP = (KS + 1) / 2;
    float4 bv[ROWP / 4];
#pragma unroll
    for (int g = 0; g < ROWP / 4; ++g) {
        const float4 *bp = reinterpret_cast<const float4 *>(st.bl + u * ROWP) + g;
        bv[g] = SB ? lds_f4_here(bp) : ldg_f4_here(bp);
    }
    float v[2 * NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) {
        const int i = 2 * j;
        const float b0 = (i & 2) ? bv[i >> 2].z : bv[i >> 2].x, b1 = (i & 2) ? bv[i >> 2].w : bv[i >> 2].y;
        const float2 s2 = __fadd2_rn(make_float2(__uint_as_float(cur[i]), __uint_as_float(cur[i + 1])), make_float2(b0, b1));
        const __half2 h = __hmax2(__floats2half2_rn(s2.x, s2.y), __float2half2_rn(0.0f));
        v[i] = __low2float(h);
        v[i + 1] = __high2float(h);       // column KS of an odd-sized row is padding: exactly 0 (zero weights, zero bias)
    }
    if (PASS == 0) {
        float rs = 0.0f;
#pragma unroll
        for (int w = 0; w < KS; ++w) rs += v[w];
        st.part += __half2float(__float2half_rn(rs));
    } else {
        // IEEE v / den from the refined reciprocal (div_rn's sequence); den = rc = 0 for an all-zero kernel -> 0
        const float2 rc2 = make_float2(st.rc, st.rc), nd2 = make_float2(-st.den, -st.den);
        float r[2 * NP];
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            const float2 v2 = make_float2(v[2 * j], v[2 * j + 1]);
            const float2 q = __fmul2_rn(v2, rc2);
            const float2 o = __ffma2_rn(rc2, __ffma2_rn(nd2, q, v2), q);
            r[2 * j] = o.x;
            r[2 * j + 1] = o.y;
        }
        __half *d = st.dst + u * KS;
        // staged position p holds column p (left) or KS - 1 - p (right)
#define SDIRT_TAIL_OUT(p) (RIGHT ? r[KS - 1 - (p)] : r[(p)])
        if (ALIGNED) {
#pragma unroll
            for (int pp = 0; pp + 1 < KS; pp += 2) *reinterpret_cast<__half2 *>(d + pp) = __floats2half2_rn(SDIRT_TAIL_OUT(pp), SDIRT_TAIL_OUT(pp + 1));
            if (KS & 1) d[KS - 1] = __float2half_rn(SDIRT_TAIL_OUT(KS - 1));
        } else {
            d[0] = __float2half_rn(SDIRT_TAIL_OUT(0));
#pragma unroll
            for (int pp = 1; pp + 1 < KS; pp += 2) *reinterpret_cast<__half2 *>(d + pp) = __floats2half2_rn(SDIRT_TAIL_OUT(pp), SDIRT_TAIL_OUT(pp + 1));
            if (!(KS & 1)) d[KS - 1] = __float2half_rn(SDIRT_TAIL_OUT(KS - 1));
        }
#undef SDIRT_TAIL_OUT
    }
}
// Kernel rows [u_lo, u_hi) of this thread's tile row; NOT unrolled.  A staged row starts at half
// (2 * px + RIGHT) * KS * KS + u * KS: on a word boundary when RIGHT + u is even (KS odd; the host refuses even KS).
template <int KS, int PASS, int RIGHT, bool SB>
__device__ __forceinline__ void tail_pass(const int u_lo, const int u_hi, TailState &st) {
    constexpr int ROWP = row_pad(KS);
    unsigned ra[ROWP], rb[ROWP];
#pragma unroll
    for (int i = 0; i < ROWP; ++i) ra[i] = rb[i] = 0u;
    int u = u_lo;
    // two kernel rows per iteration: both loads, one wait, then the two rows' arithmetic in one basic block, where the
    // scheduler interleaves their dependency chains (a warp has one partner on its scheduler to hide latency behind)
#pragma unroll 1
    for (; u + 1 < u_hi; u += 2) {
        tmem_ld_row_issue<ROWP>(st.lane_addr + (unsigned)(u * ROWP), ra, rb);
        tmem_ld_row_issue<ROWP>(st.lane_addr + (unsigned)((u + 1) * ROWP), rb, ra);
        tmem_wait_row<ROWP>(ra);
        tmem_wait_row<ROWP>(rb);
        if ((RIGHT + u) & 1) { tail_row<KS, PASS, RIGHT, 0, SB>(u, ra, st); tail_row<KS, PASS, RIGHT, 1, SB>(u + 1, rb, st); }
        else { tail_row<KS, PASS, RIGHT, 1, SB>(u, ra, st); tail_row<KS, PASS, RIGHT, 0, SB>(u + 1, rb, st); }
    }
    if (u < u_hi) {
        tmem_ld_row_issue<ROWP>(st.lane_addr + (unsigned)(u * ROWP), ra, rb);
        tmem_wait_row<ROWP>(ra);
        if ((RIGHT + u) & 1) tail_row<KS, PASS, RIGHT, 0, SB>(u, ra, st); else tail_row<KS, PASS, RIGHT, 1, SB>(u, ra, st);
    }
}

// Warp roles: warps 0-7 epilogue (thread = row (warp & 3) * 32 + lane = TMEM lane; the two warps of a lane quarter split the
// columns), warp 8 weight producer, warp 9 MMA issuer + TMEM owner (NCTA = 2: the leader CTA's warp 9 issues for the pair, the
// other CTA's warp 9 relays its CTA's "operand ready" barriers to the leader).
//
// Hand-over of the activation tile between layers (a_lo / a_hi): the epilogue drains accumulator half 0 (output columns
// 0..255 = k-blocks 0..3 of the next layer) into registers while the MMAs of half 1 still read the A tile; when the layer is
// complete it stores them and signals a_lo, and the next layer's MMAs over k-blocks 0..3 run while the epilogue converts half
// 1 (k-blocks 4..7, signalled by a_hi).  a_free lets the stores of half 0 start before the layer is complete.
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
    unsigned char *sA = fm_smem;
    unsigned char *sW = fm_smem + CF::OFF_W;
    float *sBias = reinterpret_cast<float *>(fm_smem + CF::OFF_BIAS);
    float4 *sW1 = reinterpret_cast<float4 *>(fm_smem + CF::OFF_W1);
    unsigned long long *full = reinterpret_cast<unsigned long long *>(fm_smem + CF::OFF_BAR), *empty = full + STAGES;
    unsigned long long *a_lo = empty + STAGES, *a_hi = a_lo + 1, *acc_ready = a_hi + 1;     // acc_ready[2]: one per accumulator half
    unsigned long long *a_free = acc_ready + 2;              // k-blocks 0..3 of the A tile are no longer read by this layer's MMAs
    unsigned *tmem_slot = reinterpret_cast<unsigned *>(a_free + 1);
    float *s_part = reinterpret_cast<float *>(sA + TM * KK * 2);                     // [2][TM] partial sums of the last layer, behind the packed block
    static_assert((KS & 1) == 1 && TM * KK * 2 + 2 * TM * 4 <= A_BYTES, "the packed block and its partial sums must fit in the activation tile");
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
        mbar_init(a_free, 1);
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
        if (rank == 0) {
            // ---- MMA issuer (NCTA = 2: for both CTAs of the pair); all lanes run the loop, one elected lane issues --------------
            const unsigned issue = elect_one();
            int s = 0;
            unsigned ph = 0, a_ph = 0;
            const unsigned a_base = smem_u32(sA), w_base = smem_u32(sW);
            long long t_start = clock64(), t_full = 0, t_aready = 0, t0 = 0;
            for (unsigned g = g0; g < n_groups; g += g_step) {
                for (int l = 0; l < L; ++l) {
                    const int nkb = net.K[l] >> 6, N = net.N[l], kb_lo = min(4, nkb);
                    const int h_last = N > 256 ? 1 : 0, kb_free = min(min(256, N) >> 6, nkb);
                    const bool tl = dbg && blockIdx.x == 0 && g == g0 + g_step;    // timeline of this CTA's second tile group
                    long long *tlp = dbg + 148 * 8 + l * 16;
                    if (dbg) t0 = clock64();
                    // k-blocks [0, kb_lo) of this layer's A tile are written and accumulator half 0 is drained
                    mbar_wait(a_lo, a_ph);
                    if (kb_lo == nkb) mbar_wait(a_hi, a_ph);
                    if (dbg) t_aready += clock64() - t0;
                    if (tl && issue) { tlp[0] = t0; tlp[1] = clock64(); }
                    tc_fence_after();
                    for (int h = 0; h < 2; ++h) {
                        if (h * 256 < N) {
                            const unsigned idesc = instr_desc(min(256, N - h * 256), TM * NCTA);
                            const unsigned d_tmem = tmem + (unsigned)(h * 256);
                            for (int kb = 0; kb < nkb; ++kb) {
                                if (h == 0 && kb == kb_lo) {  // the rest of the A tile, and accumulator half 1 drained
                                    if (dbg) t0 = clock64();
                                    mbar_wait(a_hi, a_ph);
                                    if (dbg) t_aready += clock64() - t0;
                                    if (tl && issue) { tlp[2] = t0; tlp[3] = clock64(); }
                                    tc_fence_after();
                                }
                                if (dbg) t0 = clock64();
                                mbar_wait(full + s, ph);
                                if (dbg) t_full += clock64() - t0;
                                tc_fence_after();
                                // operand descriptors of the four K = 16 steps of this k-block: 32 bytes apart inside the 128-byte
                                // swizzle row, i.e. +2 in the descriptor's address field
                                const unsigned long long a_desc = smem_desc(a_base + kb * A_KB_BYTES), b_desc = smem_desc(w_base + s * STAGE_BYTES);
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    tc_mma_f16<NCTA>(issue, d_tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, (unsigned)((kb | k) != 0));
                                tc_commit<NCTA>(issue, empty + s);   // the stage is free once these MMAs have read it
                                if (++s == STAGES) { s = 0; ph ^= 1u; }
                                // the last MMA that reads k-blocks [0, kb_free) of the A tile -- the ones the epilogue overwrites with
                                // accumulator half 0 -- is on its way: the epilogue may store as soon as it completes
                                if (h == h_last && kb == kb_free - 1 && l + 1 < L) tc_commit<NCTA>(issue, a_free);
                            }
                        }
                        tc_commit<NCTA>(issue, acc_ready + h);       // this half of the accumulator is complete (h = 1: the whole layer)
                        if (tl && issue) tlp[4 + h] = clock64();
                    }
                    a_ph ^= 1u;
                }
            }
            if (dbg && issue) { dbg[blockIdx.x * 8 + 0] = clock64() - t_start; dbg[blockIdx.x * 8 + 1] = t_full; dbg[blockIdx.x * 8 + 2] = t_aready; }
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
        unsigned acc_ph = 0, free_ph = 0;
        long long e_start = clock64(), e_wait = 0, e_last = 0, e_p0 = 0, e_st = 0, e0, e1;
        // first Linear (K = 3) + ReLU of tile row t (mlp_input_layer_kernel's arithmetic) into registers: tile row t is pixel
        // (t & 63) of the tile's 64, left kernel for t < 64, right kernel for t >= 64 (a warp is all-left or all-right).  Computed
        // one tile ahead -- while the epilogue warps wait for the last layer's accumulators -- so that the coordinate loads and
        // the arithmetic are off the path between two tiles.
        unsigned fl[MAX_N1 / 16][4];
        auto first_layer = [&](unsigned tile) {
            const unsigned side = (unsigned)t >> 6;
            const unsigned p = min(tile * (TM / 2) + ((unsigned)t & 63u), (n_rows >> 1) - 1);
            const unsigned q = p / (unsigned)W, x = p - q * (unsigned)W;
            const unsigned bq = q / (unsigned)nrw;
            const int y = row0 + (int)(q - bq * (unsigned)nrw), b = b0 + (int)bq;
            const float xr = __ldg(xs + x);
            const float xv = __half2float(__float2half_rn(side ? -xr : xr));
            const float yv = __half2float(__float2half_rn(__ldg(ys + y)));
            const float zv = __half2float(__float2half_rn(__ldg(z + ((int64_t)b * H + y) * W + x)));
#pragma unroll
            for (int j = 0; j < MAX_N1 / 16; ++j) {
                const int c = 8 * grp + 16 * j;
                if (c < net.n1) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float4 wa = sW1[c + 2 * k], wb = sW1[c + 2 * k + 1];
                        const float va = fmaf(zv, wa.z, fmaf(yv, wa.y, xv * wa.x)) + wa.w;
                        const float vb = fmaf(zv, wb.z, fmaf(yv, wb.y, xv * wb.x)) + wb.w;
                        fl[j][k] = pack_relu_h2(va, vb);
                    }
                }
            }
        };
        first_layer(NCTA * g0 + rank);
        for (unsigned g = g0; g < n_groups; g += g_step) {
            const unsigned tile = NCTA * g + rank;
            // first Linear + ReLU of this row (computed ahead, see first_layer below) -> the swizzled A tile; the two warps of a row
            // take alternate groups of 8 columns
#pragma unroll
            for (int j = 0; j < MAX_N1 / 16; ++j) {
                const int c = 8 * grp + 16 * j;
                if (c < net.n1) *reinterpret_cast<uint4 *>(sA + a_chunk_off(t, c)) = make_uint4(fl[j][0], fl[j][1], fl[j][2], fl[j][3]);
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
                // half 0 (columns 0..255 = k-blocks 0..3 of the next layer's A tile) goes to shared memory as soon as this layer's
                // MMAs have finished with those k-blocks, while the MMAs over k-blocks 4..7 of half 1 still run: the next layer
                // starts the moment this one ends
                if (l == 0) {                                 // the previous tile's staged block has left shared memory
                    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
                }
                e0 = clock64();
                mbar_wait(a_free, free_ph);
                free_ph ^= 1u;
                e_wait += clock64() - e0;
                if (tl) { tlp[2] = e0; tlp[3] = clock64(); }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int c = 2 * i + grp;
                    if (c < nc0) store_a_chunks(sA, t, 32 * c, held[i]);
                }
                fence_proxy_async();
                tc_fence_before();
                mbar_arrive(a_lo);
                if (tl) tlp[4] = clock64();
                e0 = clock64();
                mbar_wait(acc_ready + 1, acc_ph);
                e_wait += clock64() - e0;
                if (tl) tlp[6] = clock64();
                acc_ph ^= 1u;
                tc_fence_after();
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
            // The two warps of a tile row split the kernel rows ([0, U_SPLIT) and [U_SPLIT, KS)); their partial totals are added
            // through shared memory (fp32 sums of a few fp16 values: exact, whatever the order).
            {
                constexpr int U_SPLIT = (KS + 1) / 2;
                constexpr int U_H0 = (256 / row_pad(KS)) < KS ? (256 / row_pad(KS)) : KS;     // kernel rows that lie in accumulator half 0
                const float *bl = (SB ? sBias : bias) + net.b_off[L - 1];
                if (g + g_step < n_groups) first_layer(NCTA * (g + g_step) + rank);
                const int right = t >> 6;                     // warp-uniform: rows 64..127 are right kernels
                __half *dst = reinterpret_cast<__half *>(sA) + (size_t)(2 * (t & 63) + right) * KK;   // staged block [64 px][2][KK]
                TailState st{0.0f, 0.0f, 0.0f, dst, bl, lane_addr};
                const bool tl = dbg && blockIdx.x == 0 && threadIdx.x == 0 && g == g0 + g_step;
                long long *tlp = dbg + 148 * 8 + (L - 1) * 16 + 8;         // tail timeline: 8 stamps
                e0 = clock64();
                mbar_wait(acc_ready, acc_ph);
                e_wait += clock64() - e0;
                e0 = clock64();
                if (tl) tlp[0] = e0;
                tc_fence_after();
                // row sums of the kernel rows in half 0, while the MMAs of half 1 run
                tail_pass<KS, 0, 0, SB>(grp ? U_H0 / 2 : 0, grp ? U_H0 : U_H0 / 2, st);
                e_p0 += clock64() - e0;
                e0 = clock64();
                if (tl) tlp[1] = e0;
                mbar_wait(acc_ready + 1, acc_ph);
                acc_ph ^= 1u;
                e_wait += clock64() - e0;
                e0 = clock64();
                if (tl) tlp[2] = e0;
                tc_fence_after();
                tail_pass<KS, 0, 0, SB>(grp ? U_H0 + (KS - U_H0 + 1) / 2 : U_H0, grp ? KS : U_H0 + (KS - U_H0 + 1) / 2, st);
                const int u_lo = grp ? U_SPLIT : 0, u_hi = grp ? KS : U_SPLIT;
                e_p0 += clock64() - e0;
                if (tl) tlp[3] = clock64();
                s_part[grp * TM + t] = st.part;
                asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
                const float tot = s_part[t] + s_part[TM + t];
                st.den = __half2float(__float2half_rn(tot));
                st.den = __half2float(__float2half_rn(st.den + 1e-9f));
                if (st.den > 0.0f && st.den <= 65504.0f) {
                    st.rc = rcp_approx(st.den);
                    st.rc = fmaf(st.rc, fmaf(-st.den, st.rc, 1.0f), st.rc);
                } else {
                    st.den = 0.0f;                            // all-zero kernel: every quotient becomes 0
                    st.rc = 0.0f;
                }
                if (L == 1) {                                 // no hidden layer has waited for the previous staged block to leave
                    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
                }
                if (tl) tlp[4] = clock64();
                if (right) tail_pass<KS, 1, 1, SB>(u_lo, u_hi, st); else tail_pass<KS, 1, 0, SB>(u_lo, u_hi, st);
                e1 = clock64();
                if (tl) tlp[5] = e1;
                tc_fence_before();
                fence_proxy_async();
                asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
                if (tl) tlp[6] = clock64();
                // The staged block leaves as two bulk stores: the part under k-blocks 0 and 1 of the A tile -- which the next tile's
                // first layer overwrites at once -- and the rest, which is only waited for before the first hidden layer's stores.
                // (All CTAs reach this point together: 148 x 113 KB in one burst takes the memory system ~4 k cycles.)
                if (threadIdx.x == 0 && tile < n_tiles) {
                    const unsigned rows_here = min((unsigned)TM, n_rows - tile * TM);
                    const unsigned bytes = rows_here * (unsigned)(KK * 2), head = min(bytes, (unsigned)(2 * A_KB_BYTES));
                    unsigned char *gp = reinterpret_cast<unsigned char *>(psf + (size_t)tile * TM * KK);
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gp), "r"(smem_u32(sA)), "r"(head) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    if (bytes > head)
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gp + head), "r"(smem_u32(sA) + head), "r"(bytes - head) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                }
                asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
                if (tl) tlp[7] = clock64();
                e_last += clock64() - e0;
                e_st += clock64() - e1;
            }
        }
        if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // the last staged block has been written
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
// shared-memory image of a K-major SWIZZLE_128B operand).  ks > 0 (the last layer): packed row u * rowp + w takes output
// u * ks + w (kernel rows padded to rowp columns, see row_pad).  Rows without a source (padding) are zero.  One thread per chunk.
__device__ __forceinline__ int packed_row_source(int n, int n_true, int ks, int rowp) {
    if (ks > 0) {
        const int u = n / rowp, w = n - u * rowp;
        return (u < ks && w < ks) ? u * ks + w : -1;
    }
    return n < n_true ? n : -1;
}
__global__ void __launch_bounds__(256)
swizzle_weights_kernel(const __half *__restrict__ w, int n_true, int n_pad, int K, int ks, int rowp, unsigned char *__restrict__ out) {
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
        const int n = packed_row_source(h * 256 + r, n_true, ks, rowp), c = pch ^ (r & 7);
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (n >= 0) v = *reinterpret_cast<const uint4 *>(w + (int64_t)n * K + kb * 64 + c * 8);
        *reinterpret_cast<uint4 *>(out + i * 16) = v;
    }
}

__global__ void pad_bias_kernel(const __half *__restrict__ b, int n_true, int n_pad, int ks, int rowp, float *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    const int n = packed_row_source(i, n_true, ks, rowp);
    out[i] = n >= 0 ? __half2float(b[n]) : 0.0f;
}
}  // namespace mlpf

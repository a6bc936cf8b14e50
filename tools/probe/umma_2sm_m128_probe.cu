// Probe for the two-tile ping-pong variant of the fused PSF-MLP kernel: tcgen05.mma.cta_group::2 with UMMA M = 128
// (64 rows per CTA), N = 256, K = 16 x 4.  Checks (1) where the 64 x 256 fp32 accumulator of each CTA lands in TMEM -- expected
// from CuTe's "2x2" fragment atom: row m, column n < 128 in lane m column n; column n >= 128 in lane 64 + m column n - 128 --
// and (2) how many cycles an MMA of this shape takes back to back.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o umma_2sm_m128_probe umma_2sm_m128_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long smem_desc(unsigned addr) {   // K-major, SWIZZLE_128B, 8-row atoms of 1024 B
    return (unsigned long long)((addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

constexpr int M_CTA = 64, N = 256, K = 64;
constexpr int A_BYTES = M_CTA * 128, B_BYTES = (N / 2) * 128;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe(const unsigned char *a_img /*[2][A_BYTES]*/, const unsigned char *b_img /*[2][B_BYTES]*/, float *out /*[2][128][128]*/,
      long long *cycles, int reps) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *sA = smem, *sB = smem + A_BYTES;
    unsigned long long *done = reinterpret_cast<unsigned long long *>(smem + A_BYTES + B_BYTES);
    unsigned *slot = reinterpret_cast<unsigned *>(done + 1);
    unsigned rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < A_BYTES / 16; i += blockDim.x) reinterpret_cast<uint4 *>(sA)[i] = reinterpret_cast<const uint4 *>(a_img + rank * A_BYTES)[i];
    for (int i = threadIdx.x; i < B_BYTES / 16; i += blockDim.x) reinterpret_cast<uint4 *>(sB)[i] = reinterpret_cast<const uint4 *>(b_img + rank * B_BYTES)[i];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(done)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = *slot;
    const unsigned idesc = (1u << 4) | ((unsigned)(N >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
    long long t0 = 0, t1 = 0;
    if (rank == 0 && warp == 1 && lane == 0) {
        t0 = clock64();
        for (int r = 0; r < reps; ++r)
            for (int k = 0; k < 4; ++k) {
                const unsigned acc = k != 0;     // every repetition recomputes D = A B^T from scratch
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}"
                             ::"r"(tmem), "l"(smem_desc(smem_u32(sA) + k * 32)), "l"(smem_desc(smem_u32(sB) + k * 32)), "r"(idesc), "r"(acc) : "memory");
            }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(done)), "h"((unsigned short)3) : "memory");
    }
    mbar_wait(done, 0);
    if (rank == 0 && warp == 1 && lane == 0) { t1 = clock64(); cycles[0] = t1 - t0; }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // every warp dumps its 32 lanes x 128 columns
    for (int c = 0; c < 128; c += 32) {
        unsigned r[32];
        const unsigned taddr = tmem + ((unsigned)(warp * 32) << 16) + (unsigned)c;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                     "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\ntcgen05.wait::ld.sync.aligned;\n"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                       "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                       "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                       "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                     : "r"(taddr) : "memory");
        for (int i = 0; i < 32; ++i) out[((size_t)rank * 128 + warp * 32 + lane) * 128 + c + i] = __uint_as_float(r[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}

static void swizzle_image(const __half *src /*[rows][64]*/, int rows, unsigned char *img) {
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < 8; ++c)
            memcpy(img + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4), src + r * 64 + c * 8, 16);
}

int main() {
    std::vector<__half> A(2 * M_CTA * K), B(N * K);
    srand(1);
    for (auto &v : A) v = __float2half((rand() % 17 - 8) / 8.0f);
    for (auto &v : B) v = __float2half((rand() % 13 - 6) / 4.0f);
    std::vector<unsigned char> a_img(2 * A_BYTES), b_img(2 * B_BYTES);
    for (int r = 0; r < 2; ++r) {
        swizzle_image(A.data() + r * M_CTA * K, M_CTA, a_img.data() + r * A_BYTES);
        swizzle_image(B.data() + r * (N / 2) * K, N / 2, b_img.data() + r * B_BYTES);
    }
    unsigned char *da, *db; float *dout; long long *dc;
    cudaMalloc(&da, a_img.size()); cudaMalloc(&db, b_img.size()); cudaMalloc(&dout, 2 * 128 * 128 * 4); cudaMalloc(&dc, 8);
    cudaMemcpy(da, a_img.data(), a_img.size(), cudaMemcpyHostToDevice); cudaMemcpy(db, b_img.data(), b_img.size(), cudaMemcpyHostToDevice);
    const int smem = A_BYTES + B_BYTES + 64;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int reps : {1, 2000}) {
        cudaMemset(dout, 0, 2 * 128 * 128 * 4);
        probe<<<2, 128, smem>>>(da, db, dout, dc, reps);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
        long long cyc; cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost);
        printf("reps %d: %lld cycles for %d MMAs (M = 128 over the pair, N = 256, K = 16) = %.1f cycles / MMA incl. launch of the chain\n", reps, cyc, 4 * reps, (double)cyc / (4 * reps));
    }
    std::vector<float> out(2 * 128 * 128);
    cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
    // expected: D_rank[m][n] = sum_k A_rank[m][k] * B[n][k]; lane m holds n < 128, lane 64 + m holds n >= 128
    int bad_2x2 = 0, bad_plain = 0;
    for (int r = 0; r < 2; ++r)
        for (int m = 0; m < M_CTA; ++m)
            for (int n = 0; n < N; ++n) {
                float d = 0;
                for (int k = 0; k < K; ++k) d += __half2float(A[(r * M_CTA + m) * K + k]) * __half2float(B[n * K + k]);
                const float got_2x2 = out[((size_t)r * 128 + (n < 128 ? m : 64 + m)) * 128 + (n & 127)];
                if (fabsf(got_2x2 - d) > 1e-3f) ++bad_2x2;
                if (n < 128) { const float g = out[((size_t)r * 128 + m) * 128 + n]; if (fabsf(g - d) > 1e-3f) ++bad_plain; }
            }
    printf("layout check: %d of %d accumulators differ from the \"2x2\" placement (lanes 0..63: n < 128, lanes 64..127: n >= 128); %d of %d in lanes 0..63 alone\n",
           bad_2x2, 2 * M_CTA * N, bad_plain, 2 * M_CTA * 128);
    printf("out[rank 0][lane 0][0..3] = %g %g %g %g, [lane 64][0..3] = %g %g %g %g\n", out[0], out[1], out[2], out[3], out[64 * 128], out[64 * 128 + 1], out[64 * 128 + 2], out[64 * 128 + 3]);
    return 0;
}

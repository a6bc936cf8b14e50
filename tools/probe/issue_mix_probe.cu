#include <cstdio>
#include <cuda_runtime.h>
// What does a packed fp32 instruction cost the scheduler?  Per inner round a thread issues P packed FFMA2 (independent chains),
// S scalar FFMA, A integer ALU ops and M MUFU.RCP; the time per round in SMSP cycles (8 warps per SMSP resident, all rounds
// back to back) tells whether the second cycle of an FFMA2 is free for other instructions and whether scalar FFMA can run
// beside packed ones.
template <int P, int S, int A, int M>
__global__ void __launch_bounds__(256) k(float *out, int iters) {
    float2 a[8];
    float f[8], g[4];
    unsigned m[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 11u};
    for (int i = 0; i < 8; ++i) { a[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f); f[i] = threadIdx.x * 2e-3f + i; }
    for (int i = 0; i < 4; ++i) g[i] = 1.5f + threadIdx.x * 1e-3f + i;
    const float2 b = make_float2(1.0001f, 0.9999f), c = make_float2(1e-3f, -1e-3f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (i < P) a[i] = __ffma2_rn(a[i], b, c);
                if (i < S) f[i] = fmaf(f[i], b.x, c.x);
                if (i < A) { m[i & 3] = (m[i & 3] ^ (m[i & 3] >> 3)) + 0x9e3779b9u; }     // 2 ALU ops (LOP3 with shift folded or SHF + LOP3 + IADD)
                if (i < M) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(g[i & 3]));
            }
        }
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y + f[i];
    for (int i = 0; i < 4; ++i) s += g[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)(m[0] ^ m[1] ^ m[2] ^ m[3]);
}
template <int P, int S, int A, int M> void run(float *d) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4000, blocks = 148 * 8;
    k<P, S, A, M><<<blocks, 256>>>(d, 100); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<P, S, A, M><<<blocks, 256>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // 8 CTAs x 8 warps per SM = 16 warps per SMSP; rounds per warp = iters * 4
    const double cyc = ms * 1e-3 * 1.965e9 / (16.0 * iters * 4);
    printf("packed %d  scalar %d  alu-groups %d  mufu %d : %.3f ms  %.2f SMSP-cycles per warp-round  (%.1f TFLOP/s of FMA)\n", P, S, A, M, ms, cyc,
           (double)blocks * 256 * iters * 4 * (2 * P + S) * 2 / ms / 1e9);
}
int main() {
    float *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    run<8, 0, 0, 0>(d); run<0, 8, 0, 0>(d); run<4, 0, 0, 0>(d); run<4, 4, 0, 0>(d); run<4, 8, 0, 0>(d); run<8, 8, 0, 0>(d); run<8, 4, 0, 0>(d);
    run<8, 0, 4, 0>(d); run<8, 0, 8, 0>(d); run<0, 8, 8, 0>(d); run<0, 0, 8, 0>(d);
    run<8, 0, 0, 2>(d); run<8, 0, 0, 4>(d); run<0, 0, 0, 4>(d); run<8, 0, 4, 2>(d);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}

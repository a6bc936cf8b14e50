#include <cstdio>
#include <cuda_runtime.h>
// issue-rate probe: 8 independent chains per thread of FFMA (scalar) or FFMA2 (packed), optionally interleaved with integer ALU ops
template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, int iters) {
    float2 a[8];
    unsigned m[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 11u};
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    const float2 b = make_float2(1.0001f, 0.9999f), c = make_float2(1e-3f, -1e-3f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0 || MODE == 2) { a[i].x = fmaf(a[i].x, b.x, c.x); a[i].y = fmaf(a[i].y, b.y, c.y); }
                else a[i] = __ffma2_rn(a[i], b, c);
            }
            if (MODE >= 2) {
#pragma unroll
                for (int j = 0; j < 4; ++j) { m[j] = (m[j] ^ (m[j] >> 3)) + 0x9e3779b9u; m[j] = (m[j] << 1) | (m[j] >> 31); }   // 16 ALU ops
            }
        }
    }
    float s = 0; for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)(m[0] ^ m[1] ^ m[2] ^ m[3]);
}
template <int MODE> void run(const char *name, float *d) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4000, blocks = 148 * 8;
    k<MODE><<<blocks, 256>>>(d, 100); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<MODE><<<blocks, 256>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fma = (double)blocks * 256 * iters * 4 * 16;     // scalar FMAs
    printf("%-40s %.3f ms  %.1f TFLOP/s (FMA = 2)\n", name, ms, fma * 2 / ms / 1e9);
}
int main() {
    float *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    run<0>("FFMA x16", d); run<1>("FFMA2 x8", d); run<2>("FFMA x16 + 16 ALU", d); run<3>("FFMA2 x8 + 16 ALU", d);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}

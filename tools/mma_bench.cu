// mma_bench.cu - issue rate / latency of the legacy tensor path (mma.sync.m16n8k16 f16 -> f32, SASS HMMA.16816.F32) on sm_100a,
// in the shape the decode kernel uses it: W warps per CTA, CH independent accumulator chains per warp, chains of dependent mma.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_bench tools/mma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
template <int CH>
__global__ void k(int iters, long long *out, float *sink) {
    unsigned a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
    float c[CH][4];
#pragma unroll
    for (int i = 0; i < CH; i++) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CH; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (s == 12345.f) *sink = s;
}
// FFMA dot-product alternative: per lane 16 B of f16 weights from shared memory -> 8 cvt + 8 FFMA against x held in registers
__global__ void kf(int iters, long long *out, float *sink) {
    __shared__ uint4 w[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) w[i] = make_uint4(i, i * 3, i * 5, i * 7);
    float x[8]; for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 0.001f + i;
    float acc0 = 0.f, acc1 = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < 10; j++) {      // 10 x 16 B per lane = one chunk's share (2 rows x 1280 halfs per warp)
            const uint4 v = w[(threadIdx.x + 32 * j + it) & 2047];
            const __half2 *h = reinterpret_cast<const __half2 *>(&v);
#pragma unroll
            for (int q = 0; q < 4; q++) { const float2 f = __half22float2(h[q]); acc0 = fmaf(f.x, x[2 * q], acc0); acc1 = fmaf(f.y, x[2 * q + 1], acc1); }
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (acc0 + acc1 == 12345.f) *sink = acc0;
}
int main() {
    long long *out; float *sink; cudaMalloc(&out, 1024); cudaMalloc(&sink, 4);
    long long h[4];
    const int iters = 2000;
    for (int warps : {1, 4, 8, 16}) {
        auto rep = [&](const char *nm, int ch) {
            cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost);
            printf("%-10s warps=%2d chains=%d: %.1f cycles per mma per warp, %.2f cycles per mma per SM sub-partition\n", nm, warps, ch,
                   (double)h[0] / iters / ch, (double)h[0] / iters / ch / ((warps + 3) / 4));
        };
        k<1><<<1, warps * 32>>>(iters, out, sink); cudaDeviceSynchronize(); rep("HMMA", 1);
        k<2><<<1, warps * 32>>>(iters, out, sink); cudaDeviceSynchronize(); rep("HMMA", 2);
        k<4><<<1, warps * 32>>>(iters, out, sink); cudaDeviceSynchronize(); rep("HMMA", 4);
        k<10><<<1, warps * 32>>>(iters, out, sink); cudaDeviceSynchronize(); rep("HMMA", 10);
        kf<<<1, warps * 32>>>(iters, out, sink); cudaDeviceSynchronize();
        cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost);
        printf("FFMA dot   warps=%2d: %.0f cycles per chunk-share (10 x LDS.128 + 80 cvt/FFMA per lane)\n", warps, (double)h[0] / iters);
    }
    return 0;
}

// encoder.cu - the non-GEMM pieces of the audio encoder: LayerNorm (ggml_norm + affine, eps 1e-5,
// resources/ggml-metal.metal:571-621) producing the f16 GEMM operand, the attention itself is
// attention_sm100.cu and everything else in the encoder is a tcgen05 GEMM epilogue (gemm_sm100.cu).
#include <algorithm>
#include <cstdlib>

#include "kernels.h"

namespace ss {

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// one warp per row; NV = d / 128 float4 chunks per lane, row kept in registers (two-pass variance)
template <int NV, bool OUT_F16>
__global__ void __launch_bounds__(256) layernorm_kernel(const float *__restrict__ x, void *__restrict__ y, int rows,
                                                         const float *__restrict__ w, const float *__restrict__ b) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    pdl_trigger();
    pdl_wait();      // x is the predecessor's output, y may still be read by it
    if (row >= rows) return;
    constexpr int d = NV * 128;
    const float4 *xr = reinterpret_cast<const float4 *>(x + (size_t)row * d);
    float4 v[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; i++) { v[i] = xr[lane + 32 * i]; s += v[i].x + v[i].y + v[i].z + v[i].w; }
    const float mean = warp_sum(s) / d;
    float s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; i++) {
        v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
        s2 += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
    }
    const float scale = rsqrtf(warp_sum(s2) / d + 1e-5f);
    const float4 *w4 = reinterpret_cast<const float4 *>(w), *b4 = reinterpret_cast<const float4 *>(b);
#pragma unroll
    for (int i = 0; i < NV; i++) {
        const float4 g = w4[lane + 32 * i], be = b4[lane + 32 * i];
        float4 o;
        o.x = v[i].x * scale * g.x + be.x; o.y = v[i].y * scale * g.y + be.y;
        o.z = v[i].z * scale * g.z + be.z; o.w = v[i].w * scale * g.w + be.w;
        if (OUT_F16) {
            __half2 h0 = __floats2half2_rn(o.x, o.y), h1 = __floats2half2_rn(o.z, o.w);
            uint2 u; u.x = *reinterpret_cast<uint32_t *>(&h0); u.y = *reinterpret_cast<uint32_t *>(&h1);
            reinterpret_cast<uint2 *>(reinterpret_cast<__half *>(y) + (size_t)row * d)[lane + 32 * i] = u;
        } else {
            reinterpret_cast<float4 *>(reinterpret_cast<float *>(y) + (size_t)row * d)[lane + 32 * i] = o;
        }
    }
}

template <bool OUT_F16>
void ln_dispatch(const float *x, void *y, int rows, int d, const LNp &ln, cudaStream_t st) {
    const int grid = ceil_div(rows, 8);
    switch (d / 128) {
#define CASE(NV) case NV: launch_pdl(encoder_pdl_enabled(), layernorm_kernel<NV, OUT_F16>, dim3(grid), dim3(256), 0, st, x, y, rows, ln.w, ln.b); break;
        CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(8) CASE(10)
#undef CASE
        default: SS_THROW(-9, "unsupported model width %d", d);
    }
    CUDA_CHECK(cudaGetLastError());
}

__global__ void __launch_bounds__(256) f32_to_f16_kernel(const float *__restrict__ x, __half *__restrict__ y, size_t n) {
    pdl_trigger();
    pdl_wait();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) y[i] = __float2half_rn(x[i]);
}

}  // namespace

static thread_local bool t_pdl_scope = true;
void encoder_pdl_scope(bool on) { t_pdl_scope = on; }
bool encoder_pdl_enabled() {
    static const bool on = [] { const char *e = getenv("SS_ENC_PDL"); return !(e && e[0] == '0'); }();
    return on && t_pdl_scope;
}

void layernorm_f16_enqueue(const float *x, __half *y, int rows, int d, const LNp &ln, cudaStream_t st, int *launches) {
    ln_dispatch<true>(x, y, rows, d, ln, st); (*launches)++;
}
void layernorm_f32_enqueue(const float *x, float *y, int rows, int d, const LNp &ln, cudaStream_t st, int *launches) {
    ln_dispatch<false>(x, y, rows, d, ln, st); (*launches)++;
}
void f32_to_f16_enqueue(const float *x, __half *y, size_t n, cudaStream_t st, int *launches) {
    launch_pdl(encoder_pdl_enabled(), f32_to_f16_kernel, dim3((unsigned)std::min<size_t>(ceil_div<size_t>(n, 256), 148 * 8)), dim3(256), 0, st, x, y, n);
    CUDA_CHECK(cudaGetLastError()); (*launches)++;
}

}  // namespace ss

// mel.cu - PCM -> log-mel spectrogram, fused frame + Hann + 400-point DFT + power + mel filterbank +
// log10 in one kernel with shared-memory twiddle staging (BASELINE.json north_star stage 1;
// SURVEY.md §8a row a5, Appendix A.2 = whisper.cpp log_mel_spectrogram behind whisper.rs:75).
//
// Layout: one CTA owns 16 consecutive frames.  The windowed samples sit in shared memory frame-minor
// (xs[n][f]) so the DFT inner loop reads 16 frames with four broadcast LDS.128; thread k owns
// frequency bin k (0..200) and walks the twiddle table with stride k.  The filterbank is applied from
// shared memory with the same "groups of four float products accumulated in double" order as the
// reference implementation, restricted to each filter's non-zero range (zero groups add exactly 0).
#include <cmath>

#include "kernels.h"

namespace ss {

__constant__ float2 c_twiddle[kNFft];    // {cos, sin}(2 pi i / 400), computed on the host like the reference
static PerDeviceOnce g_twiddle_ready;      // (__constant__ tables are per device)

constexpr int kFramesPerCta = 16;
constexpr int kXsPitch = 16;
constexpr int kPwPitch = 204;

__device__ __forceinline__ int float_order_key(float x) { int b = __float_as_int(x); return b >= 0 ? b : b ^ 0x7fffffff; }
__device__ __forceinline__ float float_from_key(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }

__global__ void __launch_bounds__(256) mel_dft_kernel(const float *__restrict__ pcm, long n_samples, const float *__restrict__ filters,
                                                       const int2 *__restrict__ filt_range, int n_mels, float *__restrict__ mel,
                                                       int n_len, int n_calc, int *__restrict__ max_bits) {
    __shared__ __align__(16) float xs[kNFft * kXsPitch];
    __shared__ float2 tw[kNFft];
    __shared__ float pw[kFramesPerCta * kPwPitch];
    __shared__ float s_max[8];
    const int tid = threadIdx.x;
    const int i0 = blockIdx.x * kFramesPerCta;
    const float fill = (float)log10(1e-10);
    if (i0 >= n_calc) {   // frames past the audio: all-zero input, log10(1e-10)
        for (int o = tid; o < kFramesPerCta * n_mels; o += 256) {
            const int f = o % kFramesPerCta, j = o / kFramesPerCta;
            if (i0 + f < n_len) mel[(size_t)j * n_len + i0 + f] = fill;
        }
        if (tid == 0) atomicMax(max_bits, float_order_key(fill));
        return;
    }
    for (int i = tid; i < kNFft; i += 256) tw[i] = c_twiddle[i];
    __syncthreads();
    // stage windowed samples: padded[p] = reflect(pcm[200 - p]) for p < 200, pcm[p - 200], zeros after
    for (int e = tid; e < kFramesPerCta * kNFft; e += 256) {
        const int f = e / kNFft, j = e - f * kNFft;
        const long p = (long)(i0 + f) * kHop + j;
        float v = 0.f;
        if (i0 + f < n_calc) {
            const long s = p < kNFft / 2 ? (kNFft / 2 - p) : (p - kNFft / 2);
            if (s < n_samples) v = pcm[s];
        }
        const float hann = 0.5f * (1.0f - tw[j].x);
        xs[j * kXsPitch + f] = hann * v;
    }
    __syncthreads();
    if (tid < kNBins) {
        float re[kFramesPerCta], im[kFramesPerCta];
#pragma unroll
        for (int f = 0; f < kFramesPerCta; f++) { re[f] = 0.f; im[f] = 0.f; }
        int idx = 0;
        const int k = tid;
        for (int n = 0; n < kNFft; n++) {
            const float2 t = tw[idx];
            const float4 *xr = reinterpret_cast<const float4 *>(xs + n * kXsPitch);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const float4 x = xr[q];
                re[4 * q + 0] = fmaf(x.x, t.x, re[4 * q + 0]); im[4 * q + 0] = fmaf(-x.x, t.y, im[4 * q + 0]);
                re[4 * q + 1] = fmaf(x.y, t.x, re[4 * q + 1]); im[4 * q + 1] = fmaf(-x.y, t.y, im[4 * q + 1]);
                re[4 * q + 2] = fmaf(x.z, t.x, re[4 * q + 2]); im[4 * q + 2] = fmaf(-x.z, t.y, im[4 * q + 2]);
                re[4 * q + 3] = fmaf(x.w, t.x, re[4 * q + 3]); im[4 * q + 3] = fmaf(-x.w, t.y, im[4 * q + 3]);
            }
            idx += k; if (idx >= kNFft) idx -= kNFft;
        }
#pragma unroll
        for (int f = 0; f < kFramesPerCta; f++) pw[f * kPwPitch + k] = __fadd_rn(__fmul_rn(re[f], re[f]), __fmul_rn(im[f], im[f]));
    }
    __syncthreads();
    float lmax = -INFINITY;
    for (int o = tid; o < kFramesPerCta * n_mels; o += 256) {
        const int f = o % kFramesPerCta, j = o / kFramesPerCta;
        if (i0 + f >= n_len) continue;
        float out = fill;
        if (i0 + f < n_calc) {
            const float *fl = filters + (size_t)j * kNBins;
            const float *p = pw + f * kPwPitch;
            const int2 rg = filt_range[j];
            double sum = 0.0;
            int kb = rg.x & ~3;
            for (; kb < rg.y && kb < kNBins - 3; kb += 4) {
                float s = __fmul_rn(p[kb], fl[kb]);
                s = __fadd_rn(s, __fmul_rn(p[kb + 1], fl[kb + 1]));
                s = __fadd_rn(s, __fmul_rn(p[kb + 2], fl[kb + 2]));
                s = __fadd_rn(s, __fmul_rn(p[kb + 3], fl[kb + 3]));
                sum += (double)s;
            }
            if (rg.y > kNBins - 1) sum += (double)__fmul_rn(p[kNBins - 1], fl[kNBins - 1]);
            out = (float)log10(sum > 1e-10 ? sum : 1e-10);
        }
        mel[(size_t)j * n_len + i0 + f] = out;
        lmax = fmaxf(lmax, out);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
    if ((tid & 31) == 0) s_max[tid >> 5] = lmax;
    __syncthreads();
    if (tid == 0) {
        float m = s_max[0];
        for (int w = 1; w < 8; w++) m = fmaxf(m, s_max[w]);
        if (m > -INFINITY) atomicMax(max_bits, float_order_key(m));
    }
}

__global__ void __launch_bounds__(256) mel_norm_kernel(float *__restrict__ mel, size_t n, const int *__restrict__ max_bits) {
    const double mmax = (double)float_from_key(*max_bits) - 8.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float v = mel[i];
        if ((double)v < mmax) v = (float)mmax;
        mel[i] = (float)(((double)v + 4.0) / 4.0);
    }
}

__global__ void __launch_bounds__(256) mel_reset_max_kernel(int *max_bits) {
    if (threadIdx.x == 0 && blockIdx.x == 0) *max_bits = float_order_key(-1e20f);
}

// mel[n_mels][n_len] f32 -> win[rows = 2T + 2][n_mels] f16 with zero first / last row
__global__ void __launch_bounds__(256) mel_window_kernel(const float *__restrict__ mel, int n_len, int n_mels, int seek, int n_frames,
                                                          __half *__restrict__ win) {
    __shared__ float tile[32][33];
    const int f0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int c = c0 + r, fr = seek + f0 + tx;
        tile[r][tx] = (c < n_mels && f0 + tx < n_frames && fr < n_len) ? mel[(size_t)c * n_len + fr] : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int f = f0 + r, c = c0 + tx;
        if (f < n_frames && c < n_mels) win[(size_t)(f + 1) * n_mels + c] = __float2half_rn(tile[tx][r]);
    }
    if (blockIdx.x == 0 && threadIdx.x < 32) {
        const int c = c0 + threadIdx.x;
        if (c < n_mels) { win[c] = __float2half_rn(0.f); win[(size_t)(n_frames + 1) * n_mels + c] = __float2half_rn(0.f); }
    }
}

static void ensure_twiddles(int device) {
    if (!g_twiddle_ready.need(device)) return;
    float2 h[kNFft];
    for (int i = 0; i < kNFft; i++) {
        const double theta = (2.0 * M_PI * i) / kNFft;
        h[i].x = cosf((float)theta); h[i].y = sinf((float)theta);
    }
    CUDA_CHECK(cudaMemcpyToSymbol(c_twiddle, h, sizeof h));
    g_twiddle_ready.done(device);
}

void mel_enqueue(const Model &m, const float *d_pcm, size_t n_samples, float *d_mel, int n_len, int *d_max_bits,
                 cudaStream_t st, int *launches) {
    ensure_twiddles(m.device);
    const long n_eff = (long)n_samples + kNFft / 2;
    long n_calc = n_eff / kHop + 1;
    if (n_calc > n_len) n_calc = n_len;
    mel_reset_max_kernel<<<1, 32, 0, st>>>(d_max_bits);
    mel_dft_kernel<<<ceil_div(n_len, kFramesPerCta), 256, 0, st>>>(d_pcm, (long)n_samples, m.filters, m.filt_range, m.hp.n_mels, d_mel,
                                                                   n_len, (int)n_calc, d_max_bits);
    const size_t n = (size_t)m.hp.n_mels * n_len;
    mel_norm_kernel<<<(int)std::min<size_t>(ceil_div<size_t>(n, 256), 148 * 8), 256, 0, st>>>(d_mel, n, d_max_bits);
    *launches += 3;
}

void mel_window_enqueue(const Model &m, const float *d_mel, int n_len, int seek, __half *d_win, cudaStream_t st, int *launches) {
    const int n_frames = 2 * m.hp.n_audio_ctx;
    dim3 grid(ceil_div(n_frames, 32), ceil_div(m.hp.n_mels, 32));
    mel_window_kernel<<<grid, 256, 0, st>>>(d_mel, n_len, m.hp.n_mels, seek, n_frames, d_win);
    *launches += 1;
}

}  // namespace ss

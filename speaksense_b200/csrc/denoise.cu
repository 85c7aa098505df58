// denoise.cu - the reference's audio denoise (src/audio/mod.rs:507-735), the step immediately in front of the
// transcribe hot path on both callers (gRPC: grpc/handlers/asr.rs:196 per 5 s chunk; REST: StreamAudioProcessor per
// 2048-sample frame), on the device so that a stream chunk goes PCM -> denoise -> log-mel without leaving HBM.
// SURVEY.md §8 row f1.  Arithmetic follows the reference (f32, symmetric Hann, unnormalised inverse FFT, x10 after the
// overlap-add: SURVEY Appendix B.5); oracle: oracle/audio_oracle.c.
//
// Layout: one CTA per STFT frame, the frame lives in shared memory as interleaved complex f32; a 2048-point radix-2
// transform is 11 passes of 1024 butterflies over 256 threads with a twiddle table built once per CTA (sincospif).
// HBM traffic is trivial (a 5 s chunk is 320 KB; 153 overlapping frames): every kernel here is launch / latency bound,
// the point of the row is that the chunk stays resident.
#include "kernels.h"

namespace ss {

namespace {

constexpr int kDnThreads = 256;
constexpr int kDnMaxFrame = 4096;

__device__ __forceinline__ float hann_w(int i, int size) {      // audio/mod.rs:501-503
    return 0.5f * (1.0f - cosf(2.0f * 3.14159265358979323846f * (float)i / (float)(size - 1)));
}

struct DnSmem {
    float2 x[kDnMaxFrame];
    float2 tw[kDnMaxFrame / 2];
    float red[kDnThreads / 32];
};

__device__ __forceinline__ void build_twiddles(DnSmem &sm, int fs) {
    for (int k = threadIdx.x; k < fs / 2; k += kDnThreads) {
        float s, c;
        sincospif(-2.0f * (float)k / (float)fs, &s, &c);      // e^{-2 pi i k / fs}, exact argument reduction
        sm.tw[k] = make_float2(c, s);
    }
}

// in-place radix-2 DIT transform of sm.x (already in bit-reversed order); inverse = conjugated twiddles, no 1/N (rustfft)
__device__ void fft_passes(DnSmem &sm, int fs, int log2fs, bool inverse) {
    for (int st = 1; st <= log2fs; st++) {
        const int half = 1 << (st - 1), tw_step = fs >> st;
        __syncthreads();
        for (int b = threadIdx.x; b < fs / 2; b += kDnThreads) {
            const int k = b & (half - 1), i0 = ((b >> (st - 1)) << st) + k, i1 = i0 + half;
            float2 w = sm.tw[k * tw_step];
            if (inverse) w.y = -w.y;
            const float2 u = sm.x[i0], v = sm.x[i1];
            const float tr = v.x * w.x - v.y * w.y, ti = v.x * w.y + v.y * w.x;
            sm.x[i0] = make_float2(u.x + tr, u.y + ti);
            sm.x[i1] = make_float2(u.x - tr, u.y - ti);
        }
    }
    __syncthreads();
}
__device__ __forceinline__ int bitrev(int i, int log2fs) { return (int)(__brev((unsigned)i) >> (32 - log2fs)); }

// windowed forward transform of frame `src[0..fs)` into sm.x (natural order on return)
__device__ void load_windowed_fft(DnSmem &sm, const float *src, int fs, int log2fs) {
    for (int i = threadIdx.x; i < fs; i += kDnThreads) sm.x[bitrev(i, log2fs)] = make_float2(src[i] * hann_w(i, fs), 0.f);
    fft_passes(sm, fs, log2fs, false);
}

// ---- power spectra of the non-overlapping frames (analyze_noise_characteristics / estimate_*_spectrum) ----
__global__ void __launch_bounds__(kDnThreads) dn_power_kernel(const float *pcm, int fs, int log2fs, float *power) {
    extern __shared__ __align__(16) uint8_t dn_raw[];
    DnSmem &sm = *reinterpret_cast<DnSmem *>(dn_raw);
    build_twiddles(sm, fs);
    load_windowed_fft(sm, pcm + (size_t)blockIdx.x * fs, fs, log2fs);
    float *P = power + (size_t)blockIdx.x * fs;
    for (int i = threadIdx.x; i < fs; i += kDnThreads) { const float2 c = sm.x[i]; P[i] = c.x * c.x + c.y * c.y; }
}

// per bin: noise[i] = sum over the first 20 frames / 20, signal[i] = sum over all frames / n_frames (both in frame
// order, as the reference accumulates); per frame f >= 1: fvar[f] = sum_i (P[f][i] - P[f-1][i])^2 / fs
__global__ void __launch_bounds__(kDnThreads) dn_spectra_kernel(const float *power, int n_frames, int fs, float *noise, float *signal, float *fvar) {
    __shared__ float red[kDnThreads / 32];
    const int b = blockIdx.x;
    if (b == 0) {
        for (int i = threadIdx.x; i < fs; i += kDnThreads) {
            float nz = 0.f, sg = 0.f;
            for (int f = 0; f < n_frames; f++) {
                const float p = power[(size_t)f * fs + i];
                if (f < 20) nz += p / 20.0f;
                sg += p / (float)n_frames;
            }
            noise[i] = nz; signal[i] = sg;
        }
        return;
    }
    const int f = b;      // 1 .. n_frames-1
    float acc = 0.f;
    for (int i = threadIdx.x; i < fs; i += kDnThreads) { const float d = power[(size_t)f * fs + i] - power[(size_t)(f - 1) * fs + i]; acc += d * d; }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) { float t = 0.f; for (int w = 0; w < kDnThreads / 32; w++) t += red[w]; fvar[f] = t / (float)fs; }
}

// noise type: 0 stationary, 1 non-stationary, 2 mixed (audio/mod.rs:569-577)
__global__ void dn_classify_kernel(const float *fvar, int n_frames, size_t n_samples, int *type_out, float *nv_out) {
    float sv = 0.f;
    for (int f = 1; f < n_frames; f++) sv += fvar[f];
    const float nv = sv / (float)n_samples;
    *nv_out = nv;
    *type_out = nv < 0.1f ? 0 : nv > 0.5f ? 1 : 2;
}

// ---- one overlapping frame: window, FFT, gain (mode 0 spectral subtraction :597-616, mode 1 Wiener :644-652), inverse
//      FFT, window again -> frames[fi][j] = re * w  (the per-frame addend of overlap_add :718-727)
__global__ void __launch_bounds__(kDnThreads) dn_filter_kernel(const float *pcm, int fs, int log2fs, int step, float strength, int mode,
                                                               const float *noise, const float *signal, float *frames) {
    extern __shared__ __align__(16) uint8_t dn_raw[];
    DnSmem &sm = *reinterpret_cast<DnSmem *>(dn_raw);
    build_twiddles(sm, fs);
    load_windowed_fft(sm, pcm + (size_t)blockIdx.x * step, fs, log2fs);
    // gain in natural order, then scatter to bit-reversed order for the inverse passes (through registers: in-place permutation)
    float2 v[kDnMaxFrame / kDnThreads];
#pragma unroll
    for (int q = 0; q < kDnMaxFrame / kDnThreads; q++) {
        const int i = threadIdx.x + q * kDnThreads;
        if (i < fs) {
            const float2 c = sm.x[i];
            float gain;
            if (mode == 0) {
                const float power = c.x * c.x + c.y * c.y;
                const float freq_factor = fminf((float)i / (float)fs, 1.0f);
                const float freq_strength = strength * (1.0f - 0.3f * freq_factor);
                gain = sqrtf(fmaxf(1.0f - 1.0f * powf(noise[i] / (power + 1e-6f), freq_strength), 0.1f));
            } else {
                const float snr = signal[i] / (noise[i] + 1e-6f);
                gain = powf(snr / (1.0f + snr), strength * 0.7f);
            }
            v[q] = make_float2(c.x * gain, c.y * gain);
        }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < kDnMaxFrame / kDnThreads; q++) { const int i = threadIdx.x + q * kDnThreads; if (i < fs) sm.x[bitrev(i, log2fs)] = v[q]; }
    fft_passes(sm, fs, log2fs, true);
    float *F = frames + (size_t)blockIdx.x * fs;
    for (int j = threadIdx.x; j < fs; j += kDnThreads) F[j] = sm.x[j].x * hann_w(j, fs);
}

// overlap-add (audio/mod.rs:711-735): frames are summed in frame order (deterministic), normalised by the summed squared
// window, x10; samples no frame covers stay 0
__global__ void dn_ola_kernel(const float *frames, int n_fr, int fs, int step, size_t n, float *out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    long f_hi = (long)(i / step); if (f_hi > n_fr - 1) f_hi = n_fr - 1;
    long f_lo = (long)i - (fs - 1) <= 0 ? 0 : ((long)i - (fs - 1) + step - 1) / step;
    float o = 0.f, nm = 0.f;
    for (long f = f_lo; f <= f_hi; f++) {
        const int j = (int)((long)i - f * step);
        const float w = hann_w(j, fs);
        o += frames[(size_t)f * fs + j];
        nm += w * w;
    }
    out[i] = nm > 1e-10f ? (o / nm) * 10.0f : o;
}

// ---- StreamAudioProcessor frames (audio/mod.rs:111-141 steps 4-5), many at once.  denoise_audio on ONE frame of exactly
// frame_size samples always takes the same route: a single analysis frame has no predecessor, so the spectral variance
// is 0 -> Stationary -> spectral_subtraction with noise = |X|^2 / 20 (the frame is its own and only noise frame, divided
// by the constant 20) and a single STFT window; overlap_add of one window divides by w^2 where w^2 > 1e-10.  Frames are
// independent, so one CTA takes one frame; the noise gate (:495-499) is folded in.
__global__ void __launch_bounds__(kDnThreads) dn_frames_kernel(const float *in, int fs, int log2fs, float strength, float noise_gate, float *out) {
    extern __shared__ __align__(16) uint8_t dn_raw[];
    DnSmem &sm = *reinterpret_cast<DnSmem *>(dn_raw);
    build_twiddles(sm, fs);
    load_windowed_fft(sm, in + (size_t)blockIdx.x * fs, fs, log2fs);
    float2 v[kDnMaxFrame / kDnThreads];
#pragma unroll
    for (int q = 0; q < kDnMaxFrame / kDnThreads; q++) {
        const int i = threadIdx.x + q * kDnThreads;
        if (i < fs) {
            const float2 c = sm.x[i];
            const float power = c.x * c.x + c.y * c.y, noise = power / 20.0f;
            const float freq_factor = fminf((float)i / (float)fs, 1.0f);
            const float freq_strength = strength * (1.0f - 0.3f * freq_factor);
            const float gain = sqrtf(fmaxf(1.0f - 1.0f * powf(noise / (power + 1e-6f), freq_strength), 0.1f));
            v[q] = make_float2(c.x * gain, c.y * gain);
        }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < kDnMaxFrame / kDnThreads; q++) { const int i = threadIdx.x + q * kDnThreads; if (i < fs) sm.x[bitrev(i, log2fs)] = v[q]; }
    fft_passes(sm, fs, log2fs, true);
    float *O = out + (size_t)blockIdx.x * fs;
    for (int j = threadIdx.x; j < fs; j += kDnThreads) {
        const float w = hann_w(j, fs), o = sm.x[j].x * w, nm = w * w;
        const float r = nm > 1e-10f ? (o / nm) * 10.0f : o;
        O[j] = fabsf(r) < noise_gate ? 0.0f : r;
    }
}

}  // namespace

void denoise_frames_enqueue(const float *d_in, int n_frames, int fs, float strength, float noise_gate, float *d_out, cudaStream_t st, int *launches) {
    if (fs < 64 || fs > kDnMaxFrame || (fs & (fs - 1))) SS_THROW(-1, "denoise: frame_size must be a power of two in [64, %d]", kDnMaxFrame);
    if (n_frames <= 0) return;
    int log2fs = 0; while ((1 << log2fs) < fs) log2fs++;
    static PerDeviceOnce once;      // (function attributes are per device)
    const int dev = current_device();
    if (once.need(dev)) { CUDA_CHECK(cudaFuncSetAttribute(dn_frames_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DnSmem))); once.done(dev); }
    dn_frames_kernel<<<n_frames, kDnThreads, sizeof(DnSmem), st>>>(d_in, fs, log2fs, strength, noise_gate, d_out);
    *launches += 1;
    CUDA_CHECK(cudaGetLastError());
}

size_t denoise_scratch_floats(size_t n, int fs, float overlap) {
    const int step = (int)((float)fs * (1.0f - overlap));
    const size_t nf = n / fs, nov = n >= (size_t)fs ? (n - fs) / step + 1 : 0;
    return nf * fs + 2 * (size_t)fs + nf + 8 + nov * fs + n;      // power | noise, signal | fvar | type, nv | frames | tmp
}

// Runs denoise_audio(d_in[0..n)) -> d_out on `st`.  `scratch` holds denoise_scratch_floats(n, ...) floats.
// Returns the noise type (one 8-byte D2H read: the reference branches on it too).
int denoise_enqueue(const float *d_in, size_t n, int fs, float overlap, float strength, float *d_out, float *scratch,
                    cudaStream_t st, int *launches, float *nv_out) {
    if (fs < 64 || fs > kDnMaxFrame || (fs & (fs - 1))) SS_THROW(-1, "denoise: frame_size must be a power of two in [64, %d]", kDnMaxFrame);
    if (n < (size_t)fs) SS_THROW(-1, "denoise: fewer samples (%zu) than one frame (%d) - the reference panics here", n, fs);
    const int step = (int)((float)fs * (1.0f - overlap));
    if (step < 1) SS_THROW(-1, "denoise: overlap leaves no hop");
    int log2fs = 0; while ((1 << log2fs) < fs) log2fs++;
    const int nf = (int)(n / fs), nov = (int)((n - fs) / step + 1);
    float *power = scratch, *noise = power + (size_t)nf * fs, *signal = noise + fs, *fvar = signal + fs;
    int *d_type = reinterpret_cast<int *>(fvar + nf); float *d_nv = fvar + nf + 1;
    float *frames = fvar + nf + 8, *tmp = frames + (size_t)nov * fs;
    static PerDeviceOnce once;      // (function attributes are per device)
    const int dev = current_device();
    if (once.need(dev)) {
        CUDA_CHECK(cudaFuncSetAttribute(dn_power_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DnSmem)));
        CUDA_CHECK(cudaFuncSetAttribute(dn_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DnSmem)));
        once.done(dev);
    }
    auto spectra = [&](const float *src) {
        dn_power_kernel<<<nf, kDnThreads, sizeof(DnSmem), st>>>(src, fs, log2fs, power);
        dn_spectra_kernel<<<nf, kDnThreads, 0, st>>>(power, nf, fs, noise, signal, fvar);
        *launches += 2;
    };
    auto pass = [&](const float *src, float *dst, int mode) {
        dn_filter_kernel<<<nov, kDnThreads, sizeof(DnSmem), st>>>(src, fs, log2fs, step, strength, mode, noise, signal, frames);
        dn_ola_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(frames, nov, fs, step, n, dst);
        *launches += 2;
    };
    spectra(d_in);
    dn_classify_kernel<<<1, 1, 0, st>>>(fvar, nf, n, d_type, d_nv);
    *launches += 1;
    struct { int type; float nv; } h;
    CUDA_CHECK(cudaMemcpyAsync(&h, d_type, 8, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    if (nv_out) *nv_out = h.nv;
    if (h.type == 0) pass(d_in, d_out, 0);
    else if (h.type == 1) pass(d_in, d_out, 1);
    else { pass(d_in, tmp, 0); spectra(tmp); pass(tmp, d_out, 1); }
    CUDA_CHECK(cudaGetLastError());
    return h.type;
}

}  // namespace ss

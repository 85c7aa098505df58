/* audio_oracle.c - CPU restatement of the reference's audio pre-processing that runs immediately in
 * front of the transcribe hot path.  TEST INFRASTRUCTURE ONLY: linked into
 * oracle/_build/libwhisper_oracle.so, used by tests/ as the checker of the CUDA path, never by the product.
 *
 * Restates /root/reference/src/audio/mod.rs (Rust, f32 arithmetic, rustfft):
 *   denoise_audio                  :507-528   (called per 5 s chunk by grpc/handlers/asr.rs:196)
 *   analyze_noise_characteristics  :533-578
 *   spectral_subtraction           :581-623
 *   wiener_filter                  :626-662
 *   estimate_noise_spectrum        :665-685   (first 20 non-overlapping frames, always divided by 20)
 *   estimate_signal_spectrum       :688-708
 *   overlap_add                    :711-735   (Hann-weighted, normalised by the summed squared window, x10)
 *   hann_window                    :501-503   (symmetric, size-1 in the denominator)
 *   StreamAudioProcessor           :80-155    (REST path: per 2048-sample frame; see the quirks below)
 *   normalize_audio :408-411, preemphasis :260-269, apply_noise_gate :495-499, estimate_noise_floor :744-762
 *
 * Quirks kept on purpose: rustfft's inverse transform is unnormalised, so the
 * output is scaled by frame_size x 10; the first-frame noise floor is 0/0 = NaN, after which the VAD gain
 * is max(NaN, 0.1) = 0.1 for ever; an all-zero chunk normalises to NaN.
 *
 * The FFT here is a plain f32 radix-2 transform: rustfft's mixed-radix butterflies round differently, so
 * parity with the reference itself (not runnable offline) would be to ~1e-6 relative, not bit-exact. */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

typedef struct { float re, im; } cpx;

static float hann_w(int i, int size) {   /* mod.rs:501-503 (f32 arithmetic) */
    return 0.5f * (1.0f - cosf(2.0f * 3.14159265358979323846f * (float)i / (float)(size - 1)));
}

/* in-place radix-2 decimation-in-time FFT, n a power of two; inverse = unnormalised (rustfft convention) */
static void fft_r2(cpx *x, int n, int inverse) {
    for (int i = 1, j = 0; i < n; i++) {
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { cpx t = x[i]; x[i] = x[j]; x[j] = t; }
    }
    for (int len = 2; len <= n; len <<= 1) {
        const double ang = (inverse ? 2.0 : -2.0) * 3.14159265358979323846 / len;
        for (int i = 0; i < n; i += len)
            for (int k = 0; k < len / 2; k++) {
                const float wr = (float)cos(ang * k), wi = (float)sin(ang * k);
                const cpx u = x[i + k], v = x[i + k + len / 2];
                const float tr = v.re * wr - v.im * wi, ti = v.re * wi + v.im * wr;
                x[i + k].re = u.re + tr; x[i + k].im = u.im + ti;
                x[i + k + len / 2].re = u.re - tr; x[i + k + len / 2].im = u.im - ti;
            }
    }
}

static void windowed_fft(const float *frame, int fs, cpx *buf) {
    for (int i = 0; i < fs; i++) { buf[i].re = frame[i] * hann_w(i, fs); buf[i].im = 0.f; }
    fft_r2(buf, fs, 0);
}

/* mod.rs:533-578.  Returns 0 stationary, 1 non-stationary, 2 mixed; *nv_out = the normalised variance */
int ao_analyze_noise(const float *s, size_t n, int fs, float *nv_out) {
    cpx *buf = malloc(sizeof(cpx) * fs);
    float *prev = malloc(sizeof(float) * fs), *cur = malloc(sizeof(float) * fs);
    int have_prev = 0;
    float spectral_variance = 0.f;
    for (size_t off = 0; off + fs <= n; off += fs) {
        windowed_fft(s + off, fs, buf);
        for (int i = 0; i < fs; i++) cur[i] = buf[i].re * buf[i].re + buf[i].im * buf[i].im;
        if (have_prev) {
            float acc = 0.f;
            for (int i = 0; i < fs; i++) { const float dlt = cur[i] - prev[i]; acc += dlt * dlt; }
            spectral_variance += acc / (float)fs;
        }
        float *t = prev; prev = cur; cur = t; have_prev = 1;
    }
    free(buf); free(prev); free(cur);
    const float nv = spectral_variance / (float)n;
    if (nv_out) *nv_out = nv;
    return nv < 0.1f ? 0 : nv > 0.5f ? 1 : 2;
}

/* mod.rs:665-685 (take = 20, divide by 20) and :688-708 (all frames, divide by n / fs) */
static void mean_power(const float *s, size_t n, int fs, int take, float denom, float *out) {
    cpx *buf = malloc(sizeof(cpx) * fs);
    memset(out, 0, sizeof(float) * fs);
    int f = 0;
    for (size_t off = 0; off + fs <= n && (take < 0 || f < take); off += fs, f++) {
        windowed_fft(s + off, fs, buf);
        for (int i = 0; i < fs; i++) out[i] += (buf[i].re * buf[i].re + buf[i].im * buf[i].im) / denom;
    }
    free(buf);
}

/* shared body of spectral_subtraction (mode 0) and wiener_filter (mode 1) + overlap_add */
static void stft_filter(const float *s, size_t n, int fs, float overlap, float strength, int mode, float *out) {
    const int step = (int)((float)fs * (1.0f - overlap));
    float *noise = malloc(sizeof(float) * fs), *signal = malloc(sizeof(float) * fs);
    mean_power(s, n, fs, 20, 20.0f, noise);
    if (mode == 1) mean_power(s, n, fs, -1, (float)(n / fs), signal);
    float *norm = calloc(n, sizeof(float));
    memset(out, 0, sizeof(float) * n);
    cpx *buf = malloc(sizeof(cpx) * fs);
    size_t fi = 0;
    for (size_t start = 0; start + fs <= n; start += step, fi++) {      /* samples.windows(fs).step_by(step) */
        windowed_fft(s + start, fs, buf);
        for (int i = 0; i < fs; i++) {
            float gain;
            if (mode == 0) {
                const float power = buf[i].re * buf[i].re + buf[i].im * buf[i].im;
                const float freq_factor = fminf((float)i / (float)fs, 1.0f);
                const float freq_strength = strength * (1.0f - 0.3f * freq_factor);
                gain = sqrtf(fmaxf(1.0f - 1.0f * powf(noise[i] / (power + 1e-6f), freq_strength), 0.1f));
            } else {
                const float snr = signal[i] / (noise[i] + 1e-6f);
                gain = powf(snr / (1.0f + snr), strength * 0.7f);
            }
            buf[i].re *= gain; buf[i].im *= gain;
        }
        fft_r2(buf, fs, 1);
        for (int j = 0; j < fs; j++) {
            if (start + j < n) {
                const float w = hann_w(j, fs);
                out[start + j] += buf[j].re * w;
                norm[start + j] += w * w;
            }
        }
    }
    for (size_t i = 0; i < n; i++) if (norm[i] > 1e-10f) out[i] = (out[i] / norm[i]) * 10.0f;
    free(noise); free(signal); free(norm); free(buf);
}

/* mod.rs:507-528.  Returns the noise type that was chosen (0/1/2), or -1 when n < frame_size (the
 * reference panics there: overlap_add indexes frames[0] of an empty list). */
int ao_denoise_audio(const float *s, size_t n, int fs, float overlap, float strength, float *out) {
    if (n < (size_t)fs) return -1;
    const int type = ao_analyze_noise(s, n, fs, NULL);
    if (type == 0) stft_filter(s, n, fs, overlap, strength, 0, out);
    else if (type == 1) stft_filter(s, n, fs, overlap, strength, 1, out);
    else {
        float *tmp = malloc(sizeof(float) * n);
        stft_filter(s, n, fs, overlap, strength, 0, tmp);
        stft_filter(tmp, n, fs, overlap, strength, 1, out);
        free(tmp);
    }
    return type;
}

/* ---- StreamAudioProcessor (mod.rs:80-155): state across frames = {noise_floor, prev_energy}; the caller owns the
 * re-framing buffer (process_chunk :92-109 drains whole 2048-sample frames; finish :143-155 zero-pads the tail). */
static int cmp_f32(const void *a, const void *b) { const float x = *(const float *)a, y = *(const float *)b; return (x > y) - (x < y); }
float ao_estimate_noise_floor(const float *s, size_t n) {      /* mod.rs:744-762 */
    const size_t nf = (n + 1023) / 1024;
    float *e = malloc(sizeof(float) * (nf ? nf : 1));
    for (size_t f = 0; f < nf; f++) {
        const size_t a = f * 1024, b = a + 1024 < n ? a + 1024 : n;
        float acc = 0.f;
        for (size_t i = a; i < b; i++) acc += s[i] * s[i];
        e[f] = acc / (float)(b - a);
    }
    qsort(e, nf, sizeof(float), cmp_f32);
    const size_t cnt = (size_t)((float)nf * 0.1f);
    float sum = 0.f;
    for (size_t i = 0; i < cnt; i++) sum += e[i];
    free(e);
    return sum / (float)cnt;      /* cnt == 0 for a 2048-sample frame: 0/0 = NaN, as in the reference */
}
void ao_normalize_audio(const float *s, size_t n, float *out) {      /* mod.rs:408-411 */
    float m = n ? fabsf(s[0]) : 1.0f;
    for (size_t i = 1; i < n; i++) if (fabsf(s[i]) > m) m = fabsf(s[i]);
    for (size_t i = 0; i < n; i++) out[i] = s[i] / m;
}
/* process_frame :111-141.  state[0] = noise_floor, state[1] = prev_energy (both updated) */
void ao_process_frame(const float *frame, int fs, float overlap, float strength, float noise_gate, int enable_nr, float *state, float *out) {
    if (state[0] == 0.0f) state[0] = ao_estimate_noise_floor(frame, (size_t)fs);      /* :102-104 */
    float energy = frame[0] * frame[0];
    for (int i = 1; i < fs; i++) { const float p = frame[i] - 0.97f * frame[i - 1]; energy += p * p; }      /* preemphasis :260-269 */
    energy /= (float)fs;
    const float threshold = state[0] * 1.2f + state[1] * 0.1f;
    float gain;
    if (energy > threshold) gain = 1.0f;
    else { const float r = energy / threshold; gain = (r != r) ? 0.1f : fmaxf(r, 0.1f); }      /* f32::max ignores NaN */
    state[1] = energy;
    { const float mn = (state[0] != state[0]) ? energy : fminf(energy, state[0]); state[0] = state[0] * 0.95f + mn * 0.05f; }
    float *tmp = malloc(sizeof(float) * fs);
    for (int i = 0; i < fs; i++) tmp[i] = frame[i] * gain;
    if (enable_nr) ao_denoise_audio(tmp, (size_t)fs, fs, overlap, strength, out);
    else memcpy(out, tmp, sizeof(float) * fs);
    for (int i = 0; i < fs; i++) if (fabsf(out[i]) < noise_gate) out[i] = 0.0f;      /* :495-499 */
    free(tmp);
}

"""Independent float64 / numpy.fft restatement of the reference's denoise_audio (src/audio/mod.rs:507-735),
used only to pin oracle/audio_oracle.c (different FFT, different precision, same algorithm)."""
import numpy as np


def hann(fs):
    i = np.arange(fs)
    return 0.5 * (1 - np.cos(2 * np.pi * i / (fs - 1)))


def denoise_ref(x, fs=2048, overlap=0.75, strength=0.2):
    x = np.asarray(x, np.float64)
    n, w, step = len(x), hann(fs), int(fs * (1 - overlap))

    def powers(s):
        return [np.abs(np.fft.fft(s[o:o + fs] * w)) ** 2 for o in range(0, n - fs + 1, fs)]

    def analyze(s):
        P = powers(s)
        nv = sum(((P[i] - P[i - 1]) ** 2).sum() / fs for i in range(1, len(P))) / n
        return (0 if nv < 0.1 else 1 if nv > 0.5 else 2), nv

    def filt(s, mode):
        P = powers(s)
        noise, signal = sum(P[:20]) / 20.0, sum(P) / (n // fs)
        out, norm, i = np.zeros(n), np.zeros(n), np.arange(fs)
        for st in range(0, n - fs + 1, step):
            X = np.fft.fft(s[st:st + fs] * w)
            if mode == 0:
                fsr = strength * (1 - 0.3 * np.minimum(i / fs, 1.0))
                g = np.sqrt(np.maximum(1 - (noise / (np.abs(X) ** 2 + 1e-6)) ** fsr, 0.1))
            else:
                snr = signal / (noise + 1e-6)
                g = (snr / (1 + snr)) ** (strength * 0.7)
            y = np.fft.ifft(X * g) * fs                      # rustfft's inverse is unnormalised
            out[st:st + fs] += y.real * w
            norm[st:st + fs] += w * w
        m = norm > 1e-10
        out[m] = out[m] / norm[m] * 10
        return out

    t, nv = analyze(x)
    if t == 0:
        return filt(x, 0), t, nv
    if t == 1:
        return filt(x, 1), t, nv
    return filt(filt(x, 0), 1), t, nv

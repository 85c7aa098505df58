"""CPU: the audio oracle (oracle/audio_oracle.c = src/audio/mod.rs restated in f32) against an independent
float64 numpy restatement, for all three noise-type branches, plus the StreamAudioProcessor quirks
(SURVEY Appendix B.5).  The reference holds no golden vectors for this path (audio/mod.rs:880-1056 only
prints statistics), so this pin is the strongest available offline."""
import numpy as np
import pytest

from tests.audio_ref import denoise_ref

SCALES = {"Stationary": 0.1, "Mixed": 0.243, "NonStationary": 1.0}      # spectral variance grows with amplitude^4


@pytest.mark.parametrize("kind", list(SCALES))
@pytest.mark.parametrize("n", [80000, 2048, 20000])
def test_denoise_oracle_vs_float64(oracle_mod, audio30, kind, n):
    x = (audio30[:n] * SCALES[kind]).astype(np.float32)
    got, t = oracle_mod.denoise_audio(x)
    ref, t_ref, nv_ref = denoise_ref(x)
    _, nv = oracle_mod.analyze_noise(x)
    assert t == t_ref
    if n == 80000:
        assert ("Stationary", "NonStationary", "Mixed")[t] == kind
    assert abs(nv - nv_ref) <= 1e-4 * max(nv_ref, 1e-12) + 1e-12
    scale = np.abs(ref).max()
    assert np.abs(got - ref).max() <= 2e-3 * scale
    # the reference's scaling quirk: unnormalised inverse FFT (x frame_size) and x10 in overlap_add
    assert scale > 1000 * np.abs(x).max()
    # samples behind the last whole STFT window are never covered: they stay 0 (mod.rs:718-733)
    last = ((n - 2048) // 512) * 512 + 2048
    assert np.all(got[last:] == 0)


def test_denoise_too_short_is_an_error(oracle_mod):
    with pytest.raises(ValueError):
        oracle_mod.denoise_audio(np.zeros(100, np.float32))


def test_stream_processor_quirks(oracle_mod, audio30):
    """first-frame noise floor = 0/0 = NaN -> VAD gain 0.1 for every frame; frames are 2048 samples; the tail is zero-padded"""
    sp = oracle_mod.StreamAudioProcessor()
    chunk = audio30[:5000]
    frames = sp.process_chunk(chunk)
    assert len(frames) == 2 and all(f.shape == (2048,) for f in frames)
    assert np.isnan(sp.state[0])
    norm = (chunk / np.abs(chunk).max()).astype(np.float32)
    for k, f in enumerate(frames):
        want, _ = oracle_mod.denoise_audio((norm[k * 2048:(k + 1) * 2048] * np.float32(0.1)).astype(np.float32))
        want[np.abs(want) < 0.003] = 0
        np.testing.assert_allclose(f, want, rtol=0, atol=1e-3 * np.abs(want).max())
    tail = sp.finish()
    assert len(tail) == 1 and sp.finish() == []
    # all-zero chunk: 0/0 -> NaN samples, as in the reference
    sp2 = oracle_mod.StreamAudioProcessor()
    with np.errstate(invalid="ignore"):
        out = sp2.process_chunk(np.zeros(2048, np.float32))
    assert np.isnan(out[0]).all()

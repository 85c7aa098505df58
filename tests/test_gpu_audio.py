"""GPU: csrc/denoise.cu through the C ABI (ss_denoise_audio) against the CPU oracle (oracle/audio_oracle.c) -
SURVEY.md §8 row f1: the denoise that both callers run right before transcribe_with_state."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SCALES = {"Stationary": 0.1, "Mixed": 0.243, "NonStationary": 1.0}
TOL = 2e-3      # of max |out|: f32 FFTs with different butterfly order / twiddles (the oracle itself is 4e-4 off float64)


@pytest.fixture(scope="module")
def eng_state(tiny_en_peaked):
    from speaksense_b200 import WhisperAsr
    eng = WhisperAsr(tiny_en_peaked, device=0)
    st = eng.create_state()
    yield eng, st
    st.close(); eng.close()


@pytest.mark.parametrize("kind", list(SCALES))
@pytest.mark.parametrize("n", [80000, 2048, 20000, 480000])
def test_denoise_matches_oracle(eng_state, oracle_mod, audio30, kind, n):
    from speaksense_b200 import denoise_audio
    eng, st = eng_state
    x = (audio30[:n] * SCALES[kind]).astype(np.float32)
    want, t = oracle_mod.denoise_audio(x)
    _, nv = oracle_mod.analyze_noise(x)
    got, name, nv_gpu = denoise_audio(eng, st, x)
    assert name == ("Stationary", "NonStationary", "Mixed")[t]
    assert abs(nv_gpu - nv) <= 1e-3 * nv + 1e-12
    assert np.abs(got - want).max() <= TOL * np.abs(want).max()
    last = ((n - 2048) // 512) * 512 + 2048
    assert np.all(got[last:] == 0)
    # idempotent on the same input (no state leaks between calls; deterministic overlap-add order)
    again, _, _ = denoise_audio(eng, st, x)
    assert np.array_equal(got, again)


def test_denoise_other_frame_sizes_and_errors(eng_state, oracle_mod, audio30):
    from speaksense_b200 import DenoiseConfig, NativeError, denoise_audio
    eng, st = eng_state
    x = audio30[:30000]
    for fs, ov in ((1024, 0.75), (512, 0.5), (4096, 0.75)):
        want, t = oracle_mod.denoise_audio(x, frame_size=fs, overlap=ov, strength=0.35)
        got, name, _ = denoise_audio(eng, st, x, DenoiseConfig(frame_size=fs, overlap=ov, strength=0.35))
        assert name == ("Stationary", "NonStationary", "Mixed")[t]
        # strength 0.35 steepens the gain curve next to its 0.1 clamp: f32 differences of the two FFTs are amplified more
        assert np.abs(got - want).max() <= 2.5 * TOL * np.abs(want).max(), (fs, ov)
    with pytest.raises(NativeError):
        denoise_audio(eng, st, x[:1000])                                   # shorter than a frame: the reference panics
    with pytest.raises(NativeError):
        denoise_audio(eng, st, x, DenoiseConfig(frame_size=1000))           # not a power of two


def test_stream_processor_matches_oracle(eng_state, oracle_mod, audio30):
    """REST path: StreamAudioProcessor over ragged chunks, incl. the NaN-noise-floor quirk (gain 0.1 for ever)"""
    from speaksense_b200 import StreamAudioProcessor
    from speaksense_b200.audio import collect_frames
    eng, st = eng_state
    frames, cb = collect_frames()
    sp = StreamAudioProcessor(eng, st, None, cb)
    ref = oracle_mod.StreamAudioProcessor()
    want = []
    pos = 0
    for size in (5000, 300, 4096, 7777):
        chunk = audio30[pos:pos + size]; pos += size
        sp.process_chunk(chunk)
        want += ref.process_chunk(chunk)
    sp.finish()
    want += ref.finish()
    assert len(frames) == len(want) == (5000 + 300 + 4096 + 7777 + 2047) // 2048
    assert np.isnan(sp.noise_floor) and np.isnan(ref.state[0])
    # A single 2048-sample frame is one Hann window: overlap_add divides by w^2, so towards the frame edges (w -> 0) both
    # implementations amplify their own FFT rounding noise without bound (a property of the reference, Appendix B.5).
    # Compare where the window is not tiny; the edges only have to be finite.
    i = np.arange(2048)
    inner = 0.5 * (1 - np.cos(2 * np.pi * i / 2047)) >= 0.05
    for a, b in zip(frames, want):
        assert np.isfinite(a).all()
        m = np.abs(b[inner]).max()
        # samples within the tolerance of the noise gate may be zeroed on one side only
        near_gate = np.abs(np.abs(b) - 0.003) <= TOL * m
        assert np.all((np.abs(a - b) <= TOL * m) | near_gate | ~inner)


def test_denoised_chunk_stays_resident(eng_state, audio30):
    """gRPC flow (grpc/handlers/asr.rs:196-198): denoise_audio then transcribe_with_state.  The resident chain
    (no second upload) gives the same transcript as uploading the fetched denoised samples again."""
    from speaksense_b200 import AsrParams, denoise_audio
    eng, st = eng_state
    x = audio30[:80000]
    p = AsrParams(stream_mode=True)
    den, _, _ = denoise_audio(eng, st, x)
    r1 = eng.transcribe_resident(st, p)
    t1, _ = st.result_tokens()
    r2 = eng.transcribe_with_state(st, den, p)
    t2, _ = st.result_tokens()
    assert t1 == t2 and r1.full_text == r2.full_text

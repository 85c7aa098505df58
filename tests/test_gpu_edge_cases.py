"""Edge cases of the transcribe path through the reference-shaped API, CUDA vs oracle: ragged clip
lengths (gRPC 5 s chunks, clips that spill into a second window), the longest supported input (5-minute
stream), languages, the tinydiarize flag, several audio seeds, idempotence."""
import numpy as np
import pytest

from tests.rust_post import post_process

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pair_tiny(oracle_mod, tiny_en_peaked):
    from speaksense_b200 import WhisperAsr
    om = oracle_mod.OracleModel(tiny_en_peaked)
    eng = WhisperAsr(tiny_en_peaked)
    yield om, eng
    eng.close(); om.close()


@pytest.fixture(scope="module")
def pair_micro(oracle_mod, micro_v3_peaked):
    from speaksense_b200 import WhisperAsr
    om = oracle_mod.OracleModel(micro_v3_peaked)
    eng = WhisperAsr(micro_v3_peaked)
    yield om, eng
    eng.close(); om.close()


def _compare(om, eng, pcm, **kw):
    from speaksense_b200 import AsrParams
    ost = om.new_state()
    ref = ost.full(pcm, language=kw.get("language"), stream_mode=kw.get("stream_mode", False),
                   speaker_diarization=kw.get("speaker_diarization", False))
    ost.close()
    st = eng.create_state()
    res = eng.transcribe_with_state(st, pcm, AsrParams(language=kw.get("language"), stream_mode=kw.get("stream_mode", False),
                                                        speaker_diarization=kw.get("speaker_diarization", False)))
    toks, _ = st.result_tokens()
    assert toks == ref["tokens"]
    raw = st.raw_segments()
    assert [(s["t0"], s["t1"], s["text"], s["speaker_turn_next"]) for s in raw] == \
           [(s["t0"], s["t1"], s["text"], s["speaker_turn_next"]) for s in ref["segments"]]
    exp = post_process(ref["segments"], kw.get("stream_mode", False))
    assert [(s.text, s.speaker_id, s.start, s.end) for s in res.segments] == exp["segments"]
    assert res.full_text == exp["full_text"]
    assert st.stats()["n_windows"] == ref["n_windows"]
    st.close()
    return ref


@pytest.mark.parametrize("n_samples", [16400, 80000, 160000 + 7, 479999, 480001, 31 * 16000, 959_840])
def test_ragged_lengths(pair_tiny, n_samples):
    from speaksense_b200 import synth
    om, eng = pair_tiny
    _compare(om, eng, synth.synth_audio(n_samples, seed=n_samples % 97), stream_mode=True)


def test_five_minute_stream(pair_tiny):
    """4 800 000 samples = the per-stream size of BASELINE config 5: ten 30 s windows on one state."""
    from speaksense_b200 import synth
    om, eng = pair_tiny
    ref = _compare(om, eng, synth.synth_audio(300 * 16000, seed=5000), stream_mode=True)
    assert ref["n_windows"] >= 10


@pytest.mark.parametrize("lang", ["en", "zh", "ja"])       # the languages the REST path accepts (processors/transcribe.rs:200-204)
@pytest.mark.parametrize("seed", [1234, 1235])
def test_languages_and_seeds(pair_micro, lang, seed):
    from speaksense_b200 import synth
    om, eng = pair_micro
    _compare(om, eng, synth.synth_audio(seed=seed), language=lang, stream_mode=False)


def test_speaker_diarization_flag(pair_micro, audio30):
    """speaker_diarization=true only stops suppressing the speaker-turn token (whisper.rs:137-140); with a
    non-tinydiarize script the transcript is unchanged and every speaker_id stays 0."""
    om, eng = pair_micro
    _compare(om, eng, audio30, language="zh", speaker_diarization=True)


def test_idempotent_and_deterministic(pair_tiny, audio30):
    from speaksense_b200 import AsrParams
    om, eng = pair_tiny
    p = AsrParams(stream_mode=False)
    a, b = eng.create_state(), eng.create_state()
    r1 = eng.transcribe_with_state(a, audio30, p)
    r2 = eng.transcribe_with_state(b, audio30, p)
    assert r1 == r2 and a.result_tokens() == b.result_tokens()
    # stream_mode=false keeps context across calls on the same state (no_context=false): the second call on `a`
    # is prompted with the first call's tokens, exactly like the oracle
    ost = om.new_state()
    ost.full(audio30, stream_mode=False)
    ref2 = ost.full(audio30, stream_mode=False)
    eng.transcribe_with_state(a, audio30, p)
    assert a.result_tokens()[0] == ref2["tokens"]
    ost.close(); a.close(); b.close()


def test_non_finite_and_loud_audio_do_not_crash(pair_tiny):
    om, eng = pair_tiny
    from speaksense_b200 import AsrParams
    rng = np.random.default_rng(0)
    loud = (rng.standard_normal(48000) * 50).astype(np.float32)      # far outside [-1, 1]
    res = eng.transcribe(loud, AsrParams(stream_mode=True))
    assert isinstance(res.full_text, str)
    silent = np.zeros(64000, np.float32)
    st = eng.create_state()
    ost = om.new_state()
    ref = ost.full(silent, stream_mode=True)
    eng.transcribe_with_state(st, silent, AsrParams(stream_mode=True))
    assert st.result_tokens()[0] == ref["tokens"]
    ost.close(); st.close()

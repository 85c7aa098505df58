"""Whole-path parity through the reference-shaped API (WhisperAsr.transcribe_with_state ==
/root/reference/src/asr/whisper.rs:45-129) against the CPU oracle + a Python restatement of the
Rust post-processing."""
import numpy as np
import pytest

from tests.rust_post import post_process

pytestmark = pytest.mark.gpu


def _oracle_result(oracle_mod, path, pcm, **kw):
    om = oracle_mod.OracleModel(path)
    ost = om.new_state()
    r = ost.full(pcm, keep_logits=True, **kw)
    logits = ost.kept_logits()
    ost.close()
    om.close()
    return r, logits


@pytest.mark.parametrize("fixture,lang", [("tiny_en_peaked", None), ("micro_v3_peaked", "zh")])
@pytest.mark.parametrize("stream_mode", [False, True])
def test_transcribe_matches_oracle(request, oracle_mod, audio30, fixture, lang, stream_mode):
    from speaksense_b200 import AsrParams, WhisperAsr
    path = request.getfixturevalue(fixture)
    ref, ref_logits = _oracle_result(oracle_mod, path, audio30, language=lang, stream_mode=stream_mode)
    eng = WhisperAsr(path)
    st = eng.create_state()
    params = AsrParams(language=lang, stream_mode=stream_mode, debug_keep_logits=True)
    res = eng.transcribe_with_state(st, audio30, params)
    toks, plogs = st.result_tokens()
    assert toks == ref["tokens"]                                   # token-for-token
    raw = st.raw_segments()
    assert [(s["t0"], s["t1"], s["text"]) for s in raw] == [(s["t0"], s["t1"], s["text"]) for s in ref["segments"]]
    exp = post_process(ref["segments"], stream_mode)
    assert [(s.text, s.speaker_id, s.start, s.end) for s in res.segments] == exp["segments"]
    assert res.full_text == exp["full_text"]
    lg = st.debug_logits()
    assert lg.shape == ref_logits.shape
    assert np.abs(lg - ref_logits).max() < 1e-2                    # north_star tolerance
    stats = st.stats()
    assert stats["n_fallbacks"] == ref["n_fallbacks"] == 0
    assert stats["n_launches"] > 0
    st.close()
    eng.close()


def test_state_reuse_and_short_audio(oracle_mod, tiny_en_peaked, audio30):
    from speaksense_b200 import AsrParams, WhisperAsr
    eng = WhisperAsr(tiny_en_peaked)
    st = eng.create_state()
    p = AsrParams(stream_mode=True)
    a = eng.transcribe_with_state(st, audio30, p)
    b = eng.transcribe_with_state(st, audio30, p)            # same state, second chunk of a stream
    assert a == b
    short = eng.transcribe_with_state(st, audio30[:8000], p)  # < 1 s: whisper returns no segments
    assert short.segments == [] and short.full_text == ""
    c = eng.transcribe(audio30, p)                            # default trait method: fresh state
    assert c == a
    eng.close()


def test_unknown_language_is_an_error(micro_v3_peaked, audio30):
    from speaksense_b200 import AsrParams, NativeError, WhisperAsr
    eng = WhisperAsr(micro_v3_peaked)
    with pytest.raises(NativeError):
        eng.transcribe(audio30, AsrParams(language="xx"))
    eng.close()

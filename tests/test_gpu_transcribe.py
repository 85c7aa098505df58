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


def test_two_windows_with_context_match_oracle(oracle_mod, tiny_en_peaked):
    """45 s clip, stream_mode off: the second window is prompted with [prev] + the first window's tokens
    (no_context=false, whisper.rs:155) - exercises long prompts through the decode kernel."""
    from speaksense_b200 import AsrParams, WhisperAsr, synth
    pcm = synth.synth_audio(45 * 16000, seed=11)
    ref, _ = _oracle_result(oracle_mod, tiny_en_peaked, pcm, stream_mode=False)
    eng = WhisperAsr(tiny_en_peaked)
    st = eng.create_state()
    res = eng.transcribe_with_state(st, pcm, AsrParams(stream_mode=False))
    toks, _ = st.result_tokens()
    assert st.stats()["n_windows"] == ref["n_windows"] == 2
    assert toks == ref["tokens"]
    raw = st.raw_segments()
    assert [(s["t0"], s["t1"], s["text"]) for s in raw] == [(s["t0"], s["t1"], s["text"]) for s in ref["segments"]]
    exp = post_process(ref["segments"], False)
    assert res.full_text == exp["full_text"] and len(res.segments) == len(exp["segments"])
    eng.close()


def test_fallback_ladder_runs_on_flat_logits(oracle_mod, micro_v3_random):
    """Random weights fail the t=0 gates -> best_of=5 sampled decoders at t>0 (host-sampled path).
    Sampling is not bit-reproducible across implementations, so only the control flow is pinned."""
    from speaksense_b200 import AsrParams, WhisperAsr, synth
    pcm = synth.synth_audio(3 * 16000, seed=5)
    om = oracle_mod.OracleModel(micro_v3_random)
    ost = om.new_state()
    ref = ost.full(pcm, language="zh")
    eng = WhisperAsr(micro_v3_random)
    st = eng.create_state()
    eng.transcribe_with_state(st, pcm, AsrParams(language="zh"))
    stats = st.stats()
    assert ref["n_fallbacks"] >= 1 and stats["n_fallbacks"] >= 1
    assert stats["n_decoded"] > len(st.result_tokens()[0])
    ost.close(); om.close(); eng.close()


def test_transcribe_batch_equals_single(tiny_en_peaked):
    from speaksense_b200 import AsrParams, WhisperAsr, synth
    eng = WhisperAsr(tiny_en_peaked)
    clips = [synth.synth_audio(seed=1234 + i) for i in range(3)]
    p = AsrParams(stream_mode=True)
    singles = [eng.transcribe(c, p) for c in clips]
    states = [eng.create_state() for _ in clips]
    batch = eng.transcribe_batch(states, clips, p)
    assert batch == singles
    eng.close()


def test_large_v3_shapes_two_layers(oracle_mod, audio30):
    """Every large-v3 kernel shape (d=1280, 20 heads, 128 mels, 51866 vocab) with 2+2 layers."""
    from tests.conftest import model_path
    from speaksense_b200 import AsrParams, WhisperAsr
    path = model_path("large-v3-l2", "peaked", 0)
    ref, ref_logits = _oracle_result(oracle_mod, path, audio30, language="en", stream_mode=True)
    eng = WhisperAsr(path)
    st = eng.create_state()
    eng.transcribe_with_state(st, audio30, AsrParams(language="en", stream_mode=True, debug_keep_logits=True))
    toks, _ = st.result_tokens()
    assert toks == ref["tokens"]
    assert np.abs(st.debug_logits() - ref_logits).max() < 1e-2
    eng.close()


@pytest.mark.parametrize("fixture,lang,beam", [("tiny_en_peaked", None, 5), ("micro_v3_peaked", "zh", 5), ("tiny_en_peaked", None, 2)])
def test_beam_search_matches_oracle(request, oracle_mod, audio30, fixture, lang, beam):
    """beam_size>1 (BASELINE config 5; an extension - the reference always asks for Greedy{best_of:5}, whisper.rs:132):
    the host-driven beam loop with the device KV shuffle selects the same sequence as the oracle's whisper_full."""
    from speaksense_b200 import AsrParams, WhisperAsr
    path = request.getfixturevalue(fixture)
    om = oracle_mod.OracleModel(path)
    ost = om.new_state()
    ref = ost.full(audio30, language=lang, stream_mode=False, beam_size=beam)
    greedy = ost.full(audio30, language=lang, stream_mode=True)
    ost.close(); om.close()
    eng = WhisperAsr(path)
    st = eng.create_state()
    res = eng.transcribe_with_state(st, audio30, AsrParams(language=lang, stream_mode=False, beam_size=beam))
    toks, plogs = st.result_tokens()
    assert toks == ref["tokens"]
    assert np.allclose(plogs, ref["plogs"], atol=1e-2)
    exp = post_process(ref["segments"], False)
    assert [(s.text, s.speaker_id, s.start, s.end) for s in res.segments] == exp["segments"]
    assert st.stats()["n_fallbacks"] == ref["n_fallbacks"]
    assert toks == greedy["tokens"]          # on a peaked model the best beam is the greedy path
    # the greedy device path still works on the same state afterwards (KV buffers were swapped around by the beam shuffle)
    again = eng.transcribe_with_state(st, audio30, AsrParams(language=lang, stream_mode=True))
    assert st.result_tokens()[0] == greedy["tokens"] and again.full_text != ""
    st.close(); eng.close()


def test_beam_size_above_limit_is_rejected(tiny_en_peaked, audio30):
    from speaksense_b200 import AsrParams, WhisperAsr
    from speaksense_b200._native import NativeError
    eng = WhisperAsr(tiny_en_peaked)
    with pytest.raises(NativeError):
        eng.transcribe(audio30, AsrParams(beam_size=9))
    eng.close()


@pytest.mark.gpu
def test_turbo_like_asymmetric_layers(oracle_mod, audio30):
    """large-v3-turbo layout (script/download-ggml-model.sh:49): fewer decoder than encoder layers"""
    from speaksense_b200 import AsrParams, WhisperAsr
    from tests.conftest import model_path
    path = model_path("micro-turbo", "peaked", 2)
    eng = WhisperAsr(path)
    st = eng.create_state()
    eng.transcribe_with_state(st, audio30, AsrParams(language="en", stream_mode=True, debug_keep_logits=True))
    toks, _ = st.result_tokens()
    om = oracle_mod.OracleModel(path)
    ost = om.new_state()
    ref = ost.full(audio30, language="en", stream_mode=True, keep_logits=True)
    assert toks == ref["tokens"] and len(toks) > 10
    assert float(np.abs(st.debug_logits() - ost.kept_logits()).max()) < 1e-2
    ost.close(); om.close(); st.close(); eng.close()


def test_token_cap_then_context_window(oracle_mod, tmp_path):
    """a window that ends on the 220-token cap (not on EOT / the completion rule) followed by a window prompted with [prev] + 209
    context tokens: the device-side cap rule of decode_mega_kernel and the long prompt against the oracle, token for token (the
    oracle itself is held to an HF-built golden on this case: tests/test_oracle_golden.py)"""
    from speaksense_b200 import AsrParams, WhisperAsr, synth
    path = str(tmp_path / "ggml-tiny.en-peaked-seg175.bin")
    synth.write_model(path, "tiny.en", "peaked", 0, seg_ticks=175)
    pcm = synth.synth_audio(45 * 16000, seed=1234)
    om = oracle_mod.OracleModel(path)
    ost = om.new_state()
    ref = ost.full(pcm, stream_mode=False)
    assert ref["n_windows"] == 2 and len(ref["tokens"]) == 209 + 24
    eng = WhisperAsr(path)
    st = eng.create_state()
    eng.transcribe_with_state(st, pcm, AsrParams(stream_mode=False))
    assert st.result_tokens()[0] == ref["tokens"]
    assert [(s["t0"], s["t1"], s["text"]) for s in st.raw_segments()] == [(s["t0"], s["t1"], s["text"]) for s in ref["segments"]]
    assert st.stats()["n_windows"] == 2 and st.stats()["n_fallbacks"] == 0
    # the same through the batched path (two such clips + two ordinary ones: sequences with different prompt lengths in one round)
    clips = [pcm, synth.synth_audio(seed=7), pcm, synth.synth_audio(20 * 16000, seed=8)]
    sts = [eng.create_state() for _ in clips]
    eng.transcribe_batch(sts, clips, AsrParams(stream_mode=False))
    assert sts[0].result_tokens()[0] == ref["tokens"] and sts[2].result_tokens()[0] == ref["tokens"]
    for s in sts + [st]:
        s.close()
    eng.close(); ost.close(); om.close()

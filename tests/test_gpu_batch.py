"""Batched decoder (csrc/decoder_batch.cu + engine_batch.cc): `ss_transcribe_batch` must give, clip by clip, exactly what
the oracle-checked single-clip path gives (tokens, raw segments, post-processed result) - the two paths share the
arithmetic (f16 operands, f32 accumulate), so greedy tokens have to be identical."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture()
def batch_on():
    """batched decoder on, for every batch of two or more clips (by default batches below 4 decode clip by clip)"""
    old = {k: os.environ.get(k) for k in ("SS_BATCH_DECODE", "SS_BATCH_MIN")}
    os.environ["SS_BATCH_DECODE"] = "1"
    os.environ["SS_BATCH_MIN"] = "2"
    yield
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def _single(eng, clips, params):
    out = []
    for c in clips:
        st = eng.create_state()
        res = eng.transcribe_with_state(st, c, params)
        out.append((res, st.result_tokens()[0], st.raw_segments(), st.stats()["n_fallbacks"]))
        st.close()
    return out


def _batched(eng, clips, params):
    states = [eng.create_state() for _ in clips]
    res = eng.transcribe_batch(states, clips, params)
    out = [(r, st.result_tokens()[0], st.raw_segments(), st.stats()["n_fallbacks"]) for r, st in zip(res, states)]
    launches = [st.stats()["n_launches"] for st in states]
    for st in states:
        st.close()
    return out, launches


@pytest.mark.parametrize("n_clips", [2, 5, 9, 17])       # 1, 1, 2 and 4 n-tiles of 8 sequences
def test_batch_equals_single_tiny(batch_on, tiny_en_peaked, n_clips):
    from speaksense_b200 import AsrParams, WhisperAsr, synth
    eng = WhisperAsr(tiny_en_peaked)
    clips = [synth.synth_audio(seed=1234 + i) for i in range(n_clips)]
    p = AsrParams(stream_mode=True)
    ref = _single(eng, clips, p)
    got, launches = _batched(eng, clips, p)
    assert [g[1] for g in got] == [r[1] for r in ref]          # tokens
    assert [g[2] for g in got] == [r[2] for r in ref]          # raw segments (t0 / t1 / text)
    assert [g[0] for g in got] == [r[0] for r in ref]          # TranscribeResult after the Rust-side rules
    assert all(n > 100 for n in launches)                       # the batched kernels ran (clip by clip: ~40 launches, encoder included)
    eng.close()


def test_batch_ragged_lengths_and_context(batch_on, tiny_en_peaked):
    """Clips of different lengths in one batch: 45 s (two windows, the second prompted with [prev] + context, so the
    sequences of a round have different prompt lengths), 30 s, 12 s and one shorter than 1 s (no segments)."""
    from speaksense_b200 import AsrParams, WhisperAsr, synth
    eng = WhisperAsr(tiny_en_peaked)
    clips = [synth.synth_audio(45 * 16000, seed=11), synth.synth_audio(seed=1234), synth.synth_audio(12 * 16000, seed=5),
             synth.synth_audio(seed=77)[:8000], synth.synth_audio(60 * 16000, seed=3)]
    p = AsrParams(stream_mode=False)
    ref = _single(eng, clips, p)
    got, _ = _batched(eng, clips, p)
    for g, r in zip(got, ref):
        assert g[1] == r[1] and g[2] == r[2] and g[0] == r[0]
    assert got[3][0].segments == [] and got[3][0].full_text == ""
    eng.close()


def test_batch_multilingual_large_shapes(batch_on, audio30):
    """large-v3 kernel shapes (d = 1280, 20 heads, vocab 51866; 2 + 2 layers), multilingual prompt, 6 clips."""
    from tests.conftest import model_path
    from speaksense_b200 import AsrParams, WhisperAsr, synth
    eng = WhisperAsr(model_path("large-v3-l2", "peaked", 0))
    clips = [audio30] + [synth.synth_audio(seed=40 + i) for i in range(5)]
    p = AsrParams(language="en", stream_mode=True)
    ref = _single(eng, clips, p)
    got, _ = _batched(eng, clips, p)
    for g, r in zip(got, ref):
        assert g[1] == r[1] and g[2] == r[2] and g[0] == r[0]
    eng.close()


def test_batch_fallback_ladder_per_clip(batch_on, micro_v3_random, audio30):
    """Flat logits (random weights): the temperature-0 pass fails its thresholds, every clip walks the ladder on its
    own after the batched pass - same tokens, same fallback count as clip by clip."""
    from speaksense_b200 import AsrParams, WhisperAsr, synth
    eng = WhisperAsr(micro_v3_random)
    clips = [audio30, synth.synth_audio(seed=9), synth.synth_audio(seed=10)]
    p = AsrParams(language="zh")
    ref = _single(eng, clips, p)
    got, _ = _batched(eng, clips, p)
    for g, r in zip(got, ref):
        assert g[1] == r[1] and g[3] == r[3] and g[0] == r[0]
    assert any(r[3] >= 1 for r in ref)
    eng.close()


def test_batch_teacher_forced_logits_match_batch1(batch_on, tiny_en_peaked, audio30):
    """Sanity of the default path next to the batched one: the same state can serve both."""
    from speaksense_b200 import AsrParams, WhisperAsr
    eng = WhisperAsr(tiny_en_peaked)
    sts = [eng.create_state() for _ in range(2)]
    p = AsrParams(stream_mode=True)
    a = eng.transcribe_batch(sts, [audio30, audio30], p)
    b = eng.transcribe_with_state(sts[0], audio30, p)          # batch-1 kernel on a state the batch has used
    c = eng.transcribe_batch(sts, [audio30, audio30], p)       # and back
    assert a[0] == a[1] == b == c[0] == c[1]
    lg = eng.decode(sts[0], np.array([50257], np.int32), 0)
    assert np.isfinite(lg).all()
    eng.close()


def test_batch_resident_clips_and_front_end(batch_on, tiny_en_peaked):
    """pcm[i] == NULL takes the state's resident PCM (the stream path: denoise leaves the chunk resident); the micro-batching
    front end (speaksense_b200/batching.py) merges concurrent calls and hands every caller its own result."""
    import threading
    from speaksense_b200 import AsrParams, WhisperAsr, synth
    from speaksense_b200.batching import BatchingEngine
    eng = WhisperAsr(tiny_en_peaked)
    clips = [synth.synth_audio(seed=300 + i) for i in range(3)]
    p = AsrParams(stream_mode=True)
    ref = [eng.transcribe(c, p) for c in clips]
    sts = [eng.create_state() for _ in clips]
    for s, c in zip(sts, clips):
        eng.upload_pcm(s, c)
    assert eng.transcribe_batch(sts, [None, None, None], p) == ref
    assert eng.transcribe_batch(sts, [clips[0], None, clips[2]], p) == ref
    be = BatchingEngine(eng, linger_s=0.5)
    out = [None] * 3

    def call(i):
        out[i] = be.transcribe_resident(sts[i], p) if i == 1 else be.transcribe_with_state(sts[i], clips[i], p)
    th = [threading.Thread(target=call, args=(i,)) for i in range(3)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert out == ref and be.max_seen >= 2
    be.close()
    eng.close()


def test_batch_matches_oracle_directly(batch_on, oracle_mod, tiny_en_peaked, audio30):
    """the batched path against the CPU oracle itself (not only through the single-clip path): tokens and raw segments of every
    sequence of a batch, token for token"""
    from speaksense_b200 import AsrParams, WhisperAsr
    om = oracle_mod.OracleModel(tiny_en_peaked)
    ost = om.new_state()
    ref = ost.full(audio30, stream_mode=True)
    ost.close(); om.close()
    eng = WhisperAsr(tiny_en_peaked)
    sts = [eng.create_state() for _ in range(3)]
    eng.transcribe_batch(sts, [audio30] * 3, AsrParams(stream_mode=True))
    for st in sts:
        assert st.result_tokens()[0] == ref["tokens"]
        assert [(s["t0"], s["t1"], s["text"]) for s in st.raw_segments()] == [(s["t0"], s["t1"], s["text"]) for s in ref["segments"]]
        assert st.stats()["n_fallbacks"] == ref["n_fallbacks"] == 0
        st.close()
    eng.close()


@pytest.mark.parametrize("shape,lang,n_clips", [("tiny.en", None, 5), ("large-v3-l2", "en", 3)])
def test_batched_encoder_pass_equals_per_clip_encoders(batch_on, shape, lang, n_clips):
    """the windows of all clips as one encoder pass over [clips * 1500] rows (default) - same tokens and segments as the clip by
    clip reference AND as the per-clip encoders on their own streams (SS_BATCH_ENCODER=0)"""
    from tests.conftest import model_path
    from speaksense_b200 import AsrParams, WhisperAsr, synth
    eng = WhisperAsr(model_path(shape, "peaked", 0))
    clips = [synth.synth_audio(seed=1234 + i) for i in range(n_clips)] + [synth.synth_audio(45 * 16000, seed=11)]
    p = AsrParams(language=lang, stream_mode=True)
    ref = _single(eng, clips, p)
    got, _ = _batched(eng, clips, p)
    os.environ["SS_BATCH_ENCODER"] = "0"
    try:
        got_off, _ = _batched(eng, clips, p)
    finally:
        os.environ.pop("SS_BATCH_ENCODER", None)
    for g, o, r in zip(got, got_off, ref):
        assert g[1] == r[1] and g[2] == r[2] and g[0] == r[0]
        assert o[1] == r[1] and o[2] == r[2] and o[0] == r[0]
    eng.close()


@pytest.mark.parametrize("fixture,lang,beam", [("tiny_en_peaked", None, 5), ("micro_v3_peaked", "zh", 5), ("tiny_en_peaked", None, 2),
                                                ("micro_v3_random", "zh", 3)])
def test_beam_on_batched_step_equals_default_beam(request, audio30, fixture, lang, beam):
    """the live beams as sequences of one batched step, k candidates per beam selected on the device (default) - same tokens,
    segments and fallback count as the host-stepped beam path (SS_BATCH_BEAM=0: one batch-1 launch per beam and token, host
    top-k); tests/test_gpu_transcribe.py holds both to the oracle"""
    from speaksense_b200 import AsrParams, WhisperAsr
    eng = WhisperAsr(request.getfixturevalue(fixture))
    p = AsrParams(language=lang, stream_mode=True, beam_size=beam)
    got = _single(eng, [audio30], p)[0]
    os.environ["SS_BATCH_BEAM"] = "0"
    try:
        ref = _single(eng, [audio30], p)[0]
    finally:
        os.environ.pop("SS_BATCH_BEAM", None)
    assert got[1] == ref[1] and got[2] == ref[2] and got[3] == ref[3] and got[0] == ref[0]
    eng.close()


def _model_with_invalid_utf8_token(src: str, dst: str, token_id: int):
    """copy of the model file `src` whose vocabulary string of `token_id` is replaced by invalid UTF-8 of the same length"""
    import struct
    buf = bytearray(open(src, "rb").read())
    o = 4 + 44
    n_mel, n_fft = struct.unpack_from("<2i", buf, o); o += 8 + 4 * n_mel * n_fft
    n_tok, = struct.unpack_from("<i", buf, o); o += 4
    assert token_id < n_tok
    for i in range(n_tok):
        ln, = struct.unpack_from("<I", buf, o); o += 4
        if i == token_id:
            assert ln > 0
            buf[o:o + ln] = b"\xff" * ln
            break
        o += ln
    open(dst, "wb").write(bytes(buf))


def test_failing_clip_fails_alone_in_a_batch(batch_on, tiny_en_peaked, tmp_path):
    """whisper.rs:85: a segment that is not valid UTF-8 fails that call.  In a batch the failing clips fail alone - the others keep
    the results their own call would have produced, nothing is decoded twice.  The model's scripted transcript reaches the broken
    vocabulary entry only in its 3rd segment (after 11.8 s), so 30 s clips fail and 5 s clips do not."""
    from speaksense_b200 import AsrParams, WhisperAsr, synth
    from speaksense_b200._native import NativeError
    hp = synth.SHAPES["tiny.en"]
    tgt = synth.scripted_targets(hp, 0)
    bad_tok = int(tgt[1 + 24 + 2 + 24 + 2 + 5])          # a text token of the third segment
    assert 256 <= bad_tok < synth.special_tokens(hp.n_vocab)["eot"]
    path = str(tmp_path / "ggml-tiny.en-badutf8.bin")
    _model_with_invalid_utf8_token(tiny_en_peaked, path, bad_tok)
    eng = WhisperAsr(path)
    p = AsrParams(stream_mode=True)
    long_a, long_b = synth.synth_audio(seed=1234), synth.synth_audio(seed=99)
    short = [synth.synth_audio(5 * 16000, seed=5 + i) for i in range(4)]
    clips = [long_a, short[0], short[1], long_b, short[2], short[3]]
    ref_short = _single(eng, short, p)
    with pytest.raises(NativeError) as ei:
        eng.transcribe(long_a, p)
    assert ei.value.code == -7
    states = [eng.create_state() for _ in clips]
    got = eng.transcribe_batch(states, clips, p, return_exceptions=True)
    assert isinstance(got[0], NativeError) and got[0].code == -7 and isinstance(got[3], NativeError) and got[3].code == -7
    assert [got[i] for i in (1, 2, 4, 5)] == [r[0] for r in ref_short]
    assert all(r[0].full_text for r in ref_short)
    with pytest.raises(NativeError):                      # without return_exceptions the first error is raised, as before
        eng.transcribe_batch(states, clips, p)
    # the states of the clips that failed are usable afterwards
    ok = eng.transcribe_batch(states, [short[0]] * len(states), p)
    assert all(r == ref_short[0][0] for r in ok)
    for st in states:
        st.close()
    eng.close()


def test_batched_path_against_the_oracle_on_large_v3_shapes(batch_on, oracle_mod):
    """the batched decoder step (and the batched encoder pass) against the ORACLE directly - not through the single-clip path - on every
    large-v3 kernel shape (d = 1280, 20 heads, 128 mels, 51866 vocabulary; 2 + 2 layers): tokens, segments, teacher-forced logits"""
    from tests.conftest import model_path
    from speaksense_b200 import AsrParams, WhisperAsr, synth
    path = model_path("large-v3-l2", "peaked", 0)
    clips = [synth.synth_audio(seed=1234 + i) for i in range(3)] + [synth.synth_audio(12 * 16000, seed=5), synth.synth_audio(45 * 16000, seed=11)]
    om = oracle_mod.OracleModel(path)
    refs = []
    for c in clips:
        ost = om.new_state()
        refs.append(ost.full(c, language="en", stream_mode=False))
        ost.close()
    eng = WhisperAsr(path)
    sts = [eng.create_state() for _ in clips]
    eng.transcribe_batch(sts, clips, AsrParams(language="en", stream_mode=False))
    for st, r in zip(sts, refs):
        assert st.result_tokens()[0] == r["tokens"]
        assert [(s["t0"], s["t1"], s["text"]) for s in st.raw_segments()] == [(s["t0"], s["t1"], s["text"]) for s in r["segments"]]
        assert st.stats()["n_fallbacks"] == r["n_fallbacks"] == 0
        assert st.stats()["n_launches"] > 100      # the batched kernels ran
        st.close()
    eng.close(); om.close()

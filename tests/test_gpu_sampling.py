"""SURVEY.md §8 row a9 at temperature > 0: whisper_full's fallback ladder runs best_of = 5 SAMPLED decoders per rung
(whisper.rs:132 Greedy{best_of: 5}; thresholds :159-162).  The engine runs them as sequences of one batched decoder step and
draws on the device (bd_sample_kernel, sample == 3) from uniforms generated with each decoder's own std::mt19937 - the same
stream std::discrete_distribution consumes in whisper.cpp and in the oracle.  The sampled TOKENS are compared with the oracle:
the "soft" model family keeps the scripted transcript but with a target logit low enough that temperature 0 fails the
log-probability gate, so the ladder runs, while the draws keep a margin (oracle: min_sample_margin, the smallest distance of a
uniform to the boundary of the interval it fell into) far above what the logits tolerance can move."""
import os

import numpy as np
import pytest

from tests.conftest import model_path

pytestmark = pytest.mark.gpu

# cumulative-probability units.  A boundary at cumulative probability c moves by at most ~ delta * min(c, 1 - c) when every
# probability carries a relative error delta (~1e-3 for logits that agree to 1e-3).  On the soft models' t > 0 rungs one token
# holds > 0.999 of the mass, so every boundary sits within 1e-3 of 0 or 1: it can move by ~1e-6.
MARGIN = 1e-5


def _oracle(oracle_mod, path, pcm, **kw):
    om = oracle_mod.OracleModel(path)
    st = om.new_state()
    r = st.full(pcm, **kw)
    st.close(); om.close()
    return r


CASES = [("tiny.en", "soft10", None, 30, True), ("tiny.en", "soft4", None, 30, True), ("micro-v3", "soft10", "zh", 30, True),
         ("tiny.en", "soft10", None, 45, False)]      # 45 s, not stream mode: second window prompted with [prev] + context


@pytest.mark.parametrize("shape,family,lang,seconds,stream", CASES)
def test_sampled_decoders_match_oracle_token_for_token(oracle_mod, shape, family, lang, seconds, stream):
    from speaksense_b200 import AsrParams, WhisperAsr, synth
    path = model_path(shape, family, 0)
    pcm = synth.synth_audio(seconds * 16000, seed=1234)
    ref = _oracle(oracle_mod, path, pcm, language=lang, stream_mode=stream)
    assert ref["n_fallbacks"] >= 1 and ref["n_draws"] >= 5 * 20, ref      # the ladder ran: 5 sampled decoders
    assert ref["min_sample_margin"] > MARGIN, ref["min_sample_margin"]     # precondition of a token-for-token comparison
    eng = WhisperAsr(path)
    for env in (None, "0"):      # device-sampled batched step (default), then the host-sampled path
        if env is not None:
            os.environ["SS_BATCH_SAMPLE"] = env
        try:
            st = eng.create_state()
            eng.transcribe_with_state(st, pcm, AsrParams(language=lang, stream_mode=stream))
            toks = st.result_tokens()[0]
            assert toks == ref["tokens"], (env, [i for i, (a, b) in enumerate(zip(toks, ref["tokens"])) if a != b][:5])
            assert [(s["t0"], s["t1"], s["text"]) for s in st.raw_segments()] == [(s["t0"], s["t1"], s["text"]) for s in ref["segments"]]
            assert st.stats()["n_fallbacks"] == ref["n_fallbacks"]
            # the same state again: the decoders' generators have moved on exactly as the oracle's would on a second call
            st.close()
        finally:
            os.environ.pop("SS_BATCH_SAMPLE", None)
    eng.close()


def test_generators_advance_like_the_oracles_across_calls(oracle_mod):
    """whisper.cpp keeps one std::mt19937 per decoder for the life of the state: a second call on the same state continues the
    stream.  Two calls on one oracle state against two calls on one engine state."""
    from speaksense_b200 import AsrParams, WhisperAsr, synth
    path = model_path("tiny.en", "soft10", 0)
    clips = [synth.synth_audio(seed=1234), synth.synth_audio(20 * 16000, seed=77)]
    om = oracle_mod.OracleModel(path)
    ost = om.new_state()
    refs = [ost.full(c, stream_mode=True) for c in clips]
    assert all(r["n_fallbacks"] >= 1 and r["min_sample_margin"] > MARGIN for r in refs)
    eng = WhisperAsr(path)
    st = eng.create_state()
    for c, r in zip(clips, refs):
        eng.transcribe_with_state(st, c, AsrParams(stream_mode=True))
        assert st.result_tokens()[0] == r["tokens"]
    st.close(); eng.close(); ost.close(); om.close()


def test_ladder_inside_a_batch_equals_single_calls():
    from speaksense_b200 import AsrParams, WhisperAsr, synth
    path = model_path("tiny.en", "soft10", 0)
    eng = WhisperAsr(path)
    clips = [synth.synth_audio(seed=1234 + i) for i in range(5)]
    p = AsrParams(stream_mode=True)
    single = []
    for c in clips:
        st = eng.create_state()
        eng.transcribe_with_state(st, c, p)
        single.append((st.result_tokens()[0], st.stats()["n_fallbacks"]))
        st.close()
    assert all(f >= 1 for _, f in single)
    sts = [eng.create_state() for _ in clips]
    eng.transcribe_batch(sts, clips, p)
    assert [(s.result_tokens()[0], s.stats()["n_fallbacks"]) for s in sts] == single
    for s in sts:
        s.close()
    eng.close()


def test_flat_logits_ladder_control_flow(oracle_mod, micro_v3_random):
    """random weights: every rung fails its gates; draws land in a flat tail where neighbouring tokens are ~1e-5 apart, so only the
    control flow is comparable (the oracle reports the margin that forbids a token comparison)"""
    from speaksense_b200 import AsrParams, WhisperAsr, synth
    pcm = synth.synth_audio(3 * 16000, seed=5)
    ref = _oracle(oracle_mod, micro_v3_random, pcm, language="zh")
    assert ref["n_fallbacks"] >= 1 and ref["min_sample_margin"] < MARGIN
    eng = WhisperAsr(micro_v3_random)
    st = eng.create_state()
    eng.transcribe_with_state(st, pcm, AsrParams(language="zh"))
    assert st.stats()["n_fallbacks"] == ref["n_fallbacks"]
    st.close(); eng.close()

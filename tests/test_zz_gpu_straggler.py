"""GPU: default settings (no environment switches), a batch in which one clip outlives the others: after the first round only
that clip is still in a window, so `transcribe_batch` finishes it with the batch-1 kernel (the lone-straggler branch of
csrc/engine_batch.cc).  Written after round 1's GPU budget was spent - this branch had not been reached by a GPU test before.
(Named to sort late: under `pytest -x` a surprise here must not hide the rest of the suite.)"""
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("stream_mode", [True, False])
def test_one_clip_outlives_the_batch(tiny_en_peaked, stream_mode):
    from speaksense_b200 import AsrParams, WhisperAsr, synth
    eng = WhisperAsr(tiny_en_peaked)
    clips = [synth.synth_audio(seed=1234 + i) for i in range(4)] + [synth.synth_audio(60 * 16000, seed=3)]
    p = AsrParams(stream_mode=stream_mode)
    ref = []
    for c in clips:
        st = eng.create_state()
        ref.append((eng.transcribe_with_state(st, c, p), st.result_tokens()[0], st.raw_segments(), st.stats()["n_windows"]))
        st.close()
    assert ref[4][3] >= 2 and all(r[3] == 1 for r in ref[:4])          # only the long clip has a second window
    states = [eng.create_state() for _ in clips]
    got = eng.transcribe_batch(states, clips, p)
    for g, st, r in zip(got, states, ref):
        assert st.result_tokens()[0] == r[1] and st.raw_segments() == r[2] and g == r[0] and st.stats()["n_windows"] == r[3]
    assert states[0].stats()["n_launches"] > 100                         # the first round went through the batched step (clip by clip: ~40)
    for st in states:
        st.close()
    eng.close()

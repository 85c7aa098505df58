"""GPU: the REST processor end to end (WAV -> StreamAudioProcessor frames on the GPU in batched launches -> 30 s buffers ->
transcribe on one state) against a composition of the CPU oracle's StreamAudioProcessor and the engine per buffer."""
import wave

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_rest_processor_matches_composition(tmp_path, micro_v3_peaked, oracle_mod, audio30):
    from speaksense_b200 import AsrParams, WhisperAsr, rest
    x = np.concatenate([audio30, audio30[:100000]])            # 36.25 s -> two transcribe calls
    p = tmp_path / "clip.wav"
    with wave.open(str(p), "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000)
        w.writeframes((np.clip(x, -1, 1) * 32767).astype("<i2").tobytes())
    eng = WhisperAsr(micro_v3_peaked, device=0)
    got = rest.TranscribeProcessor(eng).process_audio(str(p), language="en")
    # composition: oracle frames -> engine
    samples = np.frombuffer((np.clip(x, -1, 1) * 32767).astype("<i2").tobytes(), "<i2").astype(np.float32) / np.float32(32768.0)
    sp = oracle_mod.StreamAudioProcessor()
    frames = []
    for o in range(0, samples.size, 4096):
        frames += sp.process_chunk(samples[o:o + 4096])
    frames += sp.finish()
    st = eng.create_state()
    params = AsrParams(language="en", stream_mode=True)
    text, segs, buf, n = "", [], [], 0
    for f in frames + [None]:
        if f is not None:
            buf.append(f); n += f.size
        if (f is None and n) or n >= 480000:
            r = eng.transcribe_with_state(st, np.concatenate(buf), params)
            text += r.full_text; segs += r.segments; buf, n = [], 0
    assert got.n_calls == 2 and got.text == text and len(got.text) > 0
    assert [(s.text, s.start_time, s.end_time) for s in got.segments] == [(s.text, s.start, s.end) for s in segs]
    st.close(); eng.close()

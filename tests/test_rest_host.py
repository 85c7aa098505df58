"""CPU: the REST task processor's host logic (speaksense_b200/rest.py == processors/transcribe.rs:62-167 +
audio/mod.rs:157-233) with a stub engine: chunking, framing, 30 s buffering (cut at 235 frames = 481280 samples),
aggregation, parameter validation."""
import wave

import numpy as np
import pytest

from speaksense_b200 import audio, rest
from speaksense_b200.asr import TranscribeResult, TranscribeSegment


class _State:
    _lock = None

    def close(self):
        pass


class _Engine:
    def __init__(self):
        self.calls = []

    def create_state(self):
        return _State()

    def transcribe_with_state(self, state, pcm, params):
        k = len(self.calls)
        self.calls.append((pcm.size, params.language, params.stream_mode, params.speaker_diarization))
        return TranscribeResult(segments=[TranscribeSegment(text="段%d" % k, speaker_id=0, start=10.0 * k, end=10.0 * k + 5)], full_text="段%d" % k)


def write_wav(path, x, channels=1, rate=16000):
    with wave.open(str(path), "wb") as w:
        w.setnchannels(channels); w.setsampwidth(2); w.setframerate(rate)
        w.writeframes((np.clip(x, -1, 1) * 32767).astype("<i2").tobytes())


def test_rest_buffering_and_aggregation(tmp_path, monkeypatch, audio30):
    seen = []
    monkeypatch.setattr(audio, "denoise_frames", lambda eng, st, fr, cfg=None: seen.append(fr.shape) or np.asarray(fr, np.float32).copy())
    x = np.concatenate([audio30, audio30, audio30[:160000]])           # 70 s
    p = tmp_path / "a.wav"
    write_wav(p, x)
    eng = _Engine()
    r = rest.TranscribeProcessor(eng).process_audio(str(p), language="zh", speaker_diarization=True)
    n_frames = (x.size + 2047) // 2048
    assert sum(s[0] for s in seen) == n_frames == 547
    assert [c[0] for c in eng.calls] == [235 * 2048, 235 * 2048, (547 - 470) * 2048]      # >= 480000 is reached after 235 frames
    assert all(c[1:] == ("zh", True, True) for c in eng.calls)
    assert r.text == "段0段1段2" and r.n_calls == 3
    assert [(s.text, s.start_time, s.end_time) for s in r.segments] == [("段0", 0.0, 5.0), ("段1", 10.0, 15.0), ("段2", 20.0, 25.0)]


def test_rest_stereo_and_validation(tmp_path, monkeypatch):
    monkeypatch.setattr(audio, "denoise_frames", lambda eng, st, fr, cfg=None: np.asarray(fr, np.float32).copy())
    t = np.arange(16000 * 2) / 16000.0
    left, right = 0.5 * np.sin(2 * np.pi * 440 * t), 0.25 * np.sin(2 * np.pi * 880 * t)
    inter = np.stack([left, right], axis=1).reshape(-1)
    p = tmp_path / "s.wav"
    write_wav(p, inter, channels=2)
    eng = _Engine()
    r = rest.TranscribeProcessor(eng).process_audio(str(p))
    assert eng.calls[0][0] == 16 * 2048 and r.n_calls == 1        # 32000 mono samples -> 16 frames (last one zero-padded)
    mono = rest.convert_to_mono(np.array([1, 3, 5, 7, 9], np.float32), 2)
    np.testing.assert_array_equal(mono, np.array([2, 6, 4.5], np.float32))      # ragged tail divided by the channel count too (mod.rs:393-396)
    with pytest.raises(ValueError):
        rest.TranscribeProcessor(eng).process_audio(str(p), language="fr")        # transcribe.rs:200-204
    p8 = tmp_path / "r.wav"
    write_wav(p8, left, rate=8000)
    with pytest.raises(ValueError):
        rest.TranscribeProcessor(eng).process_audio(str(p8))


def test_callback_payload_shape():
    """callback/mod.rs:28-33,80-97 + types.rs serde attributes, by hand"""
    import json
    r = rest.RestTranscribeResult(text="你好", segments=[rest.RestSegment("你好", 1, 0.0, 296.0)])
    p = rest.callback_payload("task-1", r)
    assert json.loads(json.dumps(p, ensure_ascii=False)) == {
        "task_id": "task-1", "status": "Completed",
        "data": {"type": "Transcribe", "result": {"text": "你好", "segments": [
            {"text": "你好", "speaker_id": 1, "start_time": 0.0, "end_time": 296.0}]}}}
    assert rest.callback_error_payload("t", "boom") == {"task_id": "t", "status": {"Failed": "boom"}, "data": "boom"}

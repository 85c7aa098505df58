"""CPU: the micro-batching front end (speaksense_b200/batching.py) with a stub engine - which calls are merged, that every
caller gets its own result, that incompatible parameters / a state used twice are never put into one batch, and that one
failing clip fails only its own call."""
import threading
import time

import numpy as np
import pytest

from speaksense_b200.asr import AsrParams, TranscribeResult, TranscribeSegment
from speaksense_b200.batching import BatchingEngine


class _State:
    def __init__(self, name):
        self.name = name
        self.resident = None

    def close(self):
        pass


class _Engine:
    """transcribe == text made of the state's name and the clip's first sample; a clip starting with NaN fails"""

    def __init__(self, delay=0.0):
        self.delay = delay
        self.batches = []          # [(kind, [state names], params key)]
        self.n_states = 0
        self._h = object()
        self.info = {"n_vocab": 1}
        self.lock = threading.Lock()

    def create_state(self):
        self.n_states += 1
        return _State("s%d" % self.n_states)

    def upload_pcm(self, state, audio):
        state.resident = np.asarray(audio, np.float32)

    def _one(self, state, pcm, params):
        x = state.resident if pcm is None else pcm
        if np.isnan(x[0]):
            raise ValueError("segment text is not valid UTF-8")
        text = "%s:%g:%s" % (state.name, float(x[0]), params.language)
        return TranscribeResult(segments=[TranscribeSegment(text, 0, 0.0, 1.0)], full_text=text)

    def transcribe_with_state(self, state, audio, params):
        time.sleep(self.delay)
        with self.lock:
            self.batches.append(("single", [state.name], params.language))
        return self._one(state, audio, params)

    def transcribe_resident(self, state, params):
        time.sleep(self.delay)
        with self.lock:
            self.batches.append(("resident", [state.name], params.language))
        return self._one(state, None, params)

    def transcribe_batch(self, states, audios, params, return_exceptions=False):
        """like WhisperAsr.transcribe_batch: a failing clip fails alone (its slot holds the exception with return_exceptions)"""
        time.sleep(self.delay)
        with self.lock:
            self.batches.append(("batch", [s.name for s in states], params.language))
        assert len({id(s) for s in states}) == len(states)
        out = []
        for s, a in zip(states, audios):
            try:
                out.append(self._one(s, a, params))
            except Exception as e:      # noqa: BLE001
                if not return_exceptions:
                    raise
                out.append(e)
        return out

    def close(self):
        pass


def _run_threads(fns):
    out = [None] * len(fns)

    def call(i):
        try:
            out[i] = fns[i]()
        except Exception as e:      # noqa: BLE001
            out[i] = e
    th = [threading.Thread(target=call, args=(i,)) for i in range(len(fns))]
    for t in th:
        t.start()
    for t in th:
        t.join(10)
    return out


def test_concurrent_calls_are_merged_and_routed():
    eng = _Engine(delay=0.02)
    be = BatchingEngine(eng, max_batch=8, linger_s=0.5)
    states = [be.create_state() for _ in range(6)]
    p = AsrParams(language="zh", stream_mode=True)
    res = _run_threads([lambda i=i: be.transcribe_with_state(states[i], np.full(4, i, np.float32), p) for i in range(6)])
    assert [r.full_text for r in res] == ["s%d:%d:zh" % (i + 1, i) for i in range(6)]
    assert sum(len(b[1]) for b in eng.batches) == 6
    assert any(b[0] == "batch" and len(b[1]) >= 2 for b in eng.batches)       # merged (the 0.5 s linger catches the six threads even on a loaded box)
    assert be.n_requests == 6 and be.max_seen >= 2
    be.close()


def test_lone_call_goes_straight_through():
    eng = _Engine()
    be = BatchingEngine(eng, linger_s=0.0)
    st = be.create_state()
    be.upload_pcm(st, np.full(3, 7, np.float32))
    r = be.transcribe_resident(st, AsrParams(language="en"))
    assert r.full_text == "s1:7:en" and eng.batches == [("resident", ["s1"], "en")]
    r = be.transcribe(np.full(3, 9, np.float32), AsrParams(language="en"))      # AsrEngine::transcribe: fresh state (mod.rs:69-72)
    assert r.full_text == "s2:9:en" and eng.batches[-1] == ("single", ["s2"], "en")
    be.close()


def test_parameters_and_states_are_never_mixed():
    eng = _Engine(delay=0.01)
    be = BatchingEngine(eng, linger_s=0.05)
    a, b, c = (be.create_state() for _ in range(3))
    zh, en = AsrParams(language="zh"), AsrParams(language="en")
    one = np.ones(2, np.float32)
    res = _run_threads([lambda: be.transcribe_with_state(a, one, zh), lambda: be.transcribe_with_state(b, one, en),
                        lambda: be.transcribe_with_state(c, one, zh), lambda: be.transcribe_with_state(a, 2 * one, zh)])
    assert sorted(r.full_text for r in res) == ["s1:1:zh", "s1:2:zh", "s2:1:en", "s3:1:zh"]
    for kind, names, lang in eng.batches:
        assert len(set(names)) == len(names)                      # a state at most once per batch
        assert all((n == "s2") == (lang == "en") for n in names)   # languages never share a batch
    be.close()


def test_one_failing_clip_fails_only_its_call():
    eng = _Engine(delay=0.01)
    be = BatchingEngine(eng, linger_s=0.05)
    sts = [be.create_state() for _ in range(3)]
    p = AsrParams(language="zh")
    clips = [np.ones(2, np.float32), np.full(2, np.nan, np.float32), 3 * np.ones(2, np.float32)]
    res = _run_threads([lambda i=i: be.transcribe_with_state(sts[i], clips[i], p) for i in range(3)])
    assert res[0].full_text == "s1:1:zh" and res[2].full_text == "s3:3:zh"
    assert isinstance(res[1], ValueError)
    be.close()
    with pytest.raises(RuntimeError):
        be.transcribe_with_state(sts[0], clips[0], p)


def test_stream_sessions_through_the_front_end():
    """AsrStreamSession only needs the engine interface: two sessions fed from two threads share the front end"""
    from speaksense_b200 import stream

    class _E(_Engine):
        def transcribe_resident(self, state, params):
            r = super().transcribe_resident(state, params)
            return r

    eng = _E()
    be = BatchingEngine(eng, linger_s=0.01)
    import speaksense_b200.stream as st_mod
    orig = st_mod.denoise_audio
    st_mod.denoise_audio = lambda engine, state, x, cfg, fetch=False: engine.upload_pcm(state, x)      # no GPU here
    try:
        pcm = np.full(16000 * 6, 0.25, np.float32)
        msgs = stream.encode_messages(pcm)

        def run():
            ses = stream.AsrStreamSession(be)
            n = 0
            for m, e in msgs:
                n += len(ses.feed(m, e, "dev"))
            ses.close()
            return n
        counts = _run_threads([run, run])
        assert counts[0] == counts[1] and counts[0] >= 1
        assert sum(1 for b in eng.batches if b[0] in ("resident", "batch")) >= 1
    finally:
        st_mod.denoise_audio = orig
        be.close()


def test_rest_tasks_through_the_front_end(tmp_path, monkeypatch):
    """two REST tasks (processors/transcribe.rs:62-167 mirror) running concurrently against one front end: their 30 s buffers
    may share batches, their aggregated texts must not mix"""
    import wave
    from speaksense_b200 import audio, rest
    monkeypatch.setattr(audio, "denoise_frames", lambda eng, st, fr, cfg=None: np.asarray(fr, np.float32).copy())
    paths = []
    for k, level in enumerate((0.25, 0.5)):
        x = np.full(16000 * 65, level, np.float32)                     # 65 s -> three buffers per task
        p = tmp_path / ("t%d.wav" % k)
        with wave.open(str(p), "wb") as w:
            w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000)
            w.writeframes((x * 32767).astype("<i2").tobytes())
        paths.append(str(p))
    eng = _Engine(delay=0.01)
    be = BatchingEngine(eng, linger_s=0.05)
    res = _run_threads([lambda p=p: rest.TranscribeProcessor(be).process_audio(p, language="zh") for p in paths])
    assert all(not isinstance(r, Exception) for r in res), res
    assert res[0].n_calls == res[1].n_calls == 3
    for r in res:                                                     # every piece of a task's text comes from ITS state
        names = {piece.split(":")[0] for piece in r.text.replace("zh", "zh|").split("|") if piece}
        assert len(names) == 1
    assert {s.text.split(":")[0] for s in res[0].segments}.isdisjoint({s.text.split(":")[0] for s in res[1].segments})
    assert sum(len(b[1]) for b in eng.batches) == 6
    be.close()


def test_failing_clip_is_not_decoded_twice_and_worker_survives():
    """a clip that fails inside the batch call fails alone: nothing is re-run clip by clip (no state is advanced twice)"""
    eng = _Engine(delay=0.02)
    be = BatchingEngine(eng, max_batch=8, linger_s=0.5)
    states = [be.create_state() for _ in range(4)]
    p = AsrParams(language="en", stream_mode=True)
    clips = [np.full(4, i, np.float32) for i in range(4)]
    clips[2][0] = np.nan
    out = [None] * 4

    def call(i):
        try:
            out[i] = be.transcribe_with_state(states[i], clips[i], p)
        except Exception as e:      # noqa: BLE001
            out[i] = e

    th = [threading.Thread(target=call, args=(i,)) for i in range(4)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert isinstance(out[2], ValueError)
    assert [o.full_text for i, o in enumerate(out) if i != 2] == ["s1:0:en", "s2:1:en", "s4:3:en"]
    merged = [b for b in eng.batches if b[0] == "batch" and len(b[1]) >= 2]
    if merged:      # the merged clips were not run again one by one
        names = [n for b in merged for n in b[1]]
        assert not any(b[0] == "single" and b[1][0] in names for b in eng.batches)
    assert be.transcribe_with_state(states[0], clips[0], p).full_text == "s1:0:en"      # the worker is still alive
    be.close()


def test_worker_death_fails_callers_instead_of_hanging():
    class _Boom(_Engine):
        def transcribe_with_state(self, state, audio, params):
            raise KeyboardInterrupt()      # not an Exception: escapes every handler of the worker loop

    be = BatchingEngine(_Boom(), max_batch=4, linger_s=0.0)
    st = be.create_state()
    with pytest.raises(BaseException):
        be.transcribe_with_state(st, np.zeros(4, np.float32), AsrParams())
    with pytest.raises(RuntimeError):
        be.transcribe_with_state(st, np.zeros(4, np.float32), AsrParams())
    be.close()


def test_close_races_with_submit():
    eng = _Engine(delay=0.001)
    be = BatchingEngine(eng, max_batch=4, linger_s=0.0)
    states = [be.create_state() for _ in range(8)]
    done = []

    def call(i):
        try:
            be.transcribe_with_state(states[i], np.full(4, i, np.float32), AsrParams())
            done.append("ok")
        except RuntimeError:
            done.append("closed")

    th = [threading.Thread(target=call, args=(i,)) for i in range(8)]
    for t in th[:4]:
        t.start()
    be.close()
    for t in th[4:]:
        t.start()
    for t in th:
        t.join(timeout=5)
    assert not any(t.is_alive() for t in th)      # nobody blocked for ever behind the sentinel
    assert len(done) == 8

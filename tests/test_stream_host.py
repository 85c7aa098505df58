"""CPU: the gRPC handler's host logic (speaksense_b200/stream.py == src/grpc/handlers/asr.rs) against values
worked out by hand from the Rust source, and its chunk accounting with a stub engine."""
import base64

import numpy as np

from speaksense_b200 import stream
from speaksense_b200.asr import TranscribeResult, TranscribeSegment


def seg(text, a=0.0, b=0.0):
    return TranscribeSegment(text=text, speaker_id=0, start=a, end=b)


def test_process_text_rules():
    pt = stream.process_text
    assert pt("你好", "", [seg("你好")]) == "你好"                                  # asr.rs:70-72
    assert pt("今天天气", "你好", [seg("今天天气")]) == "今天天气"                     # last segment is new content (:75-80)
    assert pt("你好世界", "你好世界啊", [seg("你好世界")]) == "你好世界"               # contained in the previous text -> falls to the sentence rule (:122-133)
    # from here on the last segment is contained in last_text, so the later rules decide
    assert pt("abc def", "abc", []) == "def"                                       # prefix growth, trimmed (:83-88)
    assert pt("x" * 10, "yyy", []) == "x" * 10                                     # length ratio > 2 (:91-93)
    assert pt("yyy", "x" * 10, []) == "yyy"
    assert pt("一。二。三。", "一。二。四", []) == "三。"                              # more sentences: the new ones + final mark (:107-121)
    assert pt("一。二", "一。三", []) == "二"                                        # same count, last sentence differs (:122-133)
    assert pt("一。二？", "一。三？", []) == "二？"
    assert pt("same", "same", []) is None


def test_stream_context_times():
    c = stream.StreamContext()
    assert c.calculate_segment_time(0.0, 120.0) == (0, 120000)                     # ticks treated as seconds (Appendix B.2)
    c.next_block()
    assert c.calculate_segment_time(10.0, 30.0) == (120000, 140000)                # clamped to the previous end, end shifted by the same diff
    assert c.last_end_time == 140.0
    c.next_block(); c.next_block()
    c.last_end_time = 0.0
    assert c.calculate_segment_time(1.5, 2.25) == (16500, 17250)                   # block 3 * 5.0 s base


def test_pcm16_scaling_and_tail_byte():
    raw = np.array([0, 32767, -32767, -32768, 1], "<i2").tobytes()
    x = stream.pcm16_to_f32(raw, exact=True)
    assert x.dtype == np.float32 and x[1] == 1.0 and x[2] == -1.0 and x[3] < -1.0     # 1/32767, not 1/32768 (asr.rs:192)
    y = stream.pcm16_to_f32(raw + b"\x07", exact=False)
    assert y.size == 6 and y[5] == 0.0                                              # odd trailing byte -> one 0.0 sample (asr.rs:238-243)
    assert stream.pcm16_to_f32(raw + b"\x07", exact=True).size == 5


class _StubState:
    def close(self):
        pass


class _StubEngine:
    def __init__(self):
        self.calls = []

    def create_state(self):
        return _StubState()

    def transcribe_resident(self, state, params):
        k = sum(1 for c in self.calls if c[0] == "chunk")
        self.calls.append(("chunk", params.language, params.stream_mode, params.min_segment_length))
        return TranscribeResult(segments=[seg("块%d。" % k, 0.0, 300.0)], full_text="块%d。" % k)

    def transcribe(self, audio, params):
        self.calls.append(("tail", len(audio)))
        return TranscribeResult(segments=[seg("尾。", 0.0, 50.0)], full_text="尾。")


def test_chunk_accounting_with_stub_engine(monkeypatch):
    seen = []
    monkeypatch.setattr(stream, "denoise_audio", lambda eng, st, x, cfg, fetch=True: seen.append(x.copy()) or (None, "Stationary", 0.0))
    eng = _StubEngine()
    ses = stream.AsrStreamSession(eng)
    pcm = (np.sin(np.arange(16000 * 12) * 0.01) * 0.5).astype(np.float32)          # 12 s = 384000 bytes
    msgs = stream.encode_messages(pcm)
    assert len(msgs) == 12 and msgs[-1][1] == 1 and all(e == 0 for _, e in msgs[:-1])
    assert len(base64.b64decode(msgs[0][0])) == 32768
    out = []
    for m, e in msgs:
        out += ses.feed(m, e, "dev-1")
    # 5 s chunks (160000 BYTES), 4.5 s advance: chunks start at bytes 0 and 144000; the third would need 448000 bytes
    assert [c[0] for c in eng.calls] == ["chunk", "chunk", "tail"]
    assert eng.calls[0][1:] == ("zh", True, 5)
    assert all(x.size == 80000 for x in seen)
    raw = (np.clip(pcm, -1, 1) * 32767.0).astype("<i2")
    np.testing.assert_array_equal(seen[1], raw[72000:152000].astype(np.float32) / np.float32(32767.0))
    assert eng.calls[2][1] == (384000 - 2 * 144000) // 2                            # the tail is what is left in the buffer
    assert [r.end for r in out] == [0, 0, 1] and all(r.device_id == "dev-1" for r in out)
    assert out[0].text.decode() == "块0。" and out[1].text.decode() == "块1。" and out[2].text.decode() == "尾。"
    # block 0: 0..300 "s" -> 0..300000 ms; block 1 base 5 s: starts before the previous end -> shifted
    assert (out[0].segments[0].start, out[0].segments[0].end) == (0, 300000)
    assert (out[1].segments[0].start, out[1].segments[0].end) == (300000, 600000)
    assert (out[2].segments[0].start, out[2].segments[0].end) == (600000, 650000)
    # a message that is not base64 is skipped (asr.rs:176-183)
    assert ses.feed(b"!!!not base64!!!", 0) == []


def test_grpc_transport_speaks_asr_proto(monkeypatch):
    """service asr.Asr / Transcribe over a real localhost gRPC channel, stub engine behind AsrStreamSession"""
    from speaksense_b200 import grpc_server
    monkeypatch.setattr(stream, "denoise_audio", lambda eng, st, x, cfg, fetch=True: (None, "Stationary", 0.0))
    # wire format: field numbers / types of proto/asr.proto
    req = grpc_server.TranscribeRequest(type=7, end=1, audio=b"QUJD", device_id="d")
    assert req.SerializeToString() == b"\x08\x07\x10\x01\x1a\x04QUJD\x22\x01d"
    seg_pb = grpc_server.Segment(start=3, end=300000, text="块".encode())
    assert grpc_server.Segment.FromString(seg_pb.SerializeToString()).end == 300000
    eng = _StubEngine()
    server = grpc_server.serve(eng, "127.0.0.1:0")
    try:
        pcm = (np.sin(np.arange(16000 * 12) * 0.01) * 0.5).astype(np.float32)
        out = list(grpc_server.transcribe_stream("127.0.0.1:%d" % server.bound_port, stream.encode_messages(pcm), "dev-9"))
    finally:
        server.stop(0)
    assert [c[0] for c in eng.calls] == ["chunk", "chunk", "tail"]
    assert [r.end for r in out] == [0, 0, 1] and all(r.device_id == "dev-9" for r in out)
    assert out[0].text.decode() == "块0。" and out[2].text.decode() == "尾。"
    assert (out[1].segments[0].start, out[1].segments[0].end) == (300000, 600000)

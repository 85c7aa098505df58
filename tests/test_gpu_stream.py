"""GPU: AsrStreamSession (gRPC handler semantics, src/grpc/handlers/asr.rs:146-281) end to end: every 5 s chunk is
denoised and transcribed on the stream's state; the responses equal an independent recomputation chunk by chunk
(oracle denoise -> engine on a separate state), with the handler's chunk / overlap / tail rules."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_stream_session_matches_chunkwise_recomputation(micro_v3_peaked, oracle_mod, audio30):
    from speaksense_b200 import AsrParams, WhisperAsr, stream
    eng = WhisperAsr(micro_v3_peaked, device=0)
    pcm = audio30[:16000 * 12]
    msgs = stream.encode_messages(pcm)
    ses = stream.AsrStreamSession(eng)
    got = []
    for m, e in msgs:
        got += ses.feed(m, e, "dev")
    assert ses.n_chunks == 2
    # independent recomputation
    raw = (np.clip(pcm, -1, 1) * 32767.0).astype("<i2")
    params = AsrParams(language="zh", stream_mode=True, min_segment_length=5)
    ctx = stream.StreamContext()
    want = []
    st = eng.create_state()
    for k in range(2):
        x = raw[k * 72000:k * 72000 + 80000].astype(np.float32) / np.float32(32767.0)
        den, _ = oracle_mod.denoise_audio(x)
        r = eng.transcribe_with_state(st, den, params)
        for s in r.segments:
            t = stream.process_text(s.text, ctx.last_text, [s])
            if t is not None:
                ctx.last_text = s.text
                want.append((0, t, ctx.calculate_segment_time(s.start, s.end)))
        ctx.next_block()
    tail = raw[2 * 72000:].astype(np.float32) / np.float32(32767.0)
    r = eng.transcribe(tail, params)
    t = stream.process_text(r.full_text, ctx.last_text, r.segments)
    if t is not None:
        want.append((1, t, [ctx.calculate_segment_time(s.start, s.end) for s in r.segments][-1] if r.segments else None))
    assert len(got) == len(want) and len(got) >= 1
    for g, w in zip(got, want):
        assert g.end == w[0] and g.text.decode("utf-8") == w[1]
        if w[2] is not None:
            assert (g.segments[-1].start, g.segments[-1].end) == tuple(w[2])
    st.close(); ses.close(); eng.close()


def test_grpc_server_two_concurrent_streams(micro_v3_peaked, audio30):
    """the real engine behind the gRPC transport (proto/asr.proto), two streams at once on one GPU: every stream has its own
    ss_state and CUDA stream; the responses equal those of a session driven directly"""
    import threading
    from speaksense_b200 import WhisperAsr, grpc_server, stream
    eng = WhisperAsr(micro_v3_peaked, device=0)
    pcms = [audio30[:16000 * 11], audio30[16000 * 5:16000 * 17]]
    want = []
    for pcm in pcms:
        ses = stream.AsrStreamSession(eng)
        w = []
        for m, e in stream.encode_messages(pcm):
            w += ses.feed(m, e, "x")
        ses.close()
        want.append([(r.end, r.text, [(s.start, s.end, s.text) for s in r.segments]) for r in w])
    server = grpc_server.serve(eng, "127.0.0.1:0")
    got = [None, None]

    def run(i):
        addr = "127.0.0.1:%d" % server.bound_port
        got[i] = [(r.end, r.text, [(s.start, s.end, s.text) for s in r.segments])
                  for r in grpc_server.transcribe_stream(addr, stream.encode_messages(pcms[i]), "x")]

    th = [threading.Thread(target=run, args=(i,)) for i in range(2)]
    try:
        for t in th:
            t.start()
        for t in th:
            t.join(timeout=120)
    finally:
        server.stop(0)
    assert got[0] == want[0] and got[1] == want[1] and len(want[0]) >= 1
    eng.close()

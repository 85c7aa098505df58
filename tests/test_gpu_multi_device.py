"""One process, several GPUs (the reference is one server process whose context serves every session:
src/main.rs:38-39, src/asr/whisper.rs:17,26): ss_engine_open_multi parses the file once and broadcasts the arena to every
listed device with an in-process ncclBroadcast; states are pinned to a replica.  Needs >= 2 GPUs (`gpurun --gpus 2`);
skips on a one-GPU box."""
import ctypes as C
import threading

import pytest

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


needs2 = pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs in one process")


@needs2
def test_two_devices_one_process_match_oracle(oracle_mod, tiny_en_peaked, audio30):
    from speaksense_b200 import AsrParams, WhisperAsr, _native
    om = oracle_mod.OracleModel(tiny_en_peaked)
    ost = om.new_state()
    ref = ost.full(audio30, stream_mode=True)
    eng = WhisperAsr(tiny_en_peaked, devices=[0, 1])
    assert eng.devices == [0, 1]
    # the broadcast arena is bit-identical on both devices and equals the host packer's image
    fnv = C.c_uint64()
    _native.check(_native.lib().ss_model_probe(tiny_en_peaked.encode(), None, None, C.byref(fnv), None, None, None))
    assert eng.arena_fnv1a(0) == eng.arena_fnv1a(1) == fnv.value
    s0 = eng.create_state(device=0)
    s1 = eng.create_state(device=1)
    s2 = eng.create_state()            # least loaded: both have one state, ties go to the first replica
    s3 = eng.create_state()
    assert (s0.device, s1.device, s2.device, s3.device) == (0, 1, 0, 1)
    out = {}

    def work(name, st):
        eng.transcribe_with_state(st, audio30, AsrParams(stream_mode=True))
        out[name] = st.result_tokens()[0]

    th = [threading.Thread(target=work, args=(n, s)) for n, s in (("a", s0), ("b", s1), ("c", s2), ("d", s3))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for n in "abcd":
        assert out[n] == ref["tokens"], n
    # one batch call whose states live on both devices: grouped by device, the groups run concurrently
    sts = [eng.create_state(device=i % 2) for i in range(8)]
    eng.transcribe_batch(sts, [audio30] * 8, AsrParams(stream_mode=True))
    for st in sts:
        assert st.result_tokens()[0] == ref["tokens"]
    with pytest.raises(Exception):
        eng.create_state(device=5)
    for st in sts + [s0, s1, s2, s3]:
        st.close()
    eng.close(); ost.close(); om.close()


def test_second_engine_on_same_device_and_statics(tiny_en_peaked, micro_v3_peaked, audio30):
    """two engines in one process (different models): the per-device kernel attributes / tables are set up once per device,
    not once per model, and neither engine disturbs the other"""
    from speaksense_b200 import AsrParams, WhisperAsr
    a = WhisperAsr(tiny_en_peaked)
    b = WhisperAsr(micro_v3_peaked)
    ra = a.transcribe(audio30, AsrParams(stream_mode=True))
    rb = b.transcribe(audio30, AsrParams(language="zh", stream_mode=True))
    ra2 = a.transcribe(audio30, AsrParams(stream_mode=True))
    assert ra.full_text == ra2.full_text and ra.full_text and rb.full_text
    a.close(); b.close()

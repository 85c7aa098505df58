"""ggml block-quantised model files (q4_0, q4_1, q5_0, q5_1, q8_0 - what whisper.cpp's quantize tool
writes and script/download-ggml-model.sh fetches).  Both loaders dequantise at load; the check is bit-exactness against
an f16 'twin' file that holds numpy's dequantize(quantize(w)) - i.e. the C / C++ dequantisers against the ggml block
formats restated in numpy (speaksense_b200/synth.py)."""
import os

import numpy as np
import pytest

from tests.conftest import MODEL_DIR

QT = ["q4_0", "q4_1", "q5_0", "q5_1", "q8_0"]


def quant_pair(src, qtype):
    from speaksense_b200 import synth
    base = os.path.basename(src)[:-4]
    q = os.path.join(MODEL_DIR, "%s-%s.bin" % (base, qtype))
    twin = os.path.join(MODEL_DIR, "%s-%s-twin-f16.bin" % (base, qtype))
    if not (os.path.exists(q) and os.path.exists(twin)):
        synth.quantize_model(src, q, qtype, twin)
    return q, twin


@pytest.mark.parametrize("qtype", QT)
def test_block_formats_roundtrip(qtype):
    from speaksense_b200 import synth
    rng = np.random.default_rng(7)
    x = (rng.standard_normal(32 * 257) * 0.03).astype(np.float32)
    raw = synth.quantize_blocks(x, qtype)
    assert len(raw) == 257 * synth.QTYPES[qtype][2]
    y = synth.dequantize_blocks(raw, qtype, x.size)
    bound = {"q4_0": 0.15, "q4_1": 0.15, "q5_0": 0.08, "q5_1": 0.08, "q8_0": 0.01}[qtype]
    assert np.sqrt(((x - y) ** 2).mean()) <= bound * x.std()
    # quantising the dequantised values again reproduces the same codes (fixed point of the format)
    assert synth.dequantize_blocks(synth.quantize_blocks(y, qtype), qtype, x.size).tobytes() == y.tobytes() or qtype in ("q4_1", "q5_1")


@pytest.mark.parametrize("qtype", QT)
def test_oracle_dequantises_bit_exactly(oracle_mod, micro_v3_peaked, audio30, qtype):
    q, twin = quant_pair(micro_v3_peaked, qtype)
    outs = []
    for path in (q, twin):
        m = oracle_mod.OracleModel(path)
        st = m.new_state()
        r = st.full(audio30[:16000 * 10], language="en", stream_mode=True, keep_logits=True)
        outs.append((r["tokens"], st.kept_logits().copy()))
        st.close(); m.close()
    assert outs[0][0] == outs[1][0] and len(outs[0][0]) > 4
    assert np.array_equal(outs[0][1], outs[1][1])


@pytest.mark.gpu
@pytest.mark.parametrize("qtype", QT)
def test_gpu_loader_dequantises_bit_exactly_and_matches_oracle(oracle_mod, micro_v3_peaked, audio30, qtype):
    from speaksense_b200 import AsrParams, WhisperAsr
    q, twin = quant_pair(micro_v3_peaked, qtype)
    pcm = audio30[:16000 * 10]
    res = []
    for path in (q, twin):
        eng = WhisperAsr(path, device=0)
        st = eng.create_state()
        eng.transcribe_with_state(st, pcm, AsrParams(language="en", stream_mode=True, debug_keep_logits=True))
        toks, _ = st.result_tokens()
        res.append((toks, st.debug_logits().copy()))
        st.close(); eng.close()
    assert res[0][0] == res[1][0] and np.array_equal(res[0][1], res[1][1])
    m = oracle_mod.OracleModel(q)
    ost = m.new_state()
    ref = ost.full(pcm, language="en", stream_mode=True, keep_logits=True)
    assert res[0][0] == ref["tokens"]
    assert float(np.abs(res[0][1] - ost.kept_logits()).max()) < 1e-2
    ost.close(); m.close()


@pytest.mark.parametrize("qtype", QT)
def test_cxx_loader_arena_is_bit_identical_to_the_f16_twin(micro_v3_peaked, qtype):
    """CPU-only check of csrc/model.cc's dequantiser: the packed weight arena (what would be uploaded to HBM) of the
    quantised file hashes to the same FNV-1a value as the arena of the numpy-dequantised f16 twin... except for the raw
    file prefix both arenas embed (hparams.ftype differs), so the hash is taken over the tensors only via identical
    metadata: the twin is written with the same ftype field for this comparison."""
    import ctypes as C
    import struct
    from speaksense_b200 import _native, build
    build.build()
    L = _native.lib()
    q, twin = quant_pair(micro_v3_peaked, qtype)
    # give the twin the quantised file's ftype so that the embedded prefix is byte-identical
    raw = bytearray(open(twin, "rb").read())
    ftype_q = struct.unpack_from("<i", open(q, "rb").read(48), 4 + 40)[0]
    struct.pack_into("<i", raw, 4 + 40, ftype_q)
    twin2 = twin[:-4] + "-ftype.bin"
    open(twin2, "wb").write(raw)
    hs = []
    for path in (q, twin2):
        h, ab = C.c_uint64(), C.c_int64()
        assert L.ss_model_probe(path.encode(), None, C.byref(ab), C.byref(h), None, None, None) == 0, L.ss_last_error()
        hs.append((h.value, ab.value))
    assert hs[0] == hs[1]
    os.remove(twin2)

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


# ---- pin to ggml-owned arithmetic: the ggml project's own Python package (gguf.quants, shipped in this image) ----
GGUF_T = {"q4_0": "Q4_0", "q4_1": "Q4_1", "q5_0": "Q5_0", "q5_1": "Q5_1", "q8_0": "Q8_0"}


@pytest.mark.parametrize("qtype", QT)
def test_block_formats_match_ggml_python_package(qtype):
    """synth.quantize_blocks / dequantize_blocks (the specification csrc/model.cc and the oracle are held to above) against
    gguf.quants - written by the ggml authors, the only ggml-owned code available offline: the raw block BYTES of the
    quantiser and the f32 values of the dequantiser must be identical."""
    gguf = pytest.importorskip("gguf")
    from gguf import quants
    from speaksense_b200 import synth
    qt = getattr(gguf.GGMLQuantizationType, GGUF_T[qtype])
    rng = np.random.default_rng(11)
    for scale in (0.02, 1.0, 37.5):
        x = (rng.standard_normal((64, 96)) * scale).astype(np.float32)
        x[0, :32] = 0.0                       # an all-zero block (d == 0)
        x[1, 5] = np.float32(scale * 9.0)     # an outlier that sets the block scale
        theirs = quants.quantize(x, qt)
        ours = np.frombuffer(synth.quantize_blocks(x.reshape(-1), qtype), np.uint8).reshape(theirs.shape)
        assert np.array_equal(ours, theirs), "%s quantiser differs from gguf.quants in %d bytes" % (qtype, int((ours != theirs).sum()))
        deq_theirs = quants.dequantize(theirs, qt).astype(np.float32)
        deq_ours = synth.dequantize_blocks(theirs.tobytes(), qtype, x.size).reshape(x.shape)
        assert np.array_equal(deq_ours, deq_theirs)


@pytest.mark.parametrize("qtype", QT)
def test_cxx_and_oracle_loaders_match_ggml_python_package(oracle_mod, micro_v3_peaked, qtype):
    """End to end through a model FILE: the tensors of a quantised file, dequantised by gguf.quants and written as an f16 twin,
    give the C++ loader the same arena hash and the oracle the same logits as the quantised file itself."""
    import ctypes as C
    import struct
    gguf = pytest.importorskip("gguf")
    from gguf import quants
    from speaksense_b200 import _native, build, synth
    build.build()
    q, _ = quant_pair(micro_v3_peaked, qtype)
    qt = getattr(gguf.GGMLQuantizationType, GGUF_T[qtype])
    ttype, _, bs = synth.QTYPES[qtype]
    twin = os.path.join(MODEL_DIR, os.path.basename(q)[:-4] + "-gguf-twin.bin")
    synth.write_twin_from_raw(q, twin, lambda raw, shape: quants.dequantize(
        np.frombuffer(raw, np.uint8).reshape(-1, (shape[-1] // 32) * bs), qt).astype(np.float32).reshape(shape))
    hs = []
    for path in (q, twin):
        h, ab = C.c_uint64(), C.c_int64()
        assert _native.lib().ss_model_probe(path.encode(), None, C.byref(ab), C.byref(h), None, None, None) == 0
        hs.append((h.value, ab.value))
    assert hs[0] == hs[1]
    toks = [int(t) for t in np.random.default_rng(3).integers(0, 50000, size=6)]
    outs = []
    pcm = synth.synth_audio(16000 * 3, seed=2)
    for path in (q, twin):
        om = oracle_mod.OracleModel(path)
        st = om.new_state()
        mel, _, _ = om.log_mel(pcm)
        st.encode(mel, 0)
        outs.append(st.decode(toks, 0).copy())
        st.close(); om.close()
    assert np.array_equal(outs[0], outs[1])
    os.remove(twin)
